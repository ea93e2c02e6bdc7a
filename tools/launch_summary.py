"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total, average, share."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    try:
        name, v = row['Kernel Name'], float(row['Metric Value'].replace(',', ''))
    except Exception:
        continue
    unit = row.get('Metric Unit', '')
    v = v / 1000 if unit == 'ns' else v * 1000 if unit == 'ms' else v
    agg[re.sub(r'\(.*', '', name)][0] += 1
    agg[re.sub(r'\(.*', '', name)][1] += v
skip = ('cutlass', 'at::', 'cub::', 'k_gen_pairs')          # DGEMM yardstick, torch RNG, one-off index build
tot = sum(v[1] for k, v in agg.items() if not k.startswith(tuple('void ' + s for s in skip)) and not k.startswith(skip))
print('kernel,launches,total_us,avg_us,share_of_library_kernels')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lib = not (k.startswith(tuple('void ' + s for s in skip)) or k.startswith(skip))
    print('%s,%d,%.1f,%.2f,%s' % (k.replace(',', ';'), v[0], v[1], v[1] / v[0], '%.1f%%' % (100 * v[1] / tot) if lib else 'n/a'))
