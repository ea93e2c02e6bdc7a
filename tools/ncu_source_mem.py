"""Per-instruction memory picture of one kernel from an ncu capture taken with --import-source on (SASS view):
    python tools/ncu_source_mem.py gpurun_out/r2g_kernels.ncu-rep k_point_side_obs_c
Lists the memory instructions with their L1 tag requests (global) / wavefronts (shared) and the stall samples."""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))      # first launch only
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:end]))))
tot_samples = sum(int(r['# Samples'] or 0) for r in rows)
tg = sum(int(r['L1 Tag Requests Global'] or 0) for r in rows)
ts = sum(int(r['L1 Wavefronts Shared'] or 0) for r in rows)
print('instructions', len(rows), 'samples', tot_samples, 'L1 tag requests global', tg, 'L1 wavefronts shared', ts)
print('%-6s %-58s %9s %10s %10s %8s %8s' % ('idx', 'sass', 'execs', 'tag_glob', 'wf_shared', 'samples', 'excess'))
for i, r in enumerate(rows):
    g, s2 = int(r['L1 Tag Requests Global'] or 0), int(r['L1 Wavefronts Shared'] or 0)
    smp = int(r['# Samples'] or 0)
    if g or s2 or smp > 0.01 * tot_samples:
        print('%-6d %-58s %9s %10d %10d %8d %8s' % (i, r['Source'].strip()[:58], r['Instructions Executed'], g, s2, smp,
                                                    r['L1 Wavefronts Shared Excessive']))
