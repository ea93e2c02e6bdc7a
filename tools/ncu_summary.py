"""Key metrics of an `ncu --set full` capture as one CSV row per kernel, plus the DRAM traffic per launch:
    python tools/ncu_summary.py gpurun_out/r2g_kernels.ncu-rep profiles/ncu_r2g_kernels_summary.csv profiles/r2_dram_traffic.json
(reads the report with `ncu -i ... --page raw --csv`; B200_PROFILING.md names the metrics)."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
]


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    have = [m for m in METRICS if m in col]
    seen, lines, traffic = set(), [], {}
    for r in data:
        name = r[col['Kernel Name']]
        if name in seen:
            continue
        seen.add(name)
        lines.append([name] + [r[col[m]] for m in have])
        try:
            rd, wr = float(r[col['dram__bytes_read.sum']].replace(',', '')), float(r[col['dram__bytes_write.sum']].replace(',', ''))
            f = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
            b = rd * f.get(units[col['dram__bytes_read.sum']], 1.0) + wr * f.get(units[col['dram__bytes_write.sum']], 1.0)
            traffic[re.sub(r'^void ', '', re.sub(r'[<(].*', '', name))] = b
        except Exception:
            pass
    with open(out_csv, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['Kernel Name'] + have)
        w.writerow([''] + [units[col[m]] for m in have])
        w.writerows(lines)
    if out_json:
        t = dict(traffic)
        t['k_cam_side+k_point_side'] = t.get('k_cam_side_c', 0) + t.get('k_point_side_obs_c', 0)
        t['k_schur_win+k_schur_reduce'] = t.get('k_schur_win', 0) + t.get('k_schur_reduce', 0) + t.get('k_point_minv', 0) + t.get('k_build_S', 0)
        t['source'] = 'ncu --set full capture %s: dram__bytes_read.sum + dram__bytes_write.sum per launch' % rep
        json.dump(t, open(out_json, 'w'), indent=1)
    for l in lines:
        print(l[0][:60], l[1:9])


if __name__ == '__main__':
    main()
