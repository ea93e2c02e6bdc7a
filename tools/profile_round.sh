#!/bin/bash
# Profiles of the round's final build (B200_PROFILING.md recipe), written to gpurun_out/:
#   launches_r2g.csv   launch list of `bench.py --steps 2 --warmup 1` (per-launch durations, cold cache, serialised)
#   r2g_kernels.ncu-rep  one `--set full` capture of every kernel family of the step
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/profile_round.sh'
mkdir -p gpurun_out
python - <<'PY'
import __graft_entry__ as g
g.build()
PY
# ncu serialises kernels: the factorisation must then be ONE launch (DBAT_TC_CHAIN_CTAS=0), not the chain launch + the
# bulk launch that waits on it programmatically - the chain CTAs would spin on tiles of a launch ncu holds back
export DBAT_TC_CHAIN_CTAS=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2g.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_r2g.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none \
    -k regex:"k_cam_side_c|k_point_side_obs_c|k_schur_win|k_schur_reduce|k_tchol_factor|k_tchol_bwd|k_backsub_obs|k_resid|k_build_S|k_point_minv" \
    --launch-skip 40 -c 14 -o gpurun_out/r2g_kernels python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2g.log 2>&1
tail -2 gpurun_out/ncu_r2g.log | cut -c1-300
unset DBAT_TC_CHAIN_CTAS
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2g_n1.json 2> gpurun_out/bench_r2g_n1.err
cut -c1-400 gpurun_out/bench_r2g_n1.json
