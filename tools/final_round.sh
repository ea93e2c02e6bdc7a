#!/bin/bash
# Last GPU call of a round: every -m gpu test, the bench line (with the CPU baseline), and the ncu launch list of the
# same build.   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/final_round.sh'
mkdir -p gpurun_out
python - <<'PY'
import __graft_entry__ as g
g.build()
PY
timeout 400 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -14 gpurun_out/gpu_tests.log
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
cut -c1-330 gpurun_out/bench_final_n1.json
# ncu serialises kernels: the factorisation as ONE launch (see tools/profile_round.sh)
DBAT_TC_CHAIN_CTAS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_final.log 2>&1
