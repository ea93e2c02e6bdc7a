#!/bin/bash
# Full GPU verification: every -m gpu test (no -x, so one failure cannot hide the rest), then the bench.
#   /usr/local/graft/bin/gpurun --timeout 1700 -- 'bash tools/gpu_full.sh'
mkdir -p gpurun_out
python - <<'PY'
import __graft_entry__ as g
g.build()
PY
timeout 1300 python -m pytest tests -m gpu -q --durations=15 -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -40 gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
