#!/bin/bash
# First GPU call of the next round: everything this round could only verify on CPU (last table).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_gpu_run.sh'
# Output goes to gpurun_out/ (merged back); each step has its own timeout so a hang cannot take the box.
mkdir -p gpurun_out
python - <<'PY'
import __graft_entry__ as g
g.build()
PY
# 1. the device twins that have never run (they sort after the verified parity file)
timeout 900 python -m pytest tests/test_report_golden.py tests/test_script.py tests/test_ingest.py -m gpu -q \
    > gpurun_out/first_twins.log 2>&1
tail -15 gpurun_out/first_twins.log
# 2. the verified parity suite, to see nothing moved (post-mortem try/except, CIOF from the CIO block)
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/first_parity.log 2>&1
tail -5 gpurun_out/first_parity.log
# 3. the reference's own mid-size benchmark next to MATLAB's published 7.28 s
timeout 600 python tools/roma_bench.py > gpurun_out/roma_bench.json 2> gpurun_out/roma_bench.err
cat gpurun_out/roma_bench.json
# 4. the round benchmark
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
