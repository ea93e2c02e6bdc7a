"""The point-sharded multi-GPU path against the single-GPU path, on real devices (NCCL, one rank per GPU).
Needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('nImg,nOP', [(40, 3000), (420, 12000)])
def test_sharded_solvers_match_single_gpu(built_lib, nImg, nOP):
    """Steps, f, |Jp|^2 and complete LM / GNA / LMP runs of a 2-rank point-sharded problem equal the
    single-GPU run (1e-9; identical iteration counts).  The larger case has a dissected, tiled reduced system."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.join(ROOT, 'tests', 'mg_worker.py'),
           str(nImg), str(nOP)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(out.stdout[-4000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
