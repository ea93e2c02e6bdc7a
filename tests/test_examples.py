"""The example scripts build their problem on the host and then call the device.  Without a CUDA device
they must stop at the first numerical step with the library's own error - there is no CPU fallback to
fall into - which also exercises all of their host code."""
import os
import runpy
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_cuda():
    import torch
    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_cuda(), reason='a CUDA device is present: the examples would run to the end')
@pytest.mark.parametrize('script,argv', [('camcaldemo.py', []), ('prague2016_pm.py', ['s3']), ('ps_postproc.py', [])])
def test_examples_reach_the_device_call_and_fail_loudly_without_one(script, argv, monkeypatch, tmp_path):
    from dbat_b200._lib import DbatError
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, 'argv', [script] + argv)
    with pytest.raises(DbatError) as err:
        runpy.run_path(os.path.join(ROOT, 'examples', script), run_name='__main__')
    assert 'CUDA' in str(err.value)


@pytest.mark.skipif(not _no_cuda(), reason='a CUDA device is present')
def test_script_runner_fails_loudly_without_a_device():
    from dbat_b200 import rundbatscript
    from dbat_b200._lib import DbatError
    with pytest.raises(DbatError):
        rundbatscript(os.path.join(ROOT, 'tests', 'golden', 'camcaldemo', 'camcaldemo.xml'), write=False)
