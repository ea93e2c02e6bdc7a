"""Build a variant of libdbatgpu.so with extra nvcc flags into variants/lib_<tag>.so (tuning aid)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def build(tag, extra):
    src = os.path.join(ROOT, 'dbat_b200', 'csrc')
    os.makedirs(os.path.join(ROOT, 'variants'), exist_ok=True)
    objs, procs = [], []
    for f in ['eval', 'schur', 'schur_index', 'chol', 'api', 'startval']:
        o = '/tmp/var_%s_%s.o' % (tag, f)
        cmd = ['/usr/local/cuda/bin/nvcc', '-Wno-deprecated-gpu-targets'] + extra.split() + [
            '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
            '-c', os.path.join(src, f + '.cu'), '-o', o]
        procs.append(subprocess.Popen(cmd)); objs.append(o)
    for p in procs:
        assert p.wait() == 0
    subprocess.check_call(['/usr/local/cuda/bin/nvcc', '-shared', '-o', os.path.join(ROOT, 'variants', 'lib_%s.so' % tag)] + objs + ['-ldl'])
if __name__ == '__main__':
    build(sys.argv[1], ' '.join(sys.argv[2:]))
