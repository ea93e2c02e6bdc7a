"""Dense Cholesky micro-benchmark / check (GPU box)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dbat_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(0)
M = rng.standard_normal((n, n + 16))
A = M @ M.T + n * np.eye(n)
b = rng.standard_normal(n)
x, Ai, ms = _lib.dense_chol_solve(A, b, want_inverse=(n <= 4096), repeat=int(sys.argv[2]) if len(sys.argv) > 2 else 3)
xr = np.linalg.solve(A, b)
print('n=%d  factor+solve %.3f ms  (%.2f TFLOP/s)  x rel err %.2e' % (n, ms, n ** 3 / 3 / ms / 1e9, np.abs(x - xr).max() / np.abs(xr).max()))
if Ai is not None:
    print('inverse rel err %.2e' % (np.abs(Ai - np.linalg.inv(A)).max() / np.abs(np.linalg.inv(A)).max()))
