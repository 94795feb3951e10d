"""Known-pattern probes of the tcgen05 forward layer (all-ones, identity and column-selecting weights) through
tests/test_gemm_gpu.run_gemm: prints what lands where when an operand layout or descriptor is suspected."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gemm_gpu import run_gemm
np.set_printoptions(linewidth=200, precision=3, suppress=True)
P, K, N = 128, 32, 32
# case 1: all ones
Z, st, _ = run_gemm(np.ones((P, K), np.float32), np.ones((N, K), np.float32), K)
print("ones: unique", np.unique(Z)[:10], Z.shape)
# case 2: A = one-hot column j per row pattern; W = identity-ish -> Z[r, n] = A[r, n]
A = np.zeros((P, K), np.float32); 
for r in range(P): A[r, :] = np.arange(K) + 100 * r
W = np.eye(N, K, dtype=np.float32)
Z, st, _ = run_gemm(A, W, K)
print("identity W: row0", Z[0]); print("row1", Z[1]); print("row9", Z[9]); print("row 127", Z[127])
print("max err", np.abs(Z - A[:, :N]).max())
# case 3: W selects: W[n, k] = 1 if k == 0 -> Z[r, n] = A[r, 0]
W = np.zeros((N, K), np.float32); W[:, 0] = 1
Z, st, _ = run_gemm(A, W, K)
print("W col0: row0", Z[0][:8], "row5", Z[5][:8])
W = np.zeros((N, K), np.float32); W[:, 9] = 1
Z, st, _ = run_gemm(A, W, K)
print("W col9: row0", Z[0][:8], "row5", Z[5][:8])
