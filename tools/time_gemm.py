"""Per-layer timing of the tcgen05 shared-MLP layer vs cuBLAS + streaming kernels (SSG shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import fused, lib  # noqa: E402

L = lib.load()
B = 32
SHAPES = [  # (name, P, K(lda), N)
    ("SA1.l0", B * 1024 * 32, 12, 32), ("SA1.l1", B * 1024 * 32, 32, 32), ("SA1.l2", B * 1024 * 32, 32, 64),
    ("SA2.l0", B * 256 * 32, 68, 64), ("SA2.l1", B * 256 * 32, 64, 64), ("SA2.l2", B * 256 * 32, 64, 128),
    ("SA3.l2", B * 64 * 32, 128, 256), ("SA4.l2", B * 16 * 32, 256, 512),
    ("FP0.l0", B * 8192, 136, 128), ("FP0.l1", B * 8192, 128, 128), ("head.l1", B * 8192, 128, 20),
    ("FP1.l0", B * 1024, 320, 256),
]
only = sys.argv[1:] 


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


for name, P, K, N in SHAPES:
    if only and name not in only:
        continue
    A = torch.randn(P, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    a_in = torch.rand(K, device="cuda") + 0.5
    b_in = torch.randn(K, device="cuda")
    gamma, beta = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")

    def tc():
        z, h, part, (ld, nblk) = fused._tc_layer(A, K, W, (a_in, b_in), True, True, True)
        fused._bn_finalize(P, N, part, ld, nblk, gamma, beta, None, None, 1e-5, 0.1)

    def tc_noh():
        z, h, part, (ld, nblk) = fused._tc_layer(A, K, W, (a_in, b_in), True, False, True)

    def cublas():
        h = torch.empty_like(A)
        lib.check(L.pn2_bn_relu_apply(P, K, A.data_ptr(), a_in.data_ptr(), b_in.data_ptr(), h.data_ptr(), 1,
                                      torch.cuda.current_stream().cuda_stream), "apply")
        z = torch.mm(h, W.t())   # library GEMM, for comparison only
        fused._bn_forward(z, gamma, beta, None, None, 1e-5, 0.1, True)

    t_tc, t_noh, t_cb = timed(tc), timed(tc_noh), timed(cublas)
    byt = 4.0 * P * (K + K + N)          # read A, write H, write Z
    byt_noh = 4.0 * P * (K + N)
    print("%-8s P=%8d K=%3d N=%3d | tcgen05+H %7.1f us %6.0f GB/s | tcgen05 %7.1f us %6.0f GB/s | cuBLAS path %7.1f us | %.2f GFLOP"
          % (name, P, K, N, t_tc, byt / t_tc / 1e3, t_noh, byt_noh / t_noh / 1e3, t_cb, 2e-9 * P * K * N))
