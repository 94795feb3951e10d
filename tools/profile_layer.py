"""One forward layer of the finest FP block -- pn2_mlp_gemm_fwd(262144, 128, 128, 128, 128): previous layer's
BatchNorm fold + ReLU on the A tile, statistics partials out of the epilogue, no H side output -- launched three
times, for an ncu capture of the launch bench.py names in its `roofline` object:

    ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_fwd --launch-skip 2 -c 1 \
        -o gpurun_out/fwd_fp0 python tools/profile_layer.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import fused  # noqa: E402

P, K, N = 32 * 8192, 128, 128
torch.manual_seed(0)
A = torch.randn(P, K, device="cuda")
W = torch.randn(N, K, device="cuda") / K ** 0.5
a_in, b_in = torch.rand(K, device="cuda") + 0.5, torch.randn(K, device="cuda")
for _ in range(3):
    z, h, part, _ = fused._tc_layer(A, K, W, (a_in, b_in), True, False, True)
torch.cuda.synchronize()
print("ok", tuple(z.shape), h is None, tuple(part.shape))
