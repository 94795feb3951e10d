"""dH = dZ * W (backward data GEMM): plain tcgen05 layer vs cuBLAS TF32, SSG shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import fused, lib
B = 32
SHAPES = [("SA1.l2", B*1024*32, 64, 32), ("SA1.l1", B*1024*32, 32, 32), ("SA2.l2", B*256*32, 128, 64),
          ("SA2.l1", B*256*32, 64, 64), ("FP0.l1", B*8192, 128, 128), ("FP0.l0", B*8192, 128, 136),
          ("head.l1", B*8192, 20, 128), ("SA3.l2", B*64*32, 256, 128)]
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps
for name, P, cout, cin in SHAPES:
    dz = torch.randn(P, cout, device="cuda")
    W = torch.randn(cout, cin, device="cuda")
    Wt = W.t().contiguous()                       # [cin, cout] = the [N, K] operand
    t_tc = timed(lambda: fused._tc_layer(dz, cout, Wt, None, False, False, False))
    t_cb = timed(lambda: fused._mm(dz, W))
    byt = 4.0 * P * (cout + cin)
    print("%-8s P=%8d K=%3d N=%3d | tcgen05 %7.1f us %5.0f GB/s | cuBLAS %7.1f us %5.0f GB/s" % (
        name, P, cout, cin, t_tc, byt / t_tc / 1e3, t_cb, byt / t_cb / 1e3))
