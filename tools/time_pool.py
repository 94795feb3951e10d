"""Pooled last layer of the four SSG SA blocks: tensor-core layer + group max, two ways, L2 flushed
before every timed launch sequence: (a) layer with the group extrema in its epilogue + pn2_groupmax_from_extrema,
(b) plain layer + pn2_bn_relu_groupmax (a second pass over Z)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import fused, lib
L = lib.load()
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)
s = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, P, K, N in [("SA1.l2", 1 << 20, 32, 64), ("SA2.l2", 1 << 18, 64, 128), ("SA3.l2", 1 << 16, 128, 256),
                      ("SA4.l2", 1 << 14, 256, 512)]:
    ns = 32
    A = torch.randn(P, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    a_in, b_in = torch.rand(K, device=dev) + 0.5, torch.randn(K, device=dev)
    a, b = torch.randn(N, device=dev), torch.randn(N, device=dev)
    G = P // ns
    out, arg = torch.empty(G, N, device=dev), torch.empty(G, N, dtype=torch.int32, device=dev)

    def fusedpool():
        z, h, part, geo, ext = fused._tc_layer(A, K, W, (a_in, b_in), True, False, True, None, ns)
        lib.check(L.pn2_groupmax_from_extrema(G, N, ext[0].data_ptr(), ext[1].data_ptr(), ext[2].data_ptr(),
                                              ext[3].data_ptr(), a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                              arg.data_ptr(), None, s), "x")

    def twopass():
        z, h, part, geo = fused._tc_layer(A, K, W, (a_in, b_in), True, False, True)
        lib.check(L.pn2_bn_relu_groupmax(G, ns, N, z.data_ptr(), a.data_ptr(), b.data_ptr(), out.data_ptr(),
                                         arg.data_ptr(), s), "y")

    def layer_only():
        fused._tc_layer(A, K, W, (a_in, b_in), True, False, True)

    print("%-7s P=%8d K=%3d N=%3d | layer alone %6.1f us | extrema in epilogue + finish %6.1f us | layer + groupmax pass %6.1f us"
          % (name, P, K, N, timed(layer_only), timed(fusedpool), timed(twopass)))
