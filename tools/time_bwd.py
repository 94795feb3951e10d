"""CUDA-event timing of the fused tensor-core backward layer (csrc/mlp_bwd.cu) at the SSG shapes.
Algorithmic bytes = read dH (or dOut/arg-max), Z, Z_prev; write dH_prev (when wanted)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import lib
L = lib.load()
DEV = "cuda:0"
PEAK = 6536.4
CASES = [  # name, P, N, K, lda, ns (pooled), fold, dgrad
    ("SA1 l3 pooled", 1 << 20, 64, 32, 32, 32, True, True),
    ("SA1 l2", 1 << 20, 32, 32, 32, 0, True, True),
    ("SA1 l1 (dW only)", 1 << 20, 32, 9, 12, 0, False, False),
    ("SA2 l3 pooled", 1 << 18, 128, 64, 64, 32, True, True),
    ("SA2 l2", 1 << 18, 64, 64, 64, 0, True, True),
    ("SA2 l1", 1 << 18, 64, 67, 68, 0, False, True),
    ("FP0 l2", 1 << 18, 128, 128, 128, 0, True, True),
    ("FP0 l1", 1 << 18, 128, 134, 136, 0, False, True),
    ("head l2", 1 << 18, 20, 128, 128, 0, False, True),
    ("SA3 l2", 1 << 16, 128, 128, 128, 0, True, True),
]
s = torch.cuda.current_stream().cuda_stream
sched = torch.zeros(2, dtype=torch.int32, device=DEV)
print("PN2_BWD_DBG =", os.environ.get("PN2_BWD_DBG"))
ONLY = os.environ.get("ONLY")
for name, P, N, K, lda, ns, fold, dgrad in CASES:
    if ONLY and ONLY not in name:
        continue
    g = torch.Generator(device=DEV).manual_seed(0)
    r = lambda *sh: torch.randn(*sh, device=DEV, generator=g)
    if os.environ.get("CARVE"):   # the four big tensors as consecutive pieces of ONE allocation (as the step's allocator places them)
        big = torch.empty(P * (3 * N + 2 * lda) + 1024, device=DEV)
        def carve(off, rows, cols):
            return big[off:off + rows * cols].view(rows, cols)
        z = carve(0, P, N); z.normal_(generator=g)
        aprev = carve(P * N, P, lda); aprev.normal_(generator=g)
    else:
        z, aprev = r(P, N), r(P, lda)
    a, b, mean, invstd, c1c2 = r(N), r(N), r(N), r(N).abs() + .5, 0.01 * r(2 * N)
    a_in, b_in, mean_in, invstd_in = (r(K), r(K), r(K), r(K).abs() + .5) if fold else (None,) * 4
    if os.environ.get("CARVE") and not ns:
        dh = carve(P * (N + lda), P, N); dh.normal_(generator=g)
    else:
        dh = r(P, N) if not ns else None
    dout = r(P // ns, N) if ns else None
    am = torch.randint(0, ns, (P // ns, N), device=DEV, dtype=torch.int32) if ns else None
    W = r(N, K)
    nb, kb, kpad, grid = L.pn2_mlp_bwd_nb(N), L.pn2_mlp_bwd_kb(lda), L.pn2_mlp_bwd_kpad(lda), L.pn2_mlp_bwd_grid(P)
    wtp = torch.empty(nb * kpad * 32, device=DEV) if dgrad else None
    if dgrad:
        lib.check(L.pn2_mlp_pack_weights_t(N, K, lda, K, W.data_ptr(), wtp.data_ptr(), s), "pack")
    dhp = (carve(P * (2 * N + lda), P, lda) if os.environ.get("CARVE") else torch.empty(P, lda, device=DEV)) if dgrad else None
    dwp = torch.empty(grid, nb * 32 * kb * 32, device=DEV)
    sp = torch.empty(grid, nb * 32 * kb * 32, device=DEV) if (dgrad and fold) else None
    p = lambda t: t.data_ptr() if t is not None else None
    def call():
        lib.check(L.pn2_mlp_bwd_layer(P, N, K, lda, ns, 1, 1, p(dh), p(dout), p(am), p(z), p(a), p(b), p(mean),
                                      p(invstd), p(c1c2), p(W), p(aprev), p(a_in), p(b_in), p(mean_in), p(invstd_in),
                                      p(wtp), p(dhp), p(dwp), p(sp), sched.data_ptr(), None, 1.0, None, 1.0, s), "bwd")
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    REPS = int(os.environ.get("REPS", "10"))
    for _ in range(REPS):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / REPS
    byt = 4 * P * (N + lda + (lda if dgrad else 0) + (N if not ns else 0)) + (8 * (P // ns) * N if ns else 0)
    if os.environ.get("PN2_BWD_DBG") and int(os.environ["PN2_BWD_DBG"]) & 32:
        import ctypes, numpy as np
        buf = np.zeros((148, 20), np.uint64)
        L.pn2_mlp_bwd_debug_read.argtypes = [ctypes.c_void_p]
        L.pn2_mlp_bwd_debug_read(buf.ctypes.data)
        m = buf.astype(np.float64).mean(0) / 1.9e3   # us at ~1.9 GHz
        print("   us/CTA: G0 total %.0f w_empty %.0f dz_empty %.0f h_empty %.0f loads %.0f | G1 total %.0f w %.0f dz %.0f h %.0f ld %.0f"
              " | MMA acc_empty %.0f dz_full %.0f w_full %.0f h_full %.0f | EPI acc_full %.0f | read-out %.1f set-up %.1f entry->mma done %.1f" % tuple(m[:18]))
    print("%-18s P=%7d N=%3d K=%3d  %7.1f us  %6.1f MB  %6.0f GB/s (%.0f%% of HBM peak)" % (
        name, P, N, K, us, byt / 1e6, byt / us / 1e3, 100 * byt / us / 1e3 / PEAK))
