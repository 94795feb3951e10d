"""GPU parity of the split adjoint of the WIDE shared-MLP layers through the C ABI: the weight-gradient
product on csrc/mlp_wgrad.cu and the data-gradient product on the forward layer kernel with transposed
weight tiles (pn2_mlp_pack_weights_dgrad) -- the layers of the reference's models that do not fit the
one-pass backward kernel: SSG [128,128,256], [256,256,512], FP [..,256,256]; MSG [128,196,256],
[256,384,512] (/root/reference/pointnet2/pointnet2_semseg.py:32-34, 112-114).

Reference: fp64 products of the operands reduced to TF32 the way the kernels read them (activation /
gradient operands: plain fp32 words of which the tensor core uses the upper 19 bits = truncation; packed
weight tiles: rounded to nearest).  What remains is fp32 accumulation order: 2e-5 of the output's max
magnitude for the weight gradient (sum over up to 65536 rows), 2e-5 for the data gradient.
"""
import os

import numpy as np
import pytest
import torch

from pointnet2.scannet_b200 import fused, lib
from test_bwd_gpu import tf32_rna, tf32_trunc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (P, N, K, lda): N output channels, K input channels, lda = padded row width of the input rows
WIDE = [(16384, 256, 128, 128), (4096, 256, 259, 260), (16384, 512, 256, 256), (2048, 256, 768, 768),
        (8192, 196, 128, 128), (4100, 384, 256, 256), (8192, 256, 384, 384), (32768, 256, 320, 320),
        (65536, 256, 128, 128), (1000, 128, 515, 516), (50, 20, 36, 36)]


def wgrad(dz, h, K):
    L = lib.load()
    s = torch.cuda.current_stream().cuda_stream
    P, N = dz.shape
    ldy = h.shape[1]
    S, R, C = L.pn2_mlp_wgrad_splits(P, N, ldy), L.pn2_mlp_wgrad_rows(N), L.pn2_mlp_wgrad_cols(ldy)
    assert S >= 1 and R >= N and C >= ldy
    part = torch.full((S, R, C), float("nan"), dtype=torch.float32, device=DEV)
    lib.check(L.pn2_mlp_wgrad(P, N, N, ldy, dz.data_ptr(), h.data_ptr(), part.data_ptr(), s), "wgrad")
    out = torch.full((N, K), float("nan"), dtype=torch.float32, device=DEV)
    lib.check(L.pn2_sum_partials_block(S, R, C, N, K, part.data_ptr(), out.data_ptr(), s), "sum_block")
    torch.cuda.synchronize()
    assert torch.isfinite(part).all()
    # rows past N / columns past ldy of the images are exact zeros (zero-filled operands)
    assert float(part[:, N:].abs().max()) == 0.0 if R > N else True
    assert float(part[:, :, ldy:].abs().max()) == 0.0 if C > ldy else True
    return out, part


@pytest.mark.parametrize("tma", [True, False])
@pytest.mark.parametrize("P,N,K,lda", WIDE)
def test_wgrad(P, N, K, lda, tma):
    """dW = dZ^T H on the tensor core, operands by TMA box loads in the MN-major swizzle (tma) or copied by
    the loader warps (the register path kept for comparison)."""
    rng = np.random.default_rng(P + N + K)
    dz = rng.standard_normal((P, N)).astype(np.float32)
    h = np.zeros((P, lda), np.float32)
    h[:, :K] = np.maximum(rng.standard_normal((P, K)), 0).astype(np.float32)
    ref = tf32_trunc(dz).astype(np.float64).T @ tf32_trunc(h[:, :K]).astype(np.float64)
    if tma:
        os.environ.pop("PN2_WGRAD_NO_TMA", None)
    else:
        os.environ["PN2_WGRAD_NO_TMA"] = "1"
    try:
        out, part = wgrad(torch.from_numpy(dz).to(DEV), torch.from_numpy(h).to(DEV), K)
    finally:
        os.environ.pop("PN2_WGRAD_NO_TMA", None)
    got = out.double().cpu().numpy()
    err = np.abs(got - ref).max()
    assert err <= 2e-5 * np.abs(ref).max(), (err, np.abs(ref).max())
    # the block sum equals the fp64 sum of the images
    want = part.double().sum(0).cpu().numpy()[:N, :K]
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("P,N,K,lda", WIDE)
def test_dgrad(P, N, K, lda):
    """dH_prev = dZ W on the forward layer kernel with the transposed weight tiles; the padding columns
    K..lda-1 of the rows layout come out as zeros."""
    rng = np.random.default_rng(P + N + K + 1)
    dz = rng.standard_normal((P, N)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    ref = tf32_trunc(dz).astype(np.float64) @ tf32_rna(W).astype(np.float64)
    out = fused._dgrad(torch.from_numpy(dz).to(DEV), torch.from_numpy(W).to(DEV), lda)
    torch.cuda.synchronize()
    got = out.double().cpu().numpy()
    assert got.shape == (P, lda) and np.isfinite(got).all()
    assert np.abs(got[:, K:]).max(initial=0.0) == 0.0
    err = np.abs(got[:, :K] - ref).max()
    assert err <= 2e-5 * np.abs(ref).max(), (err, np.abs(ref).max())


def test_sum_partials_block_matches_flat_sum():
    L = lib.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cpu").manual_seed(0)
    part = torch.randn(37, 64, 96, generator=g).to(DEV)
    flat = torch.empty(64, 96, device=DEV)
    lib.check(L.pn2_sum_partials(37, 64 * 96, part.data_ptr(), flat.data_ptr(), s), "sum")
    for rows, cols in ((64, 96), (20, 96), (64, 67), (33, 9), (1, 1)):
        out = torch.full((rows, cols), float("nan"), device=DEV)
        lib.check(L.pn2_sum_partials_block(37, 64, 96, rows, cols, part.data_ptr(), out.data_ptr(), s), "sum_block")
        torch.cuda.synchronize()
        assert torch.equal(out, flat[:rows, :cols])      # same order of additions: bit-identical


def run_mlp(units, x, ns, tc_bwd, dgrad_stats=False):
    import copy
    units = [copy.deepcopy(u) for u in units]
    xi = x.clone().requires_grad_(True)
    fused.TC_BWD = tc_bwd
    saved, fused.DGRAD_STATS = fused.DGRAD_STATS, dgrad_stats
    try:
        y = fused.run_mlp_rows(units, xi, ns)
        torch.manual_seed(1)
        (y * torch.randn_like(y)).sum().backward()
    finally:
        fused.TC_BWD = True
        fused.DGRAD_STATS = saved
    grads = [q.grad.clone() for u in units for q in u.parameters()]
    return y.detach(), xi.grad.clone(), grads


@pytest.mark.parametrize("P,ns,widths", [(8192, 32, [68, 64, 64, 128]), (4096, 0, [136, 128, 128]),
                                         (2048, 16, [12, 32, 32, 64])])
def test_split_adjoint_equals_one_pass_adjoint(P, ns, widths):
    """The same chain of narrow layers through both adjoints: the one-pass fused backward kernel (TC_BWD) and
    the split adjoint (streaming dZ + data-gradient + weight-gradient products).  Same TF32 operand
    reduction on both sides up to the fused kernel's dW = g1 Q + g0 S reformulation: 2e-3 in the L2 norm."""
    from pointnet2.scannet_b200 import pytorch_utils as pt_utils
    torch.manual_seed(0)
    units = [pt_utils.Conv2d(widths[i], widths[i + 1], bn=True).to(DEV).train() for i in range(len(widths) - 1)]
    for u in units:   # exercise negative BatchNorm scales
        u.bn.bn.weight.data.uniform_(-1.5, 1.5)
    x = torch.randn(P, widths[0], device=DEV)
    y1, gx1, g1 = run_mlp(units, x, ns, True)
    y2, gx2, g2 = run_mlp(units, x, ns, False)
    assert torch.equal(y1, y2)

    def l2(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-20))

    assert l2(gx1, gx2) < 2e-3, l2(gx1, gx2)
    for a, b in zip(g1, g2):
        assert l2(a, b) < 2e-3, l2(a, b)
    # ... and the split adjoint with the statistics of each layer below taken from the data-gradient epilogue
    y3, gx3, g3 = run_mlp(units, x, ns, False, dgrad_stats=True)
    assert torch.equal(y1, y3) and l2(gx3, gx2) < 1e-4
    for a, b in zip(g3, g2):
        assert l2(a, b) < 1e-4, l2(a, b)


@pytest.mark.parametrize("P,N,K,relu", [(16384, 256, 256, True), (4100, 512, 256, True), (8192, 256, 384, True),
                                        (3000, 128, 96, False), (2048, 256, 768, True), (1000, 196, 128, True)])
def test_dgrad_epilogue_statistics(P, N, K, relu):
    """pn2_mlp_dgrad_stats: the data gradient dH_prev = dZ W AND, from its epilogue, the BatchNorm-adjoint column
    sums of the layer below -- sum_rows dY_prev and sum_rows dY_prev * zhat_prev with dY_prev = dH_prev where that
    layer's ReLU passed -- finalised by pn2_bwd_stats_finalize: against fp64 sums over the kernel's own dH_prev
    (1e-5 of the largest column sum) and dgamma / dbeta / c1c2 consistency."""
    L = lib.load()
    s = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(P + N + K + 7)
    dz = rng.standard_normal((P, N)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    zp = rng.standard_normal((P, K)).astype(np.float32)
    mean, var = zp.mean(0), zp.var(0)
    invstd = (1.0 / np.sqrt(var + 1e-5)).astype(np.float32)
    gamma = rng.uniform(0.5, 1.5, K).astype(np.float32) * rng.choice([-1, 1], K).astype(np.float32)
    a = (gamma * invstd).astype(np.float32)
    b = (0.3 * rng.standard_normal(K) - mean * a).astype(np.float32)
    fold = np.stack([a, b, mean.astype(np.float32), invstd]).astype(np.float32)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    dz_d, W_d, zp_d, fold_d = t(dz), t(W), t(zp), t(fold)
    wp = fused._pack_dgrad(P, K, W_d)
    bn, nt, grid = L.pn2_mlp_gemm_bn(P, K), L.pn2_mlp_gemm_n_tiles(P, K), L.pn2_mlp_gemm_grid(P, K)
    out = torch.full((P, K), float("nan"), device=DEV)
    part = torch.full((grid, 2, nt * bn), float("nan"), device=DEV)
    lib.check(L.pn2_mlp_dgrad_stats(P, N, K, dz_d.data_ptr(), wp.data_ptr(), out.data_ptr(), zp_d.data_ptr(),
                                    fold_d.data_ptr(), int(relu), part.data_ptr(), s), "dgrad_stats")
    dgamma, dbeta, c1c2 = (torch.full((m,), float("nan"), device=DEV) for m in (K, K, 2 * K))
    lib.check(L.pn2_bwd_stats_finalize(P, K, nt * bn, grid, 1, part.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                       c1c2.data_ptr(), s), "finalize")
    torch.cuda.synchronize()
    got = out.double().cpu().numpy()
    ref = tf32_trunc(dz).astype(np.float64) @ tf32_rna(W).astype(np.float64)
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
    mask = ((a * zp + b) > 0) | (not relu)
    dy = np.where(mask, got, 0.0)
    zhat = (zp.astype(np.float64) - mean) * invstd
    want_s, want_q = dy.sum(0), (dy * zhat).sum(0)
    gs, gq = dbeta.double().cpu().numpy(), dgamma.double().cpu().numpy()
    scale = max(np.abs(want_s).max(), np.abs(want_q).max())
    # a ReLU decision taken on fmaf(a, z, b) in fp32 can differ from the fp64 one only where a z + b ~ 0: rare, tiny
    assert np.abs(gs - want_s).max() <= 2e-4 * scale, np.abs(gs - want_s).max() / scale
    assert np.abs(gq - want_q).max() <= 2e-4 * scale, np.abs(gq - want_q).max() / scale
    assert np.allclose(c1c2.cpu().numpy(), np.concatenate([gs, gq]) / P, rtol=1e-6, atol=1e-30)
