"""GPU parity of the tcgen05 shared-MLP layer (csrc/mlp_gemm.cu) through the C ABI.

Reference: the same product in fp64 on operands reduced to TF32 exactly as the kernel's are: the
activation tile stays plain fp32 in shared memory and the tensor core reads the upper 19 bits of
each word (truncation); the packed weights are rounded (cvt.rna: nearest, ties away from zero).
What remains is fp32 accumulation order, so the
tolerance is 2e-5 of the output's max magnitude; the fused BatchNorm statistics (column sums and
sums of squares) within 1e-4 relative; the side-output act(A) within 1e-6.
"""
import numpy as np
import pytest
import torch

from pointnet2.scannet_b200 import fused, lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def L_():
    return lib.load()


def tf32_rna(x):
    u = x.astype(np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_trunc(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


def run_gemm(A, W, k, a_in=None, b_in=None, relu=False, want_h=False):
    L = lib.load()
    P, lda = A.shape
    N = W.shape[0]
    s = torch.cuda.current_stream().cuda_stream
    bn, nt, kb = L.pn2_mlp_gemm_bn(P, N), L.pn2_mlp_gemm_n_tiles(P, N), L.pn2_mlp_gemm_k_blocks(k)
    grid = L.pn2_mlp_gemm_grid(P, N)
    assert A.dtype == np.float32 and W.dtype == np.float32
    Wd = torch.from_numpy(W).to(DEV)
    Wp = torch.empty(nt * kb * bn * 32, dtype=torch.float32, device=DEV)
    lib.check(L.pn2_mlp_pack_weights(P, N, k, W.shape[1], Wd.data_ptr(), Wp.data_ptr(), s), "pack")
    Ad = torch.from_numpy(A).to(DEV)
    Z = torch.full((P, N), float("nan"), dtype=torch.float32, device=DEV)
    part = torch.zeros(grid, 2, nt * bn, dtype=torch.float32, device=DEV)
    H = torch.zeros_like(Ad) if want_h else None
    ad = torch.from_numpy(a_in).to(DEV) if a_in is not None else None
    bd = torch.from_numpy(b_in).to(DEV) if b_in is not None else None
    lib.check(L.pn2_mlp_gemm_fwd(P, k, lda, N, N, Ad.data_ptr(),
                                 ad.data_ptr() if ad is not None else None,
                                 bd.data_ptr() if bd is not None else None, int(relu),
                                 H.data_ptr() if H is not None else None, Wp.data_ptr(), Z.data_ptr(),
                                 part.data_ptr(), None, 1.0, 0, None, None, None, None, s), "gemm")
    torch.cuda.synchronize()
    return Z.cpu().numpy(), part.cpu().numpy().sum(0)[:, :N], (H.cpu().numpy() if want_h else None)


@pytest.mark.parametrize("P,K,lda,N", [
    (1000, 9, 12, 32), (4096, 32, 32, 32), (5000, 64, 64, 128), (3000, 131, 132, 128),
    (2048, 256, 256, 512), (777, 64, 64, 20), (1024, 128, 128, 196), (70000, 32, 32, 64),
    (128, 8, 8, 16), (130, 260, 260, 256),
    # few row tiles: the launcher narrows the N tile to cover the machine (96 -> must not become 48, 160 -> not 80)
    (4096, 64, 64, 96), (2000, 128, 128, 160), (8192, 96, 96, 128), (3000, 64, 64, 224), (1500, 32, 32, 320)])
def test_gemm_plain(P, K, lda, N):
    rng = np.random.default_rng(P + N)
    A = np.zeros((P, lda), np.float32)
    A[:, :K] = rng.standard_normal((P, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    Z, stats, _ = run_gemm(A, W, K)
    ref = tf32_trunc(A[:, :K]).astype(np.float64) @ tf32_rna(W).astype(np.float64).T
    assert np.isfinite(Z).all()
    assert np.abs(Z - ref).max() <= 2e-5 * np.abs(ref).max()
    np.testing.assert_allclose(stats[0], ref.sum(0), rtol=1e-4, atol=1e-4 * np.abs(ref).sum(0).max())
    np.testing.assert_allclose(stats[1], (ref * ref).sum(0), rtol=1e-4)


@pytest.mark.parametrize("P,K,N,relu", [(3000, 64, 128, True), (9000, 32, 64, True), (500, 128, 256, False)])
def test_gemm_fused_bn_relu_prologue(P, K, N, relu):
    rng = np.random.default_rng(N)
    A = rng.standard_normal((P, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    a_in = rng.uniform(0.5, 1.5, K).astype(np.float32) * rng.choice([-1, 1], K).astype(np.float32)
    b_in = rng.standard_normal(K).astype(np.float32)
    Z, stats, H = run_gemm(A, W, K, a_in, b_in, relu, want_h=True)
    act = a_in.astype(np.float64) * A + b_in
    if relu:
        act = np.maximum(act, 0)
    np.testing.assert_allclose(H, act, rtol=1e-6, atol=1e-6)
    ref = tf32_trunc(H).astype(np.float64) @ tf32_rna(W).astype(np.float64).T
    assert np.abs(Z - ref).max() <= 2e-5 * np.abs(ref).max()
    np.testing.assert_allclose(stats[1], (ref * ref).sum(0), rtol=1e-4)


@pytest.mark.parametrize("G,ns,K,N", [(64, 32, 32, 64), (100, 32, 64, 128), (37, 16, 32, 32), (300, 32, 128, 256),
                                       (5, 16, 64, 20)])
def test_pooled_layer_extrema_and_groupmax(G, ns, K, N):
    """Pooled last layer of an SA block: the group extrema the layer leaves (read from its staged
    output blocks) are exactly the max / min of the Z it wrote and the first rows attaining them, and
    pn2_groupmax_from_extrema reproduces pn2_bn_relu_groupmax (the streaming kernel over Z) bit for
    bit -- values and arg-max rows -- for positive, negative and zero BatchNorm scales."""
    rng = np.random.default_rng(G * ns + N)
    P = G * ns
    A = torch.from_numpy(rng.standard_normal((P, K)).astype(np.float32)).to(DEV)
    A[: P // 3] = A[0]                                            # ties: whole groups of identical rows
    W = torch.from_numpy((rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)).to(DEV)
    assert L_().pn2_mlp_gemm_pool_supported(ns) == 1
    z, h, part, (ld, nblk), ext = fused._tc_layer(A, K, W, None, False, False, True, None, ns)
    zmax, zmin, imax, imin = ext
    torch.cuda.synchronize()
    zg = z.view(G, ns, N)
    assert torch.equal(zmax, zg.max(dim=1).values) and torch.equal(zmin, zg.min(dim=1).values)
    first_max = (zg == zmax.unsqueeze(1)).float().argmax(dim=1).int()   # first row attaining the max
    first_min = (zg == zmin.unsqueeze(1)).float().argmax(dim=1).int()
    assert torch.equal(imax, first_max) and torch.equal(imin, first_min)
    # statistics from the same column pass
    ps = part.sum(0)
    assert torch.allclose(ps[0, :N], z.sum(0), rtol=1e-4, atol=1e-3)
    assert torch.allclose(ps[1, :N], (z * z).sum(0), rtol=1e-4, atol=1e-3)
    a = torch.from_numpy(rng.uniform(0.5, 1.5, N).astype(np.float32) * rng.choice([-1.0, 1.0, 0.0], N).astype(np.float32)).to(DEV)
    b = torch.from_numpy((0.3 * rng.standard_normal(N)).astype(np.float32)).to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    out1, arg1 = torch.empty(G, N, device=DEV), torch.empty(G, N, dtype=torch.int32, device=DEV)
    out2, arg2 = torch.empty(G, N, device=DEV), torch.empty(G, N, dtype=torch.int32, device=DEV)
    lib.check(L_().pn2_bn_relu_groupmax(G, ns, N, z.data_ptr(), a.data_ptr(), b.data_ptr(), out1.data_ptr(),
                                        arg1.data_ptr(), s), "groupmax")
    zsel = torch.empty(G, N, device=DEV)
    lib.check(L_().pn2_groupmax_from_extrema(G, N, zmax.data_ptr(), zmin.data_ptr(), imax.data_ptr(), imin.data_ptr(),
                                             a.data_ptr(), b.data_ptr(), out2.data_ptr(), arg2.data_ptr(),
                                             zsel.data_ptr(), s), "from_extrema")
    torch.cuda.synchronize()
    assert torch.equal(out1, out2)
    # zsel = z at the arg-max row wherever the pooled value is positive, and the adjoint's statistics read
    # from it are bit-identical to the ones gathered over z
    zg = z.view(G, ns, N).gather(1, arg2.long().unsqueeze(1)).squeeze(1)
    live = (out2 > 0) & (a != 0).unsqueeze(0)     # a == 0: the value is b for every row, the kernels gather z instead
    assert torch.equal(zsel[live], zg[live])
    mean, invstd = torch.randn(N, device=DEV), torch.rand(N, device=DEV) + 0.5
    dout = torch.randn(G, N, device=DEV)
    res = []
    for zs in (None, zsel):
        dg, db, c12 = torch.empty(N, device=DEV), torch.empty(N, device=DEV), torch.empty(2 * N, device=DEV)
        ws = torch.empty(L_().pn2_bn_stats_workspace_floats(N) + 2 * N, device=DEV)
        lib.check(L_().pn2_pool_bn_bwd_stats(G, ns, N, 1, dout.data_ptr(), arg2.data_ptr(), z.data_ptr(),
                                             zs.data_ptr() if zs is not None else None, a.data_ptr(), b.data_ptr(),
                                             mean.data_ptr(), invstd.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                             c12.data_ptr(), ws.data_ptr(), s), "pool_bn_bwd_stats")
        res.append((dg, db, c12))
    torch.cuda.synchronize()
    for x, y in zip(*res):
        assert torch.equal(x, y)
    # the rows can only differ where two different z of a group round to the same a*z+b (then both rows carry
    # the same value); on these inputs ties are exact duplicates, so the rows agree as well
    assert torch.equal(arg1, arg2)
