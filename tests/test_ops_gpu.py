"""GPU parity: every op of libpn2_b200.so, called through the `_ext` table (ctypes ->
C ABI), against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): indices bit-exact; forward values bit-exact where
the op is a copy or an un-contracted fp32 expression; scatter-add gradients within
1e-5 relative (the reference's atomicAdd order is itself nondeterministic; the
oracle accumulates in double).
"""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# gradient tolerance: |got - want| <= GRAD_RTOL * |want| + GRAD_ATOL_REL * max|want|
GRAD_RTOL = 1e-5
GRAD_ATOL_REL = 1e-6


def g(a):
    return cases.t(a, DEV)


def assert_grad_close(got, want):
    got, want = got.cpu().numpy(), want.cpu().numpy() if torch.is_tensor(want) else want
    tol = GRAD_RTOL * np.abs(want) + GRAD_ATOL_REL * max(float(np.abs(want).max()), 1e-30)
    bad = np.abs(got - want) > tol
    assert not bad.any(), "max err %g at %d entries" % (np.abs(got - want).max(), bad.sum())


# ------------------------------------------------------------------ FPS
def test_fps_ssg_pyramid_bit_exact(ext, oracle):
    """All four SSG levels, chained, on duplicate-heavy ScanNet-shaped chunks."""
    xyz, _ = cases.chunk_xyz(4)
    cur = xyz
    for n, m, _ in cases.LEVELS:
        assert cur.shape[1] == n
        want = oracle.ext_cpu.furthest_point_sampling(cases.t(cur), m).numpy()
        got = ext.furthest_point_sampling(g(cur), m).cpu().numpy()
        assert got.dtype == np.int32 and got.shape == (4, m)
        np.testing.assert_array_equal(got, want)
        cur = cases.gather_xyz(cur, want)


@pytest.mark.parametrize("name", sorted(cases.adversarial()))
def test_fps_adversarial(ext, oracle, name):
    cloud = cases.adversarial()[name]
    n = cloud.shape[1]
    for m in (1, min(n, 16), n):
        want = oracle.ext_cpu.furthest_point_sampling(cases.t(cloud), m).numpy()
        got = ext.furthest_point_sampling(g(cloud), m).cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg="%s m=%d" % (name, m))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 37, 64, 100, 129, 257, 513, 1000, 2049, 4097, 8191])
def test_fps_sizes(ext, oracle, n):
    rng = np.random.default_rng(n)
    pts = rng.uniform(0.3, 3.0, (3, n, 3)).astype(np.float32)
    pts = np.take_along_axis(pts, rng.integers(0, max(n * 2 // 3, 1), (3, n, 1)), axis=1)  # duplicates
    m = min(n, 96)
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(pts), m).numpy()
    got = ext.furthest_point_sampling(g(pts), m).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n", [2049, 4096, 4097, 8191, 8192])
def test_fps_pruned_variant_bit_exact(oracle, n):
    """The bounding-box-skipping kernel (pn2_furthest_point_sampling_pruned) picks the same indices."""
    from pointnet2.scannet_b200 import lib
    xyz, _ = cases.chunk_xyz(3, first=40, npoints=n)
    m = 300
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(xyz), m).numpy()
    pts = g(xyz)
    out = torch.empty(3, m, dtype=torch.int32, device=DEV)
    lib.check(lib.load().pn2_furthest_point_sampling_pruned(
        3, n, m, pts.data_ptr(), None, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "fps_pruned")
    np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_fps_more_samples_than_points(ext, oracle):
    pts = np.random.default_rng(5).uniform(1, 2, (2, 50, 3)).astype(np.float32)
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(pts), 80).numpy()
    got = ext.furthest_point_sampling(g(pts), 80).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_fps_streaming_large_cloud(ext, oracle):
    """n > 8192 takes the streaming kernel (minima in the caller's temp)."""
    rng = np.random.default_rng(11)
    pts = rng.uniform(0.5, 6.0, (2, 20000, 3)).astype(np.float32)
    pts[:, 5000:9000] = pts[:, 1000:5000]
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(pts), 128).numpy()
    got = ext.furthest_point_sampling(g(pts), 128).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n,m,b", [(150000, 96, 1), (40000, 64, 3), (8200, 40, 2)])
def test_fps_cluster_whole_scene(ext, oracle, n, m, b):
    """8192 < n <= 159744: one thread-block CLUSTER per cloud (2..16 CTAs, register-resident slices,
    one distributed-shared-memory exchange + cluster barrier per iteration)."""
    rng = np.random.default_rng(n)
    pts = rng.uniform(0.5, 6.0, (b, n, 3)).astype(np.float32)
    pts[:, n // 2:n // 2 + 3000] = pts[:, 100:3100]          # exact duplicates -> ties
    pts[:, 7] = 0.01                                         # near-origin skip
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(pts), m).numpy()
    got = ext.furthest_point_sampling(g(pts), m).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    os.environ["PN2_FPS_NO_CLUSTER"] = "1"                   # the streaming kernel, same answer
    try:
        got2 = ext.furthest_point_sampling(g(pts), m).cpu().numpy()
    finally:
        del os.environ["PN2_FPS_NO_CLUSTER"]
    np.testing.assert_array_equal(got2, want)


def test_fps_streaming_beyond_cluster_capacity(ext, oracle):
    rng = np.random.default_rng(5)
    pts = rng.uniform(0.5, 9.0, (1, 170000, 3)).astype(np.float32)
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(pts), 24).numpy()
    got = ext.furthest_point_sampling(g(pts), 24).cpu().numpy()
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("split", [2, 4, 8])
def test_fps_split_over_cluster(ext, oracle, split):
    """PN2_FPS_SPLIT: resident-size clouds split over a cluster, same indices."""
    xyz, _ = cases.chunk_xyz(2)
    want = oracle.ext_cpu.furthest_point_sampling(cases.t(xyz), 300).numpy()
    os.environ["PN2_FPS_SPLIT"] = str(split)
    try:
        got = ext.furthest_point_sampling(g(xyz), 300).cpu().numpy()
    finally:
        del os.environ["PN2_FPS_SPLIT"]
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------ ball query
@pytest.mark.parametrize("nsample", [16, 32, 64])
def test_ball_query_ssg_levels(ext, oracle, nsample):
    xyz, _ = cases.chunk_xyz(3)
    cur = xyz
    for n, m, r in cases.LEVELS:
        ctr = cases.gather_xyz(cur, oracle.ext_cpu.furthest_point_sampling(cases.t(cur), m).numpy())
        for radius in (r, r / 2):
            want = oracle.ext_cpu.ball_query(cases.t(ctr), cases.t(cur), radius, nsample).numpy()
            got = ext.ball_query(g(ctr), g(cur), radius, nsample).cpu().numpy()
            np.testing.assert_array_equal(got, want)
        cur = ctr


def test_ball_query_edge_cases(ext, oracle):
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 1, (2, 1001, 3)).astype(np.float32)  # 1001*12 bytes: no bulk copy
    ctr = pts[:, :77].copy()
    for radius, ns in ((0.0, 8), (1e-4, 5), (0.2, 1), (0.3, 33), (10.0, 7)):
        want = oracle.ext_cpu.ball_query(cases.t(ctr), cases.t(pts), radius, ns).numpy()
        got = ext.ball_query(g(ctr), g(pts), radius, ns).cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg="r=%g ns=%d" % (radius, ns))
    far = ctr + 50.0  # no hit at all -> rows of zeros (ball_query.cpp:19-21)
    got = ext.ball_query(g(far), g(pts), 0.1, 16).cpu().numpy()
    assert (got == 0).all()


@pytest.mark.parametrize("offset,radius", [(100.0, 1e-3), (1000.0, 1e-3), (100.0, 0.01), (1000.0, 0.01),
                                           (30.0, 0.01), (3.0, 1e-3), (5000.0, 0.1), (0.0, 2e-3)])
def test_ball_query_grid_path_large_offsets_small_radii(ext, oracle, offset, radius):
    """The spatial-hash path (512 <= n <= 8192) with coordinates far from the origin and radii down to a
    millimetre: fp32 rounding of coordinate / cell size exceeds the cell margin once |x| / r > ~4000, and the
    cell index clamps at 2^20; the kernel must notice and scan (results stay the reference's, bit for bit).
    Points are laid on a lattice of pitch ~r/2 so that balls are well filled at every radius."""
    rng = np.random.default_rng(int(offset) + int(radius * 1e4))
    n, m, ns = 2048, 256, 32
    base = rng.integers(0, 24, (2, n, 3)).astype(np.float32) * np.float32(radius * 0.45)
    base += rng.normal(0, radius * 0.02, base.shape).astype(np.float32)
    pts = (base + np.float32(offset)).astype(np.float32)
    ctr = pts[:, rng.choice(n, m, replace=False)].copy()
    want = oracle.ext_cpu.ball_query(cases.t(ctr), cases.t(pts), radius, ns).numpy()
    got = ext.ball_query(g(ctr), g(pts), radius, ns).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert (want != want[:, :, :1]).any()          # the balls are not all singletons / empty


def test_ball_query_whole_scene_cloud(ext, oracle):
    """n > 8192: scanned from L2 instead of the shared-memory stage."""
    from pointnet2.scannet_b200 import synthetic as syn
    scene = syn.make_scene(1, npoints=30000)[None, :, :3].copy()
    ctr = scene[:, ::61][:, :300].copy()
    want = oracle.ext_cpu.ball_query(cases.t(ctr), cases.t(scene), 0.1, 32).numpy()
    got = ext.ball_query(g(ctr), g(scene), 0.1, 32).cpu().numpy()
    np.testing.assert_array_equal(got, want)


# -------------------------------------------------------------- three_nn
def test_three_nn_ssg_levels(ext, oracle):
    xyz, _ = cases.chunk_xyz(2)
    levels = [xyz]
    for n, m, _ in cases.LEVELS:
        levels.append(cases.gather_xyz(levels[-1], oracle.ext_cpu.furthest_point_sampling(
            cases.t(levels[-1]), m).numpy()))
    for unknown, known in zip(levels[:-1], levels[1:]):
        wd, wi = oracle.ext_cpu.three_nn(cases.t(unknown), cases.t(known))
        gd, gi = ext.three_nn(g(unknown), g(known))
        np.testing.assert_array_equal(gi.cpu().numpy(), wi.numpy())
        np.testing.assert_array_equal(gd.cpu().numpy(), wd.numpy())  # squared, bit-exact


@pytest.mark.parametrize("m", [1, 2, 3, 4, 5000])
def test_three_nn_few_or_many_known(ext, oracle, m):
    rng = np.random.default_rng(m)
    unknown = rng.uniform(0, 2, (2, 333, 3)).astype(np.float32)
    known = rng.uniform(0, 2, (2, m, 3)).astype(np.float32)
    if m >= 4:
        known[:, 1] = known[:, 0]  # exact tie: lower index keeps the better slot
    wd, wi = oracle.ext_cpu.three_nn(cases.t(unknown), cases.t(known))
    gd, gi = ext.three_nn(g(unknown), g(known))
    np.testing.assert_array_equal(gi.cpu().numpy(), wi.numpy())
    np.testing.assert_array_equal(gd.cpu().numpy(), wd.numpy())  # +inf where m < 3


# ------------------------------------------------------- gather / grouping
@pytest.mark.parametrize("B,C,N,npoint,nsample", [
    (2, 3, 8192, 1024, 32), (2, 6, 8192, 1024, 32), (2, 64, 1024, 256, 32),
    (2, 131, 512, 64, 16), (1, 5, 100, 7, 3), (3, 1, 33, 5, 1)])
def test_group_points_fwd_bwd(ext, oracle, B, C, N, npoint, nsample):
    rng = np.random.default_rng(B * 1000 + C)
    feats = cases.features(rng, B, C, N)
    idx = rng.integers(0, N, (B, npoint, nsample)).astype(np.int32)
    idx[:, :, nsample // 2:] = idx[:, :, :1]  # ball-query style padding with the first hit
    want = oracle.ext_cpu.group_points(cases.t(feats), cases.t(idx))
    got = ext.group_points(g(feats), g(idx))
    assert got.shape == (B, C, npoint, nsample)
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())

    go = rng.standard_normal((B, C, npoint, nsample)).astype(np.float32)
    want = oracle.ext_cpu.group_points_grad(cases.t(go), cases.t(idx), N)
    got = ext.group_points_grad(g(go), g(idx), N)
    assert got.shape == (B, C, N)
    assert_grad_close(got, want)


def test_group_points_grad_large_n_global_path(ext, oracle):
    rng = np.random.default_rng(9)
    B, C, N = 1, 2, 60000
    idx = rng.integers(0, N, (B, 500, 8)).astype(np.int32)
    go = rng.standard_normal((B, C, 500, 8)).astype(np.float32)
    want = oracle.ext_cpu.group_points_grad(cases.t(go), cases.t(idx), N)
    assert_grad_close(ext.group_points_grad(g(go), g(idx), N), want)


@pytest.mark.parametrize("B,C,N,m", [(2, 3, 8192, 1024), (2, 3, 1024, 256), (1, 7, 50, 13)])
def test_gather_points_fwd_bwd(ext, oracle, B, C, N, m):
    rng = np.random.default_rng(m)
    feats = cases.features(rng, B, C, N)
    idx = rng.integers(0, N, (B, m)).astype(np.int32)
    idx[:, 1] = idx[:, 0]
    np.testing.assert_array_equal(ext.gather_points(g(feats), g(idx)).cpu().numpy(),
                                  oracle.ext_cpu.gather_points(cases.t(feats), cases.t(idx)).numpy())
    go = rng.standard_normal((B, C, m)).astype(np.float32)
    assert_grad_close(ext.gather_points_grad(g(go), g(idx), N),
                      oracle.ext_cpu.gather_points_grad(cases.t(go), cases.t(idx), N))


# ----------------------------------------------------------- interpolation
@pytest.mark.parametrize("B,C,m,n", [(2, 128, 1024, 8192), (2, 256, 256, 1024), (2, 512, 16, 64),
                                     (1, 3, 5, 7), (2, 9, 3, 1001)])
def test_three_interpolate_fwd_bwd(ext, oracle, B, C, m, n):
    rng = np.random.default_rng(C + n)
    unknown = rng.uniform(0, 2, (B, n, 3)).astype(np.float32)
    known = rng.uniform(0, 2, (B, m, 3)).astype(np.float32)
    d2, idx = oracle.ext_cpu.three_nn(cases.t(unknown), cases.t(known))
    w = cases.fp_weights(d2.numpy())
    feats = cases.features(rng, B, C, m)
    want = oracle.ext_cpu.three_interpolate(cases.t(feats), idx, cases.t(w))
    got = ext.three_interpolate(g(feats), idx.to(DEV), g(w))
    np.testing.assert_array_equal(got.cpu().numpy(), want.numpy())  # un-contracted fp32: bit-exact

    go = rng.standard_normal((B, C, n)).astype(np.float32)
    want = oracle.ext_cpu.three_interpolate_grad(cases.t(go), idx, cases.t(w), m)
    assert_grad_close(ext.three_interpolate_grad(g(go), idx.to(DEV), g(w), m), want)


# ------------------------------------------------------------ error surface
def test_rejects_bad_tensors(ext):
    pts = torch.rand(1, 64, 3)
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(pts, 4)  # CPU tensor: no CPU path
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(pts.to(DEV).double(), 4)
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(pts.to(DEV).transpose(1, 2), 4)  # non-contiguous
    f = torch.rand(1, 4, 64, device=DEV)
    with pytest.raises(RuntimeError):
        ext.gather_points(f, torch.zeros(1, 3, dtype=torch.int64, device=DEV))  # idx must be int32


def test_launch_counter(ext):
    from pointnet2.scannet_b200 import lib
    lib.reset_launch_count()
    ext.furthest_point_sampling(torch.rand(2, 128, 3, device=DEV), 8)
    ext.three_nn(torch.rand(2, 64, 3, device=DEV), torch.rand(2, 16, 3, device=DEV))
    assert lib.launch_count() == 2
