"""CPU: the C oracle (oracle/pn2_oracle.c) against an INDEPENDENT NumPy restatement of the same
reference semantics, written from SURVEY.md 8(c)'s recipes (hit-mask rank for the ball query, stable
sort for 3-NN, index_add for the scatters) on seeded random inputs the golden fixtures do not hold:
ragged sizes, tie-heavy clouds on a coarse lattice (many equal distances and exact duplicates),
m < 3 known points, nsample larger than the cloud, empty clouds of centres.

Two restatements that were written differently and still agree bit for bit are a stronger pin than
either alone; the goldens (tests/test_oracle_golden.py) remain what ties both to the reference."""
import numpy as np
import pytest
import torch

F = np.float32


def _sqdist(a, b):
    """((dx*dx)+(dy*dy))+(dz*dz), one fp32 rounding per operation (ball_query_gpu.cu:31-32,
    interpolate_gpu.cu:33): a (..., 1, 3) against b (..., n, 3)."""
    d = (a - b).astype(F)
    d = (d * d).astype(F)
    return ((d[..., 0] + d[..., 1]).astype(F) + d[..., 2]).astype(F)


def _cloud(rng, b, n, lattice):
    if lattice:  # coordinates on a 0.125 grid: exact in fp32, distances tie all the time
        return (rng.integers(0, 6, size=(b, n, 3)) * 0.125 + 0.5).astype(F)
    return rng.random((b, n, 3), dtype=F) + F(0.5)


def ball_query_np(new_xyz, xyz, radius, nsample):  # ball_query_gpu.cu:22-43
    b, m, n = new_xyz.shape[0], new_xyz.shape[1], xyz.shape[1]
    r2 = F(F(radius) * F(radius))
    out = np.zeros((b, m, nsample), np.int32)
    for i in range(b):
        hit = _sqdist(new_xyz[i][:, None, :], xyz[i][None, :, :]) < r2   # (m, n), strict
        rank = np.cumsum(hit, axis=1) - 1
        for j in range(m):
            ks = np.nonzero(hit[j])[0]
            if len(ks) == 0:
                continue                      # the binding's zero fill stays (ball_query.cpp:19-21)
            out[i, j, :] = ks[0]              # unused slots repeat the first hit
            take = ks[rank[j, ks] < nsample]
            out[i, j, :len(take)] = take
    return out


def three_nn_np(unknown, known):  # interpolate_gpu.cu:22-58: strict <, lower index wins ties
    b, n, m = unknown.shape[0], unknown.shape[1], known.shape[1]
    d2 = np.full((b, n, 3), np.inf, F)        # (float)1e40 = +inf
    idx = np.zeros((b, n, 3), np.int32)
    for i in range(b):
        d = _sqdist(unknown[i][:, None, :], known[i][None, :, :])        # (n, m)
        order = np.argsort(d, axis=1, kind="stable")[:, :3]
        k = order.shape[1]
        idx[i, :, :k] = order
        d2[i, :, :k] = np.take_along_axis(d, order, axis=1)
    return d2, idx


@pytest.mark.parametrize("seed,b,n,m,nsample,radius,lattice", [
    (0, 2, 37, 5, 8, 0.3, False), (1, 1, 100, 17, 16, 0.25, True), (2, 3, 64, 64, 4, 0.2, True),
    (3, 1, 1, 1, 3, 0.5, False), (4, 2, 9, 4, 32, 10.0, True), (5, 1, 257, 33, 32, 0.13, False),
    (6, 2, 50, 0, 8, 0.3, False), (7, 1, 200, 40, 1, 0.126, True),
])
def test_ball_query_equals_numpy_restatement(oracle, seed, b, n, m, nsample, radius, lattice):
    rng = np.random.default_rng(seed)
    xyz = _cloud(rng, b, n, lattice)
    ctr = np.ascontiguousarray(xyz[:, rng.integers(0, n, size=m)]) if m else np.zeros((b, 0, 3), F)
    if not lattice and m:
        ctr = (ctr + rng.normal(0, 0.02, ctr.shape)).astype(F)
    got = oracle.ext_cpu.ball_query(torch.from_numpy(ctr), torch.from_numpy(xyz), radius, nsample).numpy()
    np.testing.assert_array_equal(got, ball_query_np(ctr, xyz, radius, nsample))


@pytest.mark.parametrize("seed,b,n,m,lattice", [
    (0, 2, 50, 7, False), (1, 1, 33, 1, False), (2, 2, 64, 2, True), (3, 1, 128, 3, True),
    (4, 3, 100, 40, True), (5, 1, 500, 129, False), (6, 1, 1, 5, True),
])
def test_three_nn_equals_numpy_restatement(oracle, seed, b, n, m, lattice):
    rng = np.random.default_rng(100 + seed)
    unknown, known = _cloud(rng, b, n, lattice), _cloud(rng, b, m, lattice)
    d2, idx = oracle.ext_cpu.three_nn(torch.from_numpy(unknown), torch.from_numpy(known))
    want_d2, want_idx = three_nn_np(unknown, known)
    np.testing.assert_array_equal(idx.numpy(), want_idx)
    np.testing.assert_array_equal(d2.numpy(), want_d2)   # bit-exact, +inf padding included


@pytest.mark.parametrize("seed,b,c,n,npoint,ns", [(0, 2, 5, 40, 7, 3), (1, 1, 1, 9, 1, 1), (2, 3, 16, 33, 10, 32),
                                                   (3, 1, 7, 64, 0, 4)])
def test_group_and_gather_equal_numpy_restatement(oracle, seed, b, c, n, npoint, ns):
    rng = np.random.default_rng(200 + seed)
    pts = rng.standard_normal((b, c, n)).astype(F)
    idx = rng.integers(0, n, size=(b, npoint, ns)).astype(np.int32)
    idx[:, :, 1:] = np.where(rng.random((b, npoint, ns - 1)) < 0.5, idx[:, :, :1], idx[:, :, 1:])  # padded rows
    tp, ti = torch.from_numpy(pts), torch.from_numpy(idx)
    # forward: out[b, c, j, k] = points[b, c, idx[b, j, k]]      (group_points_gpu.cu:20-27)
    want = np.stack([pts[i][:, idx[i]] for i in range(b)]) if b else pts
    np.testing.assert_array_equal(oracle.ext_cpu.group_points(tp, ti).numpy(), want)
    # adjoint: scatter-add into zeros                              (group_points_gpu.cu:55-62)
    g = rng.standard_normal((b, c, npoint, ns)).astype(F)
    acc = np.zeros((b, c, n), np.float64)
    for i in range(b):
        np.add.at(acc[i], (slice(None), idx[i].reshape(-1)), g[i].reshape(c, -1).astype(np.float64))
    got = oracle.ext_cpu.group_points_grad(torch.from_numpy(g), ti, n).numpy()
    np.testing.assert_allclose(got, acc.astype(F), rtol=1e-6, atol=1e-6)
    # gather = grouping with one sample per centre                 (sampling_gpu.cu:12-19, 38-46)
    gi = np.ascontiguousarray(idx[:, :, 0])
    np.testing.assert_array_equal(oracle.ext_cpu.gather_points(tp, torch.from_numpy(gi)).numpy(),
                                  np.stack([pts[i][:, gi[i]] for i in range(b)]))
    gg = np.ascontiguousarray(g[:, :, :, 0])
    acc = np.zeros((b, c, n), np.float64)
    for i in range(b):
        np.add.at(acc[i], (slice(None), gi[i]), gg[i].astype(np.float64))
    np.testing.assert_allclose(oracle.ext_cpu.gather_points_grad(torch.from_numpy(gg), torch.from_numpy(gi), n).numpy(),
                               acc.astype(F), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("seed,b,c,m,n", [(0, 2, 6, 5, 30), (1, 1, 1, 3, 1), (2, 2, 33, 64, 100), (3, 1, 4, 1, 9)])
def test_three_interpolate_equals_numpy_restatement(oracle, seed, b, c, m, n):
    rng = np.random.default_rng(300 + seed)
    pts = rng.standard_normal((b, c, m)).astype(F)
    idx = rng.integers(0, m, size=(b, n, 3)).astype(np.int32)
    w = rng.random((b, n, 3), dtype=F)
    w = (w / w.sum(axis=2, keepdims=True)).astype(F)
    # ((p1*w1) + (p2*w2)) + (p3*w3), each operation rounded to fp32  (interpolate_gpu.cu:95-98)
    want = np.empty((b, c, n), F)
    for i in range(b):
        t = [(pts[i][:, idx[i, :, k]] * w[i, :, k][None, :]).astype(F) for k in range(3)]
        want[i] = ((t[0] + t[1]).astype(F) + t[2]).astype(F)
    got = oracle.ext_cpu.three_interpolate(torch.from_numpy(pts), torch.from_numpy(idx), torch.from_numpy(w))
    np.testing.assert_array_equal(got.numpy(), want)
    # adjoint: grad_points[b, c, idx_k] += grad_out[b, c, j] * w_k   (interpolate_gpu.cu:128-141)
    g = rng.standard_normal((b, c, n)).astype(F)
    acc = np.zeros((b, c, m), np.float64)
    for i in range(b):
        for k in range(3):
            np.add.at(acc[i], (slice(None), idx[i, :, k]), (g[i] * w[i, :, k][None, :]).astype(F).astype(np.float64))
    got = oracle.ext_cpu.three_interpolate_grad(torch.from_numpy(g), torch.from_numpy(idx), torch.from_numpy(w), m)
    np.testing.assert_allclose(got.numpy(), acc.astype(F), rtol=1e-6, atol=1e-6)


def fps_np(p, m, key):
    """sampling_gpu.cu:69-173 with the thread layout folded into `key` (oracle.fps_tiebreak_key): the next
    sample is the live point of maximal running minimum, ties to the smallest key; skipped points
    ((double)|p|^2 <= 1e-3, :100-101) never update nor win; no live point -> 0."""
    n = len(p)
    sq = (p * p).astype(F)
    live = ((sq[:, 0] + sq[:, 1]).astype(F) + sq[:, 2]).astype(F).astype(np.float64) > 1e-3
    temp = np.full(n, 1e10, F)
    out, cur = np.zeros(m, np.int32), 0
    for j in range(1, m):
        temp[live] = np.minimum(_sqdist(p[cur][None, :], p)[live], temp[live])
        cur = 0
        if live.any():
            ks = np.nonzero(live)[0]
            cur = ks[np.lexsort((key[ks], -temp[ks].astype(np.float64)))[0]]
        out[j] = cur
    return out


@pytest.mark.parametrize("seed,n,m,lattice,origin", [
    (0, 37, 37, True, False), (1, 64, 20, True, True), (2, 100, 64, True, False), (3, 200, 50, False, True),
    (4, 300, 128, True, True), (5, 513, 40, True, False), (6, 1000, 64, False, False), (7, 5, 5, True, True),
    (8, 1, 1, False, False), (9, 8191, 16, True, False),
])
def test_fps_equals_numpy_restatement(oracle, seed, n, m, lattice, origin):
    rng = np.random.default_rng(400 + seed)
    p = _cloud(rng, 1, n, lattice)[0]
    if origin:  # some points inside the skip radius around the origin, the first one included
        p[rng.integers(0, n, size=max(1, n // 7))] = (rng.random((max(1, n // 7), 3)) * 0.015).astype(F)
        p[0] = 0.0
    got = oracle.ext_cpu.furthest_point_sampling(torch.from_numpy(p[None]), m)[0].numpy()
    np.testing.assert_array_equal(got, fps_np(p, m, oracle.fps_tiebreak_key(n)))
