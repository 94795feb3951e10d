"""CPU: the oracle (oracle/pn2_oracle.c) against the golden outputs of the reference's
own kernels (tests/golden/ref_nofma_ops.npz, made by tests/golden/make_golden.py on a
B200 from /root/reference compiled with -fmad=false).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import cases
import golden_util


@pytest.mark.skipif(not golden_util.have_golden(), reason="golden vectors not generated yet")
def test_oracle_matches_reference_goldens(oracle):
    assert golden_util.check_backend(oracle.ext_cpu, "cpu") > 40


def test_reference_known_answer(oracle):
    """The one KAT derivable from the reference's own test (pointnet2_test.py:18-30):
    idx [[0,1,2],[1,2,3]], weight [[1,1,1],[2,2,2]] -> out[..,0] = f0+f1+f2, out[..,1] = 2(f1+f2+f3)."""
    f = torch.arange(8, dtype=torch.float32).view(1, 2, 4)
    idx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32)
    w = torch.tensor([[[1, 1, 1], [2, 2, 2]]], dtype=torch.float32)
    out = oracle.ext_cpu.three_interpolate(f, idx, w)
    assert out.tolist() == [[[3.0, 12.0], [15.0, 36.0]]]
    gp = oracle.ext_cpu.three_interpolate_grad(torch.ones(1, 2, 2), idx, w, 4)
    assert gp.tolist() == [[[1.0, 3.0, 3.0, 2.0]] * 2]


def test_fps_literal_emulation_equals_closed_form(oracle):
    """The oracle emulates the reference's thread layout literally; the CUDA kernel uses the
    closed-form tie-break key.  They must agree, especially on duplicate-heavy clouds."""
    def closed_form(p, m):
        n = len(p)
        key = oracle.fps_tiebreak_key(n)
        temp = np.full(n, 1e10, np.float32)
        mag = (p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]) + p[:, 2] * p[:, 2]
        live = ~(mag.astype(np.float64) <= 1e-3)
        out = np.zeros(m, np.int32)
        cur = 0
        for j in range(1, m):
            d = p - p[cur]
            d = d * d
            temp = np.where(live, np.minimum((d[:, 0] + d[:, 1]) + d[:, 2], temp), temp)
            if live.any():
                cand = np.where(live & (temp == temp[live].max()))[0]
                cur = cand[np.argmin(key[cand])]
            else:
                cur = 0
            out[j] = cur
        return out

    clouds = dict(cases.adversarial())
    xyz, _ = cases.chunk_xyz(1, npoints=2048)
    clouds["chunk_2048"] = xyz
    for name, c in clouds.items():
        m = min(c.shape[1], 64)
        want = closed_form(c[0], m)
        got = oracle.ext_cpu.furthest_point_sampling(cases.t(c), m)[0].numpy()
        np.testing.assert_array_equal(got, want, err_msg=name)


def test_block_threads_matches_reference_formula(oracle):
    import math
    for n in list(range(1, 70)) + [100, 255, 256, 511, 512, 513, 1000, 1024, 8191, 8192, 150000]:
        want = max(min(1 << int(math.log(n) / math.log(2.0)), 512), 1)  # cuda_utils.h:15-19
        assert oracle.block_threads(n) == want


def test_oracle_edge_semantics(oracle):
    pts = torch.rand(1, 50, 3)
    ctr = pts[:, :4].contiguous() + 10.0
    assert (oracle.ext_cpu.ball_query(ctr, pts, 0.1, 8) == 0).all()  # empty ball -> zeros
    idx = oracle.ext_cpu.ball_query(pts[:, :4].contiguous(), pts, 1e-6, 8)
    assert (idx == torch.arange(4, dtype=torch.int32).view(1, 4, 1)).all()  # first hit repeated
    d2, i3 = oracle.ext_cpu.three_nn(pts, pts[:, :2].contiguous())
    assert torch.isinf(d2[..., 2]).all() and (i3[..., 2] == 0).all()  # m < 3 -> (+inf, 0)
    near = torch.full((1, 8, 3), 0.01)
    assert (oracle.ext_cpu.furthest_point_sampling(near, 5) == 0).all()  # every point skipped
