"""GPU, opt-in (PN2_EXPERIMENTAL=1): kernels written after round 1's GPU budget was spent and therefore
NOT yet run on a B200 -- the shared-memory-gather variants of the channel-major grouping and
three-point interpolation forward kernels (PN2_GROUP_SMEM / PN2_INTERP_SMEM in csrc/group.cu,
csrc/interpolate.cu).  They are off by default, so these tests are skipped by default; run

    PN2_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -q

on the GPU box before switching either on.  Each case compares the variant with the oracle AND with the
default kernel bit for bit (a gather is a copy; the interpolation is the same three un-contracted
multiplies and two adds)."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PN2_EXPERIMENTAL") != "1",
                                 reason="experimental kernels: set PN2_EXPERIMENTAL=1")]
DEV = "cuda"


def g(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


class _env(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        os.environ[self.name] = "1"   # putenv: the library reads it with getenv on every call

    def __exit__(self, *exc):
        os.environ.pop(self.name, None)


# E >= 2 n (the variant's own precondition), vectorised and scalar tails, one channel tile and many,
# the SSG shapes of SA2 (64 channels, 1024 points -> 256 x 32) and SA1's multiview width
@pytest.mark.parametrize("B,C,N,npoint,nsample", [(2, 64, 1024, 256, 32), (1, 3, 100, 50, 5), (3, 7, 33, 40, 3),
                                                  (2, 131, 2048, 256, 32), (1, 1, 4, 8, 1), (2, 20, 8192, 1024, 32),
                                                  (1, 5, 14000, 1024, 32)])
def test_experimental_group_smem_equals_default(ext, oracle, B, C, N, npoint, nsample):
    rng = np.random.default_rng(B * 1000 + C)
    feats = cases.features(rng, B, C, N)
    idx = rng.integers(0, N, (B, npoint, nsample)).astype(np.int32)
    idx[:, :, nsample // 2:] = idx[:, :, :1]
    base = ext.group_points(g(feats), g(idx))
    with _env("PN2_GROUP_SMEM"):
        got = ext.group_points(g(feats), g(idx))
    torch.cuda.synchronize()
    assert torch.equal(got, base)
    np.testing.assert_array_equal(got.cpu().numpy(),
                                  oracle.ext_cpu.group_points(cases.t(feats), cases.t(idx)).numpy())


@pytest.mark.parametrize("B,C,m,n", [(2, 128, 1024, 8192), (2, 256, 256, 1024), (1, 512, 16, 64), (3, 5, 7, 30),
                                     (1, 33, 100, 1001), (2, 1, 3, 4), (1, 9, 14000, 9400)])
def test_experimental_interp_smem_equals_default(ext, oracle, B, C, m, n):
    rng = np.random.default_rng(C + n)
    unknown = rng.uniform(0, 2, (B, n, 3)).astype(np.float32)
    known = rng.uniform(0, 2, (B, m, 3)).astype(np.float32)
    d2, idx = oracle.ext_cpu.three_nn(cases.t(unknown), cases.t(known))
    w = cases.fp_weights(d2.numpy())
    feats = cases.features(rng, B, C, m)
    base = ext.three_interpolate(g(feats), idx.to(DEV), g(w))
    with _env("PN2_INTERP_SMEM"):
        got = ext.three_interpolate(g(feats), idx.to(DEV), g(w))
    torch.cuda.synchronize()
    assert torch.equal(got, base)
    np.testing.assert_array_equal(got.cpu().numpy(),
                                  oracle.ext_cpu.three_interpolate(cases.t(feats), idx, cases.t(w)).numpy())
