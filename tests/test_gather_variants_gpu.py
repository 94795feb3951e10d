"""GPU: the two gather kernels behind group_points / three_interpolate of the channel-major op surface --
the shared-memory gather (default where the rows fit and are reused, csrc/group.cu, csrc/interpolate.cu) and
the L1 gather it replaced (PN2_GROUP_SMEM=0 / PN2_INTERP_SMEM=0, still used for the other shapes).  Each case
compares both with the oracle AND with each other bit for bit (a gather is a copy; the interpolation is the
same three un-contracted multiplies and two adds)."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
DEV = "cuda"


def g(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


class _env(object):
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        os.environ[self.name] = "0"   # putenv: the library reads it with getenv on every call

    def __exit__(self, *exc):
        os.environ.pop(self.name, None)


# E >= 2 n (the variant's own precondition), vectorised and scalar tails, one channel tile and many,
# the SSG shapes of SA2 (64 channels, 1024 points -> 256 x 32) and SA1's multiview width
@pytest.mark.parametrize("B,C,N,npoint,nsample", [(2, 64, 1024, 256, 32), (1, 3, 100, 50, 5), (3, 7, 33, 40, 3),
                                                  (2, 131, 2048, 256, 32), (1, 1, 4, 8, 1), (2, 20, 8192, 1024, 32),
                                                  (1, 5, 14000, 1024, 32)])
def test_group_smem_equals_l1_gather(ext, oracle, B, C, N, npoint, nsample):
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
def test_interp_smem_equals_l1_gather(ext, oracle, B, C, m, n):
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
