"""Comparison of an `_ext`-shaped backend against tests/golden/ref_nofma_ops.npz."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_nofma_ops.npz")
INFO = os.path.join(HERE, "golden", "ref_info.json")
sys.path.insert(0, os.path.join(HERE, "golden"))

# scatter-add gradients: the reference's own result depends on atomicAdd order
GRAD_KEYS = {"group_grad", "gather_grad", "interp_grad"}
GRAD_RTOL, GRAD_ATOL_REL = 1e-5, 1e-6


def have_golden():
    return os.path.exists(GOLDEN) and os.path.exists(INFO)


def check_backend(backend, dev):
    """Run the golden call sequence on `backend` and compare with the reference's outputs."""
    import make_golden
    got, hashes = make_golden.run_ops(backend, dev)
    want = np.load(GOLDEN)
    info = json.load(open(INFO))
    assert hashes == info["input_sha256"], "seeded inputs differ from the ones the goldens were made on"
    assert sorted(got) == sorted(want.files)
    for k in want.files:
        w, g = want[k], got[k]
        assert g.shape == w.shape and g.dtype == w.dtype, k
        if k in GRAD_KEYS:
            tol = GRAD_RTOL * np.abs(w) + GRAD_ATOL_REL * np.abs(w).max()
            assert (np.abs(g - w) <= tol).all(), "%s: max err %g" % (k, np.abs(g - w).max())
        else:  # indices, copies, squared distances, un-contracted fp32 expressions: bit-exact
            assert np.array_equal(g.view(np.uint8), w.view(np.uint8)), "%s differs from the reference" % k
    return len(want.files)
