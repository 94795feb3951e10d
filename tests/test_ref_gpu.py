"""GPU: the product against the reference's golden outputs, and -- when oracle/_ref was
built (it ships with the snapshot, the sources do not) -- against the reference's own
kernels executed live on the same device."""
import numpy as np
import pytest

import cases
import golden_util

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not golden_util.have_golden(), reason="golden vectors not generated yet")
def test_product_matches_reference_goldens(ext):
    assert golden_util.check_backend(ext, "cuda:0") > 40


def _ref():
    from oracle import build_ref
    if not build_ref.available("nofma"):
        pytest.skip("oracle/_ref/nofma not built")
    return build_ref.load_ref("nofma")


def test_product_vs_live_reference_full_size(ext):
    """B=8 chunks of 8192 points through FPS -> ball query -> three_nn, live vs reference."""
    ref = _ref()
    xyz, _ = cases.chunk_xyz(8, first=100)
    cur = cases.t(xyz, "cuda:0")
    levels = [cur]
    for n, m, r in cases.LEVELS:
        a = ref.furthest_point_sampling(cur, m)
        b = ext.furthest_point_sampling(cur, m)
        assert (a == b).all(), "FPS differs at level N=%d" % n
        ctr = cur.gather(1, a.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        assert (ref.ball_query(ctr, cur, r, 32) == ext.ball_query(ctr, cur, r, 32)).all()
        levels.append(ctr)
        cur = ctr
    for unknown, known in zip(levels[:-1], levels[1:]):
        (d_a, i_a), (d_b, i_b) = ref.three_nn(unknown, known), ext.three_nn(unknown, known)
        assert (i_a == i_b).all()
        assert np.array_equal(d_a.cpu().numpy().view(np.uint32), d_b.cpu().numpy().view(np.uint32))


def test_oracle_vs_live_reference(oracle):
    """Pins the CPU oracle on fresh inputs too (not only on the committed goldens)."""
    ref = _ref()
    xyz, _ = cases.chunk_xyz(2, first=200)
    want = ref.furthest_point_sampling(cases.t(xyz, "cuda:0"), 512).cpu()
    got = oracle.ext_cpu.furthest_point_sampling(cases.t(xyz), 512)
    assert (want == got).all()


@pytest.mark.parametrize("flags", [[], ["--msg", "--multiview"]])
def test_reference_checkpoint_round_trip(flags):
    """model.pth saved by the reference's own get_model (its Python + its CUDA extension),
    loaded by the mirror, evaluated unfused (bit-identical logits) and on the fused tcgen05 path (TF32 class
    bound): tools/checkpoint_roundtrip.py, in its own process."""
    import os
    import subprocess
    import sys
    from oracle import build_ref
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not build_ref.available("nofma") or not os.path.exists(os.path.join(build_ref.python_dir(), "pointnet2_semseg.py")):
        pytest.skip("oracle/_ref (reference extension + staged Python) not built")
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "checkpoint_roundtrip.py")] + flags,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "bit for bit" in p.stdout
