"""Generate the golden vectors that pin the oracle: outputs of the REFERENCE'S OWN
kernels (oracle/_ref, built by oracle/build_ref.py from /root/reference, unmodified,
-fmad=false) on the seeded inputs of tests/cases.py.

Needs a GPU (the reference has no CPU path), so it runs on the B200 box:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

and the two files it writes are then committed under tests/golden/:

    ref_nofma_ops.npz   outputs (indices in full; float outputs in full for the
                        small cases, as fp32) of ref_nofma
    ref_info.json       provenance + how often the STOCK build (fmad on) differs

Inputs are not stored: tests regenerate them from the same seeds and check their
sha256 against the ones recorded here.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import build_ref  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_ops(ref, dev="cuda:0"):
    """-> dict name -> numpy array.  Shared with the tests (same call sequence)."""
    out, hashes = {}, {}
    g = lambda a: cases.t(a, dev)  # noqa: E731

    # ---- SSG pyramid on 2 chunks: FPS, ball query, three_nn at every level
    xyz, full = cases.chunk_xyz(2)
    hashes["chunk_xyz"] = sha(xyz)
    levels = [xyz]
    for li, (n, m, r) in enumerate(cases.LEVELS):
        cur = levels[-1]
        idx = ref.furthest_point_sampling(g(cur), m).cpu().numpy()
        out["fps_l%d" % li] = idx
        ctr = cases.gather_xyz(cur, idx)
        out["bq_l%d" % li] = ref.ball_query(g(ctr), g(cur), r, 32).cpu().numpy()
        out["bq16_l%d" % li] = ref.ball_query(g(ctr), g(cur), r / 2, 16).cpu().numpy()
        levels.append(ctr)
    for li, (unknown, known) in enumerate(zip(levels[:-1], levels[1:])):
        d2, i3 = ref.three_nn(g(unknown), g(known))
        out["nn_d2_l%d" % li] = d2.cpu().numpy()
        out["nn_idx_l%d" % li] = i3.cpu().numpy()

    # ---- adversarial clouds: full-length FPS, ball query, three_nn with tiny m
    for name, cloud in sorted(cases.adversarial().items()):
        n = cloud.shape[1]
        hashes["adv_" + name] = sha(cloud)
        out["adv_fps_" + name] = ref.furthest_point_sampling(g(cloud), n).cpu().numpy()
        ctr = cloud[:, : min(n, 16)].copy()
        out["adv_bq_" + name] = ref.ball_query(g(ctr), g(cloud), 0.25, 8).cpu().numpy()
        for m in (1, 2, 3):
            d2, i3 = ref.three_nn(g(cloud), g(cloud[:, :m].copy()))
            out["adv_nn_d2_%d_%s" % (m, name)] = d2.cpu().numpy()
            out["adv_nn_idx_%d_%s" % (m, name)] = i3.cpu().numpy()

    # ---- value ops on level-1 geometry (N=1024 -> 256), C = 16
    rng = np.random.default_rng(1234)
    feats = cases.features(rng, 2, 16, 1024)
    hashes["feats"] = sha(feats)
    bq = out["bq_l1"]
    grouped = ref.group_points(g(feats), g(bq)).cpu().numpy()
    out["group_sha"] = np.frombuffer(bytes.fromhex(sha(grouped)), dtype=np.uint8)
    go = rng.standard_normal(grouped.shape).astype(np.float32)
    out["group_grad"] = ref.group_points_grad(g(go), g(bq), 1024).cpu().numpy()
    fidx = out["fps_l1"]
    out["gather"] = ref.gather_points(g(feats), g(fidx)).cpu().numpy()
    go = rng.standard_normal(out["gather"].shape).astype(np.float32)
    out["gather_grad"] = ref.gather_points_grad(g(go), g(fidx), 1024).cpu().numpy()
    # interpolation 256 known -> 1024 unknown
    kfeats = cases.features(rng, 2, 16, 256)
    w = cases.fp_weights(out["nn_d2_l1"])
    out["interp"] = ref.three_interpolate(g(kfeats), g(out["nn_idx_l1"]), g(w)).cpu().numpy()
    go = rng.standard_normal(out["interp"].shape).astype(np.float32)
    out["interp_grad"] = ref.three_interpolate_grad(g(go), g(out["nn_idx_l1"]), g(w), 256).cpu().numpy()
    # the one known-answer test the reference ships (pointnet2_test.py:18-30)
    f = torch.arange(8, dtype=torch.float32, device=dev).view(1, 2, 4)
    kidx = torch.tensor([[[0, 1, 2], [1, 2, 3]]], dtype=torch.int32, device=dev)
    kw = torch.tensor([[[1, 1, 1], [2, 2, 2]]], dtype=torch.float32, device=dev)
    out["kat_interp"] = ref.three_interpolate(f, kidx, kw).cpu().numpy()
    return out, hashes


def main(dst):
    os.makedirs(dst, exist_ok=True)
    nofma = build_ref.load_ref("nofma")
    out, hashes = run_ops(nofma)
    np.savez_compressed(os.path.join(dst, "ref_nofma_ops.npz"), **out)

    info = {
        "generated_by": "tests/golden/make_golden.py",
        "reference": "ero1311/Pointnet2.ScanNet pointnet2/_ext_src, unmodified, nvcc -O3 -fmad=false, sm_100a",
        "gpu": torch.cuda.get_device_name(0),
        "torch": torch.__version__,
        "numpy": np.__version__,
        "input_sha256": hashes,
    }
    # informational: the stock build (fmad on) evaluates distances in a
    # compiler-chosen contracted order and may pick different indices
    try:
        stock = build_ref.load_ref("stock")
        sout, _ = run_ops(stock)
        diff = {}
        for k in out:
            if out[k].dtype == np.int32:
                diff[k] = float((out[k] != sout[k]).mean())
        info["stock_vs_nofma_index_mismatch_fraction"] = diff
    except Exception as e:  # noqa: BLE001
        info["stock_vs_nofma_index_mismatch_fraction"] = "unavailable: %s" % e
    with open(os.path.join(dst, "ref_info.json"), "w") as f:
        json.dump(info, f, indent=1, sort_keys=True)
    print("wrote", dst, {k: v.shape for k, v in list(out.items())[:4]})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
