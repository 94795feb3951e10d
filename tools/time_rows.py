"""CUDA-event timing of the four gather / scatter rows kernels (SA grouping, FP interpolation and
their adjoints) at the SSG level shapes, B = 32, L2 flushed before every timed launch.

    python tools/time_rows.py            # sweeps PN2_ROWS_U = 1, 2, 4 in subprocesses
    python tools/time_rows.py --one      # times the current setting only

Bytes are SURVEY.md 8(d)'s algorithmic bytes in rows layout: indices (+ weights) read, every
gathered row read once (or the table if smaller), every output row written once.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PEAK = 6536.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def main_one():
    import torch
    from pointnet2.scannet_b200 import lib as _lib
    L = _lib.load()
    dev = torch.device("cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)   # 1 GiB > L2
    B = 32
    g = torch.Generator(device="cpu").manual_seed(0)

    def timed(fn, reps=8):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    out = []
    # SA levels: (N, npoint, ns, ct)
    for (N, npnt, ns, ct) in [(8192, 1024, 32, 12), (1024, 256, 32, 68), (256, 64, 32, 132), (64, 16, 32, 260)]:
        table = torch.randn(B, N, ct, generator=g).to(dev)
        new_xyz = torch.randn(B, npnt, 3, generator=g).to(dev)
        idx = torch.randint(0, N, (B, npnt, ns), generator=g, dtype=torch.int32).to(dev)
        rows = torch.empty(B * npnt * ns, ct, device=dev)
        gt = torch.empty(B, N, ct, device=dev)
        P = B * npnt * ns
        us = timed(lambda: _lib.check(L.pn2_sa_group_rows(B, N, npnt, ns, ct, table.data_ptr(), new_xyz.data_ptr(),
                                                           idx.data_ptr(), rows.data_ptr(), s), "g"))
        by = 4 * P + 4 * min(B * N, P) * ct + 4 * P * ct
        out.append(("sa_group_rows(%d,%d,%d,%d)" % (N, npnt, ns, ct), us, by))
        us = timed(lambda: _lib.check(L.pn2_sa_group_rows_grad(B, N, npnt, ns, ct, rows.data_ptr(), idx.data_ptr(),
                                                                gt.data_ptr(), s), "gg"))
        by = 4 * P + 4 * P * ct + 4 * B * N * ct
        out.append(("sa_group_rows_grad(%d,%d,%d,%d)" % (N, npnt, ns, ct), us, by))
        if L.pn2_inverse_map_supported(N, npnt * ns):
            off = torch.empty(B, N + 1, dtype=torch.int32, device=dev)
            items = torch.empty(B, npnt * ns, dtype=torch.int32, device=dev)
            us = timed(lambda: _lib.check(L.pn2_inverse_map(B, N, npnt * ns, idx.data_ptr(), off.data_ptr(),
                                                             items.data_ptr(), s), "inv"))
            out.append(("inverse_map(%d,%d)" % (N, npnt * ns), us, 8 * P + 4 * B * N))
            us = timed(lambda: _lib.check(L.pn2_sa_group_rows_grad_inv(B, N, npnt, ns, ct, rows.data_ptr(), off.data_ptr(),
                                                                        items.data_ptr(), gt.data_ptr(), s), "ggi"))
            out.append(("sa_group_rows_grad_inv(%d,%d,%d,%d)" % (N, npnt, ns, ct), us, by))
    # FP levels: (n, m, c2, c1)
    for (n, m, c2, c1) in [(8192, 1024, 128, 6), (1024, 256, 256, 64), (256, 64, 256, 128), (64, 16, 512, 256)]:
        known = torch.randn(B, m, c2, generator=g).to(dev)
        skip = torch.randn(B, n, c1, generator=g).to(dev)
        idx = torch.randint(0, m, (B, n, 3), generator=g, dtype=torch.int32).to(dev)
        w = torch.rand(B, n, 3, generator=g).to(dev)
        ct = (c2 + c1 + 3) // 4 * 4
        rows = torch.empty(B * n, ct, device=dev)
        gk = torch.empty(B, m, c2, device=dev)
        gs = torch.empty(B, n, c1, device=dev)
        P = B * n
        us = timed(lambda: _lib.check(L.pn2_fp_interp_rows(B, n, m, c2, c1, known.data_ptr(), skip.data_ptr(),
                                                            idx.data_ptr(), w.data_ptr(), rows.data_ptr(), s), "i"))
        by = 24 * P + 4 * B * min(m, 3 * n) * c2 + 4 * P * c1 + 4 * P * ct
        out.append(("fp_interp_rows(%d,%d,%d,%d)" % (n, m, c2, c1), us, by))
        us = timed(lambda: _lib.check(L.pn2_fp_interp_rows_grad(B, n, m, c2, c1, rows.data_ptr(), idx.data_ptr(),
                                                                 w.data_ptr(), gk.data_ptr(), gs.data_ptr(), s), "ig"))
        by = 24 * P + 4 * P * ct + 4 * B * m * c2 + 4 * P * c1
        out.append(("fp_interp_rows_grad(%d,%d,%d,%d)" % (n, m, c2, c1), us, by))
        if L.pn2_inverse_map_supported(m, 3 * n):
            off = torch.empty(B, m + 1, dtype=torch.int32, device=dev)
            items = torch.empty(B, 3 * n, dtype=torch.int32, device=dev)
            us = timed(lambda: _lib.check(L.pn2_inverse_map(B, m, 3 * n, idx.data_ptr(), off.data_ptr(),
                                                             items.data_ptr(), s), "inv"))
            out.append(("inverse_map(%d,%d)" % (m, 3 * n), us, 24 * P + 4 * B * m))
            us = timed(lambda: _lib.check(L.pn2_fp_interp_rows_grad_inv(
                B, n, m, c2, c1, rows.data_ptr(), off.data_ptr(), items.data_ptr(), w.data_ptr(), gk.data_ptr(),
                gs.data_ptr(), s), "igi"))
            out.append(("fp_interp_rows_grad_inv(%d,%d,%d,%d)" % (n, m, c2, c1), us, by))
            us = timed(lambda: _lib.check(L.pn2_fp_interp_rows_grad_inv(
                B, n, m, c2, c1, rows.data_ptr(), off.data_ptr(), items.data_ptr(), w.data_ptr(), gk.data_ptr(),
                None, s), "igi"))
            out.append(("fp_interp_rows_grad_inv(%d,%d,%d,%d) no skip grad" % (n, m, c2, c1), us,
                        24 * P + 4 * P * c2 + 4 * B * m * c2))
    res = {"PN2_ROWS_U": os.environ.get("PN2_ROWS_U", "default"), "hbm_peak_gbs": PEAK, "kernels": [
        {"call": k, "us": round(us, 2), "MB": round(by / 1e6, 1), "GBs": round(by / us / 1e3, 1),
         "frac_of_hbm_peak": round(by / us / 1e3 / PEAK, 3)} for k, us, by in out]}
    for r in res["kernels"]:
        print("U=%-7s %-42s %8.1f us %8.1f MB %8.1f GB/s  %.3f" % (
            res["PN2_ROWS_U"], r["call"], r["us"], r["MB"], r["GBs"], r["frac_of_hbm_peak"]))
    print("JSON " + json.dumps(res))


if __name__ == "__main__":
    if "--one" in sys.argv:
        main_one()
    else:
        for u in ("1", "2", "4"):
            env = dict(os.environ, PN2_ROWS_U=u)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, check=False)
