"""Per-op CUDA-event timing of the `_ext` table and the algorithmic-byte model of
SURVEY.md section 8(d) (compulsory traffic: every input read once -- or only the gathered
elements if fewer -- every output written once, 4 bytes per element)."""
import collections

import torch

from . import pointnet2_utils

HBM_BOUND = ("gather_points", "gather_points_grad", "ball_query", "group_points",
             "group_points_grad", "three_nn", "three_interpolate", "three_interpolate_grad")


def algorithmic_bytes(name, args, out):
    """Bytes one call must move."""
    if name == "furthest_point_sampling":
        B, N, _ = args[0].shape
        return 12 * B * N + 4 * B * int(args[1])
    if name == "gather_points":
        B, C, N = args[0].shape
        m = args[1].shape[1]
        return 4 * B * m + 4 * B * C * min(N, m) + 4 * B * C * m
    if name == "gather_points_grad":
        B, C, m = args[0].shape
        return 4 * B * m + 4 * B * C * m + 4 * B * C * int(args[2])
    if name == "ball_query":
        B, m, _ = args[0].shape
        N = args[1].shape[1]
        return 12 * B * N + 12 * B * m + 4 * B * m * int(args[3])
    if name == "group_points":
        B, C, N = args[0].shape
        _, npoint, ns = args[1].shape
        return 4 * B * npoint * ns + 4 * B * C * min(N, npoint * ns) + 4 * B * C * npoint * ns
    if name == "group_points_grad":
        B, C, npoint, ns = args[0].shape
        return 4 * B * npoint * ns + 4 * B * C * npoint * ns + 4 * B * C * int(args[2])
    if name == "three_nn":
        B, n, _ = args[0].shape
        m = args[1].shape[1]
        return 12 * B * n + 12 * B * m + 24 * B * n
    if name == "three_interpolate":
        B, C, m = args[0].shape
        n = args[1].shape[1]
        return 24 * B * n + 4 * B * C * min(m, 3 * n) + 4 * B * C * n
    if name == "three_interpolate_grad":
        B, C, n = args[0].shape
        return 24 * B * n + 4 * B * C * n + 4 * B * C * int(args[3])
    raise KeyError(name)


def _shape_key(name, args):
    parts = []
    for a in args:
        parts.append("x".join(map(str, a.shape)) if torch.is_tensor(a) else repr(a))
    return name + "(" + ", ".join(parts) + ")"


class ExtTimer(object):
    """Context manager: while active, every call through ``pointnet2_utils._ext`` is
    bracketed by CUDA events on the current stream (the stream the kernels launch on)."""

    def __init__(self):
        self.records = []
        self._orig = None

    def __enter__(self):
        self._orig = pointnet2_utils._ext
        timer, orig = self, self._orig

        class Proxy(object):
            def __getattr__(self, name):
                fn = getattr(orig, name)

                def wrapped(*args):
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    out = fn(*args)
                    e1.record()
                    timer.records.append((name, _shape_key(name, args),
                                          algorithmic_bytes(name, args, out), e0, e1))
                    return out
                return wrapped

        pointnet2_utils._ext = Proxy()
        return self

    def __exit__(self, *exc):
        pointnet2_utils._ext = self._orig
        return False

    def summary(self, steps):
        """-> list of dicts per (op, shape): calls/step, us per call, bytes, GB/s; sorted by time."""
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for name, key, nbytes, e0, e1 in self.records:
            d = agg.setdefault(key, dict(op=name, call=key, calls=0, ms=0.0, bytes=nbytes))
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
        rows = []
        for d in agg.values():
            us = 1e3 * d["ms"] / d["calls"]
            rows.append(dict(op=d["op"], call=d["call"], calls_per_step=d["calls"] / float(steps),
                             us_per_call=round(us, 2), bytes_per_call=d["bytes"],
                             gbs=round(d["bytes"] / (us * 1e-6) / 1e9, 1) if us > 0 else None,
                             ms_per_step=round(d["ms"] / steps, 4)))
        rows.sort(key=lambda r: -r["ms_per_step"])
        return rows
