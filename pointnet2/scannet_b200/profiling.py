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


# ---------------------------------------------------------------------------------------------
# every launch through libpn2_b200.so (the fused-block kernels are called straight from fused.py,
# not through the nine-function table): bracket the C-ABI calls themselves
# ---------------------------------------------------------------------------------------------
def _lib_bytes(name, a):
    """Algorithmic bytes (and flops) of one C-ABI call from its scalar arguments; None = not modelled.
    Same convention as SURVEY.md 8(d): inputs read once, outputs written once, 4 bytes per element."""
    if name == "pn2_mlp_gemm_fwd":
        p, k, lda, n = a[0], a[1], a[2], a[3]
        h_out, z = a[9], a[11]
        return 4 * p * lda + (4 * p * lda if h_out else 0) + (4 * p * n if z else 0) + 4 * n * k, 2 * p * k * n
    if name == "pn2_bn_stats":
        return 4 * a[0] * a[1], 0
    if name == "pn2_bn_relu_apply":
        return 8 * a[0] * a[1], 0
    if name == "pn2_bn_relu_groupmax":
        g, ns, c = a[0], a[1], a[2]
        return 4 * g * ns * c + 8 * g * c, 0
    if name == "pn2_pool_bn_bwd":
        g, ns, c = a[0], a[1], a[2]
        return 8 * g * ns * c + 12 * g * c, 0                 # read z, write dz (+ dout, argmax, z at argmax)
    if name == "pn2_relu_bn_bwd":
        return 20 * a[0] * a[1], 0                            # stats pass reads dh, z; dz pass reads both, writes dz
    if name in ("pn2_sa_group_rows", "pn2_sa_group_rows_grad"):
        b, n, npoint, ns, ct = a[:5]
        rows = b * npoint * ns
        return 4 * rows + 4 * ct * min(b * n, rows) + 4 * rows * ct, 0
    if name in ("pn2_fp_interp_rows", "pn2_fp_interp_rows_grad"):
        b, n, m, c2, c1 = a[:5]
        ct = (c2 + c1 + 3) // 4 * 4
        return 24 * b * n + 4 * b * min(m, 3 * n) * c2 + 4 * b * n * c1 + 4 * b * n * ct, 0
    if name == "pn2_furthest_point_sampling":
        return 12 * a[0] * a[1] + 4 * a[0] * a[2], 0
    if name == "pn2_query_ball_point":
        b, n, m, ns = a[0], a[1], a[2], a[4]
        return 12 * b * n + 12 * b * m + 4 * b * m * ns, 0
    if name == "pn2_three_nn":
        b, n, m = a[0], a[1], a[2]
        return 36 * b * n + 12 * b * m, 0
    if name in ("pn2_gather_points", "pn2_gather_points_grad"):
        b, c, n, m = a[:4]
        return 4 * b * m + 4 * b * c * min(n, m) + 4 * b * c * m, 0
    if name == "pn2_adam_flat":
        return 28 * a[0], 0
    if name == "pn2_mlp_bwd_layer":
        # read dh (or dout + arg-max), z, a_prev; write dh_prev when wanted; two products of 2*p*n*k flops
        p, n, k, lda, ns = a[0], a[1], a[2], a[3], a[4]
        dh, dh_prev = a[7], a[23]
        up = 4 * p * n if dh else 8 * (p // max(ns, 1)) * n
        return up + 4 * p * n + 4 * p * lda + (4 * p * lda if dh_prev else 0), (4 if dh_prev else 2) * p * n * k
    if name == "pn2_relu_bn_bwd_stats":
        return 8 * a[0] * a[1], 0
    if name == "pn2_pool_bn_bwd_stats":   # dout, arg-max and z at the arg-max rows (gathered, or kept by the forward pass)
        return 12 * a[0] * a[2], 0
    if name == "pn2_sa_lead_fwd":      # idx, u rows (or the table if smaller), coordinates; z out
        b, n, npoint, ns, c = a[:5]
        rows = b * npoint * ns
        return 4 * rows + (4 * c + 12) * min(b * n, rows) + 12 * b * npoint + 4 * rows * c, 0
    if name == "pn2_sa_lead_bwd":      # dh, z, idx, coordinates in; du out
        b, n, npoint, ns, c = a[:5]
        rows = b * npoint * ns
        return 8 * rows * c + 4 * rows + 12 * min(b * n, rows) + 12 * b * npoint + 4 * b * n * c, 0
    if name == "pn2_fp_lead_fwd":      # idx + weights, gk rows (or the table), skip; z out
        b, n, m, c, c1 = a[:5]
        return 24 * b * n + 4 * b * min(m, 3 * n) * c + 4 * b * n * c1 + 4 * b * n * c, 0
    if name == "pn2_fp_lead_bwd":      # dh, z, idx + weights, skip in; dg out
        b, n, m, c, c1 = a[:5]
        return 8 * b * n * c + 24 * b * n + 4 * b * n * c1 + 4 * b * m * c, 0
    if name == "pn2_pad_rows":
        return 4 * a[0] * (a[1] + a[2]), 0
    if name == "pn2_weighted_ce_fwd":                          # logits, targets (int64), weights; lse out
        return 4 * a[0] * a[1] + 16 * a[0], 0
    if name == "pn2_weighted_ce_bwd":                          # logits in, d logits out, targets, weights, lse
        return 8 * a[0] * a[1] + 16 * a[0], 0
    return None


class LibTimer(object):
    """Context manager: CUDA events around EVERY call into libpn2_b200.so (on the stream the kernels
    launch on).  summary() aggregates by (entry point, scalar arguments)."""

    def __init__(self):
        from . import lib
        self.lib = lib.load()
        self.records = []
        self._saved = {}

    def __enter__(self):
        from . import lib
        names = [n for n in lib.SIGNATURES if not n.endswith(("_bn", "_n_tiles", "_k_blocks", "_grid", "_nb", "_kb",
                                                              "_kpad", "_supported", "_workspace_floats"))]
        for name in names:
            fn = getattr(self.lib, name)
            self._saved[name] = fn

            def wrapped(*args, _fn=fn, _name=name):
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                st = _fn(*args)
                e1.record()
                scal = tuple(x for x in args[:6] if isinstance(x, float) or (isinstance(x, int) and 0 <= x < 2 ** 31))
                self.records.append((_name, scal, _lib_bytes(_name, args), e0, e1))
                return st
            setattr(self.lib, name, wrapped)
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(self.lib, name, fn)
        return False

    def summary(self, steps):
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for name, scal, model, e0, e1 in self.records:
            key = "%s%s" % (name, scal)
            d = agg.setdefault(key, dict(kernel=name, call=key, calls=0, ms=0.0, model=model))
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
        rows = []
        for d in agg.values():
            us = 1e3 * d["ms"] / d["calls"]
            row = dict(kernel=d["kernel"], call=d["call"], calls_per_step=d["calls"] / float(steps),
                       us_per_call=round(us, 2), ms_per_step=round(d["ms"] / steps, 4))
            if d["model"] is not None and us > 0:
                nbytes, flops = d["model"]
                row["bytes_per_call"] = int(nbytes)
                row["gbs"] = round(nbytes / (us * 1e-6) / 1e9, 1)
                if flops:
                    row["tflops"] = round(flops / (us * 1e-6) / 1e12, 2)
            rows.append(row)
        rows.sort(key=lambda r: -r["ms_per_step"])
        return rows
