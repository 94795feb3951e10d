"""Fused shared-MLP blocks in position-major ("rows") layout.

What the reference's SA / FP modules compute around their 1x1 convolutions
(/root/reference/pointnet2/pointnet2_modules.py:60-75 and 395-416: QueryAndGroup ->
SharedMLP [Conv2d 1x1 + BatchNorm2d + ReLU]* -> max_pool2d over nsample), restated on
[P, C] row-major activations so that each conv is one GEMM and BatchNorm / ReLU / max /
their adjoints are the streaming kernels of csrc/mlp_rows.cu (C ABI: include/pn2_b200.h,
"Shared-MLP block").  The module parameters and BatchNorm buffers are used as they are
(no new parameters: state dicts stay the reference's); running statistics and
num_batches_tracked are updated exactly like torch.nn.BatchNorm2d does in train mode.

Numerics: BatchNorm statistics fp32 partials + fp64 finalise; every 1x1 convolution and both of its
adjoint products run on the hand-written tcgen05 kernels (csrc/mlp_gemm.cu, mlp_bwd.cu, mlp_wgrad.cu)
with TF32 operands and fp32 accumulation -- the arithmetic class of the reference's own cuDNN convs,
whose ``cudnn.allow_tf32`` defaults to True.  There is no library GEMM on this path.

In training mode the FIRST conv of a block does not see gathered rows at all: it is linear, so it is applied to
the source points and the gather / interpolation kernel writes its output directly (``SALead`` / ``FPLead``,
csrc/mlp_lead.cu); the rows kernels (``SAGroupRows`` / ``FPInterpRows``) remain the inference / fallback path.
"""
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from . import lib as _lib

ENABLED = True     # module forward()s take the fused path when eligible
TC_BWD = True      # backward layers on the ONE-PASS fused tcgen05 kernel (csrc/mlp_bwd.cu: BN adjoint + dH +
                   # dW + next statistics) where the shape fits it (N <= 128, lda <= 160); the wider layers --
                   # and every layer when this is off (tests) -- take the split adjoint: dZ written once by a
                   # streaming kernel, then the data-gradient product on the forward layer kernel and the
                   # weight-gradient product on csrc/mlp_wgrad.cu
POOL_IN_GEMM = True  # pooled last layer: group extrema from the tensor-core layer's epilogue (no pass over Z)


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


_sched = {}


def _sched_counters(dev):
    """Two zeroed ints per (device, stream): the tile counter of the persistent kernels' dynamic
    scheduler (the kernels re-arm it themselves; kernels of one stream run one after another)."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    buf = _sched.get(key)
    if buf is None:
        buf = _sched[key] = torch.zeros(64, dtype=torch.int32, device=dev)
    return buf


# ---- weight-tile packing beside the step ------------------------------------------------------------
# The tensor-core layers read their weights as pre-packed UMMA tile images (pn2_mlp_pack_weights[_t]):
# ~33 tiny launches per SSG step that depend on nothing but the weights.  Inside ``pack_branch`` they
# run on a forked side stream (a parallel branch of the captured step graph) into PERSISTENT tile
# buffers -- persistent because a buffer from the per-step allocator could be a block whose previous
# tenant is still in flight on the main stream -- and the consuming layer waits for its own pack event.
PACK_BRANCH = os.environ.get("PN2_PACK_BRANCH", "1") != "0"
_pack_state = {"stream": None, "active": False, "bufs": {}}


class pack_branch(object):
    """with pack_branch(dev): forward pass -- fork at entry, join at exit (capture-legal)."""

    def __init__(self, dev):
        self.dev = torch.device(dev)

    def __enter__(self):
        self.on = PACK_BRANCH and ENABLED and self.dev.type == "cuda" and not _pack_state["active"]
        if self.on:
            if _pack_state["stream"] is None:
                _pack_state["stream"] = torch.cuda.Stream(self.dev)
            _pack_state["stream"].wait_stream(torch.cuda.current_stream(self.dev))
            _pack_state["active"] = True
            _pack_state["n_drop"] = 0
            _pack_state["n_zero"] = 0
        return self

    def __exit__(self, *exc):
        if self.on:
            _pack_state["active"] = False
            torch.cuda.current_stream(self.dev).wait_stream(_pack_state["stream"])
        return False


def _packed(kind, W2, dims, floats, launch):
    """Tile image of W2 for this step: launch(ptr, stream) packs it.  Inside a pack_branch the launch
    goes to the branch stream and the current stream waits for it; outside, it is an ordinary launch
    into a per-step buffer."""
    dev = W2.device
    if not _pack_state["active"]:
        buf = torch.empty(floats, dtype=torch.float32, device=dev)
        launch(buf.data_ptr(), _stream(dev))
        return buf
    key = (kind, W2.data_ptr(), dims, floats)
    buf = _pack_state["bufs"].get(key)
    if buf is None:
        if torch.cuda.is_current_stream_capturing():
            # never allocate a persistent buffer from a graph's private pool: this layer packs in line
            # (callers that want the branch inside a graph run one eager step first, as the trainers do)
            buf = torch.empty(floats, dtype=torch.float32, device=dev)
            launch(buf.data_ptr(), _stream(dev))
            return buf
        buf = _pack_state["bufs"][key] = torch.empty(floats, dtype=torch.float32, device=dev)
    side = _pack_state["stream"]
    launch(buf.data_ptr(), side.cuda_stream)
    ev = torch.cuda.Event()
    ev.record(side)
    torch.cuda.current_stream(dev).wait_event(ev)
    return buf


# ---- weight gradients beside the backward pass -------------------------------------------------------
# dW of a layer is needed by nobody before the optimiser, while dH_prev is needed by the next kernel.  Inside
# ``dw_branch`` (the trainers wrap loss.backward() in it) the dW work of the layers that do not run on the
# fused backward kernel -- the split-K cuBLAS product dZ^T H and its partial sum -- and the dW-only partial
# sums of the fused layers go to a forked side stream and are joined once, before the gradients are
# packed.  Their operands are kept referenced until the join (no allocator reuse under the branch's feet).
DW_BRANCH = os.environ.get("PN2_DW_BRANCH", "1") != "0"
_dw_state = {"stream": None, "active": False, "keep": [], "warmed": False}


class dw_branch(object):
    def __init__(self, dev):
        self.dev = torch.device(dev)

    def __enter__(self):
        self.on = DW_BRANCH and ENABLED and self.dev.type == "cuda" and not _dw_state["active"]
        if self.on and torch.cuda.is_current_stream_capturing() and not _dw_state["warmed"]:
            self.on = False   # first use must be eager (cuBLAS handle / workspace of the branch stream)
        if self.on:
            if _dw_state["stream"] is None:
                _dw_state["stream"] = torch.cuda.Stream(self.dev)
            _dw_state["active"] = True
        return self

    def __exit__(self, *exc):
        if self.on:
            _dw_state["active"] = False
            torch.cuda.current_stream(self.dev).wait_stream(_dw_state["stream"])
            _dw_state["keep"].clear()
            if not torch.cuda.is_current_stream_capturing():
                _dw_state["warmed"] = True
        return False


def _on_dw_branch(dev, operands, fn):
    """fn() on the dW branch stream, after everything queued on the current stream so far."""
    if not _dw_state["active"]:
        return fn()
    side = _dw_state["stream"]
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        out = fn()
    _dw_state["keep"].append((operands, out))
    return out


def _ws(c, dev):
    n = _lib.load().pn2_bn_stats_workspace_floats(c) + 2 * c
    return torch.empty(n, dtype=torch.float32, device=dev)


class _BNLayer(object):
    """Per-layer state shared between forward and backward."""
    __slots__ = ("z", "a", "b", "mean", "invstd", "cout", "fold4")


def _bn_forward(z, gamma, beta, rm, rv, eps, momentum, training, nbt=None):
    """-> (a, b, mean, invstd); updates rm / rv / num_batches_tracked in place when training."""
    L = _lib.load()
    P, C = z.shape
    dev = z.device
    a, b, mean, invstd = torch.empty(4, C, dtype=torch.float32, device=dev).unbind(0)   # rows of one [4, C] buffer
    if training:
        _lib.check(L.pn2_bn_stats(P, C, z.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, momentum,
                                  rm.data_ptr() if rm is not None else None,
                                  rv.data_ptr() if rv is not None else None,
                                  nbt.data_ptr() if nbt is not None else None,
                                  a.data_ptr(), b.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                  _ws(C, dev).data_ptr(), _stream(dev)), "bn_stats")
    else:
        _lib.check(L.pn2_bn_fold_eval(C, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                      eps, a.data_ptr(), b.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                                      _stream(dev)), "bn_fold_eval")
    return a, b, mean, invstd


def _tc_layer(A, k, W2, fold, relu_in, want_h, want_stats, drop=None, pool_ns=0):
    """One layer on the tcgen05 kernel: z = act(A)[:, :k] @ W2^T.  fold = (a_in, b_in) or None;
    drop = (keep mask uint8 [P, k], scale) of a dropout between act(A) and this layer, or None;
    pool_ns > 0: also the per-group extrema of z (zmax, zmin, imax, imin), see pn2_mlp_gemm_fwd.
    -> (z, h_out or None, partials or None, (ld, nblk)[, extrema])"""
    L = _lib.load()
    dev = A.device
    P, lda = A.shape
    N = W2.shape[0]
    s = _stream(dev)
    bn, nt, kb = L.pn2_mlp_gemm_bn(P, N), L.pn2_mlp_gemm_n_tiles(P, N), L.pn2_mlp_gemm_k_blocks(k)
    grid = L.pn2_mlp_gemm_grid(P, N)
    W2 = W2.contiguous()
    Wp = _packed("fwd", W2, (P, N, k), nt * kb * bn * 32, lambda ptr, st: _lib.check(
        L.pn2_mlp_pack_weights(P, N, k, W2.shape[1], W2.data_ptr(), ptr, st), "mlp_pack_weights"))
    z = torch.empty(P, N, dtype=torch.float32, device=dev)
    h = torch.empty_like(A) if (want_h and fold is not None) else None
    part = torch.empty(grid, 2, nt * bn, dtype=torch.float32, device=dev) if want_stats else None
    ext = None
    if pool_ns:
        G = P // pool_ns
        ext = (torch.empty(G, N, dtype=torch.float32, device=dev), torch.empty(G, N, dtype=torch.float32, device=dev),
               torch.empty(G, N, dtype=torch.int32, device=dev), torch.empty(G, N, dtype=torch.int32, device=dev))
    _lib.check(L.pn2_mlp_gemm_fwd(
        P, k, lda, N, N, A.data_ptr(),
        fold[0].data_ptr() if fold is not None else None,
        fold[1].data_ptr() if fold is not None else None, int(relu_in),
        h.data_ptr() if h is not None else None, Wp.data_ptr(), z.data_ptr(),
        part.data_ptr() if part is not None else None,
        drop[0].data_ptr() if drop is not None else None, float(drop[1]) if drop is not None else 1.0,
        int(pool_ns), *((t.data_ptr() for t in ext) if ext is not None else (None, None, None, None)), s),
        "mlp_gemm_fwd")
    if pool_ns:
        return z, h, part, (nt * bn, grid), ext
    return z, h, part, (nt * bn, grid)


def _pack_dgrad(P, n_out, W2c):
    """Tile images of W^T for dH_prev[P, n_out] = dZ[P, N] @ W2c[N, K] on the forward layer kernel
    (columns K..n_out-1, the zero padding of the rows layout, get zero weights)."""
    L = _lib.load()
    N, K = W2c.shape
    bn, nt, kb = L.pn2_mlp_gemm_bn(P, n_out), L.pn2_mlp_gemm_n_tiles(P, n_out), L.pn2_mlp_gemm_k_blocks(N)
    return _packed("dgrad", W2c, (P, n_out, N, K), nt * kb * bn * 32, lambda ptr, st: _lib.check(
        L.pn2_mlp_pack_weights_dgrad(P, n_out, N, K, W2c.shape[1], W2c.data_ptr(), ptr, st),
        "mlp_pack_weights_dgrad"))


def _dgrad(dz, W2c, n_out, wp=None):
    """dH_prev [P, n_out] = dz [P, N] @ W2c [N, K] (zero columns past K): the data-gradient product of a
    wide layer, on the tcgen05 forward layer kernel with the transposed weight tiles."""
    L = _lib.load()
    dev = dz.device
    P, N = dz.shape
    if wp is None:
        wp = _pack_dgrad(P, n_out, W2c)
    out = torch.empty(P, n_out, dtype=torch.float32, device=dev)
    _lib.check(L.pn2_mlp_gemm_fwd(P, N, N, n_out, n_out, dz.data_ptr(), None, None, 0, None, wp.data_ptr(),
                                  out.data_ptr(), None, None, 1.0, 0, None, None, None, None, _stream(dev)),
               "mlp_gemm_fwd (dgrad)")
    return out


# BatchNorm-adjoint statistics of the layer below from the data-gradient product's epilogue (pn2_mlp_dgrad_stats)
# instead of a separate pass over (dH_prev, Z_prev).  Measured in one call (round 2): it removes 6 of the 11 statistics
# launches of an SSG step but the epilogue's per-thread row reads of Z_prev slow the products down by more:
# 3.68 vs 3.64 ms per step -- off by default, kept for the tests that pin its arithmetic.
DGRAD_STATS = os.environ.get("PN2_DGRAD_STATS", "0") == "1"


def _dgrad_stats(dz, W2c, n_out, wp, prev, relu_prev, training):
    """_dgrad plus (dgamma, dbeta, c1c2) of the layer below (``prev``: its z and [a | b | mean | invstd]) from the
    product's epilogue: column sums of dY_prev and dY_prev * zhat_prev per CTA, finalised in fixed order."""
    L = _lib.load()
    dev = dz.device
    P, N = dz.shape
    s = _stream(dev)
    if wp is None:
        wp = _pack_dgrad(P, n_out, W2c)
    bn, nt, grid = L.pn2_mlp_gemm_bn(P, n_out), L.pn2_mlp_gemm_n_tiles(P, n_out), L.pn2_mlp_gemm_grid(P, n_out)
    out = torch.empty(P, n_out, dtype=torch.float32, device=dev)
    part = torch.empty(grid, 2, nt * bn, dtype=torch.float32, device=dev)
    _lib.check(L.pn2_mlp_dgrad_stats(P, N, n_out, dz.data_ptr(), wp.data_ptr(), out.data_ptr(), prev.z.data_ptr(),
                                     prev.fold4.data_ptr(), relu_prev, part.data_ptr(), s), "mlp_dgrad_stats")
    Kc = prev.cout
    pg, pb, pc = (torch.empty(m, dtype=torch.float32, device=dev) for m in (Kc, Kc, 2 * Kc))
    _lib.check(L.pn2_bwd_stats_finalize(P, Kc, nt * bn, grid, training, part.data_ptr(), pg.data_ptr(), pb.data_ptr(),
                                        pc.data_ptr(), s), "bwd_stats_finalize")
    return out, (pg, pb, pc)


def _wgrad(dz, h, K):
    """(dz^T @ h)[:, :K] -> dense [N, K]: csrc/mlp_wgrad.cu, row-split partial images summed in fixed order
    straight into the parameter's layout (autograd takes the tensor as .grad without a copy)."""
    L = _lib.load()
    dev = dz.device
    P, N = dz.shape
    ldy = h.shape[1]
    S, R, Cc = L.pn2_mlp_wgrad_splits(P, N, ldy), L.pn2_mlp_wgrad_rows(N), L.pn2_mlp_wgrad_cols(ldy)
    part = torch.empty(S, R, Cc, dtype=torch.float32, device=dev)
    _lib.check(L.pn2_mlp_wgrad(P, N, N, ldy, dz.data_ptr(), h.data_ptr(), part.data_ptr(), _stream(dev)), "mlp_wgrad")
    out = torch.empty(N, K, dtype=torch.float32, device=dev)
    _lib.check(L.pn2_sum_partials_block(S, R, Cc, N, K, part.data_ptr(), out.data_ptr(), _stream(dev)),
               "sum_partials_block")
    return out


def _bn_finalize(P, C, part, ld, nblk, gamma, beta, rm, rv, eps, momentum, nbt=None):
    L = _lib.load()
    dev = part.device
    a, b, mean, invstd = torch.empty(4, C, dtype=torch.float32, device=dev).unbind(0)   # rows of one [4, C] buffer
    _lib.check(L.pn2_bn_finalize(P, C, ld, nblk, part.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps,
                                 momentum, rm.data_ptr() if rm is not None else None,
                                 rv.data_ptr() if rv is not None else None,
                                 nbt.data_ptr() if nbt is not None else None, a.data_ptr(), b.data_ptr(),
                                 mean.data_ptr(), invstd.data_ptr(), _stream(dev)), "bn_finalize")
    return a, b, mean, invstd


# ---- first layer of a block applied BEFORE the gather (csrc/mlp_lead.cu) ------------------------------
# A block's first 1x1 convolution acts on gathered rows (QueryAndGroup / three_interpolate + cat,
# /root/reference/pointnet2/pointnet2_modules.py:60-75, 395-412); it is linear, so it commutes with the gather:
# the product runs over the N source points (4-8x fewer rows than gathered positions) and the gather kernel
# writes the layer's pre-BatchNorm output directly -- the gathered rows tensor is never written or re-read,
# forward or backward.  Training mode only: inference keeps the layer-by-layer / one-kernel paths (whose
# results are bit-identical to each other).
LEAD = os.environ.get("PN2_LEAD", "1") != "0"


def _prezeroed(rows, cols, dev):
    """A zero [rows, cols] reduction target for the BACKWARD pass, zero-filled now -- during the forward pass, on the
    packing branch -- so that the memset is off the backward pass's critical path (5 memsets, ~18 us per SSG step).
    Persistent buffers, one per call site of the step (as the dropout masks); None outside a pack_branch."""
    if not _pack_state["active"]:
        return None
    n = _pack_state["n_zero"] = _pack_state.get("n_zero", 0) + 1
    key = ("zero", n, rows, cols, dev)
    buf = _pack_state["bufs"].get(key)
    if buf is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        buf = _pack_state["bufs"][key] = torch.empty(rows, cols, dtype=torch.float32, device=dev)
    with torch.cuda.stream(_pack_state["stream"]):
        buf.zero_()
    return buf      # the branch is joined when the forward pass ends (pack_branch.__exit__)


def _gemm_plain(A, k, wp, n_out):
    """[P, n_out] = A[:, :k] @ (packed weight tiles)^T on the tcgen05 layer kernel: no fold, no statistics."""
    L = _lib.load()
    P, lda = A.shape
    out = torch.empty(P, n_out, dtype=torch.float32, device=A.device)
    _lib.check(L.pn2_mlp_gemm_fwd(P, k, lda, n_out, n_out, A.data_ptr(), None, None, 0, None, wp.data_ptr(),
                                  out.data_ptr(), None, None, 1.0, 0, None, None, None, None, _stream(A.device)),
               "mlp_gemm_fwd (lead)")
    return out


def _sum_partials(part):
    """[S, ...] -> [...]: fixed-order sum over the leading axis (pn2_sum_partials)."""
    out = torch.empty(part.shape[1:], dtype=torch.float32, device=part.device)
    _lib.check(_lib.load().pn2_sum_partials(part.shape[0], out.numel(), part.data_ptr(), out.data_ptr(),
                                            _stream(part.device)), "sum_partials")
    return out


class SALead(object):
    """QueryAndGroup + the first conv of an SA scale: z[g, s] = U[idx[g, s]] + W_x (xyz[idx[g, s]] - new_xyz[g]),
    U = F W_f^T over the points; the coordinate term is evaluated in fp32 on the centred differences.  The
    differentiable source ``x0`` (B, N, ld) holds the point features in columns [c_off, c_off + Cf): the feature
    rows themselves (c_off = 0, ld = Cf; levels whose Cf is a multiple of 4) or the [xyz | features | pad] table
    (c_off = 3; the coordinate columns then get zero weights and a zero gradient)."""

    def __init__(self, xyz, new_xyz, idx, c_off):
        # xyz may be a column view of a wider point table (row stride ldx): read in place, no copy
        if not (xyz.stride(2) == 1 and xyz.stride(0) == xyz.shape[1] * xyz.stride(1) and xyz.stride(1) >= 3
                and xyz.data_ptr() % 4 == 0):
            xyz = xyz.contiguous()
        self.xyz, self.ldx, self.new_xyz, self.idx = xyz, xyz.stride(1), new_xyz.contiguous(), idx
        self.B, self.N = xyz.shape[0], xyz.shape[1]
        self.npoint, self.ns = idx.shape[1], idx.shape[2]
        self.P = self.B * self.npoint * self.ns
        self.c_off = c_off
        self.mark = False

    def forward(self, x0, W2, need_dx):
        L = _lib.load()
        dev = x0.device
        B, N, ld = x0.shape
        C, K3 = W2.shape                        # K3 = 3 + Cf
        PN, c_off = B * N, self.c_off
        W2c = W2.contiguous()
        ldw = W2c.shape[1]
        self.x2d, self.W2c = x0.reshape(PN, ld), W2c
        self.K = K = c_off + K3 - 3             # columns of x2d the product reads
        assert K <= ld
        # column k >= c_off of x2d meets W2[:, 3 + (k - c_off)]: the weight pointer is shifted accordingly
        self.w_ptr = W2c.data_ptr() + 4 * (3 - c_off)
        bn, nt, kb = L.pn2_mlp_gemm_bn(PN, C), L.pn2_mlp_gemm_n_tiles(PN, C), L.pn2_mlp_gemm_k_blocks(K)
        w_ptr = self.w_ptr
        wp = _packed("lead_sa", W2c, (PN, C, K, c_off), nt * kb * bn * 32, lambda ptr, st: _lib.check(
            L.pn2_mlp_pack_weights_cols(PN, C, K, c_off, ldw, w_ptr, ptr, st), "mlp_pack_weights_cols"))
        self.wp_dgrad = None
        if need_dx and _pack_state["active"]:
            self.wp_dgrad = self._pack_dgrad()
        self.du0 = _prezeroed(PN, C, dev) if getattr(self, "want_grad", False) else None
        u = _gemm_plain(self.x2d, K, wp, C)
        nblk = L.pn2_lead_blocks(self.P, C, 0)
        z = torch.empty(self.P, C, dtype=torch.float32, device=dev)
        part = torch.empty(nblk, 2, C, dtype=torch.float32, device=dev)
        _lib.check(L.pn2_sa_lead_fwd(B, N, self.npoint, self.ns, C, u.data_ptr(), self.xyz.data_ptr(), self.ldx,
                                     self.new_xyz.data_ptr(), self.idx.data_ptr(), W2c.data_ptr(), ldw,
                                     z.data_ptr(), part.data_ptr(), _stream(dev)), "sa_lead_fwd")
        return z, part, (C, nblk)

    def _pack_dgrad(self):
        L = _lib.load()
        W2c, K, c_off, w_ptr = self.W2c, self.K, self.c_off, self.w_ptr
        PN, ld = self.x2d.shape
        C = W2c.shape[0]
        bn, nt, kb = L.pn2_mlp_gemm_bn(PN, ld), L.pn2_mlp_gemm_n_tiles(PN, ld), L.pn2_mlp_gemm_k_blocks(C)
        return _packed("lead_sa_dgrad", W2c, (PN, ld, C, K, c_off), nt * kb * bn * 32, lambda ptr, st: _lib.check(
            L.pn2_mlp_pack_weights_dgrad_cols(PN, ld, C, K, c_off, W2c.shape[1], w_ptr, ptr, st),
            "mlp_pack_weights_dgrad_cols"))

    def backward(self, dh, st, c1c2, relu, need_dx):
        """-> (dW [C, 3 + Cf] dense, d x0 (B, N, ld) or None)"""
        L = _lib.load()
        dev = dh.device
        C, K, c_off = self.W2c.shape[0], self.K, self.c_off
        PN, ld = self.x2d.shape
        nblk = L.pn2_lead_blocks(self.P, C, 1)
        du = self.du0 if self.du0 is not None else torch.empty(PN, C, dtype=torch.float32, device=dev)
        part = torch.empty(nblk, 3, C, dtype=torch.float32, device=dev)
        _lib.check(L.pn2_sa_lead_bwd(self.B, self.N, self.npoint, self.ns, C, relu, dh.data_ptr(), st.z.data_ptr(),
                                     st.a.data_ptr(), st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(),
                                     c1c2.data_ptr(), self.xyz.data_ptr(), self.ldx, self.new_xyz.data_ptr(),
                                     self.idx.data_ptr(), du.data_ptr(), int(self.du0 is None), part.data_ptr(),
                                     _stream(dev)), "sa_lead_bwd")
        dx = None
        if need_dx:   # the data gradient first: the next kernel of the chain waits for it
            wp = self.wp_dgrad if self.wp_dgrad is not None else self._pack_dgrad()
            dx = _gemm_plain(du, C, wp, ld).view(self.B, self.N, ld)
        x2d = self.x2d

        def weight_grad():
            wg = _wgrad(du, x2d, K)              # [C, c_off + Cf]: columns >= c_off are dW_f
            return torch.cat([_sum_partials(part).t(), wg[:, c_off:]], dim=1)
        dW = _on_dw_branch(dev, (du, x2d, part), weight_grad)
        return dW, dx


class FPLead(object):
    """three_interpolate + cat + the first conv of an FP block: z[r] = sum_t w[r, t] G[idx[r, t]] + W_s skip[r],
    G = known W_k^T over the m known points; the (<= pn2_fp_lead_max_skip()) skip channels in fp32."""

    def __init__(self, skip, idx, weight, mark=False):
        # skip may be a column view of a wider row tensor (the network's input): read in place, no copy
        if skip is not None and not (skip.stride(2) == 1 and skip.stride(0) == skip.shape[1] * skip.stride(1)
                                     and skip.stride(1) >= skip.shape[2] and skip.data_ptr() % 4 == 0):
            skip = skip.contiguous()
        self.skip, self.lds = skip, (skip.stride(1) if skip is not None else 0)
        self.idx, self.weight, self.mark = idx, weight.contiguous(), mark
        self.B, self.n = idx.shape[0], idx.shape[1]
        self.c1 = skip.shape[2] if skip is not None else 0
        self.P = self.B * self.n

    def forward(self, known, W2, need_dx):
        L = _lib.load()
        dev = known.device
        B, m, c2 = known.shape
        C = W2.shape[0]
        PM = B * m
        W2c = W2.contiguous()
        ldw = W2c.shape[1]
        assert ldw == c2 + self.c1
        self.known2d, self.W2c, self.m, self.c2 = known.reshape(PM, c2), W2c, m, c2
        bn, nt, kb = L.pn2_mlp_gemm_bn(PM, C), L.pn2_mlp_gemm_n_tiles(PM, C), L.pn2_mlp_gemm_k_blocks(c2)
        wp = _packed("lead_fp", W2c, (PM, C, c2), nt * kb * bn * 32, lambda ptr, st: _lib.check(
            L.pn2_mlp_pack_weights(PM, C, c2, ldw, W2c.data_ptr(), ptr, st), "mlp_pack_weights"))
        self.wp_dgrad = None
        if need_dx and _pack_state["active"]:
            self.wp_dgrad = self._pack_dgrad()
        self.dg0 = _prezeroed(PM, C, dev) if getattr(self, "want_grad", False) else None
        gk = _gemm_plain(self.known2d, c2, wp, C)
        nblk = L.pn2_lead_blocks(self.P, C, 2)
        z = torch.empty(self.P, C, dtype=torch.float32, device=dev)
        part = torch.empty(nblk, 2, C, dtype=torch.float32, device=dev)
        _lib.check(L.pn2_fp_lead_fwd(B, self.n, m, C, self.c1, gk.data_ptr(),
                                     self.skip.data_ptr() if self.skip is not None else None, self.lds,
                                     self.idx.data_ptr(), self.weight.data_ptr(), W2c.data_ptr() + 4 * c2, ldw, z.data_ptr(),
                                     part.data_ptr(), _stream(dev)), "fp_lead_fwd")
        return z, part, (C, nblk)

    def _pack_dgrad(self):
        L = _lib.load()
        W2c, c2 = self.W2c, self.c2
        PM = self.known2d.shape[0]
        C = W2c.shape[0]
        bn, nt, kb = L.pn2_mlp_gemm_bn(PM, c2), L.pn2_mlp_gemm_n_tiles(PM, c2), L.pn2_mlp_gemm_k_blocks(C)
        return _packed("lead_fp_dgrad", W2c, (PM, c2, C), nt * kb * bn * 32, lambda ptr, st: _lib.check(
            L.pn2_mlp_pack_weights_dgrad(PM, c2, C, c2, W2c.shape[1], W2c.data_ptr(), ptr, st),
            "mlp_pack_weights_dgrad"))

    def backward(self, dh, st, c1c2, relu, need_dx):
        """-> (dW [C, c2 + c1] dense, d known (B, m, c2) or None)"""
        L = _lib.load()
        dev = dh.device
        C, c2, c1 = self.W2c.shape[0], self.c2, self.c1
        PM = self.known2d.shape[0]
        nblk = L.pn2_lead_blocks(self.P, C, 3)
        dg = self.dg0 if self.dg0 is not None else torch.empty(PM, C, dtype=torch.float32, device=dev)
        part = torch.empty(nblk, L.pn2_fp_lead_max_skip(), C, dtype=torch.float32, device=dev)
        _lib.check(L.pn2_fp_lead_bwd(self.B, self.n, self.m, C, c1, relu, dh.data_ptr(), st.z.data_ptr(),
                                     st.a.data_ptr(), st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(),
                                     c1c2.data_ptr(), self.skip.data_ptr() if self.skip is not None else None, self.lds,
                                     self.idx.data_ptr(), self.weight.data_ptr(), dg.data_ptr(), int(self.dg0 is None),
                                     part.data_ptr(), _stream(dev)), "fp_lead_bwd")
        dknown = None
        if need_dx:
            wp = self.wp_dgrad if self.wp_dgrad is not None else self._pack_dgrad()
            dknown = _gemm_plain(dg, C, wp, c2).view(self.B, self.m, c2)
        if self.mark and BWD_MARK_EVENT is not None:
            BWD_MARK_EVENT.record(torch.cuda.current_stream(dev))
        known2d = self.known2d

        def weight_grad():
            dWk = _wgrad(dg, known2d, c2)        # [C, c2]
            if c1 == 0:
                return dWk
            return torch.cat([dWk, _sum_partials(part)[:c1].t()], dim=1)
        dW = _on_dw_branch(dev, (dg, known2d, part), weight_grad)
        return dW, dknown


class SharedMLPRows(Function):
    """x0 [P, Cin] -> out [P/ns, C_L] (max over each run of ns rows) or [P, C_L] (ns == 0).

    args: x0, ns, training, eps, momentum, buffers, relus, drops, then (W, gamma, beta) per layer.
    ``drops``: per layer None or (keep mask uint8 [P, C_l], scale): dropout on that layer's activated
    output, applied inside the NEXT layer's kernels (tensor-core path with fused backward only).
    ``buffers`` is a list of (running_mean, running_var, num_batches_tracked) per layer (updated
    in place, as torch.nn.BatchNorm does in train mode);
    ``relus`` one flag per layer (the pooled last layer always has its ReLU).
    """

    @staticmethod
    def forward(ctx, x0, ns, training, eps, momentum, buffers, relus, drops, lead, *params):
        L = _lib.load()
        dev = x0.device
        nl = len(params) // 3
        P = x0.shape[0] if lead is None else lead.P   # lead: x0 is the SOURCE tensor (point table / known rows)
        layers, hs = [], [x0]
        out = argmax = zsel = None
        want_grad = any(ctx.needs_input_grad)
        bwd_fused = [False] * nl
        wt_packed = [None] * nl   # transposed tile images for the fused backward layers, packed beside the forward pass
        with torch.cuda.device(dev):
            for l in range(nl):
                W, gamma, beta = params[3 * l:3 * l + 3]
                W2 = W.reshape(W.shape[0], -1)
                if l == 0 and lead is not None:
                    # the block's first conv, applied before the gather (SALead / FPLead): z and its statistics
                    st = _BNLayer()
                    st.cout = W2.shape[0]
                    rm, rv, nbt = buffers[0]
                    lead.want_grad = want_grad     # (a backward pass will follow: reserve its reduction target now)
                    st.z, part, (ld, nblk) = lead.forward(x0, W2, ctx.needs_input_grad[0])
                    st.a, st.b, st.mean, st.invstd = _bn_finalize(P, st.cout, part, ld, nblk, gamma, beta,
                                                                  rm, rv, eps, momentum, nbt)
                    st.fold4 = st.a._base if st.a._base is not None else None
                    layers.append(st)
                    continue
                lda = x0.shape[1] if l == 0 else layers[l - 1].cout
                bwd_fused[l] = bool(TC_BWD and want_grad and
                                    L.pn2_mlp_bwd_supported(P, W2.shape[0], W2.shape[1], lda, int(l > 0)) and
                                    not (l == nl - 1 and ns > 0 and ns % 4 != 0))
                want_h = want_grad and not bwd_fused[l]   # the fused adjoint recomputes H_{l-1} from Z_{l-1}
                st = _BNLayer()
                st.cout = W2.shape[0]
                rm, rv, nbt = buffers[l]
                # tensor-core layer: the previous layer's BN fold + ReLU is applied to the A tile
                # on chip (its normalised output is only written out, once, where the split adjoint's
                # weight-gradient product needs it); this layer's BN statistics come out of the epilogue
                prev = layers[l - 1] if l > 0 else None
                A = prev.z if prev is not None else x0
                fold = (prev.a, prev.b) if prev is not None else None
                drop = drops[l - 1] if l > 0 else None
                assert drop is None or (bwd_fused[l] and bwd_fused[l - 1]), "dropout needs the fused backward layers"
                # pooled last layer: the group extrema of z come out of the same kernel
                ext = None
                pool_here = (l == nl - 1 and ns > 0 and POOL_IN_GEMM and P % ns == 0
                             and bool(L.pn2_mlp_gemm_pool_supported(ns)))
                res = _tc_layer(A, W2.shape[1], W2, fold, relus[l - 1] if l > 0 else False, want_h, training,
                                drop, ns if pool_here else 0)
                st.z, h, part, (ld, nblk) = res[:4]
                if pool_here:
                    ext = res[4]
                if _pack_state["active"] and want_grad and (l > 0 or ctx.needs_input_grad[0]):
                    W2c, Cn, Kk = W2.contiguous(), W2.shape[0], W2.shape[1]
                    if bwd_fused[l]:
                        wt_packed[l] = _packed(
                            "bwd", W2c, (Cn, Kk, lda), L.pn2_mlp_bwd_nb(Cn) * L.pn2_mlp_bwd_kpad(lda) * 32,
                            lambda ptr, strm: _lib.check(L.pn2_mlp_pack_weights_t(
                                Cn, Kk, lda, W2c.shape[1], W2c.data_ptr(), ptr, strm), "mlp_pack_weights_t"))
                    else:
                        wt_packed[l] = _pack_dgrad(P, lda, W2c)
                if l > 0:
                    hs.append(h)          # None when the fused backward layer will recompute it
                if training:
                    st.a, st.b, st.mean, st.invstd = _bn_finalize(P, st.cout, part, ld, nblk, gamma, beta,
                                                                  rm, rv, eps, momentum, nbt)
                else:
                    st.a, st.b, st.mean, st.invstd = _bn_forward(st.z, gamma, beta, rm, rv, eps, momentum, False)
                st.fold4 = st.a._base if st.a._base is not None else None   # [a | b | mean | invstd], [4, C]
                layers.append(st)
                C = st.cout
                if l < nl - 1:
                    continue
                if l < nl - 1 or ns == 0:
                    h = torch.empty_like(st.z)
                    _lib.check(L.pn2_bn_relu_apply(P, C, st.z.data_ptr(), st.a.data_ptr(), st.b.data_ptr(),
                                                   h.data_ptr(), int(relus[l]), _stream(dev)), "bn_relu_apply")
                    if l < nl - 1:
                        hs.append(h)
                    else:
                        out = h
                elif ext is not None:
                    G = P // ns
                    out = torch.empty(G, C, dtype=torch.float32, device=dev)
                    argmax = torch.empty(G, C, dtype=torch.int32, device=dev)
                    if want_grad and bwd_fused[l] and C % 4 == 0:   # z at the arg-max rows, for the adjoint's statistics
                        zsel = torch.empty(G, C, dtype=torch.float32, device=dev)
                    _lib.check(L.pn2_groupmax_from_extrema(G, C, ext[0].data_ptr(), ext[1].data_ptr(),
                                                           ext[2].data_ptr(), ext[3].data_ptr(), st.a.data_ptr(),
                                                           st.b.data_ptr(), out.data_ptr(), argmax.data_ptr(),
                                                           zsel.data_ptr() if zsel is not None else None,
                                                           _stream(dev)), "groupmax_from_extrema")
                else:
                    G = P // ns
                    out = torch.empty(G, C, dtype=torch.float32, device=dev)
                    argmax = torch.empty(G, C, dtype=torch.int32, device=dev)
                    _lib.check(L.pn2_bn_relu_groupmax(G, ns, C, st.z.data_ptr(), st.a.data_ptr(),
                                                      st.b.data_ptr(), out.data_ptr(), argmax.data_ptr(),
                                                      _stream(dev)), "bn_relu_groupmax")
        ctx.layers, ctx.hs, ctx.argmax, ctx.bwd_fused, ctx.wt_packed = layers, hs, argmax, bwd_fused, wt_packed
        ctx.zsel = zsel
        ctx.lead = lead
        ctx.ns, ctx.training, ctx.nl, ctx.relus, ctx.drops = ns, training, nl, relus, drops
        ctx.weights = [params[3 * l] for l in range(nl)]
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.load()
        layers, hs, nl, ns = ctx.layers, ctx.hs, ctx.nl, ctx.ns
        dev = dout.device
        dout = dout.contiguous()
        grads = [None] * (3 * nl)
        training = 1 if ctx.training else 0
        s = _stream(dev)

        def f32(*shape):
            return torch.empty(*shape, dtype=torch.float32, device=dev)

        with torch.cuda.device(dev):
            dh = dout
            pending = None   # (dgamma, dbeta, c1c2) of layer l, produced by layer l+1's fused kernel
            for l in range(nl - 1, -1, -1):
                st = layers[l]
                P, C = st.z.shape
                W = ctx.weights[l]
                W2 = W.reshape(W.shape[0], -1)
                K = W2.shape[1]
                pooled = l == nl - 1 and ns > 0
                need_dh = l > 0 or ctx.needs_input_grad[0]
                if l == 0 and ctx.lead is not None:
                    # ---- first conv applied before the gather: dZ on the fly, scattered onto the source points ----
                    if pending is None:
                        dgamma, dbeta, c1c2 = f32(C), f32(C), f32(2 * C)
                        _lib.check(L.pn2_relu_bn_bwd_stats(
                            P, C, training, int(ctx.relus[0]), dh.data_ptr(), st.z.data_ptr(), st.a.data_ptr(),
                            st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(), dgamma.data_ptr(),
                            dbeta.data_ptr(), c1c2.data_ptr(), _ws(C, dev).data_ptr(), s), "relu_bn_bwd_stats")
                    else:
                        dgamma, dbeta, c1c2 = pending
                    dW, dh = ctx.lead.backward(dh, st, c1c2, int(ctx.relus[0]), need_dh)
                    grads[0] = dW.reshape(W.shape)
                    grads[1], grads[2] = dgamma, dbeta
                    continue
                if ctx.bwd_fused[l]:
                    # ---- one tensor-core pass: BN adjoint + dH_prev + dW + the next layer's statistics ----
                    if pending is None:
                        dgamma, dbeta, c1c2 = f32(C), f32(C), f32(2 * C)
                        if pooled:
                            _lib.check(L.pn2_pool_bn_bwd_stats(
                                P // ns, ns, C, training, dh.data_ptr(), ctx.argmax.data_ptr(), st.z.data_ptr(),
                                ctx.zsel.data_ptr() if ctx.zsel is not None else None,
                                st.a.data_ptr(), st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(),
                                dgamma.data_ptr(), dbeta.data_ptr(), c1c2.data_ptr(), _ws(C, dev).data_ptr(), s),
                                "pool_bn_bwd_stats")
                        else:
                            _lib.check(L.pn2_relu_bn_bwd_stats(
                                P, C, training, int(ctx.relus[l]), dh.data_ptr(), st.z.data_ptr(), st.a.data_ptr(),
                                st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(), dgamma.data_ptr(),
                                dbeta.data_ptr(), c1c2.data_ptr(), _ws(C, dev).data_ptr(), s), "relu_bn_bwd_stats")
                    else:
                        dgamma, dbeta, c1c2 = pending
                    prev = layers[l - 1] if l > 0 else None
                    a_prev = prev.z if prev is not None else hs[0]
                    lda = a_prev.shape[1]
                    nb, kb, kpad = L.pn2_mlp_bwd_nb(C), L.pn2_mlp_bwd_kb(lda), L.pn2_mlp_bwd_kpad(lda)
                    grid = L.pn2_mlp_bwd_grid(P)
                    wtp = dh_prev = sp = None
                    if need_dh:
                        W2c = W2.contiguous()
                        wtp = ctx.wt_packed[l]
                        if wtp is None:
                            wtp = f32(nb * kpad * 32)
                            _lib.check(L.pn2_mlp_pack_weights_t(C, K, lda, W2c.shape[1], W2c.data_ptr(),
                                                                wtp.data_ptr(), s), "mlp_pack_weights_t")
                        dh_prev = f32(P, lda)
                        if prev is not None:   # BatchNorm below: the kernel leaves S in dwp and Q in sp
                            sp = f32(grid, nb * 32 * kb * 32)
                    dwp = f32(grid, nb * 32 * kb * 32)
                    d_out, d_in = ctx.drops[l], (ctx.drops[l - 1] if l > 0 else None)
                    _lib.check(L.pn2_mlp_bwd_layer(
                        P, C, K, lda, ns if pooled else 0, 1 if pooled else int(ctx.relus[l]),
                        int(ctx.relus[l - 1]) if prev is not None else 0,
                        None if pooled else dh.data_ptr(), dh.data_ptr() if pooled else None,
                        ctx.argmax.data_ptr() if pooled else None, st.z.data_ptr(), st.a.data_ptr(),
                        st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(), c1c2.data_ptr(),
                        W2c.data_ptr() if sp is not None else None, a_prev.data_ptr(), prev.a.data_ptr() if prev is not None else None,
                        prev.b.data_ptr() if prev is not None else None,
                        prev.mean.data_ptr() if prev is not None else None,
                        prev.invstd.data_ptr() if prev is not None else None,
                        wtp.data_ptr() if wtp is not None else None,
                        dh_prev.data_ptr() if dh_prev is not None else None, dwp.data_ptr(),
                        sp.data_ptr() if sp is not None else None, _sched_counters(dev).data_ptr(),
                        d_out[0].data_ptr() if d_out is not None else None, float(d_out[1]) if d_out is not None else 1.0,
                        d_in[0].data_ptr() if d_in is not None else None, float(d_in[1]) if d_in is not None else 1.0, s),
                        "mlp_bwd_layer")
                    pending = None
                    if sp is not None:   # weight-gradient partials + the statistics of the layer below: one launch
                        dW = f32(nb * 32, kb * 32)
                        Kc = prev.cout
                        pg, pb, pc = f32(Kc), f32(Kc), f32(2 * Kc)
                        scratch = f32(L.pn2_mlp_bwd_finalize_scratch_floats(C, lda))
                        _lib.check(L.pn2_mlp_bwd_finalize(grid, C, Kc, lda, dwp.data_ptr(), sp.data_ptr(),
                                                          W2c.data_ptr(), prev.a.data_ptr(), prev.b.data_ptr(),
                                                          prev.mean.data_ptr(), prev.invstd.data_ptr(), P, training,
                                                          dW.data_ptr(), pg.data_ptr(), pb.data_ptr(), pc.data_ptr(),
                                                          scratch.data_ptr(),
                                                          _sched_counters(dev)[8:].data_ptr(), s),
                                   "mlp_bwd_finalize")
                        pending = (pg, pb, pc)
                        grads[3 * l] = dW[:C, :K].reshape(W.shape)
                    else:
                        # dense [C, K] written on the branch: autograd takes it as .grad as it is (a strided
                        # view of a padded image would be COPIED by AccumulateGrad on the main stream, beside
                        # the branch that is still summing it)
                        dWd = f32(C, K)
                        _on_dw_branch(dev, (dwp, dWd), lambda: _lib.check(L.pn2_sum_partials_block(
                            grid, nb * 32, kb * 32, C, K, dwp.data_ptr(), dWd.data_ptr(), _stream(dev)),
                            "sum_partials_block"))
                        grads[3 * l] = dWd.reshape(W.shape)
                    grads[3 * l + 1], grads[3 * l + 2] = dgamma, dbeta
                    dh = dh_prev
                    continue
                # ---- split adjoint (wide layers): dZ by a streaming kernel, then two tensor-core products ----
                dz = torch.empty_like(st.z)
                if pending is not None and not pooled:
                    # this layer's statistics came out of the data-gradient product of the layer above: dZ only
                    dgamma, dbeta, c1c2 = pending
                    _lib.check(L.pn2_relu_bn_bwd_dz(P, C, int(ctx.relus[l]), dh.data_ptr(), st.z.data_ptr(),
                                                    st.a.data_ptr(), st.b.data_ptr(), st.mean.data_ptr(),
                                                    st.invstd.data_ptr(), c1c2.data_ptr(), dz.data_ptr(), s),
                               "relu_bn_bwd_dz")
                else:
                    dgamma, dbeta = f32(C), f32(C)
                    ws = _ws(C, dev)
                    if pooled:
                        _lib.check(L.pn2_pool_bn_bwd(P // ns, ns, C, training, dh.data_ptr(),
                                                     ctx.argmax.data_ptr(), st.z.data_ptr(), st.a.data_ptr(),
                                                     st.b.data_ptr(), st.mean.data_ptr(), st.invstd.data_ptr(),
                                                     dz.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                                     ws.data_ptr(), s), "pool_bn_bwd")
                    else:
                        _lib.check(L.pn2_relu_bn_bwd(P, C, training, int(ctx.relus[l]), dh.data_ptr(), st.z.data_ptr(),
                                                     st.a.data_ptr(), st.b.data_ptr(), st.mean.data_ptr(),
                                                     st.invstd.data_ptr(), dz.data_ptr(), dgamma.data_ptr(),
                                                     dbeta.data_ptr(), ws.data_ptr(), s), "relu_bn_bwd")
                pending = None
                h_prev = hs[l]
                # the data gradient first: the next kernel of the chain waits for it, nobody waits for dW
                # (measured neutral for the step, 3.72 vs 3.74 ms: the two compete for the same L2 bandwidth)
                dh = None
                if need_dh:
                    prev = layers[l - 1] if l > 0 else None
                    if prev is not None and prev.fold4 is not None and DGRAD_STATS:
                        # ... and the BatchNorm-adjoint statistics of the layer below from its epilogue
                        dh, pending = _dgrad_stats(dz, W2.contiguous(), h_prev.shape[1], ctx.wt_packed[l], prev,
                                                   int(ctx.relus[l - 1]), training)
                    else:
                        dh = _dgrad(dz, W2.contiguous(), h_prev.shape[1], ctx.wt_packed[l])
                dW = _on_dw_branch(dev, (dz, h_prev), lambda: _wgrad(dz, h_prev, K))   # dense [C, K]
                grads[3 * l] = dW.reshape(W.shape)
                grads[3 * l + 1], grads[3 * l + 2] = dgamma, dbeta
        return (dh if ctx.needs_input_grad[0] else None, None, None, None, None, None, None, None, None) + tuple(grads)


# Adjoints of the two row gathers through an inverse map (csrc/inverse_map.cu): deterministic, atomic-free --
# and, measured, not faster than the vector-reduction scatters (see that file's header), so off by default.
INVERSE_ADJOINT = os.environ.get("PN2_INVERSE_ADJOINT", "0") == "1"


def _inv_or_build(inv, keys, n_src, needed):
    """(offsets, items) of ``keys``: the one handed in (geometry.build_plan made it ahead of time) or, when a
    gradient will be wanted, built here; None -> the adjoint scatters with atomics."""
    if not INVERSE_ADJOINT:
        return None
    if inv is not None or not needed:
        return inv
    from . import geometry
    return geometry.inverse_map(keys, n_src)


class SAGroupRows(Function):
    """QueryAndGroup in rows layout: table (B, N, ct) [xyz | features | 0 pad], new_xyz (B, np, 3),
    idx (B, np, ns) -> rows (B*np*ns, ct) with the xyz columns centred on their query point.
    inv_off / inv_items: the inverse map of idx (geometry.inverse_map), or None."""

    @staticmethod
    def forward(ctx, table, new_xyz, idx, inv_off=None, inv_items=None):
        B, N, ct = table.shape
        _, npoint, ns = idx.shape
        dev = table.device
        rows = torch.empty(B * npoint * ns, ct, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pn2_sa_group_rows(B, N, npoint, ns, ct, table.data_ptr(),
                                                     new_xyz.data_ptr(), idx.data_ptr(), rows.data_ptr(),
                                                     _stream(dev)), "sa_group_rows")
            ctx.inv = _inv_or_build((inv_off, inv_items) if inv_off is not None else None, idx, N,
                                    ctx.needs_input_grad[0])
        ctx.idx, ctx.shape = idx, (B, N, ct, npoint, ns)
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        B, N, ct, npoint, ns = ctx.shape
        dev = grad_rows.device
        grad_rows = grad_rows.contiguous()
        grad_table = torch.empty(B, N, ct, dtype=torch.float32, device=dev)
        L = _lib.load()
        with torch.cuda.device(dev):
            if ctx.inv is not None:   # segmented gather-sum over the inverse map: no atomics, fixed order
                _lib.check(L.pn2_sa_group_rows_grad_inv(B, N, npoint, ns, ct, grad_rows.data_ptr(),
                                                        ctx.inv[0].data_ptr(), ctx.inv[1].data_ptr(),
                                                        grad_table.data_ptr(), _stream(dev)), "sa_group_rows_grad_inv")
            else:
                _lib.check(L.pn2_sa_group_rows_grad(B, N, npoint, ns, ct, grad_rows.data_ptr(),
                                                    ctx.idx.data_ptr(), grad_table.data_ptr(),
                                                    _stream(dev)), "sa_group_rows_grad")
        # new_xyz is a gather of input coordinates; like the reference's semseg models this path
        # does not propagate a gradient into it
        return grad_table, None, None, None, None


# A trainer can ask to be told when the backward pass leaves the finest level: the event (created with
# external=True, so that it may be recorded inside a captured graph and waited for outside it) is recorded
# right after the adjoint of the FINEST FP block's interpolation -- from there to the SA2 adjoint the step
# is ~60 short, latency-bound kernels of the coarse levels that leave most SMs idle
# (train.PipelinedTrainStep releases the next batch's FPS there).
BWD_MARK_EVENT = None


class FPInterpRows(Function):
    """three_interpolate + concat of PointnetFPModule in rows layout: known (B, m, C2) and skip
    (B, n, C1) point-major, idx / weight (B, n, 3) -> rows (B*n, pad4(C2 + C1)).  ``mark``: record
    BWD_MARK_EVENT after this call's adjoint."""

    @staticmethod
    def forward(ctx, known, skip, idx, weight, mark=False, inv_off=None, inv_items=None):
        ctx.mark = mark
        B, m, c2 = known.shape
        n = idx.shape[1]
        c1 = skip.shape[2] if skip is not None else 0
        dev = known.device
        rows = torch.empty(B * n, pad4(c2 + c1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pn2_fp_interp_rows(
                B, n, m, c2, c1, known.data_ptr(), skip.data_ptr() if skip is not None else None,
                idx.data_ptr(), weight.data_ptr(), rows.data_ptr(), _stream(dev)), "fp_interp_rows")
            ctx.inv = _inv_or_build((inv_off, inv_items) if inv_off is not None else None, idx, m,
                                    ctx.needs_input_grad[0])
        ctx.saved = (idx, weight, B, n, m, c2, c1)
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        idx, weight, B, n, m, c2, c1 = ctx.saved
        dev = grad_rows.device
        grad_rows = grad_rows.contiguous()
        gk = torch.empty(B, m, c2, dtype=torch.float32, device=dev)
        gs = None
        if c1 > 0 and ctx.needs_input_grad[1]:
            gs = torch.empty(B, n, c1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if ctx.inv is not None:   # segmented gather-sum over the inverse map: no atomics, fixed order
                _lib.check(_lib.load().pn2_fp_interp_rows_grad_inv(
                    B, n, m, c2, c1, grad_rows.data_ptr(), ctx.inv[0].data_ptr(), ctx.inv[1].data_ptr(),
                    weight.data_ptr(), gk.data_ptr(), gs.data_ptr() if gs is not None else None, _stream(dev)),
                    "fp_interp_rows_grad_inv")
            else:
                _lib.check(_lib.load().pn2_fp_interp_rows_grad(
                    B, n, m, c2, c1, grad_rows.data_ptr(), idx.data_ptr(), weight.data_ptr(), gk.data_ptr(),
                    gs.data_ptr() if gs is not None else None, _stream(dev)), "fp_interp_rows_grad")
        if ctx.mark and BWD_MARK_EVENT is not None:
            BWD_MARK_EVENT.record(torch.cuda.current_stream(dev))
        return gk, gs, None, None, None, None, None


CHAIN_EVAL = True   # inference: a whole SA block (gather + 3 layers + max) in one kernel where it fits
_chain_cache = {}


def sa_chain_eval(mlp, table, new_xyz, idx):
    """QueryAndGroup + SharedMLP + max-pool of one SA scale in ONE kernel (csrc/sa_chain.cu), inference only:
    table (B, N, ct), new_xyz (B, np, 3), idx (B, np, ns) -> (B * np, C3), or None when the block does not fit
    the kernel (then the caller runs it layer by layer).  BatchNorm is folded from the running statistics."""
    layers = list(mlp)
    if not CHAIN_EVAL or len(layers) != 3 or torch.is_grad_enabled() or any(l.training for l in layers):
        return None
    if not all(getattr(l, "activation", None) is not None for l in layers):
        return None
    L = _lib.load()
    B, N, ct = table.shape
    _, npoint, ns = idx.shape
    Ws = [l.conv.weight.reshape(l.conv.weight.shape[0], -1) for l in layers]
    cs = [w.shape[0] for w in Ws]
    if Ws[0].shape[1] > ct or not L.pn2_sa_chain_supported(ns, ct, cs[0], cs[1], cs[2]):
        return None
    dev = table.device
    s = _stream(dev)
    # weight tiles and folds depend on the parameters / running statistics only: kept per SharedMLP until any
    # of them changes (inference evaluates many batches with the same weights)
    state = [t for layer in layers for t in (layer.conv.weight, layer.bn.bn.weight, layer.bn.bn.bias,
                                             layer.bn.bn.running_mean, layer.bn.bn.running_var)]
    key = (ct, dev) + tuple((t.data_ptr(), t._version) for t in state)
    hit = _chain_cache.get(id(mlp))
    if hit is not None and hit[0] == key and not torch.cuda.is_current_stream_capturing():
        packed, fold = hit[1], hit[2]
        out = torch.empty(B * npoint, cs[2], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.pn2_sa_chain_eval(B, N, npoint, ns, ct, table.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(),
                                           cs[0], cs[1], cs[2], packed[0].data_ptr(), packed[1].data_ptr(),
                                           packed[2].data_ptr(), fold.data_ptr(), out.data_ptr(), s), "sa_chain_eval")
        return out
    with torch.cuda.device(dev):
        packed, folds = [], []
        for l, (layer, W2) in enumerate(zip(layers, Ws)):
            W2 = W2.contiguous()
            k = ct if l == 0 else cs[l - 1]
            wp = torch.empty(L.pn2_sa_chain_weight_floats(cs[l], k), dtype=torch.float32, device=dev)
            _lib.check(L.pn2_mlp_pack_weights_bn(cs[l], W2.shape[1], W2.shape[1], W2.data_ptr(), wp.data_ptr(), s),
                       "mlp_pack_weights_bn")
            packed.append(wp)
            norm = layer.bn.bn
            ab = torch.empty(4, cs[l], dtype=torch.float32, device=dev)   # a | b | mean | invstd
            _lib.check(L.pn2_bn_fold_eval(cs[l], norm.weight.data_ptr(), norm.bias.data_ptr(),
                                          norm.running_mean.data_ptr(), norm.running_var.data_ptr(), float(norm.eps),
                                          ab[0].data_ptr(), ab[1].data_ptr(), ab[2].data_ptr(), ab[3].data_ptr(), s),
                       "bn_fold_eval")
            folds.append(ab[:2].reshape(-1))
        fold = torch.cat(folds)
        if not torch.cuda.is_current_stream_capturing():
            _chain_cache[id(mlp)] = (key, packed, fold)
        out = torch.empty(B * npoint, cs[2], dtype=torch.float32, device=dev)
        _lib.check(L.pn2_sa_chain_eval(B, N, npoint, ns, ct, table.data_ptr(), new_xyz.data_ptr(), idx.data_ptr(),
                                       cs[0], cs[1], cs[2], packed[0].data_ptr(), packed[1].data_ptr(),
                                       packed[2].data_ptr(), fold.data_ptr(), out.data_ptr(), s), "sa_chain_eval")
    return out


def pad_rows(x, width):
    """(B, N, c) contiguous rows -> (B, N, width) with zero columns appended (width % 4 == 0), one launch; ``x``
    itself when it already has that width.  No gradient (the network's input)."""
    B, N, c = x.shape
    if c == width:
        return x
    out = torch.empty(B, N, width, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().pn2_pad_rows(B * N, c, width, x.data_ptr(), out.data_ptr(), _stream(x.device)), "pad_rows")
    return out


def rows_of(t):
    """Point-major (B, N, C) contiguous alias of a channel-major (B, C, N) feature tensor: the
    rows tensor a fused module produced it from when there is one, else a transposed copy."""
    r = getattr(t, "_pn2_rows", None)
    if r is not None:
        return r
    return t.transpose(1, 2).contiguous()


LAZY_CHANNEL_MAJOR = False  # set by the semseg models around their forward (see channel_major)


def channel_major(rows3):
    """(B, N, C) rows -> the (B, C, N) contiguous tensor the op surface promises, remembering the
    rows it came from so that the next fused block can skip the transpose back.

    Inside the semseg models every consumer of a module output is another fused block, which takes
    the rows alias: there (``LAZY_CHANNEL_MAJOR``) the channel-major tensor is returned as a
    transposed VIEW of the rows -- same shape and values, no copy (and no copy of its gradient on
    the way back); ``.contiguous()`` materialises it if some other consumer needs that."""
    out = rows3.transpose(1, 2)
    if not LAZY_CHANNEL_MAJOR:
        out = out.contiguous()
    out._pn2_rows = rows3
    return out


# ------------------------------------------------------------------------ module glue
def mlp_eligible(mlp, x_like):
    """True when ``mlp`` (a pytorch_utils.SharedMLP) is conv1x1(no bias) + BatchNorm2d + ReLU
    throughout and the data is fp32 on CUDA -- the configuration of every semseg block."""
    if not ENABLED or not x_like.is_cuda or x_like.dtype != torch.float32:
        return False
    for layer in mlp:
        conv = getattr(layer, "conv", None)
        bn = getattr(layer, "bn", None)
        act = getattr(layer, "activation", None)
        if conv is None or bn is None or not (act is None or isinstance(act, nn.ReLU)):
            return False
        if conv.bias is not None or conv.weight.shape[0] % 4 != 0 or conv.weight.shape[0] > 1024:
            return False
        if list(layer._modules.keys())[0] != "conv":  # preact ordering is not handled here
            return False
        norm = bn.bn
        if norm.momentum is None or not norm.affine or not norm.track_running_stats:
            return False
    return True


def lead_ok(units, min_width=16):
    """May the first unit of ``units`` be applied before the gather?  Training mode (batch statistics), at least one
    more unit behind it (the first is never the pooled one), and a width the kernels tile exactly: a power of
    two in [16, 1024] (SA) / [64, 1024] (FP) -- every first layer of the reference's semseg models."""
    units = list(units)
    if not (LEAD and ENABLED and len(units) >= 2 and units[0].training):
        return False
    c = units[0].conv.weight.shape[0]
    return (c & (c - 1)) == 0 and min_width <= c <= 1024


def fp_lead_max_skip():
    return int(_lib.load().pn2_fp_lead_max_skip())


def run_mlp_rows(mlp, x0, ns, dropouts=None, lead=None):
    """Apply a SharedMLP (or any sequence of conv+BN[+ReLU] units) to rows ``x0`` [P, Cin(_pad)];
    ns > 0 pools every run of ns rows after the last layer.  ``lead`` (SALead / FPLead): ``x0`` is the SOURCE
    tensor (point table (B, N, ct) / known rows (B, m, C2)) and the first unit is applied before the gather.  ``dropouts``: {unit index: p} (or
    {unit index: (p, keep mask)}) -- nn.Dropout(p) on that unit's output, applied inside the next
    unit's kernels (see ``chain_supports_dropout``); ignored in eval mode."""
    params, buffers, relus = [], [], []
    layers = list(mlp)
    training = layers[0].training
    norms = [layer.bn.bn for layer in layers]
    for layer, norm in zip(layers, norms):
        params += [layer.conv.weight, norm.weight, norm.bias]
        buffers.append((norm.running_mean, norm.running_var, norm.num_batches_tracked))
        relus.append(getattr(layer, "activation", None) is not None)
    drops = [None] * len(layers)
    if dropouts and training:
        for i, spec in dropouts.items():
            p, mask = spec if isinstance(spec, tuple) else (spec, None)
            if p <= 0.0:
                continue
            if mask is None:  # keep mask, one byte per element
                mask = _dropout_mask(x0.shape[0] if lead is None else lead.P, layers[i].conv.weight.shape[0],
                                     1.0 - p, x0.device)
            drops[i] = (mask, 1.0 / (1.0 - p))
    # running statistics AND num_batches_tracked are updated inside the finalise kernels
    return SharedMLPRows.apply(x0, ns, training, float(norms[0].eps), float(norms[0].momentum), buffers,
                               tuple(relus), tuple(drops), lead, *params)


_drop_state = {}


def _dropout_mask(rows, cols, keep, dev):
    """uint8 keep mask [rows, cols] from the library's counter-based generator (pn2_dropout_mask): its
    state (seed drawn once from torch's CPU generator, launch counter) lives on the device, so the launch
    is CUDA-graph replayable and every replay draws a fresh mask."""
    st = _drop_state.get(dev)
    if st is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        # data-parallel ranks seed torch identically (same initial weights): without the rank every replica
        # would draw the same keep masks
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            seed = (seed ^ (0x9E3779B97F4A7C15 * (dist.get_rank() + 1))) & (2 ** 62 - 1)
        st = _drop_state[dev] = torch.tensor([seed, 0], dtype=torch.int64, device=dev)
    def launch(ptr, strm):
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pn2_dropout_mask(rows * cols, float(keep), st.data_ptr(), ptr, strm), "dropout_mask")

    if _pack_state["active"]:
        # the mask depends on nothing: drawn on the packing branch, into a persistent buffer (kept for the
        # backward pass; the next step's branch forks after it).  One buffer per dropout site of the step.
        n = _pack_state["n_drop"] = _pack_state.get("n_drop", 0) + 1
        key = ("drop", n, rows, cols, dev)
        buf = _pack_state["bufs"].get(key)
        if buf is None and not torch.cuda.is_current_stream_capturing():
            buf = _pack_state["bufs"][key] = torch.empty(rows, cols, dtype=torch.uint8, device=dev)
        if buf is not None:
            side = _pack_state["stream"]
            launch(buf.data_ptr(), side.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(side)
            torch.cuda.current_stream(dev).wait_event(ev)
            return buf
    mask = torch.empty(rows, cols, dtype=torch.uint8, device=dev)
    launch(mask.data_ptr(), _stream(dev))
    return mask


def chain_supports_dropout(units, x0, i):
    """Can nn.Dropout after unit ``i`` of the chain be applied inside the kernels of unit i + 1?  Needs the
    tensor-core forward layer and the fused backward layers for units i and i + 1."""
    if not (TC_BWD and ENABLED) or i + 1 >= len(units):
        return False
    L = _lib.load()
    shape = tuple(x0.shape) if hasattr(x0, "shape") else tuple(x0)   # rows [P, width] or just their shape
    P = shape[0]
    for l in (i, i + 1):
        W = units[l].conv.weight
        cout, cin = W.shape[0], W.reshape(W.shape[0], -1).shape[1]
        lda = shape[1] if l == 0 else units[l - 1].conv.weight.shape[0]
        if not L.pn2_mlp_bwd_supported(P, cout, cin, lda, int(l > 0)):
            return False
    return units[i].conv.weight.shape[0] % 4 == 0 and any(p.requires_grad for p in units[0].parameters())


def pad4(c):
    return (c + 3) // 4 * 4
