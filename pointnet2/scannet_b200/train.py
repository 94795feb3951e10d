"""Training step for the semseg models: forward + weighted cross-entropy + backward +
Adam, data-parallel over the GPUs of one box.

Mirrors what the reference's solver does per iteration
(/root/reference/lib/solver.py:161-176: ``model(cat([coords, feats], 2))`` ->
``WeightedCrossEntropyLoss`` (/root/reference/lib/loss.py:10-18) -> ``zero_grad`` /
``backward`` / ``optimizer.step``; Adam lr 1e-3, scripts/train.py:44) minus its
per-iteration host round trips (``.cpu().numpy()`` metrics, CUDA_LAUNCH_BLOCKING=1).

The reference is single-GPU.  Data parallelism here is the only multi-GPU structure the
path admits: every kernel is per-cloud, so each rank runs a full replica
on its own chunks and the ONLY exchange is one all-reduce of the gradients -- all of them
live in a single flat fp32 bucket (3.8 MB SSG / 12.1 MB MSG) so it is one NCCL call per
step.  BatchNorm statistics stay per replica, as in the reference (no SyncBN).
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

# graph capture runs backward on the capture stream while the AccumulateGrad nodes were created
# on the warm-up stream: intended here, so silence the per-step warning
if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)


def weighted_cross_entropy(logits, targets, weights=None):
    """logits (P, classes), targets (P,), weights (P,) -> scalar  (lib/loss.py:10-18; loss.py)."""
    from . import loss as _loss
    return _loss.weighted_cross_entropy(logits, targets, weights)


# timing experiment only (results invalid): skip the geometry graph, i.e. train every batch on the first
# batch's sampling / grouping indices -- shows how much of the overlapped geometry work is NOT hidden
import os as _os
_STALE_GEOMETRY = _os.environ.get("PN2_STALE_GEOMETRY") == "1"


class GradBucket(object):
    """All gradients of a module as views into one contiguous fp32 buffer."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def zero(self):
        self.flat.zero_()

    def attached(self):
        """True while every .grad still aliases the bucket (zero_grad(set_to_none=True) breaks it)."""
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base
                   for p in self.params)

    def all_reduce_mean(self, group=None):
        """Sum over ranks, then 1/world: one collective for the whole model."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))


class FlatAdam(object):
    """torch.optim.Adam's update over ONE flat buffer, CUDA only (libpn2_b200.so: pn2_adam_flat).

    All parameters are re-pointed at views of a single fp32 buffer (state dicts are unaffected);
    gradients are handed over by autograd as it produces them (``.grad = None`` before backward,
    so there is no accumulate kernel per tensor), packed into one flat gradient with a single
    concatenation, all-reduced once when data-parallel, and applied by one kernel whose step
    counter lives on the device (replayable from a CUDA graph).  This replaces ~75 accumulate
    launches, the per-tensor-list fused-Adam launches and the gradient bucket zero-fill."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        dev = self.params[0].device
        with torch.no_grad():
            self.flat = torch.cat([p.detach().reshape(-1) for p in self.params])
            off = 0
            for p in self.params:
                p.data = self.flat[off:off + p.numel()].view_as(p)
                off += p.numel()
        self.grad = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=dev)

    @property
    def nbytes(self):
        return self.flat.numel() * 4

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def pack_grads(self):
        torch.cat([p.grad.reshape(-1) for p in self.params], out=self.grad)

    def step(self, grad_scale=1.0):
        from . import lib
        dev = self.flat.device
        with torch.cuda.device(dev):
            lib.check(lib.load().pn2_adam_flat(
                self.flat.numel(), self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.lr, self.betas[0],
                self.betas[1], self.eps, self.weight_decay, grad_scale,
                torch.cuda.current_stream(dev).cuda_stream), "adam_flat")


class TrainStep(object):
    """One optimisation step on a (B, N, 3+C) batch; call it with device tensors.

    On CUDA the update runs through ``FlatAdam``; on the CPU (the benchmark's cpu_baseline leg and
    the gloo tests) through ``GradBucket`` + ``torch.optim.Adam``."""

    def __init__(self, model, lr=1e-3, weight_decay=0.0, group=None, fused_adam=None, capturable=False):
        self.model = model
        self.group = group
        params = [p for p in model.parameters() if p.requires_grad]
        self.flat_mode = params[0].is_cuda and fused_adam is not False
        if self.flat_mode:
            self.opt = FlatAdam(params, lr=lr, weight_decay=weight_decay)
            self.bucket = None
        else:
            self.bucket = GradBucket(params)
            self.optimizer = torch.optim.Adam(self.bucket.params, lr=lr, weight_decay=weight_decay)

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def forward_loss(self, pointcloud, targets, weights, plan=None):
        if self.flat_mode:
            self.opt.zero_grad()
        else:
            self.bucket.zero()
        from . import fused
        with fused.pack_branch(pointcloud.device):   # weight-tile packing on a forked branch (CUDA only)
            logits = self.model(pointcloud, plan) if plan is not None else self.model(pointcloud)
        return weighted_cross_entropy(logits.reshape(-1, logits.size(-1)),
                                      targets.reshape(-1), weights.reshape(-1))

    def backward_pack(self, loss):
        from . import fused
        with fused.dw_branch(loss.device):   # weight-gradient GEMMs of the wide layers on a forked branch
            loss.backward()
        if self.flat_mode:
            self.opt.pack_grads()

    def forward_backward(self, pointcloud, targets, weights, plan=None):
        loss = self.forward_loss(pointcloud, targets, weights, plan)
        self.backward_pack(loss)
        return loss

    def reduce_and_update(self):
        if not self.flat_mode:
            self.bucket.all_reduce_mean(self.group)
            self.optimizer.step()
            return
        world = self._world()
        if world > 1:  # the one collective of the step: the whole gradient, summed
            dist.all_reduce(self.opt.grad, op=dist.ReduceOp.SUM, group=self.group)
        self.opt.step(grad_scale=1.0 / world)

    def __call__(self, pointcloud, targets, weights):
        loss = self.forward_backward(pointcloud, targets, weights)
        self.reduce_and_update()
        return loss


class _Snapshot(object):
    """Parameters, optimiser state and BatchNorm buffers of a TrainStep, to be put back after the warm-up
    passes a graph capture needs (allocator, lazy initialisations): building a graphed step must not train
    the model, move its running statistics or advance Adam's step count."""

    def __init__(self, step):
        self.step = step
        self.buffers = [b.detach().clone() for b in step.model.buffers()]
        if step.flat_mode:
            o = step.opt
            self.opt = [t.detach().clone() for t in (o.flat, o.exp_avg, o.exp_avg_sq, o.step_count)]
        else:
            import copy
            self.params = [p.detach().clone() for p in step.bucket.params]
            self.opt = copy.deepcopy(step.optimizer.state_dict())

    def restore(self):
        step = self.step
        with torch.no_grad():
            for b, saved in zip(step.model.buffers(), self.buffers):
                b.copy_(saved)
            if step.flat_mode:
                o = step.opt
                for t, saved in zip((o.flat, o.exp_avg, o.exp_avg_sq, o.step_count), self.opt):
                    t.copy_(saved)
            else:
                for p, saved in zip(step.bucket.params, self.params):
                    p.copy_(saved)
                step.optimizer.load_state_dict(self.opt)


class GraphedTrainStep(object):
    """The same step replayed from a CUDA graph: the ~470 kernel launches of one SSG step
    (libpn2_b200.so kernels, cuBLAS GEMMs, the fused Adam) become one graph launch, which removes
    the host launch gaps between the many short kernels of the coarse levels.

    ``scope="step"`` captures forward + backward + all-reduce + Adam (single GPU, or NCCL in the
    graph); ``scope="fwd_bwd"`` captures forward + backward only and runs the gradient all-reduce
    and Adam eagerly after the replay.  Inputs are copied into static device buffers."""

    def __init__(self, step, example, scope="step", warmup=3):
        assert scope in ("step", "fwd_bwd")
        self.step, self.scope = step, scope
        self.static = tuple(torch.empty_like(t) for t in example)
        for s, t in zip(self.static, example):
            s.copy_(t)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        snap = _Snapshot(step)
        with torch.cuda.stream(side):  # allocator / workspaces / lazy inits outside capture
            for _ in range(warmup):
                self._body()
            snap.restore()             # the warm-up passes leave no trace in the model or the optimiser
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()

    def _body(self):
        loss = self.step.forward_backward(*self.static)
        if self.scope == "step":
            self.step.reduce_and_update()
        return loss

    def __call__(self, pointcloud, targets, weights):
        for s, t in zip(self.static, (pointcloud, targets, weights)):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        if self.scope == "fwd_bwd":
            self.step.reduce_and_update()
        return self.loss


class PipelinedTrainStep(object):
    """GraphedTrainStep + geometry prefetch: while the graph of batch i runs on the main stream,
    the coordinates-only part of batch i+1 (geometry.build_plan: FPS chain, ball queries, 3-NN) is
    computed by a second graph on a side stream.  Every batch still gets its own geometry,
    computed by the same kernels; only WHEN it is computed changes.  Call ``prefetch(batch)`` with
    the next batch before (or right after) calling the step on the current one."""

    def __init__(self, step, example, scope="step", warmup=3, split_fwd_bwd=True):
        from . import geometry
        assert scope in ("step", "fwd_bwd")
        self.step, self.scope = step, scope
        self.static = tuple(torch.empty_like(t) for t in example)        # batch being trained on
        self.next = tuple(torch.empty_like(t) for t in example)          # batch being prefetched
        self.side = torch.cuda.Stream()
        self.copied = torch.cuda.Event()
        self.ready = torch.cuda.Event()
        for s, t in zip(self.next, example):
            s.copy_(t)
        for s, t in zip(self.static, example):
            s.copy_(t)
        model = step.model
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            # both plans live in flat buffers: handing the prefetched plan over is one device copy
            self.plan = geometry.build_plan(model, self.static[0][..., 0:3].contiguous()).pack()
            self.plan_next = geometry.build_plan(model, self.next[0][..., 0:3].contiguous()).pack()
            snap = _Snapshot(step)
            for _ in range(warmup):
                self._body()
            snap.restore()             # the warm-up passes leave no trace in the model or the optimiser
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        # graph 1 (side stream): next batch's coordinates -> plan_next
        self.geo_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.geo_graph, stream=self.side):
            fresh = geometry.build_plan(model, self.next[0][..., 0:3].contiguous())
            self.plan_next.copy_from(fresh)
        # graphs 2a / 2b (main stream): the training step on (static, plan), split after the forward pass.
        # The geometry graph of the NEXT batch is released by an event recorded between the two, so that it
        # overlaps the backward pass -- whose tensor-core layers hand out their row tiles dynamically and
        # keep their speed when FPS holds 32 SMs -- instead of the forward pass, whose statically tiled
        # layers wait for the CTAs that cannot become resident (measured: 0.48 ms of the 1.0 ms of geometry
        # work stayed exposed when it started together with the step).
        # Where in the step the geometry graph starts (PN2_GEO_RELEASE): "fwd" (default) = after the forward
        # pass, beside the finest level's backward layers; "mid" = after the finest FP block's interpolation
        # adjoint, beside the ~60 short kernels of the coarse levels.  Measured in one run on one box (round 2):
        # "mid" gives the FP0 / head backward layers all 148 SMs (-66 us) but moves 3-NN / ball query beside
        # the SA2 / SA1 backward layers: 3.77-3.80 vs 3.72-3.74 ms per step, so "fwd" stays.
        # "fwd1": the same release point as "fwd", but forward and backward stay ONE graph and the release is
        # an external event-record node after the forward pass (no second graph launch, whose ~30 us gap sat
        # on the critical path between the loss and the first backward kernel).
        self.release = _os.environ.get("PN2_GEO_RELEASE", "fwd1")
        self.split = split_fwd_bwd
        self.fwd_done = torch.cuda.Event()
        self.mid_bwd = torch.cuda.Event(external=True) if (self.split and self.release in ("mid", "fwd1")) else None
        self.graph = torch.cuda.CUDAGraph()
        self.one_graph = self.split and self.release == "fwd1"
        if self.one_graph:
            with torch.cuda.graph(self.graph):
                self.loss = self.step.forward_loss(*self.static, plan=self.plan)
                self.mid_bwd.record(torch.cuda.current_stream())
                self.step.backward_pack(self.loss)
                if self.scope == "step":
                    self.step.reduce_and_update()
        elif self.split:
            with torch.cuda.graph(self.graph):
                self.loss = self.step.forward_loss(*self.static, plan=self.plan)
            self.graph_bwd = torch.cuda.CUDAGraph()
            from . import fused
            fused.BWD_MARK_EVENT = self.mid_bwd
            try:
                with torch.cuda.graph(self.graph_bwd, pool=self.graph.pool()):
                    self.step.backward_pack(self.loss)
                    if self.scope == "step":
                        self.step.reduce_and_update()
            finally:
                fused.BWD_MARK_EVENT = None
        else:
            with torch.cuda.graph(self.graph):
                self.loss = self._body()
        self._have_next = False
        self._fwd_recorded = False
        torch.cuda.synchronize()

    def _body(self):
        loss = self.step.forward_backward(*self.static, plan=self.plan)
        if self.scope == "step":
            self.step.reduce_and_update()
        return loss

    def prefetch(self, pointcloud, targets, weights):
        """Stage the NEXT batch (host or device tensors) and compute its geometry on the side stream."""
        main = torch.cuda.current_stream()
        self.side.wait_event(self.copied)        # the previous contents of `next` were consumed
        with torch.cuda.stream(self.side):
            for s, t in zip(self.next, (pointcloud, targets, weights)):
                s.copy_(t, non_blocking=True)
            if self.split and self._fwd_recorded:
                # start beside the backward pass of the running step (the event node of the replayed graph)
                self.side.wait_event(self.mid_bwd if self.mid_bwd is not None else self.fwd_done)
            if not _STALE_GEOMETRY:
                self.geo_graph.replay()
            self.ready.record(self.side)
        self._have_next = True
        del main

    def __call__(self, pointcloud=None, targets=None, weights=None):
        """Train on the prefetched batch (arguments are only used when nothing was prefetched)."""
        main = torch.cuda.current_stream()
        if not self._have_next:
            self.prefetch(pointcloud, targets, weights)
        main.wait_event(self.ready)
        for d, s in zip(self.static, self.next):   # next -> current (device copies, ~12 MB)
            d.copy_(s, non_blocking=True)
        self.plan.copy_from(self.plan_next)
        self.copied.record(main)
        self._have_next = False
        self.graph.replay()
        if self.one_graph:
            self._fwd_recorded = True
        elif self.split:
            self.fwd_done.record(main)
            self._fwd_recorded = True
            self.graph_bwd.replay()
        if self.scope == "fwd_bwd":
            self.step.reduce_and_update()
        return self.loss
