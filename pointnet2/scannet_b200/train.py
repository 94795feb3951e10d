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


def weighted_cross_entropy(logits, targets, weights=None):
    """logits (P, classes), targets (P,), weights (P,) -> scalar  (lib/loss.py:10-18)."""
    loss = F.cross_entropy(logits, targets, reduction="none")
    return torch.mean(loss * weights) if weights is not None else torch.mean(loss)


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


class TrainStep(object):
    """One optimisation step on a (B, N, 3+C) batch; call it with device tensors."""

    def __init__(self, model, lr=1e-3, weight_decay=0.0, group=None, fused_adam=None, capturable=False):
        self.model = model
        self.group = group
        self.bucket = GradBucket(model.parameters())
        if fused_adam is None:
            fused_adam = self.bucket.flat.is_cuda
        self.optimizer = torch.optim.Adam(self.bucket.params, lr=lr, weight_decay=weight_decay,
                                          fused=fused_adam, capturable=capturable)

    def forward_backward(self, pointcloud, targets, weights):
        self.bucket.zero()
        logits = self.model(pointcloud)
        loss = weighted_cross_entropy(logits.reshape(-1, logits.size(-1)),
                                      targets.reshape(-1), weights.reshape(-1))
        loss.backward()
        return loss

    def reduce_and_update(self):
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()

    def __call__(self, pointcloud, targets, weights):
        self.bucket.zero()
        logits = self.model(pointcloud)                                   # (B, N, classes)
        loss = weighted_cross_entropy(logits.reshape(-1, logits.size(-1)),
                                      targets.reshape(-1), weights.reshape(-1))
        loss.backward()
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()
        return loss


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
        with torch.cuda.stream(side):  # allocator / cuBLAS workspaces / lazy inits outside capture
            for _ in range(warmup):
                self._body()
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
