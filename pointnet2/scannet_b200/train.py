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

    def __init__(self, model, lr=1e-3, weight_decay=0.0, group=None, fused_adam=None):
        self.model = model
        self.group = group
        self.bucket = GradBucket(model.parameters())
        if fused_adam is None:
            fused_adam = self.bucket.flat.is_cuda
        self.optimizer = torch.optim.Adam(self.bucket.params, lr=lr, weight_decay=weight_decay,
                                          fused=fused_adam)

    def __call__(self, pointcloud, targets, weights):
        self.bucket.zero()
        logits = self.model(pointcloud)                                   # (B, N, classes)
        loss = weighted_cross_entropy(logits.reshape(-1, logits.size(-1)),
                                      targets.reshape(-1), weights.reshape(-1))
        loss.backward()
        self.bucket.all_reduce_mean(self.group)
        self.optimizer.step()
        return loss
