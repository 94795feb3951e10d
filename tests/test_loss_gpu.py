"""GPU parity of the weighted cross-entropy kernels (csrc/loss.cu, through the C ABI) against the
reference's formulation (/root/reference/lib/loss.py:10-18: F.cross_entropy(reduction="none",
ignore_index) -> mean(loss * weights)) evaluated by torch in fp64 on the same inputs.
Tolerance: loss 1e-6 relative, d logits 1e-5 of the gradient's max magnitude (fp32 exp/log)."""
import pytest
import torch
import torch.nn.functional as F

from pointnet2.scannet_b200 import loss as L

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(logits, targets, weights, ignore_index, upstream):
    x = logits.detach().double().requires_grad_(True)
    per = F.cross_entropy(x, targets, reduction="none", ignore_index=ignore_index)
    loss = torch.mean(per * weights.double()) if weights is not None else torch.mean(per)
    (loss * upstream).backward()
    return loss.detach(), x.grad


@pytest.mark.parametrize("P,C", [(1, 4), (255, 20), (257, 20), (8192 * 4, 20), (1000, 64), (262144, 20)])
@pytest.mark.parametrize("weighted", [True, False])
def test_weighted_ce_matches_torch(P, C, weighted):
    g = torch.Generator(device="cpu").manual_seed(P * 31 + C)
    logits = (torch.randn(P, C, generator=g) * 4.0).to(DEV).requires_grad_(True)
    targets = torch.randint(0, C, (P,), generator=g).to(DEV)
    if P > 8:
        targets[::7] = -100                                  # ignored rows (reference default ignore_index)
    weights = (torch.rand(P, generator=g) * 3.0).to(DEV) if weighted else None
    assert L.kernel_eligible(logits, targets, weights)
    crit = L.WeightedCrossEntropyLoss()
    loss = crit(logits, targets, weights)
    (loss * 2.5).backward()                                  # upstream gradient != 1
    ref_loss, ref_grad = _reference(logits, targets, weights, -100, 2.5)
    assert abs(float(loss) - float(ref_loss)) <= 1e-6 * max(1.0, abs(float(ref_loss)))
    err = float((logits.grad.double() - ref_grad).abs().max())
    assert err <= 1e-5 * float(ref_grad.abs().max().clamp_min(1e-30)), err
    if P > 8:
        assert float(logits.grad[::7].abs().max()) == 0.0    # ignored rows get exactly zero


def test_weighted_ce_is_deterministic_and_rearms_its_counter():
    g = torch.Generator(device="cpu").manual_seed(5)
    logits = torch.randn(100000, 20, generator=g).to(DEV)
    targets = torch.randint(0, 20, (100000,), generator=g).to(DEV)
    weights = torch.rand(100000, generator=g).to(DEV)
    vals = [float(L.weighted_cross_entropy(logits, targets, weights)) for _ in range(5)]
    assert len(set(vals)) == 1, vals


def test_weighted_ce_ineligible_inputs_take_the_torch_formulation():
    logits = torch.randn(64, 21, device=DEV)                 # 21 classes: not a multiple of 4
    targets = torch.randint(0, 21, (64,), device=DEV)
    weights = torch.rand(64, device=DEV)
    assert not L.kernel_eligible(logits, targets, weights)
    ref = torch.mean(F.cross_entropy(logits, targets, reduction="none") * weights)
    assert torch.allclose(L.weighted_cross_entropy(logits, targets, weights), ref)
