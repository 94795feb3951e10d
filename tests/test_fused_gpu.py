"""GPU parity of the fused shared-MLP blocks (rows layout, tcgen05 layers) against the SAME module run
through the reference's formulation (torch.nn Conv2d + BatchNorm2d + ReLU + max_pool2d on
(B, C, npoint, nsample)).

Every 1x1 convolution of the fused path and both of its adjoint products run on the hand-written
tensor-core kernels with TF32 operands and fp32 accumulation -- there is no fp32 or library GEMM mode.
The stated tolerance: the fused path's deviation from the fp32 result of the
reference formulation (``allow_tf32 = False``: the truth) is at most 4x the deviation the reference
formulation ITSELF shows under its default arithmetic (``cudnn.allow_tf32 = True``), with a floor of
5e-3 of the tensor's max magnitude for outputs (entry by entry) and 1e-2 in the L2 norm for gradients.
Gradients are compared in the L2 norm because under TF32 a pooled value can change its arg-max row or
cross the ReLU threshold, which re-routes ONE gradient entry completely (the reference formulation shows
4-40 % max-entry deviations on these inputs for the same reason).  (4x: the tensor core truncates the
activation operands to TF32 where cuDNN's kernels may round them; truncation doubles the operand error.)
The kernels' arithmetic itself is pinned to 2e-5 .. 4e-4 against fp64 products of the TF32-reduced
operands in test_gemm_gpu.py / test_bwd_gpu.py / test_wide_gpu.py.
"""
import copy

import numpy as np
import pytest
import torch

import cases
from pointnet2.scannet_b200 import fused, pointnet2_modules as M, pointnet2_semseg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def l2_err(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


class torch_tf32(object):
    """arithmetic of the torch.nn formulation: TF32 convs / matmuls on (the reference's default) or off"""

    def __init__(self, tf32):
        self.tf32 = tf32

    def __enter__(self):
        self.saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = self.tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.saved


def run(module, enabled, *inputs, train=True):
    module = copy.deepcopy(module)
    module.train(train)
    ins = [i.clone().requires_grad_(r) for i, r in inputs]
    fused.ENABLED = enabled
    try:
        out = module(*ins)
        out = out[1] if isinstance(out, tuple) else out
        torch.manual_seed(1)
        w = torch.randn_like(out)
        (out * w).sum().backward()
    finally:
        fused.ENABLED = True
    grads = {n: p.grad for n, p in module.named_parameters()}
    bufs = {n: b.clone() for n, b in module.named_buffers()}
    in_grads = [i.grad if i.requires_grad else None for i in ins]
    return out.detach(), grads, bufs, in_grads


def class_bound(module, inputs, factor=4.0, out_floor=5e-3, grad_floor=1e-2, stat_tol=2e-3):
    """fused (tcgen05, TF32) vs the torch.nn formulation in fp32, bounded by what the torch.nn
    formulation itself loses under TF32.  Returns the measured deviations (printed on failure)."""
    with torch_tf32(False):
        out_t, g_t, b_t, ig_t = run(module, False, *inputs)
    with torch_tf32(True):
        out_r, g_r, _, ig_r = run(module, False, *inputs)
    out_f, g_f, b_f, ig_f = run(module, True, *inputs)
    assert out_f.shape == out_t.shape
    report = {"out": (rel_err(out_f, out_t), rel_err(out_r, out_t))}
    assert report["out"][0] <= max(factor * report["out"][1], out_floor), report
    pairs = [(n, g_f[n], g_r[n], g_t[n]) for n in g_t]
    pairs += [("input%d" % i, f, r, t) for i, (f, r, t) in enumerate(zip(ig_f, ig_r, ig_t)) if t is not None]
    for n, f, r, t in pairs:
        assert f is not None, n
        report[n] = (l2_err(f, t), l2_err(r, t))
        assert report[n][0] <= max(factor * report[n][1], grad_floor), (n, report[n])
    for n in b_t:
        if b_t[n].dtype == torch.long:
            assert torch.equal(b_f[n], b_t[n]), n
        else:
            assert rel_err(b_f[n], b_t[n]) < stat_tol, "buffer %s %g" % (n, rel_err(b_f[n], b_t[n]))
    return report


def sa_inputs(B, N, C, seed=0):
    xyz, full = cases.chunk_xyz(B, npoints=N, first=seed)
    rng = np.random.default_rng(seed)
    feats = rng.standard_normal((B, C, N)).astype(np.float32)
    return [(cases.t(xyz, DEV), False), (cases.t(feats, DEV), True)]


def test_sa_module_ssg():
    torch.manual_seed(0)
    sa = M.PointnetSAModule(npoint=256, radius=0.2, nsample=32, mlp=[6, 32, 32, 64]).to(DEV)
    class_bound(sa, sa_inputs(2, 2048, 6))


def test_sa_module_ssg_wide():
    """SA3 / SA4 of the SSG model (pointnet2_semseg.py:112-114): the pooled 256- and 512-wide layers and the
    259-channel first layer take the split adjoint (mlp_wgrad.cu + data gradient on the forward kernel)."""
    torch.manual_seed(0)
    sa3 = M.PointnetSAModule(npoint=64, radius=0.4, nsample=32, mlp=[128, 128, 128, 256]).to(DEV)
    # (8 / 16 clouds: a pooled channel's bias gradient sums one term per group, and with only ~100 groups a
    # single ReLU decision that TF32 flips moves it by a percent)
    class_bound(sa3, sa_inputs(8, 256, 128, seed=1), grad_floor=2e-2)
    sa4 = M.PointnetSAModule(npoint=16, radius=0.8, nsample=32, mlp=[256, 256, 256, 512]).to(DEV)
    class_bound(sa4, sa_inputs(16, 64, 256, seed=2), grad_floor=2e-2)


def test_sa_module_msg_odd_widths():
    torch.manual_seed(0)
    sa = M.PointnetSAModuleMSG(npoint=128, radii=[0.1, 0.2], nsamples=[16, 32],
                               mlps=[[64, 64, 96, 128], [64, 128, 196, 256]]).to(DEV)
    class_bound(sa, sa_inputs(2, 1024, 64, seed=3))


def fp_inputs(n, m, c1, c2, seed=1):
    xyz, _ = cases.chunk_xyz(2, npoints=n)
    unknown = cases.t(xyz, DEV)
    known = unknown[:, :m].contiguous()
    rng = np.random.default_rng(seed)
    uf = cases.t(rng.standard_normal((2, c1, n)).astype(np.float32), DEV)
    kf = cases.t(rng.standard_normal((2, c2, m)).astype(np.float32), DEV)
    return [(unknown, False), (known, False), (uf, True), (kf, True)]


def test_fp_module():
    torch.manual_seed(0)
    fp = M.PointnetFPModule(mlp=[128 + 6, 128, 128]).to(DEV)
    class_bound(fp, fp_inputs(2048, 256, 6, 128))


def test_fp_module_wide():
    """FP1 / FP3 of the SSG model ([320, 256, 128] and [768, 256, 256]): split adjoint with 320- / 768-column
    data gradients."""
    torch.manual_seed(0)
    class_bound(M.PointnetFPModule(mlp=[256 + 64, 256, 128]).to(DEV), fp_inputs(1024, 256, 64, 256, seed=2))
    class_bound(M.PointnetFPModule(mlp=[512 + 256, 256, 256]).to(DEV), fp_inputs(64, 16, 256, 512, seed=3))


def test_eval_mode_uses_running_stats():
    torch.manual_seed(0)
    sa = M.PointnetSAModule(npoint=64, radius=0.3, nsample=16, mlp=[6, 16, 32]).to(DEV)
    for m in sa.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(), m.running_var.uniform_(0.5, 2.0)
    ins = sa_inputs(2, 512, 6)
    with torch.no_grad():
        with torch_tf32(False):
            out_t = run_noback(sa, False, ins)
        with torch_tf32(True):
            out_r = run_noback(sa, False, ins)
        out_f = run_noback(sa, True, ins)
    assert rel_err(out_f, out_t) <= max(4 * rel_err(out_r, out_t), 5e-3), (rel_err(out_f, out_t), rel_err(out_r, out_t))


def run_noback(module, enabled, inputs):
    module = copy.deepcopy(module).eval()
    fused.ENABLED = enabled
    try:
        return module(*[i for i, _ in inputs])[1]
    finally:
        fused.ENABLED = True


def model_step(model, pc, enabled):
    m = copy.deepcopy(model).train()
    m.cls_layer[1].p = 0.0  # the dropout mask is drawn per layout; keep the comparison deterministic
    fused.ENABLED = enabled
    try:
        logits = m(pc)
        torch.manual_seed(6)
        (logits * torch.randn_like(logits)).sum().backward()  # (a mean of squares after the
        # final BatchNorm would be constant: every gradient would be rounding noise)
    finally:
        fused.ENABLED = True
    return logits.detach(), {n: p.grad for n, p in m.named_parameters()}


@pytest.mark.parametrize("is_msg,c0", [(False, 6), (True, 6), (False, 131), (True, 131)])
def test_full_model_step_matches_torch_nn(is_msg, c0):
    """Whole SSG / MSG network (BASELINE configs 2, 3, 5: 6 or 131 feature channels), one forward/backward on
    the tcgen05 path against the torch.nn formulation in fp32, bounded by the torch.nn formulation's own
    TF32 deviation.  Through 4 SA + 4 FP levels TF32 differences flip ReLU / max-pool decisions, so the
    gradient is compared in the L2 norm over all parameters and single parameters loosely."""
    from pointnet2.scannet_b200 import synthetic
    torch.manual_seed(0)
    model = pointnet2_semseg.get_model(20, is_msg=is_msg, input_channels=c0).to(DEV)
    x, _, _ = synthetic.make_batch(2, 2, 8192, use_multiview=(c0 == 131), use_color=(c0 != 131))
    pc = cases.t(x, DEV)
    assert pc.shape[2] == 3 + c0
    with torch_tf32(False):
        lt, gt = model_step(model, pc, False)
    with torch_tf32(True):
        lr, gr = model_step(model, pc, False)
    lf, gf = model_step(model, pc, True)
    e_f, e_r = rel_err(lf, lt), rel_err(lr, lt)
    assert e_f <= max(4 * e_r, 2e-2), (e_f, e_r)

    def flat(g):
        return torch.cat([g[n].flatten() for n in gt])

    l2_f, l2_r = l2_err(flat(gf), flat(gt)), l2_err(flat(gr), flat(gt))
    assert l2_f <= max(4 * l2_r, 5e-2), (l2_f, l2_r)
    worst = max((l2_err(gf[n], gt[n]), n) for n in gt)
    worst_r = max((l2_err(gr[n], gt[n]), n) for n in gt)
    assert worst[0] <= max(4 * worst_r[0], 0.2), (worst, worst_r)


def test_flat_adam_matches_torch_adam():
    """pn2_adam_flat over one flat buffer == torch.optim.Adam (eps outside the root, bias correction,
    L2 weight decay) step for step: parameters within 1e-6 after 5 steps."""
    from pointnet2.scannet_b200 import train
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(17, 33), torch.nn.ReLU(), torch.nn.Linear(33, 5)).to(DEV)
    mine = copy.deepcopy(ref)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-3)
    opt = train.FlatAdam(mine.parameters(), lr=1e-2, weight_decay=1e-3)
    for i in range(5):
        torch.manual_seed(i)
        x = torch.randn(64, 17, device=DEV)
        opt_ref.zero_grad()
        ref(x).square().sum().backward()
        opt_ref.step()
        opt.zero_grad()
        mine(x).square().sum().backward()
        opt.pack_grads()
        opt.step()
    for a, b in zip(mine.parameters(), ref.parameters()):
        assert (a - b).abs().max() < 1e-6, (a - b).abs().max()
    assert int(opt.step_count.item()) == 5
    assert set(mine.state_dict()) == set(ref.state_dict())


@pytest.mark.parametrize("warmup", [0, 1])
def test_pipelined_step_equals_eager_step(warmup):
    """CUDA-graph replay with the geometry of the next batch prefetched on a side stream: (1) the plan
    a step trains with is, bit for bit, the plan of ITS batch (not a stale one); (2) losses follow the
    eager step's (1e-2: Adam's first steps amplify the summation-order noise of the backward pass --
    atomic scatter-adds, and weight gradients accumulated over dynamically scheduled row tiles)."""
    from pointnet2.scannet_b200 import geometry, synthetic, train
    batches = []
    for i in range(3):
        x, y, w = synthetic.make_batch(2, 2, 4096, first_chunk=10 * i)
        batches.append(tuple(torch.from_numpy(a).to(DEV) for a in (x, y, w)))

    def make():
        torch.manual_seed(0)
        m = pointnet2_semseg.get_model(20, is_msg=False, input_channels=6).to(DEV).train()
        m.cls_layer[1].p = 0.0
        return train.TrainStep(m, lr=1e-3)

    eager = make()                   # (the pipelined step's warm-up passes are undone: parameters, Adam state
    want = [float(eager(*batches[i % 3]).item()) for i in range(5)]   # and BatchNorm buffers are restored)
    # warmup = 1: the weight-tile packing runs as a forked branch of the captured graph (fused.pack_branch);
    # warmup = 0: no eager step allocated its persistent buffers, so the graph packs in line
    piped = train.PipelinedTrainStep(make(), batches[2], warmup=warmup)
    fresh_model = make().model               # building the graphed step must not have trained the model
    for (n, a), (_, b) in zip(piped.step.model.state_dict().items(), fresh_model.state_dict().items()):
        assert torch.equal(a, b), "warm-up changed %s" % n
    assert int(piped.step.opt.step_count.item()) == 0 and float(piped.step.opt.exp_avg.abs().max()) == 0.0
    piped.prefetch(*batches[0])
    got = []
    for i in range(5):
        loss = piped()
        piped.prefetch(*batches[(i + 1) % 3])
        got.append(float(loss.item()))
        torch.cuda.synchronize()
        assert torch.equal(piped.static[0], batches[i % 3][0])
        fresh = geometry.build_plan(piped.step.model, batches[i % 3][0][..., 0:3].contiguous())
        for a, b in zip(piped.plan.tensors(), fresh.tensors()):
            assert torch.equal(a, b), "step %d trained with a plan that is not its batch's" % i
    assert np.allclose(got, want, rtol=1e-2), (got, want)


def test_dropout_inside_the_chain_matches_materialised_dropout():
    """nn.Dropout between two units (the classifier head, pointnet2_semseg.py:152-159) applied inside the
    second unit's kernels, against the same units run one by one with the dropped activation
    materialised in between -- same keep mask, same kernels otherwise."""
    from pointnet2.scannet_b200 import pytorch_utils as pt_utils
    torch.manual_seed(0)
    P, p_drop = 5000, 0.5
    u0 = pt_utils.Conv1d(64, 128, bn=True).to(DEV).train()
    u1 = pt_utils.Conv1d(128, 128, bn=True).to(DEV).train()
    u2 = pt_utils.Conv1d(128, 20, activation=None, bn=True).to(DEV).train()
    x = torch.randn(P, 64, device=DEV)
    mask = torch.empty(P, 128, dtype=torch.uint8, device=DEV).bernoulli_(1.0 - p_drop)
    w = torch.randn(P, 20, device=DEV)
    assert fused.chain_supports_dropout([u0, u1, u2], x, 1)

    def run_chain(chained):
        units = [copy.deepcopy(u) for u in (u0, u1, u2)]
        xi = x.clone().requires_grad_(True)
        if chained:
            y = fused.run_mlp_rows(units, xi, 0, {1: (p_drop, mask)})
        else:
            h = fused.run_mlp_rows(units[:2], xi, 0)
            h = h * mask.float() * (1.0 / (1.0 - p_drop))
            y = fused.run_mlp_rows(units[2:], h, 0)
        (y * w).sum().backward()
        grads = [q.grad for u in units for q in u.parameters()]
        bufs = [b.clone() for u in units for b in u.buffers() if b.dtype != torch.long]
        return y.detach(), xi.grad, grads, bufs

    y1, gx1, g1, b1 = run_chain(True)
    y2, gx2, g2, b2 = run_chain(False)
    assert rel_err(y1, y2) < 1e-3, rel_err(y1, y2)
    assert l2_err(gx1, gx2) < 5e-3, l2_err(gx1, gx2)
    for a, b in zip(g1, g2):
        assert l2_err(a, b) < 5e-3, l2_err(a, b)
    for a, b in zip(b1, b2):
        assert rel_err(a, b) < 1e-3


def test_dropout_mask_generator():
    """pn2_dropout_mask: keep rate, byte values, a fresh mask per launch (device-side counter), tails that are
    not a multiple of 16 bytes, reproducibility from (seed, counter)."""
    from pointnet2.scannet_b200 import lib
    L = lib.load()
    s = torch.cuda.current_stream().cuda_stream
    state = torch.tensor([1234, 0], dtype=torch.int64, device=DEV)
    n = 1 << 20
    m1 = torch.full((n + 16,), 7, dtype=torch.uint8, device=DEV)
    m2 = torch.empty(n, dtype=torch.uint8, device=DEV)
    lib.check(L.pn2_dropout_mask(n, 0.5, state.data_ptr(), m1.data_ptr(), s), "mask")
    lib.check(L.pn2_dropout_mask(n, 0.5, state.data_ptr(), m2.data_ptr(), s), "mask")
    torch.cuda.synchronize()
    assert state.tolist() == [1234, 2]
    assert int(m1[n:].min()) == 7 and int(m1[:n].max()) == 1           # nothing written past n; bytes are 0 / 1
    for m in (m1[:n], m2):
        assert abs(float(m.float().mean()) - 0.5) < 5e-3
    assert float((m1[:n] != m2).float().mean()) > 0.4                   # a fresh draw per launch
    state2 = torch.tensor([1234, 0], dtype=torch.int64, device=DEV)
    m3 = torch.empty(n, dtype=torch.uint8, device=DEV)
    lib.check(L.pn2_dropout_mask(n, 0.5, state2.data_ptr(), m3.data_ptr(), s), "mask")
    assert torch.equal(m3, m1[:n])                                      # same (seed, counter) -> same mask
    odd = torch.full((64,), 9, dtype=torch.uint8, device=DEV)
    lib.check(L.pn2_dropout_mask(37, 0.8, state2.data_ptr(), odd.data_ptr(), s), "mask")
    torch.cuda.synchronize()
    assert int(odd[37:].min()) == 9 and int(odd[:37].max()) <= 1
    big = torch.empty(1 << 22, dtype=torch.uint8, device=DEV)
    lib.check(L.pn2_dropout_mask(big.numel(), 0.8, state2.data_ptr(), big.data_ptr(), s), "mask")
    assert abs(float(big.float().mean()) - 0.8) < 2e-3


@pytest.mark.parametrize("B,n,m,c2,c1", [(3, 8192, 1024, 128, 6), (2, 1024, 256, 256, 64), (2, 256, 64, 256, 128),
                                         (2, 64, 16, 512, 256), (1, 1000, 100, 64, 0), (2, 777, 50, 96, 3),
                                         (1, 300, 40, 36, 5)])
def test_fp_interp_rows_bit_exact(B, n, m, c2, c1):
    """pn2_fp_interp_rows (whichever of the gather kernels -- 1, 2 or 4 elements per thread -- the launcher
    picks for the shape) against the reference's arithmetic evaluated by torch elementwise ops:
    ((k1*w1) + (k2*w2)) + (k3*w3), one rounding per operation (interpolate_gpu.cu:95-98) -- bit-exact --
    followed by the skip columns and zero padding (pointnet2_modules.py:407-412)."""
    g = torch.Generator(device="cpu").manual_seed(n + m)
    known = torch.randn(B, m, c2, generator=g).to(DEV)
    skip = torch.randn(B, n, c1, generator=g).to(DEV) if c1 else None
    idx = torch.randint(0, m, (B, n, 3), generator=g, dtype=torch.int32).to(DEV)
    w = torch.rand(B, n, 3, generator=g).to(DEV)
    rows = fused.FPInterpRows.apply(known, skip, idx, w)
    ct = fused.pad4(c2 + c1)
    assert rows.shape == (B * n, ct)
    gath = [torch.gather(known, 1, idx[:, :, t].long().unsqueeze(-1).expand(-1, -1, c2)) for t in range(3)]
    want = (gath[0] * w[:, :, 0:1] + gath[1] * w[:, :, 1:2]) + gath[2] * w[:, :, 2:3]
    got = rows.view(B, n, ct)
    assert torch.equal(got[:, :, :c2], want)
    if c1:
        assert torch.equal(got[:, :, c2:c2 + c1], skip)
    assert float(got[:, :, c2 + c1:].abs().max()) == 0.0 if ct > c2 + c1 else True
