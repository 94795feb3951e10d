"""CPU-side tests: the drop-in boundary, the host mirror against the reference's own Python
(imported from /root/reference where it exists), the C ABI's symbol table, the DP gradient bucket
over gloo with world_size 2, and the synthetic workload."""
import builtins
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pointnet2"
have_ref = os.path.isdir(REF)


def _import_reference(name):
    """Import one of the reference's top-level modules (pointnet2_utils, ...) from /root/reference."""
    builtins.__POINTNET2_SETUP__ = False
    if REF not in sys.path:
        sys.path.append(REF)
    return importlib.import_module(name)


# ------------------------------------------------------------------ boundary
def test_abi_exports_every_declared_symbol():
    import ctypes
    from pointnet2.scannet_b200 import build
    build.build()
    header = open(os.path.join(ROOT, "include", "pn2_b200.h")).read()
    declared = set(re.findall(r"PN2_API[^;(]*?\b(pn2_\w+)\s*\(", header))
    assert len(declared) >= 30, declared
    L = ctypes.CDLL(build.LIB)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert L.pn2_abi_version() == 1
    L.pn2_status_string.restype = ctypes.c_char_p
    assert L.pn2_status_string(0) == b"ok" and b"invalid" in L.pn2_status_string(-1)
    # argument validation happens before any CUDA call: no GPU needed
    assert L.pn2_furthest_point_sampling(-1, 8, 2, None, None, None, None) == -1
    assert L.pn2_three_nn(1, 8, 2, None, None, None, None, None) == -1
    assert L.pn2_query_ball_point(0, 8, 2, ctypes.c_float(0.1), 4, None, None, None, None) == 0


def test_ctypes_signatures_cover_the_header():
    from pointnet2.scannet_b200 import lib
    header = open(os.path.join(ROOT, "include", "pn2_b200.h")).read()
    declared = set(re.findall(r"PN2_API[^;(]*?\b(pn2_\w+)\s*\(", header))
    assert declared == set(lib.SIGNATURES) | set(lib.LIBRARY_SYMBOLS)


def test_no_cpu_fallback_and_loud_failure():
    from pointnet2.scannet_b200 import ext
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.furthest_point_sampling(torch.rand(1, 16, 3), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ext.ball_query(torch.rand(1, 4, 3), torch.rand(1, 16, 3), 0.1, 4)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pointnet2")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f
                assert "pn2o_" not in src, f


def test_lead_path_eligibility_and_byte_models():
    """Host logic of csrc/mlp_lead.cu's path: the first conv of a block is applied before the gather only in
    training mode, with at least one more unit behind it and a power-of-two width the kernels tile exactly;
    the per-kernel table models the lead kernels' algorithmic bytes."""
    from pointnet2.scannet_b200 import fused, profiling, pytorch_utils as pt_utils
    ok = pt_utils.SharedMLP([9, 32, 32, 64], bn=True).train()
    assert fused.lead_ok(ok) and fused.lead_ok(ok, 16) and not fused.lead_ok(ok, 64)
    assert not fused.lead_ok(pt_utils.SharedMLP([9, 32, 32, 64], bn=True).eval())
    assert not fused.lead_ok(pt_utils.SharedMLP([9, 24, 32], bn=True).train())       # 24: not a power of two
    assert not fused.lead_ok(pt_utils.SharedMLP([9, 32], bn=True).train())           # nothing behind the first unit
    assert fused.lead_ok(pt_utils.SharedMLP([134, 128, 128], bn=True).train(), 64)
    saved = fused.LEAD
    try:
        fused.LEAD = False
        assert not fused.lead_ok(ok)
    finally:
        fused.LEAD = saved
    b, n, m, c, c1 = 32, 8192, 1024, 128, 6
    fwd, _ = profiling._lib_bytes("pn2_fp_lead_fwd", (b, n, m, c, c1))
    bwd, _ = profiling._lib_bytes("pn2_fp_lead_bwd", (b, n, m, c, c1, 1))
    assert fwd == 24 * b * n + 4 * b * m * c + 4 * b * n * c1 + 4 * b * n * c
    assert bwd == 8 * b * n * c + 24 * b * n + 4 * b * n * c1 + 4 * b * m * c
    sa, _ = profiling._lib_bytes("pn2_sa_lead_bwd", (32, 8192, 1024, 32, 32, 1))
    assert sa > 8 * 32 * 1024 * 32 * 32              # dH and z of 2^20 rows x 32 channels, read once each


def test_programmatic_launches_wait_first():
    """Kernels launched through pn2::launch (programmatic dependent launch, pn2_common.cuh) may be scheduled
    before the previous kernel of their stream has finished: each of them must call pdl_enter() --
    griddepcontrol.wait -- as its FIRST statement, and a source file must not mix the two launch forms."""
    csrc = os.path.join(ROOT, "pointnet2", "scannet_b200", "csrc")
    seen = 0
    for f in sorted(os.listdir(csrc)):
        if not f.endswith(".cu"):
            continue
        src = open(os.path.join(csrc, f)).read()
        if "pn2::launch(" not in src:
            continue
        seen += 1
        assert "<<<" not in src, "%s mixes <<< >>> launches with pn2::launch" % f
        code = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"__global__", code):
            depth, i = 0, code.index("(", m.end())
            if "__launch_bounds__" in code[m.end():i]:       # skip its argument list
                i = code.index(")", i) + 1
                i = code.index("(", i)
            for j in range(i, len(code)):
                depth += code[j] == "("
                depth -= code[j] == ")"
                if depth == 0:
                    break
            body = code[code.index("{", j) + 1:]
            assert body.lstrip().startswith("pdl_enter();"), "%s: %s" % (f, code[m.start():m.start() + 120])
    assert seen >= 6


@pytest.mark.skipif(not have_ref, reason="/root/reference not present")
def test_reference_op_layer_resolves_to_the_dropin():
    """`import pointnet2._ext` inside the reference's unchanged pointnet2_utils.py finds this repo's."""
    ref_utils = _import_reference("pointnet2_utils")
    import pointnet2._ext as dropin
    assert ref_utils._ext is dropin
    from pointnet2.scannet_b200 import ext
    for name in ext.__all__:
        assert getattr(dropin, name) is getattr(ext, name)


@pytest.mark.skipif(not have_ref, reason="/root/reference not present")
@pytest.mark.parametrize("msg,channels", [(False, 6), (True, 6), (False, 131), (True, 131)])
def test_state_dict_matches_reference(msg, channels):
    from pointnet2.scannet_b200 import pointnet2_semseg as mine
    ref = _import_reference("pointnet2_semseg")
    torch.manual_seed(0)
    a = mine.get_model(20, msg, channels).state_dict()
    torch.manual_seed(0)
    b = ref.get_model(20, msg, channels).state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    expect = {(False, 6): 949384, (False, 131): 969384, (True, 6): 3027752, (True, 131): 3049752}
    assert sum(v.numel() for k, v in a.items() if "running" not in k and "tracked" not in k) == expect[(msg, channels)]


@pytest.mark.skipif(not have_ref, reason="/root/reference not present")
def test_mirror_modules_equal_reference_modules_on_cpu(oracle):
    """Same ops table (the oracle's CPU one), same seed: the mirror's SA / FP modules and whole SSG
    model reproduce the reference's Python modules bit for bit, forward and backward."""
    from pointnet2.scannet_b200 import pointnet2_semseg as mine, pointnet2_utils as my_utils, synthetic
    ref = _import_reference("pointnet2_semseg")
    ref_utils = _import_reference("pointnet2_utils")
    saved = (my_utils._ext, ref_utils._ext)
    my_utils._ext = ref_utils._ext = oracle.ext_cpu
    try:
        x, _, _ = synthetic.make_batch(2, 2, npoints=2048)
        pc = torch.from_numpy(x)
        outs = []
        for mod in (mine, ref):
            torch.manual_seed(0)
            model = mod.get_model(20, is_msg=False, input_channels=6).train()
            torch.manual_seed(1)
            logits = model(pc)
            torch.manual_seed(2)
            (logits * torch.randn_like(logits)).sum().backward()
            outs.append((logits.detach(), [p.grad for p in model.parameters()], model.state_dict()))
    finally:
        my_utils._ext, ref_utils._ext = saved
    (la, ga, sa), (lb, gb, sb) = outs
    assert torch.equal(la, lb)
    assert all(torch.equal(x, y) for x, y in zip(ga, gb))
    assert all(torch.equal(sa[k], sb[k]) for k in sa)


@pytest.mark.skipif(not have_ref, reason="/root/reference not present")
@pytest.mark.parametrize("msg,channels", [(False, 6), (True, 131)])
def test_reference_checkpoint_round_trip_on_cpu(oracle, tmp_path, msg, channels):
    """a model.pth in the reference's format -- ``torch.save(model.state_dict(), ...)`` of the
    reference's own ``get_model`` (lib/solver.py:142, 325) -- loads into the mirror with ``strict=True`` as
    scripts/eval.py:183-188 loads it, and the mirror in eval mode reproduces the reference's logits bit for
    bit (both on the oracle's CPU ops table)."""
    from pointnet2.scannet_b200 import pointnet2_semseg as mine, pointnet2_utils as my_utils, synthetic
    ref = _import_reference("pointnet2_semseg")
    ref_utils = _import_reference("pointnet2_utils")
    saved = (my_utils._ext, ref_utils._ext)
    my_utils._ext = ref_utils._ext = oracle.ext_cpu
    try:
        torch.manual_seed(3)
        ref_model = ref.get_model(20, is_msg=msg, input_channels=channels)
        g = torch.Generator().manual_seed(4)
        for m in ref_model.modules():          # a trained-looking checkpoint: non-trivial BatchNorm state
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.running_mean.normal_(0, 0.2, generator=g), m.running_var.uniform_(0.5, 2.0, generator=g)
                m.weight.data.uniform_(0.5, 1.5, generator=g), m.bias.data.normal_(0, 0.1, generator=g)
                m.num_batches_tracked.fill_(1234)
        path = str(tmp_path / "model.pth")
        torch.save(ref_model.state_dict(), path)                      # lib/solver.py:325
        mirror = mine.get_model(20, is_msg=msg, input_channels=channels)
        missing = mirror.load_state_dict(torch.load(path), strict=True)   # scripts/eval.py:183-188
        assert not missing.missing_keys and not missing.unexpected_keys
        x, _, _ = synthetic.make_batch(2, 1, npoints=1024, use_multiview=(channels == 131), use_color=(channels != 131))
        pc = torch.from_numpy(x)
        with torch.no_grad():
            want = ref_model.eval()(pc)
            got = mirror.eval()(pc)
    finally:
        my_utils._ext, ref_utils._ext = saved
    assert got.shape == want.shape == (1, 1024, 20)
    assert torch.equal(got, want)


def test_groupall_and_surface_names():
    from pointnet2.scannet_b200 import pointnet2_utils as U
    for name in ("furthest_point_sample", "gather_operation", "three_nn", "three_interpolate",
                 "grouping_operation", "ball_query", "QueryAndGroup", "GroupAll", "RandomDropout",
                 "FurthestPointSampling", "GatherOperation", "ThreeNN", "ThreeInterpolate",
                 "GroupingOperation", "BallQuery"):
        assert hasattr(U, name), name
    xyz, feats = torch.rand(2, 10, 3), torch.rand(2, 4, 10)
    out = U.GroupAll()(xyz, None, feats)
    assert out.shape == (2, 7, 1, 10)
    out, gx = U.GroupAll(ret_grouped_xyz=True)(xyz, None, feats)  # the reference raises here (ref :422)
    assert gx.shape == (2, 3, 1, 10)


# ------------------------------------------------------------------ workload / accounting
def test_synthetic_chunks_are_scannet_shaped():
    from pointnet2.scannet_b200 import synthetic
    x, y, w = synthetic.make_batch(2, 2)
    assert x.shape == (2, 8192, 9) and x.dtype == np.float32 and y.shape == (2, 8192)
    x2, _, _ = synthetic.make_batch(2, 2)
    assert np.array_equal(x, x2)  # seeded
    uniq = len(np.unique(x[0, :, :3], axis=0)) / 8192.0
    assert 0.6 < uniq < 0.9, uniq  # drawn with replacement: many exact duplicates (dataset.py:248)
    assert np.abs(np.linalg.norm(x[0, :, 6:9], axis=1) - 1).max() < 1e-5  # unit normals
    xm, _, _ = synthetic.make_batch(5, 1, use_color=False, use_multiview=True)
    assert xm.shape == (1, 8192, 3 + 128 + 3)
    scene = synthetic.make_scene(0, npoints=20000)
    tiles = synthetic.tile_scene(scene)
    assert tiles.shape[1:] == (8192, 9) and tiles.shape[0] >= 12


def test_algorithmic_bytes_match_survey_totals():
    """compulsory traffic of the custom ops of one SSG 32x8192 forward = 437 MB."""
    from pointnet2.scannet_b200 import profiling
    B = 32
    lv = [(8192, 1024, 6), (1024, 256, 64), (256, 64, 128), (64, 16, 256)]
    t = lambda *s: torch.empty(*s, device="meta")  # noqa: E731
    total = 0
    for n, m, c in lv:
        total += profiling.algorithmic_bytes("furthest_point_sampling", (t(B, n, 3), m), None)
        total += profiling.algorithmic_bytes("gather_points", (t(B, 3, n), t(B, m)), None)
        total += profiling.algorithmic_bytes("ball_query", (t(B, m, 3), t(B, n, 3), 0.1, 32), None)
        total += profiling.algorithmic_bytes("group_points", (t(B, 3, n), t(B, m, 32)), None)
        total += profiling.algorithmic_bytes("group_points", (t(B, c, n), t(B, m, 32)), None)
    for n, m, c in [(64, 16, 512), (256, 64, 256), (1024, 256, 256), (8192, 1024, 128)]:
        total += profiling.algorithmic_bytes("three_nn", (t(B, n, 3), t(B, m, 3)), None)
        total += profiling.algorithmic_bytes("three_interpolate", (t(B, c, m), t(B, n, 3), t(B, n, 3)), None)
    assert abs(total / 1e6 - 437) < 15, total / 1e6


# ------------------------------------------------------------------ data parallel (gloo, 2 ranks)
_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from pointnet2.scannet_b200.train import GradBucket, TrainStep
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
torch.manual_seed(0)
model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
bucket = GradBucket(model.parameters())
assert bucket.attached()
torch.manual_seed(100 + rank)
x = torch.randn(4, 5)
model(x).square().sum().backward()
local = bucket.flat.clone()
bucket.all_reduce_mean()
gathered = [torch.zeros_like(local) for _ in range(2)]
dist.all_gather(gathered, local)
assert torch.allclose(bucket.flat, (gathered[0] + gathered[1]) / 2, atol=1e-7)
assert bucket.attached() and all(torch.equal(p.grad.flatten(), bucket.flat[o:o + p.numel()])
                                 for p, o in zip(bucket.params, [0, 35, 42, 63]))
# a full TrainStep keeps the replicas identical
class Tiny(torch.nn.Module):
    def __init__(s):
        super().__init__(); s.l = torch.nn.Linear(9, 20)
    def forward(s, pc):
        return s.l(pc)
torch.manual_seed(0)
step = TrainStep(Tiny(), lr=1e-2, fused_adam=False)
for i in range(3):
    torch.manual_seed(10 * rank + i)
    pc = torch.randn(2, 16, 9); y = torch.randint(0, 20, (2, 16)); w = torch.ones(2, 16)
    step(pc, y, w)
flat = torch.cat([p.detach().flatten() for p in step.model.parameters()])
both = [torch.zeros_like(flat) for _ in range(2)]
dist.all_gather(both, flat)
assert torch.equal(both[0], both[1]), "replicas diverged"
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_grad_bucket_all_reduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o


# ------------------------------------------------------------------ loss mirror
def test_weighted_ce_mirror_equals_reference_on_cpu():
    """CPU tensors take the reference's torch formulation: same values as lib/loss.py, bit for bit."""
    from pointnet2.scannet_b200 import loss as L
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(500, 20, generator=g)
    targets = torch.randint(0, 20, (500,), generator=g)
    targets[::9] = -100
    weights = torch.rand(500, generator=g)
    assert not L.kernel_eligible(logits, targets, weights)
    mine = L.WeightedCrossEntropyLoss()(logits, targets, weights)
    per = torch.nn.functional.cross_entropy(logits, targets, reduction="none", ignore_index=-100)
    assert torch.equal(mine, torch.mean(per * weights))
    if os.path.isfile("/root/reference/lib/loss.py"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_loss", "/root/reference/lib/loss.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert torch.equal(mine, mod.WeightedCrossEntropyLoss()(logits, targets, weights))


def test_geometry_plan_pack_and_single_copy():
    """GeometryPlan.pack(): every tensor becomes a view into one flat buffer (values kept) and
    copy_from between two packed plans is one copy; packed <-> unpacked falls back to per-tensor copies."""
    from pointnet2.scannet_b200.geometry import GeometryPlan

    def make(fill):
        sa = [(fill(2, 4, 3), [fill(2, 4, 8).to(torch.int32), fill(2, 4, 5).to(torch.int32)]) for _ in range(2)]
        fp = [(fill(2, 9, 3).to(torch.int32), fill(2, 9, 3)) for _ in range(3)]
        return GeometryPlan(sa, fp)

    g = torch.Generator().manual_seed(0)
    src = make(lambda *s: torch.randint(0, 100, s, generator=g).float())
    want = [t.clone() for t in src.tensors()]
    packed = src.pack()
    assert packed._flat is not None and all(torch.equal(a, b) for a, b in zip(packed.tensors(), want))
    base = packed._flat.untyped_storage().data_ptr()
    assert all(t.untyped_storage().data_ptr() == base for t in packed.tensors())
    dst = make(lambda *s: torch.zeros(*s)).pack()
    dst.copy_from(packed)
    assert all(torch.equal(a, b) for a, b in zip(dst.tensors(), want))
    loose = make(lambda *s: torch.zeros(*s))
    loose.copy_from(packed)
    assert all(torch.equal(a, b) for a, b in zip(loose.tensors(), want))


def test_chunker_augmentation_draws_follow_the_reference_order():
    """The product's host-side draws (chunker._draw_augmentation) consume np.random exactly like the
    oracle's restatement of _augment / _translate / _rotate / _scale (pinned against the reference's own
    classes in test_chunker_cpu.py) and yield the same parameters, for every branch of _augment."""
    import numpy as np
    from oracle import chunker_np as O
    from pointnet2.scannet_b200 import chunker
    seen = set()
    for seed in range(60):
        np.random.seed(seed)
        ops, t, R, f = O.draw_augmentation(np.random)
        after_oracle = np.random.random()
        np.random.seed(seed)
        mask, params = chunker._draw_augmentation()
        assert np.random.random() == after_oracle          # same number of draws
        assert mask == sum({"t": 1, "r": 2, "s": 4}[o] for o in ops)
        assert np.array_equal(params[:3], t) and np.array_equal(params[3:12], R.ravel()) and params[12] == f
        seen.add(ops)
    assert seen == {"t", "r", "s", "tr", "ts", "rs", "trs", ""}


def test_chunker_descriptor_matches_the_library():
    from pointnet2.scannet_b200 import chunker, lib
    assert lib.load().pn2_chunk_desc_bytes() == chunker.DESC.itemsize


def test_bench_prints_exactly_one_json_line():
    """The driver parses bench.py's stdout: ONE JSON line, whatever the libraries print.  Exercised on the arm
    that runs without a GPU (the reference arm's CPU-port fallback, bounded sample)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference-cpu", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-chunks", "1", "--npoints", "512"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.strip().splitlines()
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun the reference arm runs on rank 0 alone; every other rank exits 0 without output."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def _product_line(with_per_op):
    """What run_product hands to the formatter, at its longest (8 GPUs, long kernel name, throttle reasons)."""
    import bench
    res = {"metric": bench.METRIC, "value": 123456.7, "unit": bench.UNIT, "n_gpus": 8, "steps": 200, "warmup": 50,
           "ms_per_step": 12.3456, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": "configs[2] MSG train step 32x8192/GPU C=134", "per_gpu_batch": 32, "npoints": 8192,
                      "parallelism": "dp8", "launch": "graph(fwd_bwd)+geometry prefetch,split", "mlp": "tcgen05 tf32",
                      "settle": 50, "l2": "inputs>L2"},
           "e2e": {"value": 123456.7, "unit": bench.UNIT, "h2d_bytes_per_step": 146800640, "d2h_bytes_per_step": 4},
           "gpu_launches": 1234567, "loss": 2.99123,
           "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0,
                      "reasons": ["hw_slowdown", "hw_thermal_slowdown", "sw_power_cap", "sw_thermal_slowdown"]}}
    if with_per_op:
        res["roofline"] = {"bound": "hbm", "kernel": "pn2_mlp_bwd_layer(1048576,128,128,128,32)", "achieved": 3518.25,
                           "peak": 6536.4, "unit": "GB/s", "frac": 0.5367, "traffic": 524600000,
                           "peak_source": "MEASURED_PEAKS", "share": 0.312}
        res["fps_ns_per_iter"] = 631.2
    res["cpu_baseline"] = {"value": 13.37, "unit": bench.UNIT, "cores": 128, "kind": "port",
                           "sample": "2x8192 pts fwd+bwd+Adam, oracle C ops+torch CPU, median of 10"}
    return bench.format_line(res)


@pytest.mark.parametrize("with_per_op", [True, False])
def test_bench_product_line_fits_the_drivers_tail(with_per_op):
    """Round 1's line was 32 KB and the driver, which parses the tail of stdout, recorded nothing."""
    import json
    line = _product_line(with_per_op)
    assert len(line) < 1000 and "\n" not in line
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "cpu_baseline"):
        assert k in d, k
    assert "workload" in d["config"] and set(d["e2e"]) == {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    if with_per_op:
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "cores", "kind"} <= set(d["cpu_baseline"])


def test_reference_arm_runs_the_references_own_python():
    """bench.py --impl reference imports the reference's pointnet2_semseg / lib.loss from oracle/_ref/py with
    `pointnet2._ext` set to the extension handed in -- exercised here on the CPU with the oracle's ops table."""
    from oracle import build_ref
    if not os.path.exists(os.path.join(build_ref.python_dir(), "pointnet2_semseg.py")):
        pytest.skip("oracle/_ref/py not staged (needs /root/reference)")
    import bench
    import oracle
    import torch
    saved = {k: sys.modules.get(k) for k in ("pointnet2._ext", "pointnet2_utils", "pointnet2_modules",
                                              "pointnet2_semseg", "pytorch_utils")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        semseg, loss = bench.load_reference_python(oracle.ext_cpu)
        assert os.path.realpath(semseg.__file__).startswith(os.path.realpath(build_ref.python_dir()))
        assert sys.modules["pointnet2_utils"]._ext is oracle.ext_cpu
        if os.path.isdir("/root/reference/pointnet2"):
            for f in build_ref.REF_PY:   # verbatim
                assert open(os.path.join(build_ref.python_dir(), f)).read() == \
                    open(os.path.join("/root/reference/pointnet2", f)).read()
        torch.manual_seed(0)
        m = semseg.get_model(num_classes=20, is_msg=False, input_channels=6, use_xyz=True, bn=True).train()
        y = m(torch.rand(2, 1024, 9))
        l = loss.WeightedCrossEntropyLoss()(y.view(-1, 20), torch.zeros(2048, dtype=torch.long), torch.ones(2048))
        l.backward()
        assert y.shape == (2, 1024, 20) and torch.isfinite(l)
    finally:
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
        import pointnet2
        if saved["pointnet2._ext"] is not None:
            pointnet2._ext = saved["pointnet2._ext"]
        elif hasattr(pointnet2, "_ext"):
            del pointnet2._ext
