#!/usr/bin/env python
"""Benchmark of the PointNet++ SA/FP hot path on B200 (BASELINE.json metric:
8192-point chunks/sec, forward+backward, SSG semseg training step, batch 32 x 8192).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one optimisation step -- forward, weighted
cross-entropy, backward, gradient all-reduce (N > 1), Adam -- on one batch of synthetic
ScanNet-shaped chunks.  `value` is timed with inputs resident in HBM; `e2e` is the same
step called with HOST buffers (pinned), the host->device copies of the batch and the
device->host read of the loss inside the timed region.

The line stays below 1000 characters (the driver parses the tail of stdout); the per-kernel
table of the step goes to a side file (--per-kernel-out, default gpurun_out/).

--impl reference times the UNMODIFIED REFERENCE on the same box: its own Python
(pointnet2_semseg.py / pointnet2_modules.py / pointnet2_utils.py / pytorch_utils.py and
lib/loss.py, staged verbatim into git-ignored oracle/_ref/py by oracle/build_ref.py), its own
CUDA extension (oracle/_ref/stock: pointnet2/_ext_src compiled unmodified with the
reference's flags for sm_100a), torch.optim.Adam and the iteration of lib/solver.py:161-176
+ 219-231.  No module of this repository's package is imported on that arm.  (The reference
has no CPU implementation -- every binding asserts "CPU not supported", e.g.
ball_query.cpp:27-29 -- so its own implementation of the path IS CUDA; north_star's target
is a ratio against exactly these kernels.)  If that build is not present the arm falls back
to the oracle's CPU port on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "8192-pt chunks/sec fwd+bwd (SSG semseg training step, 32x8192 per GPU)"
UNIT = "chunks/s"
CONFIG_ID = 2  # BASELINE.json configs[1]
NUM_CLASSES = 20
LINE_LIMIT = 1000   # characters; the driver reads only the tail of stdout
SETTLE_STEPS = 20   # both arms run at least this many untimed steps before the clock starts


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="product", choices=["product", "reference", "reference-cpu"])
    p.add_argument("--batch", type=int, default=32, help="chunks per GPU")
    p.add_argument("--npoints", type=int, default=8192)
    p.add_argument("--msg", action="store_true", help="MSG model (BASELINE configs[2]) instead of SSG")
    p.add_argument("--multiview", action="store_true", help="xyz+normal+128 features (configs[4])")
    p.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of from a CUDA graph")
    p.add_argument("--no-prefetch", action="store_true",
                   help="compute each batch's sampling/grouping geometry inline instead of one batch ahead")
    p.add_argument("--no-split", action="store_true",
                   help="one CUDA graph for the whole step (the next batch's geometry then starts beside the forward pass)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-per-op", action="store_true")
    p.add_argument("--cpu-sample-chunks", type=int, default=2)
    p.add_argument("--settle", type=int, default=SETTLE_STEPS,
                   help="untimed steps before the clock starts: max(--warmup, this)")
    p.add_argument("--per-kernel-out", default=None,
                   help="where the per-kernel table goes (default gpurun_out/bench_per_kernel_<cfg>_<N>gpu.json)")
    return p.parse_args()


# --------------------------------------------------------------------------- plumbing
def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif torch.cuda.is_available():
        torch.cuda.set_device(0)
    return world, rank, local


def barrier_sync(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9 or c[0] != str(self.index):
                continue
            try:
                sm.append(float(c[1])), mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower() == "active":
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def _synthetic(standalone=False):
    """The workload generator (numpy only).  standalone: loaded by file path, so that the reference
    arm does not import this repository's package at all."""
    if not standalone:
        from pointnet2.scannet_b200 import synthetic
        return synthetic
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "pn2_bench_synthetic", os.path.join(ROOT, "pointnet2", "scannet_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def host_batches(args, rank, count, standalone=False):
    """`count` distinct pinned host batches for this rank (weak scaling: own chunks per rank)."""
    synthetic = _synthetic(standalone)
    kw = dict(use_color=not args.multiview, use_normal=True, use_multiview=args.multiview)
    out = []
    for i in range(count):
        x, y, w = synthetic.make_batch(CONFIG_ID, args.batch, args.npoints,
                                       first_chunk=(rank * count + i) * args.batch, **kw)
        out.append(tuple(torch.from_numpy(a).pin_memory() if torch.cuda.is_available()
                         else torch.from_numpy(a) for a in (x, y, w)))
    return out, synthetic.input_channels(**kw)


def build_step(args, in_ch, device, capturable=False):
    from pointnet2.scannet_b200 import pointnet2_semseg, train
    torch.manual_seed(0)  # same replica on every rank
    model = pointnet2_semseg.get_model(NUM_CLASSES, is_msg=args.msg, input_channels=in_ch).to(device)
    model.train()
    return train.TrainStep(model, lr=1e-3)


def time_steps(step, batches, steps, world, e2e):
    """K steps bracketed by barrier+synchronize; CUDA events on the launching stream."""
    dev_batches = None
    if not e2e:
        dev_batches = [tuple(t.cuda(non_blocking=True) for t in b) for b in batches]
    elif not hasattr(step, "prefetch"):  # a graphed step owns static input buffers: copy straight into them
        stage = getattr(step, "static", None) or tuple(torch.empty_like(t, device="cuda") for t in batches[0])
    pipelined = hasattr(step, "prefetch")
    src = batches if e2e else dev_batches
    if pipelined:  # batch 0 is staged before the clock starts; every timed step stages the next one,
        step.prefetch(*src[0])  # so K batches are staged + K geometries computed inside the region
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(steps):
        if pipelined:
            last = step()                                   # trains on the batch staged one step ago
            step.prefetch(*src[(i + 1) % len(src)])         # host->device copy (e2e) + geometry, side stream
            if e2e:
                last = float(last.item())                   # device -> host read of the loss
        elif e2e:
            for d, h in zip(stage, batches[i % len(batches)]):
                d.copy_(h, non_blocking=True)          # host -> device, pinned source
            last = float(step(*stage).item())          # device -> host read of the loss
        else:
            last = step(*dev_batches[i % len(dev_batches)])
    if pipelined:   # the geometry of the batch staged by the last step belongs to the timed region
        torch.cuda.current_stream().wait_event(step.ready)
    e1.record()
    barrier_sync(world)
    ms = e0.elapsed_time(e1)
    if torch.is_tensor(last):
        last = float(last.item())
    return max_over_ranks(ms, world), last


# ------------------------------------------------------------------- cpu baseline (port)
def cpu_baseline(args, in_ch):
    """The oracle's C port of the nine ops (all host threads over clouds) under this
    package's model code + torch CPU layers: forward+backward+Adam on a bounded sample."""
    import oracle
    from pointnet2.scannet_b200 import pointnet2_semseg, pointnet2_utils, synthetic, train
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle.set_threads(cores)
    nb = args.cpu_sample_chunks
    x, y, w = synthetic.make_batch(CONFIG_ID, nb, args.npoints, use_color=not args.multiview,
                                   use_normal=True, use_multiview=args.multiview)
    saved = pointnet2_utils._ext
    pointnet2_utils._ext = oracle.ext_cpu
    try:
        torch.manual_seed(0)
        model = pointnet2_semseg.get_model(NUM_CLASSES, is_msg=args.msg, input_channels=in_ch)
        model.train()
        step = train.TrainStep(model, lr=1e-3, fused_adam=False)
        xs, ys, ws = torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(w)
        step(xs, ys, ws)  # warm-up
        times = []
        t_end = time.time() + 20.0
        while len(times) < 3 or (time.time() < t_end and len(times) < 10):
            t0 = time.time()
            step(xs, ys, ws)
            times.append(time.time() - t0)
    finally:
        pointnet2_utils._ext = saved
    sec = statistics.median(times)
    return {"value": round(nb / sec, 3), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%dx%d pts fwd+bwd+Adam, oracle C ops+torch CPU, median of %d" % (nb, args.npoints, len(times))}


# ------------------------------------------------------------------------------ arms
def run_product(args, world, rank, local):
    from pointnet2.scannet_b200 import lib, profiling
    lib.load()  # fail loudly if the CUDA extension is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    from pointnet2.scannet_b200 import train
    batches, in_ch = host_batches(args, rank, 4)
    eager = build_step(args, in_ch, "cuda", capturable=not args.no_graph)
    dev = [tuple(t.cuda() for t in b) for b in batches]
    step, launch_mode = eager, "eager"
    if not args.no_graph:
        try:  # single GPU: the whole step; DP: forward+backward in the graph, NCCL + Adam after it
            cls = train.GraphedTrainStep if args.no_prefetch else train.PipelinedTrainStep
            kw = {} if args.no_prefetch else {"split_fwd_bwd": not args.no_split}
            # PN2_DP_SCOPE=step: the gradient all-reduce and Adam captured inside the step graph as well
            scope = "step" if world == 1 else os.environ.get("PN2_DP_SCOPE", "fwd_bwd")
            step = cls(eager, dev[0], scope=scope, **kw)
            launch_mode = "graph(%s)%s" % (step.scope, "" if args.no_prefetch else
                                           "+geometry prefetch" + ("" if args.no_split else ",split"))
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("CUDA graph capture failed, running eagerly: %r\n" % (e,))
            torch.cuda.synchronize()
            step, launch_mode = eager, "eager (graph capture failed)"
    settle = max(args.warmup, args.settle)
    for i in range(settle):
        step(*dev[i % len(dev)])
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.reset_launch_count()
    ms, loss = time_steps(step, batches, args.steps, world, e2e=False)
    launches = lib.launch_count()
    if launch_mode.startswith("graph"):
        # replayed kernels do not pass through the library's launchers: count them once, eagerly
        lib.reset_launch_count()
        eager.forward_backward(*dev[0])
        torch.cuda.synchronize()
        launches = lib.launch_count() * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = time_steps(step, batches, args.steps, world, e2e=True)

    chunks = args.batch * world * args.steps
    res = {
        "metric": METRIC, "value": round(chunks / (ms * 1e-3), 1), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": "configs[%d] %s train step %dx%d/GPU C=%d" % (
                2 if args.msg else 1, "MSG" if args.msg else "SSG", args.batch, args.npoints, in_ch),
            "per_gpu_batch": args.batch, "npoints": args.npoints, "parallelism": "dp%d" % world,
            "launch": launch_mode, "mlp": "tcgen05 tf32", "settle": settle,
            "l2": "inputs>L2",
        },
        "e2e": {"value": round(chunks / (ms_e2e * 1e-3), 1), "unit": UNIT,
                "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in batches[0])),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "loss": round(loss, 5),
    }
    if rank != 0:
        return None
    res["clocks"] = {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")}

    # per-kernel timing: CUDA events around every call into libpn2_b200.so -> roofline
    if not args.no_per_op:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        which = "MEASURED_PEAKS" if "hbm_gbs" in peaks else "fallback"
        nprof = min(args.steps, 5)

        def local_step(*batch):
            # the step WITHOUT its collective: only rank 0 runs this pass (the other ranks are already waiting
            # at the final barrier), so the gradient all-reduce of TrainStep.__call__ would never be matched
            loss = eager.forward_backward(*batch)
            if eager.flat_mode:
                eager.opt.step(grad_scale=1.0)
            else:
                eager.optimizer.step()
            return loss

        # The eager pass is bound by the host (one C-ABI call per launch): with an idle GPU every pair of events
        # would also time the wait for its own launch to arrive.  A busy-wait kernel in front of each step lets the
        # host run ahead, so the launches are already queued and the events bracket the kernels alone.
        sleep_cycles = int(6e-3 * 1e6 * float(clocks.get("sm_max_mhz") or 1965.0))   # ~6 ms
        with profiling.LibTimer() as timer:
            for i in range(nprof):
                torch.cuda._sleep(sleep_cycles)
                local_step(*dev[i % len(dev)])
        rows = timer.summary(nprof)
        for r in rows:
            if r.get("gbs"):
                r["frac_of_hbm_peak"] = round(r["gbs"] / peak, 4)
        by_kernel = {}
        for r in rows:
            by_kernel[r["kernel"]] = by_kernel.get(r["kernel"], 0.0) + r["ms_per_step"]
        # dominant kernel of the step among those bounded by HBM bandwidth (FPS, ball query and
        # three_nn are latency / issue bound: reported in the side file and "fps", not as a roofline)
        not_hbm = ("pn2_furthest_point_sampling", "pn2_query_ball_point", "pn2_three_nn")
        modelled = {r["kernel"] for r in rows if r.get("gbs")}
        dom = max((k for k in by_kernel if k not in not_hbm and k in modelled), key=lambda k: by_kernel[k])
        top = max((r for r in rows if r["kernel"] == dom and r.get("gbs")), key=lambda r: r["ms_per_step"])
        traffic = None
        try:  # DRAM bytes of this very call from the committed ncu --set full capture, if there is one
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            traffic = tr.get(top["call"])
        except Exception:  # noqa: BLE001
            pass
        # algorithmic bytes (every input read once, every output written once: profiling._lib_bytes) of the
        # dominant kernel's largest launch / its CUDA-event duration on the launching stream
        res["roofline"] = {"bound": "hbm", "kernel": top["call"].replace(" ", ""), "achieved": top["gbs"],
                           "peak": peak, "unit": "GB/s", "frac": round(top["gbs"] / peak, 4), "traffic": traffic,
                           "peak_source": which,
                           "share": round(by_kernel[dom] / sum(by_kernel.values()), 3)}
        fps = [r for r in rows if r["kernel"] == "pn2_furthest_point_sampling"]
        if fps:
            f = fps[0]
            m = int(f["call"].split("(")[1].split(",")[2].strip(" )"))
            res["fps_ns_per_iter"] = round(1e3 * f["us_per_call"] / max(m - 1, 1), 1)
        side = {"line": dict(res), "per_kernel": rows,
                "lib_ms_per_step_eager": round(sum(r["ms_per_step"] for r in rows), 4),
                "lib_ms_per_step_by_kernel": {k: round(v, 4) for k, v in
                                              sorted(by_kernel.items(), key=lambda kv: -kv[1])}}
        path = args.per_kernel_out or os.path.join(
            ROOT, "gpurun_out", "bench_per_kernel_%s%s_%dgpu.json" % (
                "msg" if args.msg else "ssg", "_mv" if args.multiview else "", world))
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "w") as f:
                json.dump(side, f, indent=1)
        except OSError as e:
            sys.stderr.write("per-kernel table not written: %r\n" % (e,))
    if not args.no_cpu_baseline and world == 1:   # reported at N = 1 only
        try:
            res["cpu_baseline"] = cpu_baseline(args, in_ch)
        except Exception as e:  # noqa: BLE001
            res["cpu_baseline"] = {"unavailable": repr(e)[:80]}
    return res


# ------------------------------------------------------------------ the reference, unmodified
def load_reference_python(ext_module):
    """Import the reference's OWN model code (staged verbatim by oracle/build_ref.stage_python into
    git-ignored oracle/_ref/py) with `pointnet2._ext` resolving to `ext_module` -- the seam the reference
    itself uses (pointnet2_utils.py:25-33).  Returns (pointnet2_semseg, lib.loss)."""
    import importlib
    import types
    from oracle import build_ref
    pydir = build_ref.python_dir()
    if not os.path.exists(os.path.join(pydir, "pointnet2_semseg.py")):
        raise ImportError("oracle/_ref/py not staged (run oracle/build_ref.py where /root/reference exists)")
    pkg = sys.modules.get("pointnet2")
    if pkg is None:   # a bare stand-in for the reference's `pointnet2` package: only `_ext` is looked up in it
        pkg = types.ModuleType("pointnet2")
        pkg.__path__ = []
        sys.modules["pointnet2"] = pkg
    sys.modules["pointnet2._ext"] = ext_module
    pkg._ext = ext_module
    for d in (pydir, os.path.join(pydir, "solver")):
        if d not in sys.path:
            sys.path.insert(0, d)
    semseg = importlib.import_module("pointnet2_semseg")      # as scripts/train.py:37-38 imports it
    loss = importlib.import_module("lib.loss")
    return semseg, loss


def run_reference(args, world, rank, local):
    """Reference arm: rank 0 only; the reference is single-GPU."""
    if rank != 0:
        return None
    from oracle import build_ref
    use_cuda_ref = args.impl == "reference" and torch.cuda.is_available() and build_ref.available("stock")
    if not use_cuda_ref:
        batches, in_ch = host_batches(args, 0, 1)
        base = cpu_baseline(args, in_ch)
        return {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(1e3 * args.cpu_sample_chunks / base["value"], 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": "SSG training step, CPU port on a bounded sample"},
                "cpu_baseline": base,
                "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}

    batches, in_ch = host_batches(args, 0, 4, standalone=True)
    ref = build_ref.load_ref("stock")
    semseg, ref_loss = load_reference_python(ref)
    torch.manual_seed(0)
    # scripts/train.py:38-44: get_model(...).cuda(), WeightedCrossEntropyLoss(), Adam(lr=1e-3, weight_decay=0)
    model = semseg.get_model(num_classes=NUM_CLASSES, is_msg=args.msg, input_channels=in_ch,
                             use_xyz=True, bn=True).cuda()
    criterion = ref_loss.WeightedCrossEntropyLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0)
    model.train()

    def iteration(pc, seg, w):
        """lib/solver.py:219-231 without the host-side metrics: _forward (:161-166; the batch arrives as
        coords|feats already concatenated), _compute_loss (:174-177), _backward (:168-172)."""
        preds = model(pc)
        nc = preds.size(-1)
        loss = criterion(preds.contiguous().view(-1, nc), seg.view(-1), w.view(-1))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        return loss

    dev = [tuple(t.cuda() for t in b) for b in batches]
    settle = max(args.warmup, args.settle)
    for i in range(settle):
        iteration(*dev[i % len(dev)])
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ms, loss = time_steps(iteration, batches, args.steps, 1, e2e=False)
    clocks = sampler.stop()
    ms_e2e, _ = time_steps(iteration, batches, args.steps, 1, e2e=True)
    chunks = args.batch * args.steps
    v, ve = round(chunks / (ms * 1e-3), 1), round(chunks / (ms_e2e * 1e-3), 1)
    blocking = os.environ.get("CUDA_LAUNCH_BLOCKING") == "1"
    sample = "reference's own python + its CUDA ext (unmodified, -O3, sm_100a) on this B200, full workload"
    return {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[%d] %s train step %dx%d/GPU C=%d" % (
                       2 if args.msg else 1, "MSG" if args.msg else "SSG", args.batch, args.npoints, in_ch),
                   "per_gpu_batch": args.batch, "npoints": args.npoints, "parallelism": "single GPU",
                   "launch": "eager" + (", CUDA_LAUNCH_BLOCKING=1 (scripts/train.py:133)" if blocking else ""),
                   "settle": settle},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 0, "kind": "reference", "sample": sample},
        "e2e": {"value": ve, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in batches[0])),
                "d2h_bytes_per_step": 4},
        "clocks": {k: clocks[k] for k in ("sm_mhz", "sm_max_mhz", "reasons")},
        "loss": round(loss, 5),
    }


def format_line(res):
    """The ONE stdout line.  Kept under LINE_LIMIT characters: optional keys are dropped, least
    important first, until it fits (the driver parses the tail of stdout only)."""
    res = dict(res)
    for victim in (None, "loss", ("roofline", "share"), ("roofline", "peak_source"), ("config", "mlp"),
                   "fps_ns_per_iter", ("config", "l2"), ("cpu_baseline", "sample")):
        if isinstance(victim, tuple):
            if isinstance(res.get(victim[0]), dict):
                res[victim[0]] = {k: v for k, v in res[victim[0]].items() if k != victim[1]}
        elif victim is not None:
            res.pop(victim, None)
        line = json.dumps(res, separators=(",", ":"))
        if len(line) < LINE_LIMIT:
            return line
    raise RuntimeError("bench line does not fit %d characters: %d" % (LINE_LIMIT, len(line)))


def main():
    args = parse()
    # stdout carries exactly ONE line, the JSON: libraries that chat on fd 1 (NCCL prints its version there
    # when NCCL_DEBUG is set) are sent to stderr, the result goes to a private copy of the original stdout
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl != "product":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        res = run_reference(args, 1, 0, int(os.environ.get("LOCAL_RANK", "0")))
    else:
        world, rank, local = dist_setup(args)
        res = run_product(args, world, rank, local)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
    if res is not None:
        out.write(format_line(res) + "\n")
        out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
