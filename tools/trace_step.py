"""Ordered kernel trace of ONE replayed training step (CUPTI via torch.profiler): for every kernel its
stream, start, duration and the idle gap in front of it on the same stream -- what the per-name totals of
tools/timeline.py cannot show: the critical path of the step and where the main stream waits.

    python tools/trace_step.py [tag] [--msg]      -> gpurun_out/trace_<tag>.txt
    torchrun --nproc-per-node N tools/trace_step.py [tag]     data parallel: rank 0 writes its trace (where the gradient
                                                              all-reduce and Adam sit relative to the last backward kernel)
"""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import pointnet2_semseg, synthetic, train

tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "step"
msg = "--msg" in sys.argv
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
batches = [tuple(torch.from_numpy(a).cuda() for a in synthetic.make_batch(2, 32, first_chunk=32 * (i + 2 * rank)))
           for i in range(2)]
torch.manual_seed(0)
model = pointnet2_semseg.get_model(20, is_msg=msg, input_channels=6).cuda().train()
# (as bench.py: one GPU captures the whole step; data parallel captures forward + backward, NCCL + Adam run after it)
step = train.PipelinedTrainStep(train.TrainStep(model, lr=1e-3), batches[0], scope="step" if world == 1 else "fwd_bwd")
for i in range(8):
    step(*batches[i % 2])
torch.cuda.synchronize()
step.prefetch(*batches[0])
STEPS = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(STEPS):
        step()
        step.prefetch(*batches[(i + 1) % 2])
    torch.cuda.synchronize()
if rank != 0:
    dist.barrier()
    sys.exit(0)
import json, tempfile
tmp = os.path.join(tempfile.gettempdir(), "pn2_trace.json")
prof.export_chrome_trace(tmp)
evs = []
for e in json.load(open(tmp))["traceEvents"]:
    if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e:
        evs.append((float(e["ts"]), float(e["ts"]) + float(e.get("dur", 0)), e["name"], e.get("args", {}).get("stream", 0)))
evs.sort(key=lambda e: e[0])
adam = [i for i, e in enumerate(evs) if "adam_flat_kernel" in e[2]]
assert len(adam) >= 3, "expected one Adam kernel per step"
lo, hi = adam[1] + 1, adam[2] + 1          # one whole step: after Adam of step 1 up to Adam of step 2
t0 = evs[lo][0]
last_end, busy, lines = {}, {}, []
for s, e, name, sid in evs[lo:hi]:
    gap = s - last_end.get(sid, s)
    last_end[sid] = max(e, last_end.get(sid, e))
    busy[sid] = busy.get(sid, 0.0) + (e - s)
    short = name.replace("pn2::", "").replace("void ", "").split("(")[0][:70]
    lines.append("%4s %9.1f %8.1f %7.1f  %s" % (sid, s - t0, e - s, gap, short))
span = evs[hi - 1][1] - t0
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "trace_%s.txt" % tag)
with open(path, "w") as f:
    f.write("# one step: %d kernels, span %.1f us; busy per stream (us): %s\n" % (
        hi - lo, span, {k: round(v, 1) for k, v in busy.items()}))
    f.write("# strm  start_us   dur_us  gap_us  kernel\n")
    f.write("\n".join(lines) + "\n")
print(open(path).read()[:1500])
if world > 1:
    dist.barrier()
