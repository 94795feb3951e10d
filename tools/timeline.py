"""In-situ kernel timeline of the benchmark step (CUDA graph + geometry prefetch), via CUPTI
(torch.profiler): per-kernel-name totals over the profiled steps, so that the share of every
kernel INSIDE the replayed step is known (CUDA-event timings of eager calls overstate the small
kernels; ncu serialises).  Output: gpurun_out/timeline_<tag>.json

    python tools/timeline.py [tag] [--msg]
"""
import json, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import pointnet2_semseg, synthetic, train

tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "step"
msg = "--msg" in sys.argv
STEPS = 5
batches = [tuple(torch.from_numpy(a).cuda() for a in synthetic.make_batch(2, 32, first_chunk=32 * i)) for i in range(2)]
torch.manual_seed(0)
model = pointnet2_semseg.get_model(20, is_msg=msg, input_channels=6).cuda().train()
step = train.PipelinedTrainStep(train.TrainStep(model, lr=1e-3), batches[0], scope="step")
for i in range(5):
    step(*batches[i % 2])
torch.cuda.synchronize()
step.prefetch(*batches[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(STEPS):
    step()
    step.prefetch(*batches[(i + 1) % 2])
e1.record()
torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1) / STEPS
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(STEPS):
        step()
        step.prefetch(*batches[(i + 1) % 2])
    torch.cuda.synchronize()
agg = {}
t_min, t_max = None, None
for ev in prof.events():
    if ev.device_type != torch.autograd.DeviceType.CUDA:
        continue
    d = agg.setdefault(ev.name, [0, 0.0])
    d[0] += 1
    d[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    tr = ev.time_range
    t_min = tr.start if t_min is None else min(t_min, tr.start)
    t_max = tr.end if t_max is None else max(t_max, tr.end)
rows = sorted(([n, c / STEPS, t / STEPS] for n, (c, t) in agg.items()), key=lambda r: -r[2])
out = {"ms_per_step_unprofiled": ms_plain, "span_ms_per_step_profiled": (t_max - t_min) / 1e3 / STEPS,
       "sum_kernel_ms_per_step": sum(r[2] for r in rows) / 1e3,
       "kernels": [{"name": n[:110], "calls_per_step": c, "us_per_step": round(t, 2)} for n, c, t in rows]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "timeline_%s.json" % tag), "w") as f:
    json.dump(out, f, indent=1)
print("ms/step %.3f (profiled span %.3f), sum of kernel time %.3f ms" % (
    ms_plain, out["span_ms_per_step_profiled"], out["sum_kernel_ms_per_step"]))
for n, c, t in rows[:45]:
    print("%8.1f us %6.1f x  %s" % (t, c, n[:100]))
