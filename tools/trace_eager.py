"""Durations of the backward layer kernels inside an EAGER step (natural launch gaps between kernels) vs
inside the replayed graph (back to back): same kernels, same data.  Experiment for DESIGN section 9."""
import json, os, sys, tempfile
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import pointnet2_semseg, synthetic, train

batch = tuple(torch.from_numpy(a).cuda() for a in synthetic.make_batch(2, 32))
torch.manual_seed(0)
model = pointnet2_semseg.get_model(20, is_msg=False, input_channels=6).cuda().train()
step = train.TrainStep(model, lr=1e-3)
mode = sys.argv[1] if len(sys.argv) > 1 else "eager"
if mode == "graph":
    step = train.GraphedTrainStep(step, batch, scope="step")
for _ in range(5):
    step(*batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(*batch)
    torch.cuda.synchronize()
tmp = os.path.join(tempfile.gettempdir(), "pn2_trace_e.json")
prof.export_chrome_trace(tmp)
rows = [(float(e["ts"]), float(e.get("dur", 0)), e["name"]) for e in json.load(open(tmp))["traceEvents"]
        if e.get("cat") == "kernel"]
rows.sort()
names = ("mlp_bwd_layer_kernel", "gemm_tf32_fwd_kernel")
adam = [i for i, r in enumerate(rows) if "adam_flat_kernel" in r[2]]
lo, hi = adam[0] + 1, adam[1] + 1
print("mode", mode, "step span %.1f us" % (rows[hi - 1][0] + rows[hi - 1][1] - rows[lo][0]))
for ts, dur, name in rows[lo:hi]:
    if any(n in name for n in names) and dur > 40:
        print("%9.1f %8.1f  %s" % (ts - rows[lo][0], dur, name.replace("pn2::", "").split("(")[0][:60]))
