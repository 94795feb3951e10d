"""BASELINE configs[3]: whole-scene evaluation on a synthetic 150k-point ScanNet-shaped scene.

(i)  reference semantics (lib/dataset.py:316-345, scripts/eval.py:27-38): tile the scene into
     1.5 m columns, resample each to 8192 points, run the SSG network in eval() mode in batches
     of 32 chunks -> scenes/s;
(ii) extension: furthest point sampling 150000 -> 8192 / 1024 and a ball query (r = 0.1, 32) over
     the WHOLE scene as one cloud (streaming FPS kernel, L2-scanned ball query).
"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import ext, pointnet2_semseg, synthetic


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


scene = synthetic.make_scene(0, npoints=150000)
tiles = torch.from_numpy(synthetic.tile_scene(scene)).cuda()          # (K, 8192, 9)
torch.manual_seed(0)
model = pointnet2_semseg.get_model(20, is_msg=False, input_channels=6).cuda().eval()


@torch.no_grad()
def eval_scene():
    return torch.cat([model(t) for t in torch.split(tiles, 32)], dim=0).argmax(-1)


ms_eval, pred = timed(eval_scene)
# the same evaluation without the one-kernel SA blocks (layer by layer), and both replayed from a CUDA graph
from pointnet2.scannet_b200 import fused
fused.CHAIN_EVAL = False
ms_eval_layers, pred_layers = timed(eval_scene)
fused.CHAIN_EVAL = True
assert torch.equal(pred, pred_layers)


def graphed(chain):
    fused.CHAIN_EVAL = chain
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eval_scene()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = eval_scene()
    ms, _ = timed(lambda: g.replay(), reps=5)
    fused.CHAIN_EVAL = True
    return ms, out


ms_graph_chain, pred_g = graphed(True)
ms_graph_layers, _ = graphed(False)
assert torch.equal(pred_g, pred)
xyz = torch.from_numpy(np.ascontiguousarray(scene[None, :, :3])).cuda()
ms_fps8k, idx8k = timed(lambda: ext.furthest_point_sampling(xyz, 8192), reps=1)
ms_fps1k, idx1k = timed(lambda: ext.furthest_point_sampling(xyz, 1024), reps=2)
ctr = xyz.gather(1, idx1k.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
ms_bq, bq = timed(lambda: ext.ball_query(ctr, xyz, 0.1, 32))
print(json.dumps({
    "scene_points": int(scene.shape[0]), "tiles": int(tiles.shape[0]),
    "ssg_eval_ms_per_scene": round(ms_eval, 3), "scenes_per_s": round(1e3 / ms_eval, 2),
    "chunks_per_s_eval": round(tiles.shape[0] / ms_eval * 1e3, 1),
    "ssg_eval_ms_per_scene_layer_by_layer": round(ms_eval_layers, 3),
    "ssg_eval_ms_per_scene_graph_replay": round(ms_graph_chain, 3),
    "ssg_eval_ms_per_scene_graph_replay_layer_by_layer": round(ms_graph_layers, 3),
    "scenes_per_s_graph_replay": round(1e3 / ms_graph_chain, 2),
    "fps_150k_to_8192_ms": round(ms_fps8k, 2), "fps_150k_to_8192_us_per_iteration": round(1e3 * ms_fps8k / 8191, 2),
    "fps_150k_to_1024_ms": round(ms_fps1k, 2),
    "ball_query_150k_x_1024_ms": round(ms_bq, 3),
}))
