"""Inference-time set-abstraction block at the SSG shapes (B = 32): the one-kernel chain (csrc/sa_chain.cu:
gather + three 1x1 convs with folded BatchNorm + ReLU + max over nsample) against the same block run layer by
layer (row gather, three tensor-core layers each writing its [P, C] activation, fold, pooling).  CUDA events,
L2 flushed before every repetition; geometry (centres, ball query) precomputed.  Bytes of the chain =
index table + gathered rows (or the point table if smaller) + pooled output."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet2.scannet_b200 import ext, fused, pointnet2_modules as M, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    PEAK = 6536.4


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


x, _, _ = synthetic.make_batch(2, B)
xyz0 = torch.from_numpy(np.ascontiguousarray(x[:, :, :3])).to(dev)
feat0 = torch.from_numpy(np.ascontiguousarray(x[:, :, 3:].transpose(0, 2, 1))).to(dev)
levels = [("SA1", 1024, 0.1, 32, [6, 32, 32, 64]), ("SA2", 256, 0.2, 32, [64, 64, 64, 128])]
xyz, feat = xyz0, feat0
torch.manual_seed(0)
with torch.no_grad():
    for name, npoint, r, ns, mlp in levels:
        sa = M.PointnetSAModule(npoint=npoint, radius=r, nsample=ns, mlp=list(mlp)).to(dev).eval()
        idx = ext.furthest_point_sampling(xyz, npoint)
        new_xyz = xyz.gather(1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
        bq = ext.ball_query(new_xyz, xyz, r, ns)
        res = {}
        for chain in (True, False):
            fused.CHAIN_EVAL = chain
            out = sa._forward_rows(xyz, new_xyz, feat, [bq])
            res[chain] = (timed(lambda: sa._forward_rows(xyz, new_xyz, feat, [bq])), out)
        fused.CHAIN_EVAL = True
        assert torch.equal(res[True][1], res[False][1])
        # the kernel alone (point table already concatenated)
        cols = [xyz, fused.rows_of(feat)]
        width = sum(c.size(2) for c in cols)
        if fused.pad4(width) != width:
            cols.append(xyz.new_zeros(B, xyz.shape[1], fused.pad4(width) - width))
        table = torch.cat(cols, dim=2)
        t_kernel = timed(lambda: fused.sa_chain_eval(sa.mlps[0], table, new_xyz, bq))
        N, ct, P = xyz.shape[1], fused.pad4(3 + feat.shape[1]), B * npoint * ns
        nbytes = 4 * P + 4 * ct * min(B * N, P) + 4 * B * npoint * mlp[3]
        layer_bytes = nbytes + 4 * P * (2 * ct + 2 * mlp[1] + 2 * mlp[2] + mlp[3])
        flops = 2.0 * P * (fused.pad4(3 + feat.shape[1]) * mlp[1] + mlp[1] * mlp[2] + mlp[2] * mlp[3])
        print("%s B=%d: block with the chain kernel %7.1f us (kernel alone %6.1f us = %5.1f TFLOP/s TF32, %5.1f MB algorithmic) | "
              "layer by layer %7.1f us (%6.1f MB) | %.1fx, bit-identical" % (
                  name, B, res[True][0], t_kernel, flops / t_kernel / 1e6, nbytes / 1e6, res[False][0], layer_bytes / 1e6,
                  res[False][0] / res[True][0]))
        xyz, feat = new_xyz, res[True][1]
