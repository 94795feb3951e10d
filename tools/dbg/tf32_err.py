import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_fused_gpu as T
from pointnet2.scannet_b200 import fused, pointnet2_modules as M
torch.manual_seed(0)
sa = M.PointnetSAModule(npoint=256, radius=0.2, nsample=32, mlp=[6, 32, 32, 64]).to("cuda:0")
ins = T.sa_inputs(2, 2048, 6)
with T.precision(False):
    out_t, g_t, _, ig_t = T.run(sa, False, *ins)
for bwd in (True,):
    fused.TC_BWD = bwd
    with T.precision(True):
        out_f, g_f, _, ig_f = T.run(sa, True, *ins)
        out_r, g_r, _, ig_r = T.run(sa, False, *ins)
    print("TC_BWD", bwd)
    print("out", T.rel_err(out_f, out_t), T.rel_err(out_r, out_t))
    for n in g_t:
        print(n, T.rel_err(g_f[n], g_t[n]), T.rel_err(g_r[n], g_t[n]), "L2", T.l2_err(g_f[n], g_t[n]), T.l2_err(g_r[n], g_t[n]))
    print("in_grad", T.rel_err(ig_f[1], ig_t[1]), T.rel_err(ig_r[1], ig_t[1]), "L2", T.l2_err(ig_f[1], ig_t[1]), T.l2_err(ig_r[1], ig_t[1]))
