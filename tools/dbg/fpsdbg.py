import ctypes, sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
from pointnet2.scannet_b200 import synthetic
L = ctypes.CDLL("/root/repo/tools/dbg/libfpsdbg.so")
x, _, _ = synthetic.make_batch(2, 32)
xyz = torch.from_numpy(np.ascontiguousarray(x[:, :, :3])).cuda()
out = torch.empty(32, 1024, dtype=torch.int32, device="cuda")
vp = ctypes.c_void_p
L.pn2_furthest_point_sampling.argtypes = [ctypes.c_int]*3 + [vp]*4
st = L.pn2_furthest_point_sampling(32, 8192, 1024, xyz.data_ptr(), None, out.data_ptr(), None)
buf = (ctypes.c_ulonglong * 1024)()
L.pn2_debug_fps_active(buf)
a = np.array(buf[:], dtype=np.float64) / 32.0   # per cloud, of 32 warps
print("status", st, "active warps per iteration (of 32): j=1..8", a[1:9], " j=50", a[50], " j=100", a[100], "j=200", a[200], " j=500", a[500], " j=1000", a[1000])
print("mean over all iterations", a[1:1024].mean())
