"""Calibration for write-dominated kernels: CUDA-event time of a pure fill and of a contiguous copy
at the output sizes of the rows gather kernels (L2 flushed before every launch)."""
import torch
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.float32, device=dev)


def timed(fn, reps=8):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for mb in (19, 50, 71, 143, 268, 537):
    n = mb * 1000 * 1000 // 4
    a = torch.empty(n, device=dev)
    b = torch.empty(n, device=dev)
    t_fill = timed(lambda: a.zero_())
    t_copy = timed(lambda: b.copy_(a))
    print("%4d MB  fill %7.1f us = %6.0f GB/s written   copy %7.1f us = %6.0f GB/s read+written" % (
        mb, t_fill, mb * 1e6 / t_fill / 1e3, t_copy, 2 * mb * 1e6 / t_copy / 1e3))
