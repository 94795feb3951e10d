import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_bwd_gpu as T
np.set_printoptions(precision=3, suppress=True, linewidth=200)
for (P, N, K, lda) in [(128, 32, 32, 32), (256, 32, 32, 32), (4096, 32, 32, 32), (512, 128, 64, 64)]:
    rng = np.random.default_rng(1)
    c, ref = T.make_case(rng, P, N, K, lda, 0, True, True, True)
    out = T.run_layer(c)
    for k in ("dh_prev", "dw", "s", "q"):
        g, w = out[k], ref[k]
        if k == "dh_prev": g = g[:, :K]
        print(P, N, K, k, "err", np.abs(g - w).max(), "max", np.abs(w).max(), "nan", np.isnan(g).sum())
    g, w = out["dw"], ref["dw"]
    print("got\n", g[:6, :8]); print("want\n", w[:6, :8])
    print("got^T vs want err", np.abs(g.T - w).max() if N == K else None)
    # which 8-row groups contribute? compare against partial sums
    sys.stdout.flush()
