import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_util import debug_gemm, bf16_round
rng = np.random.default_rng(0)
for (M, N, K) in [(256, 256, 64), (64, 64, 64), (700, 384, 1536)]:
    A = rng.standard_normal((M, K), dtype=np.float32); W = rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)
    b = rng.standard_normal(N, dtype=np.float32)
    ref = bf16_round(A).astype(np.float64) @ bf16_round(W).astype(np.float64).T + b
    x0 = rng.standard_normal((M, N), dtype=np.float32)
    out = debug_gemm(3, A, W, b, 2, out0=x0)
    err = np.abs(out - (x0 + ref))
    print(M, N, K, "max err", err.max(), "n bad", (err > 1e-3).sum(), "of", err.size)
    bad = np.argwhere(err > 1e-3)
    if len(bad):
        print(" first bad", bad[:8].tolist(), "rows bad", np.unique(bad[:, 0])[:20], "cols bad", np.unique(bad[:, 1])[:40])
        r, c = bad[0]
        print(" out", out[r, c], "x0", x0[r, c], "ref", ref[r, c], "out-x0", out[r, c] - x0[r, c])
        # is out - x0 equal to some other position's ref?
        d = (out - x0)[r]
        j = np.argmin(np.abs(ref[r] - d[c])); print(" value matches ref col", j, ref[r, j])
