"""Debug helper: tensor Gram vs exact on a few shapes, printing where they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L

for (F, S, D) in [(128, 64, 1), (256, 1024, 2), (300, 5000, 3)]:
    rng = np.random.default_rng(1)
    Z = np.asfortranarray(rng.standard_normal((F, S)).astype(np.float32))
    Y = np.asfortranarray(rng.standard_normal((D, S)).astype(np.float32))
    dev = R.default_device()
    l0 = dev.launch_count()
    gt = R.gram(Z, Y, gram_mode=L.GRAM_TENSOR)
    print(F, S, D, "launches", dev.launch_count() - l0, "gram ms", dev.kernel_ms(L.TIMER_GRAM))
    Z64 = Z.astype(np.float64)
    G = Z64 @ Z64.T
    Gt = gt[:, :F]
    print("  nonzero frac lower:", np.count_nonzero(np.tril(Gt)) / (F * (F + 1) / 2), " upper nnz:", np.count_nonzero(np.triu(Gt, 1)))
    err = np.abs(np.tril(Gt) - np.tril(G)) / np.sqrt(np.outer(np.diag(G), np.diag(G)))
    print("  max rel err", err.max(), " C err", np.abs(gt[:, F:] - Z64 @ Y.astype(np.float64).T).max())
    print("  G[0:3,0:3] got\n", Gt[:3, :3], "\n  want\n", G[:3, :3])
    if err.max() > 1e-5:
        bad = np.argwhere(err > 1e-5)
        print("  bad count", len(bad), "first", bad[:5].tolist(), "rows range", bad[:, 0].min(), bad[:, 0].max(), "cols", bad[:, 1].min(), bad[:, 1].max())
        r = np.tril(Gt)[np.tril(G) != 0] / np.tril(G)[np.tril(G) != 0]
        print("  ratio got/want percentiles", np.percentile(r, [0, 10, 50, 90, 100]))
