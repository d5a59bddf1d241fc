import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L
for (F, S, sub) in [(8192, 3000, None), (8192, 3000, 1024), (4096, 3000, 1024), (8448, 1500, None), (6144, 2000, None)]:
    rng = np.random.default_rng(1)
    Z = np.asfortranarray(rng.standard_normal((F, S)).astype(np.float32))
    Y = np.asfortranarray(rng.standard_normal((3, S)).astype(np.float32))
    if sub: os.environ["RCB_GRAM_SUB_SAMPLES"] = str(sub)
    else: os.environ.pop("RCB_GRAM_SUB_SAMPLES", None)
    g = R.gram(Z, Y, gram_mode=L.GRAM_TENSOR)
    C = Z.astype(np.float64) @ Y.astype(np.float64).T
    err = np.abs(g[:, F:] - C)
    bad = np.argwhere(err > 1e-6)
    print(F, S, sub, "cross max err", err.max(), "bad rows", (bad[:, 0].min(), bad[:, 0].max(), len(bad)) if len(bad) else None, flush=True)
