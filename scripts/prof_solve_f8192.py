import os, sys, numpy as np
ROOT = "/root/repo" if os.path.exists("/root/repo/bench.py") else os.getcwd()
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))
import rcb200 as R
F, D = 8192, 3
rng = np.random.default_rng(0)
A = rng.standard_normal((F, 2 * F)); G = A @ A.T
GC = np.asfortranarray(np.concatenate([G, rng.standard_normal((F, D))], axis=1))
for _ in range(2):
    W = R.ridge_solve(GC, D, 1e-3)
print("ok", np.isfinite(W).all())
