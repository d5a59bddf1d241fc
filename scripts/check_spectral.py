"""Arbitration run for the GPU spectral radius: repeated GPU calls, ARPACK with several start vectors, dense eigvals (N=4096)."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))
import rcb200 as R
n = 4096
rng = np.random.default_rng(17)
rows = np.stack([rng.choice(n, size=6, replace=False) for _ in range(n)], axis=1)
W = sp.csc_matrix((rng.standard_normal(6 * n).astype(np.float32), rows.T.ravel(), np.arange(0, 6 * n + 1, 6)), shape=(n, n))
for i in range(3):
    print("gpu", R.spectral_radius(W, return_dim=True), flush=True)
for tol in (1e-6, 1e-12):
    print("gpu tol", tol, R.spectral_radius(W, tol=tol, return_dim=True), flush=True)
for seed in range(4):
    v0 = np.random.default_rng(seed).standard_normal(n)
    for k in (6, 20):
        vals = spla.eigs(W.astype(np.float64), k=k, which="LM", v0=v0, return_eigenvectors=False, tol=1e-10, maxiter=20000)
        print("arpack seed", seed, "k", k, np.sort(np.abs(vals))[::-1][:4], flush=True)
t = time.perf_counter()
ev = np.linalg.eigvals(W.toarray().astype(np.float64))
print("dense", np.sort(np.abs(ev))[::-1][:6], time.perf_counter() - t, flush=True)
