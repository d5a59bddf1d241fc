"""Host-setup timing: spectral radius for scale_radius! — GPU Arnoldi vs scipy ARPACK vs dense eigvals."""
import os, sys, time
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))
import rcb200 as R

for n in [int(a) for a in sys.argv[1:]] or (1024, 2048, 4096, 8192, 16384):
    rng = np.random.default_rng(17)
    rows = np.stack([rng.choice(n, size=6, replace=False) for _ in range(n)], axis=1)
    W = sp.csc_matrix((rng.standard_normal(6 * n).astype(np.float32), rows.T.ravel(), np.arange(0, 6 * n + 1, 6)), shape=(n, n))
    R.spectral_radius(W)  # warm-up (context, allocations)
    t = time.perf_counter(); rho, dim = R.spectral_radius(W, return_dim=True); tg = time.perf_counter() - t
    t = time.perf_counter()
    ref = np.max(np.abs(spla.eigs(W.astype(np.float64), k=6, which="LM", return_eigenvectors=False, tol=1e-10, maxiter=20000)))
    ta = time.perf_counter() - t
    td, dd = float("nan"), float("nan")
    if n <= 2048:
        t = time.perf_counter(); refd = np.max(np.abs(np.linalg.eigvals(W.toarray().astype(np.float32)))); td = time.perf_counter() - t
        dd = abs(rho - refd) / refd
    print(f"N={n}: gpu arnoldi {tg*1e3:.0f} ms (krylov dim {dim}), rel diff vs ARPACK {abs(rho-ref)/ref:.1e}, vs dense eigvals (Float32 LAPACK) {dd:.1e}; "
          f"scipy ARPACK {ta*1e3:.0f} ms; dense eigvals {td:.2f} s", flush=True)
