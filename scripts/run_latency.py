"""Per-step latency of the single-sequence (latency-bound) shapes: collectstates and autoregressive predict."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
T = 20000
dev = R.default_device()
series = O.lorenz_series(T + 1).astype(np.float32) / 20
for N in (300, 1024, 2048, 4096, 8192):
    for mods in ((), R.NLAT2):
        rng = np.random.default_rng(17)
        esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), state_modifiers=mods)
        ps, st = R.setup(rng, esn)
        data = np.asfortranarray(series[:, :T])
        for _ in range(2):
            s, st2 = R.collectstates(esn, data, ps, st)
        k1 = dev.kernel_ms(0)
        ps2, st3 = R.train(esn, data, np.asfortranarray(series[:, 1:T + 1]), ps, st, objective=R.RidgeRegression(1e-4), washout=100)
        for _ in range(2):
            Y, _ = R.predict(esn, 2000, ps2, st3, initialdata=data[:, 0])
        k4 = dev.kernel_ms(0)
        print(f"N={N} mods={'nlat2' if mods else 'none'}: collect {k1 / T * 1e3:.2f} us/step ({dev.last_kernel() if False else ''}) predict(auto) {k4 / 2000 * 1e3:.2f} us/step", flush=True)
