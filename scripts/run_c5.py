"""BASELINE config 5 for the record: ensemble sweep, members = (radius x leak) recurrences of one N=1024 reservoir,
16 lambdas each, Lorenz T=5000 train, washout 100.  argv: number of (radius, leak) members (default 256)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N, T, w = 1024, 5000, 100
rng = np.random.default_rng(17)
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=1.0, sparsity=6 / N), leak_coefficient=1.0)
ps, st = R.setup(rng, esn)
series = O.lorenz_series(T + 1)
data = np.asfortranarray(series[:, :T].astype(np.float32)); target = np.asfortranarray(series[:, 1:T + 1].astype(np.float32))
g = int(round(M ** 0.5))
radii, leaks = np.meshgrid(np.linspace(0.5, 1.4, g), np.linspace(0.1, 1.0, M // g))
lams = np.logspace(-8, -1, 16)
st0 = R.resetcarry(None, esn, st, init_carry=lambda r, n: np.zeros((n, radii.size), np.float32))
dev = R.default_device()
for rep in range(2):
    t0 = time.time()
    W, hT = R.train_ensemble(esn, data, target, ps, st0, w_scale=radii.ravel().astype(np.float32), leak=leaks.ravel().astype(np.float32),
                             lambdas=lams, washout=w)
    dt = time.time() - t0
print(f"{radii.size} recurrences x {len(lams)} lambdas = {W.shape[0] * W.shape[1]} readouts (N={N}, T={T}): {dt:.2f} s wall; "
      f"K1 {dev.kernel_ms(0):.1f} ms ({N * T * radii.size / dev.kernel_ms(0) * 1e3:.3e} neuron-steps/s), K2 {dev.kernel_ms(1):.1f} ms, "
      f"K3 {dev.kernel_ms(2):.1f} ms; launches {dev.launch_count()}  last K1: {dev.last_kernel()}")
print("finite:", bool(np.isfinite(W).all()))
