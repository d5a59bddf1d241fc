"""BASELINE config 4 for the record: DeepESN 4 x 2048 with per-layer modifiers, Lorenz T=20000, ridge readout."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
T, w = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, 200
rng = np.random.default_rng(17)
esn = R.DeepESN(3, [2048] * 4, 3, depth=4, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / 2048),
                leak_coefficient=(1.0, 0.8, 0.6, 0.4), state_modifiers=(R.NLAT2, R.NLAT2, R.NLAT2, R.Pad(1.0)))
ps, st = R.setup(rng, esn)
series = O.lorenz_series(T + 1)
data = np.asfortranarray(series[:, :T].astype(np.float32)); target = np.asfortranarray(series[:, 1:T + 1].astype(np.float32))
dev = R.default_device()
for rep in range(2):
    t0 = time.time()
    ps2, st2 = R.train(esn, data, target, ps, st, objective=R.RidgeRegression(1e-6 * T), washout=w)
    dt = time.time() - t0
print(f"DeepESN 4x2048, T={T}: train {dt:.2f} s wall; K1 (4 layers + 3 projections) {dev.kernel_ms(0):.1f} ms, K2 {dev.kernel_ms(1):.1f} ms, "
      f"K3 {dev.kernel_ms(2):.1f} ms; W_out {ps2.readout.weight.shape}  last K1: {dev.last_kernel()}")
Y, _ = R.predict(esn, data[:, :2000], ps2, st)
print("one-step NRMSE on the first 2000 steps:", float(np.sqrt(np.mean((Y[:, w:] - target[:, w:2000]) ** 2)) / np.std(target[:, w:2000])))
