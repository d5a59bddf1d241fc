"""BASELINE config 3 through ONE public call: rcb_train on 1024 sequences x 10000 steps, N=8192, NLAT2, pooled ridge readout
(the library cuts the 335 GB state stream into time windows itself)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))
import bench
import rcb200 as R
N, B, T, w = 8192, 1024, 10000, 100
rng = np.random.default_rng(17)
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=True), leak_coefficient=0.5, state_modifiers=R.NLAT2)
ps, st = R.setup(rng, esn)
lor = bench.lorenz_batch(B, T + 1)
data, tgt = np.asfortranarray(lor[:, :T]), np.asfortranarray(lor[:, 1:])
dev = R.default_device()
for rep in range(2):
    t0 = time.time()
    ps2, st2 = R.train(esn, data, tgt, ps, st, objective=R.RidgeRegression(1e-6 * (T - w) * B), washout=w)
    dt = time.time() - t0
print(f"C3 train through R.train / rcb_train: {dt:.2f} s wall (K1 {dev.kernel_ms(0):.0f} ms, K2 {dev.kernel_ms(1):.0f} ms, K3 {dev.kernel_ms(2):.0f} ms); "
      f"W_out {ps2.readout.weight.shape}, finite {np.all(np.isfinite(ps2.readout.weight))}; last K1: {dev.last_kernel()}")
