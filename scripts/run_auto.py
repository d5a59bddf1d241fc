import os, sys, numpy as np
sys.path.insert(0, "/root/repo/reservoircomputing.jl_b200"); sys.path.insert(0, "/root/repo")
import rcb200 as R
from oracle import esn_oracle as O
N = int(sys.argv[1])
series = O.lorenz_series(3001).astype(np.float32) / 20
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), state_modifiers=R.NLAT2)
ps, st = R.setup(np.random.default_rng(17), esn)
data = np.asfortranarray(series[:, :3000]); tgt = np.asfortranarray(series[:, 1:3001])
ps2, st2 = R.train(esn, data, tgt, ps, st, objective=R.RidgeRegression(1e-4), washout=100)
for _ in range(2):
    Y, _ = R.predict(esn, 2000, ps2, st2, initialdata=data[:, 0])
print(R.default_device().kernel_ms(0) / 2000 * 1e3, R.default_device().last_kernel())
