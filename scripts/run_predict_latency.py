import os, sys, numpy as np
sys.path.insert(0, "/root/repo/reservoircomputing.jl_b200"); sys.path.insert(0, "/root/repo")
import rcb200 as R
from oracle import esn_oracle as O
dev = R.default_device()
series = O.lorenz_series(5001).astype(np.float32) / 20
for N in (300, 1024, 4096):
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), state_modifiers=R.NLAT2)
    ps, st = R.setup(np.random.default_rng(17), esn)
    data = np.asfortranarray(series[:, :5000]); tgt = np.asfortranarray(series[:, 1:5001])
    for name, obj in (("f64", R.RidgeRegression(1e-4)), ("f32", R.RidgeRegression(np.float32, 1e-4))):
        ps2, st2 = R.train(esn, data, tgt, ps, st, objective=obj, washout=100)
        for _ in range(2):
            Y, _ = R.predict(esn, 2000, ps2, st2, initialdata=data[:, 0])
        ka = dev.kernel_ms(0) / 2000 * 1e3
        for _ in range(2):
            Y, _ = R.predict(esn, np.asfortranarray(data[:, :2000]), ps2, st2)
        kt = dev.kernel_ms(0) / 2000 * 1e3
        print(f"N={N} W_out {name}: autoregressive {ka:.2f} us/step ({Y.dtype}), teacher-forced {kt:.2f} us/step", flush=True)
