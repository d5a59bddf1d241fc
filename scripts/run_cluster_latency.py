import os, sys, time
import numpy as np
sys.path[:0]=['/root/repo','/root/repo/reservoircomputing.jl_b200']
import rcb200 as R
from oracle import esn_oracle as O
T = 20000
dev = R.default_device()
series = O.lorenz_series(T + 1).astype(np.float32) / 20
data = np.asfortranarray(series[:, :T])
for N in (1024, 2048, 4096):
    for cs in ("single", "2", "4", "8"):
        if cs == "single":
            os.environ["RCB_K1_MODE"] = "single"
        else:
            os.environ["RCB_K1_MODE"] = "cluster"; os.environ["RCB_CLUSTER_SIZE"] = cs
        rng = np.random.default_rng(17)
        esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), state_modifiers=R.NLAT2)
        ps, st = R.setup(rng, esn)
        for _ in range(2):
            s, st2 = R.collectstates(esn, data, ps, st)
        print(f"N={N} {cs}: {dev.kernel_ms(0) / T * 1e3:.2f} us/step  {dev.last_kernel()}", flush=True)
