"""Single-sequence collectstates (latency-bound shape) for profiling: N, T from argv."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
N, T = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(17)
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), state_modifiers=R.NLAT2)
ps, st = R.setup(rng, esn)
data = np.asfortranarray(O.lorenz_series(T).astype(np.float32))
dev = R.default_device()
for _ in range(2):
    s, st2 = R.collectstates(esn, data, ps, st)
print(f"N={N} T={T}: K1 {dev.kernel_ms(0):.2f} ms = {dev.kernel_ms(0) / T * 1e3:.2f} us/step  {dev.last_kernel()}")
