"""ES2N sequences (dense orthogonal skip operator) through collectstates: N, T [, B] from argv."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
N, T = int(sys.argv[1]), int(sys.argv[2])
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rc = R.ES2N(3, N, 3, proximity=0.4, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), state_modifiers=R.NLAT2)
ps, st = R.setup(np.random.default_rng(1), rc)
data = np.asfortranarray(O.lorenz_series(T)[:, :T].astype(np.float32) / 20)
if B > 1:
    data = np.asfortranarray(np.repeat(data[:, :, None], B, axis=2) * np.linspace(0.5, 1.0, B, dtype=np.float32))
dev = R.default_device()
for _ in range(2):
    s, _ = R.collectstates(rc, data, ps, st)
print(f"ES2N N={N} T={T} B={B}: K1 {dev.kernel_ms(0):.2f} ms = {dev.kernel_ms(0) / T * 1e3:.2f} us/step  {dev.last_kernel()}")
