"""K1 experiments on the C3 shape: time one 250-step launch of k1_fast7 for every RCB_F7_EXP value given on the command
line (a testing hook of recurrence_fast7.cu that selects kernel variants) and compare carry / last states with EXP=0.
Usage: python scripts/k1_exp.py 0 1 2 3"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "reservoircomputing.jl_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
import rcb200 as R  # noqa: E402
from rcb200 import _lib as L  # noqa: E402
from rcb200.api import _Handle  # noqa: E402

w = bench.WORKLOADS["c3"]
N, d_in, B, Tc = w["N"], w["d_in"], w["B"], w["chunk"]
torch.cuda.set_device(0)
stream = torch.cuda.current_stream()
dev = R.Device(0, stream.cuda_stream)
W, Win = bench.make_weights(w)
esn = R.ESN(d_in, N, d_in, init_reservoir=lambda rng, *d: W, init_input=lambda rng, *d: Win, state_modifiers=getattr(R, w["mod"]),
            leak_coefficient=w["leak"])
ps, st = R.setup(np.random.default_rng(0), esn)
h = _Handle(esn, ps, np.float32, dev)
lib = R.lib()
u = torch.from_numpy(np.ascontiguousarray(bench.lorenz_batch(B, Tc, seed=42).transpose(2, 1, 0))).cuda()
h0 = torch.from_numpy(np.random.default_rng(1).standard_normal((B, N)).astype(np.float32)).cuda()
ring = torch.empty((B, Tc, h.F), dtype=torch.float32, device="cuda")
carry = torch.empty_like(h0)
ref = None
for exp in sys.argv[1:] or ["0"]:
    os.environ["RCB_F7_EXP"] = exp
    ms = []
    for it in range(6):
        L.check(lib.rcb_collect(h.handle, L.ptr(u), Tc, B, L.ptr(h0), L.ptr(ring), L.ptr(carry)))
        torch.cuda.synchronize()
        ms.append(dev.kernel_ms(L.TIMER_RECURRENCE))
    t = float(np.median(ms[2:]))
    cur = (carry.clone(), ring[:, -1, :].clone(), ring[:, 0, :].clone())
    if ref is None:
        ref, note = cur, "reference"
    else:
        note = "carry bit-identical %s, first/last states identical %s" % (bool(torch.equal(cur[0], ref[0])), bool(torch.equal(cur[1], ref[1]) and torch.equal(cur[2], ref[2])))
    print(f"RCB_F7_EXP={exp}: {t:.3f} ms per {Tc}-step launch = {t * w['T'] / Tc:.1f} ms per C3 pass; {dev.last_kernel()}; {note}", flush=True)
