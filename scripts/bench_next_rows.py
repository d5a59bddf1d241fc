"""Measurements for the SURVEY 8f rows (device-resident inputs, CUDA-event timing on the library's stream where the
library exposes a timer, wall-clock around synchronised calls otherwise).  Output is copied to profiles/."""
import os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L
from oracle import esn_oracle as O

dev = R.Device(0, stream=torch.cuda.current_stream().cuda_stream)
lib = L.lib()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- f-2: DelayLayer / Extend / readout over features: HBM-bound copies ----------------------------------------------
n, T, B, D = 4096, 20000, 8, 2
x = torch.randn(B, T, n, device="cuda")            # memory image of an (n, T, B) column-major array
out = torch.empty(B, T, (D + 1) * n, device="cuda")
hist = torch.empty(B, D, n, device="cuda")
ms = timed(lambda: L.check(lib.rcb_delay_features(dev.handle, L.ptr(x), 0, n, T, B, D, 3, None, 0, L.ptr(out), L.ptr(hist))))
byt = 4 * n * T * B * (1 + D + 1)
print(f"f-2 rcb_delay_features n={n} T={T} B={B} D={D} stride=3: {ms:.2f} ms, {byt / ms / 1e6:.0f} GB/s (read 1x + write {D + 1}x)")
u = torch.randn(B, T, 3, device="cuda")
oute = torch.empty(B, T, n + 3, device="cuda")
ms = timed(lambda: L.check(lib.rcb_extend_features(dev.handle, L.ptr(u), 3, L.ptr(x), n, T * B, 0, L.ptr(oute))))
print(f"f-2 rcb_extend_features n={n} d_in=3 cols={T * B}: {ms:.2f} ms, {4 * (2 * n + 6) * T * B / ms / 1e6:.0f} GB/s")
W = torch.randn(n, 3, device="cuda", dtype=torch.float64)
y = torch.empty(B, T, 3, device="cuda", dtype=torch.float64)
ms = timed(lambda: L.check(lib.rcb_readout_apply(dev.handle, L.ptr(W), None, 0, L.ptr(x), 0, n, T * B, 3, L.ptr(y))))
print(f"f-2 rcb_readout_apply F={n} d_out=3 cols={T * B}: {ms:.2f} ms, {4 * n * T * B / ms / 1e6:.0f} GB/s of features")

# ---- f-4: correlation matrix (K2 Gram + symmetric fill) ---------------------------------------------------------------
for F, S in ((1024, 200000), (4096, 200000)):
    Z = torch.randn(S, F, device="cuda")
    Rm = torch.empty(F, F, device="cuda")
    ms = timed(lambda: L.check(lib.rcb_correlation_matrix(dev.handle, L.ptr(Z), 0, F, S, F, 0, L.ptr(Rm))), reps=3)
    print(f"f-4 rcb_correlation_matrix F={F} S={S}: {ms:.1f} ms, {F * (F + 1) * S / ms / 1e9:.1f} algorithmic TFLOP/s")

# ---- f-1: sibling cells, single sequence and batched ------------------------------------------------------------------
series = O.lorenz_series(4001).astype(np.float32) / 20
for name, N, Bq, Tq in (("es2n", 300, 1, 4000), ("es2n", 1024, 1, 4000), ("es2n", 1024, 148, 500), ("eusn", 1024, 1, 4000),
                        ("eusn", 8192, 1024, 250), ("resesn", 2048, 148 * 2, 250)):
    kw = dict(init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500))
    rc = {"es2n": lambda: R.ES2N(3, N, 3, proximity=0.4, **kw), "resesn": lambda: R.ResESN(3, N, 3, alpha=0.8, beta=0.6, **kw),
          "eusn": lambda: R.EuSN(3, N, 3, leak_coefficient=0.5, diffusion=0.2, state_modifiers=R.NLAT2, **kw)}[name]()
    ps, st = R.setup(np.random.default_rng(1), rc)
    h = R.api._handle(rc, ps, np.float32, dev)
    ud = torch.from_numpy(np.ascontiguousarray(np.tile(series[:, :Tq].T[None], (Bq, 1, 1)))).cuda()
    zs = torch.empty(Bq, Tq, h.F, device="cuda")
    hT = torch.empty(Bq, h.sumN, device="cuda")
    L.check(lib.rcb_collect(h.handle, L.ptr(ud), Tq, Bq, None, L.ptr(zs), L.ptr(hT)))
    L.check(lib.rcb_collect(h.handle, L.ptr(ud), Tq, Bq, None, L.ptr(zs), L.ptr(hT)))
    ms = dev.kernel_ms(0)
    print(f"f-1 {name} N={N} B={Bq} T={Tq}: {ms:.2f} ms = {ms / Tq * 1e3:.2f} us/step, {N * Bq * Tq / ms / 1e6:.2f} G neuron-steps/s  [{dev.last_kernel()}]")
