"""BASELINE config 2 for the record: single ESN N=4096, 1e6-step Mackey-Glass-shaped series, ridge readout.
(a) sequential semantics on a prefix (single sequence = one CTA: latency-bound), (b) time-chunked full series."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
N, T, w = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 1000
rng = np.random.default_rng(17)
esn = R.ESN(1, N, 1, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=True), leak_coefficient=1.0)
ps, st = R.setup(rng, esn)
t0 = time.time(); mg = O.mackey_glass_series(T + 1); print(f"series {time.time() - t0:.1f}s", flush=True)
data = np.asfortranarray(mg[:, :T].astype(np.float32)); target = np.asfortranarray(mg[:, 1:T + 1].astype(np.float32))
st0 = R.resetcarry(None, esn, st, init_carry=lambda r, n: np.zeros(n, np.float32))
dev = R.default_device()
Ts = 20000
t0 = time.time(); ps_s, _ = R.train(esn, data[:, :Ts], target[:, :Ts], ps, st0, objective=R.RidgeRegression(1e-6 * Ts), washout=w)
dt = time.time() - t0
k1 = dev.kernel_ms(0)
print(f"sequential prefix T={Ts}: {dt:.2f}s wall; K1 {k1:.1f} ms = {k1 / Ts * 1e3:.2f} us/step ({N * Ts / k1 * 1e3:.3e} neuron-steps/s), K2 {dev.kernel_ms(1):.1f} ms, K3 {dev.kernel_ms(2):.1f} ms  kernel {dev.last_kernel()}", flush=True)
for nch, ov in ((1024, 500), (4096, 500)):
    t0 = time.time()
    ps_c, _ = R.train_time_chunked(esn, data, target, ps, st0, objective=R.RidgeRegression(1e-6 * T), washout=w, n_chunks=nch, overlap=ov)
    dt = time.time() - t0
    print(f"time-chunked T={T}, {nch} chunks, overlap {ov}: {dt:.2f}s wall ({N * T / dt:.3e} neuron-steps/s incl. host staging)  "
          f"K1 {dev.kernel_ms(0):.1f} ms K2 {dev.kernel_ms(1):.1f} ms  kernel {dev.last_kernel()}", flush=True)
# readout quality on a held-out stretch (teacher-forced one-step prediction)
Y, _ = R.predict(esn, data[:, :3000], ps_c, st0)
print("one-step prediction NRMSE (chunked readout, first 3000 steps):", float(np.sqrt(np.mean((Y[:, w:] - target[:, w:3000]) ** 2)) / np.std(target[:, w:3000])))
