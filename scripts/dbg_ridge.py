"""Debug: chunked pooled Gram through rcb_collect + rcb_gram_accumulate (tensor vs exact), PD check."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L
from rcb200.api import _Handle
import bench as Bn
N, B, T, Tc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 250
NCH = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
w = dict(N=N, d_in=3, k=6, radius=0.9, leak=0.5)
torch.cuda.set_device(0)
dev = R.Device(0, torch.cuda.current_stream().cuda_stream)
W, Win = Bn.make_weights(w)
esn = R.ESN(3, N, 3, init_reservoir=lambda rng, *d: W, init_input=lambda rng, *d: Win, state_modifiers=R.NLAT2, leak_coefficient=0.5)
ps, st = R.setup(np.random.default_rng(0), esn)
h = _Handle(esn, ps, np.float32, dev)
lib = R.lib()
u = Bn.lorenz_batch(B, T, seed=42)
chunks = [torch.from_numpy(np.ascontiguousarray(u[:, c * Tc:(c + 1) * Tc, :].transpose(2, 1, 0))).cuda() for c in range(T // Tc)]
h0 = torch.from_numpy(np.random.default_rng(1).standard_normal((B, N)).astype(np.float32)).cuda()
ring = torch.empty((B, Tc, N), dtype=torch.float32, device="cuda")
res = {}
for name, mode in (("tensor", L.GRAM_TENSOR), ("fp64", L.GRAM_FP64)):
    GC = torch.zeros((N + 3, N), dtype=torch.float64, device="cuda")
    carry = h0.clone()
    for c, uc in enumerate(chunks[:NCH]):
        L.check(lib.rcb_collect(h.handle, L.ptr(uc), Tc, B, L.ptr(carry), L.ptr(ring), L.ptr(carry)))
        L.check(lib.rcb_gram_accumulate(dev.handle, L.ptr(ring), 0, N, Tc * B, N, L.ptr(uc), 0, 3, 3, mode, 1 if c == 0 else 0, L.ptr(GC)))
    torch.cuda.synchronize()
    G = GC[:N].T.tril()
    G = G + G.tril(-1).T
    ev = torch.linalg.eigvalsh(G)
    print(name, "diag mean", float(G.diag().mean()), "eig min", float(ev[0]), "eig max", float(ev[-1]), flush=True)
    res[name] = G
d = (res["tensor"] - res["fp64"])
S = min(NCH, len(chunks)) * Tc * B
rel = d.abs() / torch.sqrt(torch.outer(res["fp64"].diag(), res["fp64"].diag()))
print("S", S, "max rel (to sqrt GiiGjj)", float(rel.max()), "implied per-sample rel noise", float(rel.max()) * S ** 0.5 / 4.5)
print("max abs diff", float(d.abs().max()), "rel to max diag", float(d.abs().max() / res["fp64"].diag().max()), "||E||2", float(torch.linalg.matrix_norm(d, 2)))
