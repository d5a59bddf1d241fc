"""K2 timing: tensor-core Gram (and optionally the exact FP64 kernel) on device-resident states.
usage: python scripts/bench_gram.py F S [D] [--fp64]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L

F, S = int(sys.argv[1]), int(sys.argv[2])
D = int(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("-") else 3
modes = [("tensor", L.GRAM_TENSOR)] + ([("fp64", L.GRAM_FP64)] if "--fp64" in sys.argv else [])
torch.cuda.set_device(0)
dev = R.Device(0, torch.cuda.current_stream().cuda_stream)
lib = R.lib()
Z = torch.tanh(torch.randn((S, F), device="cuda", dtype=torch.float32))      # memory order == column-major F x S
Y = torch.randn((S, D), device="cuda", dtype=torch.float32)
GC = torch.zeros((F + D, F), device="cuda", dtype=torch.float64)             # column-major F x (F + D)
res = {}
for name, mode in modes:
    for rep in range(3):
        L.check(lib.rcb_gram_accumulate(dev.handle, L.ptr(Z), 0, F, S, F, L.ptr(Y), 0, D, D, mode, 1, L.ptr(GC)))
        ms = dev.kernel_ms(L.TIMER_GRAM)
        if os.environ.get('RCB_GRAM_TIMING'): print('   rep', rep, 'total', ms, 'mma kernel only', dev.kernel_ms(L.TIMER_SOLVE))
    alg = (F * (F + 1) + 2 * F * D) * S            # lower triangle + cross term
    print(f"{name}: F={F} S={S} D={D}  {ms:.3f} ms  algorithmic {alg / ms / 1e9:.1f} TFLOP/s  "
          f"full-square-equivalent {2 * F * F * S / ms / 1e9:.1f} TFLOP/s", flush=True)
    res[name] = GC.clone()
Cref = (Z.double().T @ Y.double()).T.contiguous()    # (D, F) == rows F..F+D of the packed array
for name in res:
    e = (res[name][F:] - Cref).abs()
    bad = (e > 1e-6).nonzero()
    print(f"{name}: cross term max abs err vs torch fp64: {float(e.max()):.3e}  bad entries {len(bad)}"
          + (f" rows(d) {bad[:, 0].unique().tolist()} f range {int(bad[:, 1].min())}..{int(bad[:, 1].max())}" if len(bad) else ""))
if len(res) == 2:
    a, b = res["tensor"][:F].T.tril(), res["fp64"][:F].T.tril()
    d = torch.sqrt(torch.outer(torch.diag(b), torch.diag(b)))
    print("max rel diff tensor vs fp64:", float(((a - b).abs() / d).max()), " cross:", float((res["tensor"][F:] - res["fp64"][F:]).abs().max()))
