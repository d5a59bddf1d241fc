"""K3 timing: FP64 Cholesky ridge solve of a random SPD [G | C]."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L
torch.cuda.set_device(0)
dev = R.Device(0, torch.cuda.current_stream().cuda_stream)
lib = R.lib()
for F in [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192]:
    D = 3
    A = torch.randn((F, 2 * F), device="cuda", dtype=torch.float64)
    G = A @ A.T
    C = torch.randn((D, F), device="cuda", dtype=torch.float64)
    GC = torch.cat([G, C], dim=0).contiguous()       # (F + D, F) row-major == column-major F x (F + D)
    W = torch.empty((F, D), device="cuda", dtype=torch.float64)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        L.check(lib.rcb_ridge_solve(dev.handle, L.ptr(GC), F, D, 1e-3, L.ptr(W)))
        e1.record(); torch.cuda.synchronize()
    ref = torch.linalg.solve(G + 1e-3 * torch.eye(F, device="cuda", dtype=torch.float64), C.T)
    err = float((W - ref).norm() / ref.norm())
    print(f"F={F}: ridge solve {e0.elapsed_time(e1):.2f} ms  ({F**3 / 3 / e0.elapsed_time(e1) / 1e9:.2f} TFLOP/s FP64)  rel err vs torch {err:.2e}", flush=True)
