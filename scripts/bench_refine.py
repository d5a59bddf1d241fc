"""Timing of the iterative-refinement kernels (exact residual of the normal equations) on device-resident states.
usage: python scripts/bench_refine.py F S [lambda_per_sample]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from rcb200 import _lib as L

F, S = int(sys.argv[1]), int(sys.argv[2])
lps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
torch.cuda.set_device(0)
dev = R.Device(0, torch.cuda.current_stream().cuda_stream)
lib = R.lib()
g = torch.Generator(device="cuda").manual_seed(1)
Z = torch.tanh(torch.randn(S, 48, device="cuda", generator=g) @ torch.randn(48, F, device="cuda", generator=g) * 0.35
               + 0.05 * torch.randn(S, F, device="cuda", generator=g)).contiguous()
Y = (Z[:, :64] @ torch.randn(64, 3, device="cuda", generator=g)).contiguous()
W = np.empty((3, F), dtype=np.float64, order="F")
os.environ["RCB_GRAM"] = "tensor"
for rep in range(2):
    L.check(lib.rcb_fit_readout(dev.handle, L.ptr(Z), 0, F, S, F, L.ptr(Y), 0, 3, S, 3, lps * S, L.GRAM_AUTO, L.ptr(W)))
    print(f"F={F} S={S} lambda/S={lps:g}: gram+residual {dev.kernel_ms(L.TIMER_GRAM):.2f} ms, solve {dev.kernel_ms(L.TIMER_SOLVE):.2f} ms, "
          f"refinement steps {dev.last_refine()}", flush=True)
