"""BASELINE config 5 across GPUs: the (radius x leak) members shard over ranks (shard_range), every rank sweeps its share
on its own GPU, no collective on the data path.  Weak scaling: 256 recurrences x 16 lambdas per GPU.
    python scripts/run_c5_dist.py                      # 1 GPU
    torchrun --nproc-per-node N scripts/run_c5_dist.py # N GPUs"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
from oracle import esn_oracle as O
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl")
torch.cuda.set_device(local)
dev = R.Device(local)
PER = 256
M = PER * world
N, T, w = 1024, 5000, 100
rng = np.random.default_rng(17)
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=1.0, sparsity=6 / N), leak_coefficient=1.0)
ps, st = R.setup(rng, esn)
series = O.lorenz_series(T + 1)
data = np.asfortranarray(series[:, :T].astype(np.float32)); target = np.asfortranarray(series[:, 1:T + 1].astype(np.float32))
radii = np.linspace(0.5, 1.4, M).astype(np.float32); leaks = np.tile(np.linspace(0.1, 1.0, 16), M // 16 + 1)[:M].astype(np.float32)
lo, hi = R.shard_range(M, rank, world)
lams = np.logspace(-8, -1, 16)
st0 = R.resetcarry(None, esn, st, init_carry=lambda r, n: np.zeros((n, hi - lo), np.float32))
def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
for rep in range(2):
    sync(); t0 = time.perf_counter()
    W, hT = R.train_ensemble(esn, data, target, ps, st0, w_scale=radii[lo:hi], leak=leaks[lo:hi], lambdas=lams, washout=w, device=dev)
    sync(); dt = time.perf_counter() - t0
ok = bool(np.isfinite(W).all())
t = torch.tensor([dt], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"C5 on {world} GPU(s): {M} recurrences x 16 lambdas = {M * 16} readouts in {float(t):.2f} s wall (max over ranks) = "
          f"{M * 16 / float(t):.0f} readouts/s; finite {ok}; per-rank device time K1 {dev.kernel_ms(0):.0f} ms, K2 {dev.kernel_ms(1):.0f} ms, K3 {dev.kernel_ms(2):.0f} ms")
if world > 1:
    dist.destroy_process_group()
