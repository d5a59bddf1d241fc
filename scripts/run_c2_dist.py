"""BASELINE config 2 across GPUs: one N=4096 reservoir, a 1e6-step Mackey-Glass-shaped series, time-chunked after the washout.
Ranks take contiguous slices of the chunk list (rcb200.train_time_chunked), accumulate partial Grams, all-reduce the packed
[G|C] once (NCCL) and solve redundantly.  Strong scaling (the series is fixed).
    python scripts/run_c2_dist.py                      # 1 GPU
    torchrun --nproc-per-node N scripts/run_c2_dist.py # N GPUs"""
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
N, T, w = 4096, 1_000_000, 1000
rng = np.random.default_rng(17)
esn = R.ESN(1, N, 1, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=True), leak_coefficient=1.0)
ps, st = R.setup(rng, esn)
mg = O.mackey_glass_series(T + 1)
data = np.asfortranarray(mg[:, :T].astype(np.float32)); target = np.asfortranarray(mg[:, 1:T + 1].astype(np.float32))
st0 = R.resetcarry(None, esn, st, init_carry=lambda r, n: np.zeros(n, np.float32))
def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
for rep in range(2):
    sync(); t0 = time.perf_counter()
    ps_c, _ = R.train_time_chunked(esn, data, target, ps, st0, objective=R.RidgeRegression(1e-6 * T), washout=w, n_chunks=4096, overlap=500,
                                   device=dev)
    sync(); dt = time.perf_counter() - t0
t = torch.tensor([dt], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
wsum = float(np.abs(ps_c.readout.weight).sum())
if rank == 0:
    print(f"C2 on {world} GPU(s): T={T}, 4096 chunks, overlap 500: train {float(t):.3f} s wall (max over ranks, host staging included); "
          f"per-rank device time K1 {dev.kernel_ms(0):.1f} ms, K2 {dev.kernel_ms(1):.1f} ms, K3 {dev.kernel_ms(2):.1f} ms; |W_out|_1 = {wsum:.6f}")
if world > 1:
    dist.destroy_process_group()
