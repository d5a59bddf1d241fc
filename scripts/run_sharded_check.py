"""Pooled readout over sequences sharded across GPUs (rcb200.train_sharded): every rank trains on its shard, the packed
[G|C] is all-reduced over NCCL, every rank solves; the result must equal the single-process fit of all sequences."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200")); sys.path.insert(0, ROOT)
import rcb200 as R
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
import torch.distributed as dist
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29555")
dist.init_process_group("nccl", rank=rank, world_size=world)
torch.cuda.set_device(local)
dev = R.Device(local)
N, B, T, w = 1024, 96, 400, 50
rng = np.random.default_rng(5)
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), leak_coefficient=0.6, state_modifiers=R.NLAT2)
ps, st = R.setup(rng, esn)
data = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
tgt = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
h0 = rng.standard_normal((N, B)).astype(np.float32)
lo, hi = R.shard_range(B, rank, world)
st_r = R.resetcarry(None, esn, st, init_carry=lambda r, n: h0[:, lo:hi])
ps_s, _ = R.train_sharded(esn, np.asfortranarray(data[:, :, lo:hi]), np.asfortranarray(tgt[:, :, lo:hi]), ps, st_r,
                          objective=R.RidgeRegression(1e-2), washout=w, device=dev)
st_all = R.resetcarry(None, esn, st, init_carry=lambda r, n: h0)
ps_a, _ = R.train(esn, data, tgt, ps, st_all, objective=R.RidgeRegression(1e-2), washout=w, device=dev)
err = np.linalg.norm(ps_s.readout.weight - ps_a.readout.weight) / np.linalg.norm(ps_a.readout.weight)
t = torch.tensor([err], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"train_sharded on {world} GPU(s) vs one-process train of all {B} sequences: max relative weight difference {float(t):.2e}")
dist.destroy_process_group()
