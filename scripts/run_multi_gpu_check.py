"""Multi-GPU paths behind the C-ABI, checked on N GPUs (one process per GPU; torch.distributed carries only the NCCL id):
    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/run_multi_gpu_check.py
  1. rcb_train_sharded (Cholesky-on-Gram and the QR path) == the one-process rcb_train of all sequences;
  2. rcb_train_time_chunked gives the same readout on N ranks as on one;
  3. rcb_allreduce_gc / rcb_comm_allgather move what they should.
Prints one PASS/FAIL line per check from rank 0; exit code 1 on any failure."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))
sys.path.insert(0, ROOT)
import rcb200 as R  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29555")
dist.init_process_group(os.environ.get("RCB_BENCH_BACKEND", "nccl"), rank=rank, world_size=world)
torch.cuda.set_device(local)
dev = R.Device(local)
fails = 0


def report(name, err, tol):
    global fails
    t = torch.tensor([err], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = float(t) <= tol
    fails += 0 if ok else 1
    if rank == 0:
        print(f"{'PASS' if ok else 'FAIL'} {name}: max over ranks {float(t):.3e} (tol {tol:g}) on {world} GPU(s)", flush=True)


# ---- 1. sequences sharded, pooled readout ---------------------------------------------------------------------------
N, B, T, w = 1024, 96, 400, 50
rng = np.random.default_rng(5)
esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), leak_coefficient=0.6, state_modifiers=R.NLAT2)
ps, st = R.setup(rng, esn)
data = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
tgt = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
h0 = rng.standard_normal((N, B)).astype(np.float32)
lo, hi = R.shard_range(B, rank, world)
st_r = R.resetcarry(None, esn, st, init_carry=lambda r, n: h0[:, lo:hi])
st_all = R.resetcarry(None, esn, st, init_carry=lambda r, n: h0)
# (the exact FP64 Gram on both sides: under AUTO the shard sizes decide between the FP64 and the tensor-core kernel)
for lam, solver, name, tol in ((1e-2, R.NormalCholeskyFactorization(), "train_sharded (FP64 Gram + all-reduce)", 1e-8),
                               (1e-2, R.QRFactorization(), "train_sharded (QR factors merged)", 1e-8),
                               (0.0, None, "train_sharded lambda = 0 (AUTO -> QR)", 1e-6)):
    ps_s, st_s = R.train_sharded(esn, np.asfortranarray(data[:, :, lo:hi]), np.asfortranarray(tgt[:, :, lo:hi]), ps, st_r,
                                 objective=R.RidgeRegression(lam), solver=solver, washout=w, device=dev)
    ps_a, st_a = R.train(esn, data, tgt, ps, st_all, objective=R.RidgeRegression(lam), solver=solver, washout=w, device=dev)
    report(name, np.linalg.norm(ps_s.readout.weight - ps_a.readout.weight) / np.linalg.norm(ps_a.readout.weight), tol)
    report(name + " carry", float(np.max(np.abs(st_s.reservoir.carry[0] - st_a.reservoir.carry[0][:, lo:hi]))), 0.0)

# ---- 2. one long sequence, time-chunked --------------------------------------------------------------------------------
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (series generator)

T2, w2 = 60000, 600
sig = bench.mackey_glass(T2 + 1)
esn2 = R.ESN(1, 512, 1, init_reservoir=R.rand_sparse(radius=0.8, sparsity=6 / 512))
ps2, st2 = R.setup(np.random.default_rng(3), esn2)
st20 = R.resetcarry(None, esn2, st2, init_carry=lambda r, n: np.zeros(n, np.float32))
X, Y = np.asfortranarray(sig[:, :T2]), np.asfortranarray(sig[:, 1:])
psc, _ = R.train_time_chunked(esn2, X, Y, ps2, st20, objective=R.RidgeRegression(1e-3), washout=w2, n_chunks=37, overlap=400, device=dev)
# the same call on ONE rank: a private single-rank run (group of one is "no comm")
import rcb200.api as A  # noqa: E402

saved = A.comm_for
A.comm_for = lambda d=None, g=None: None
ps1, _ = R.train_time_chunked(esn2, X, Y, ps2, st20, objective=R.RidgeRegression(1e-3), washout=w2, n_chunks=37, overlap=400, device=dev)
A.comm_for = saved
report("train_time_chunked N ranks vs 1 rank", np.linalg.norm(psc.readout.weight - ps1.readout.weight) / np.linalg.norm(ps1.readout.weight), 1e-8)
pss, _ = R.train(esn2, X, Y, ps2, st20, objective=R.RidgeRegression(1e-3), washout=w2, device=dev)
states, _ = R.collectstates(esn2, X, ps2, st20, device=dev)
ys, yc = pss.readout.weight @ states[:, w2:].astype(np.float64), psc.readout.weight @ states[:, w2:].astype(np.float64)
report("train_time_chunked vs sequential (predictions)", np.linalg.norm(ys - yc) / np.linalg.norm(ys), 1e-4)

# ---- 3. raw collectives ---------------------------------------------------------------------------------------------------
comm = R.comm_for(dev)
F, D = 37, 3
GC = np.asfortranarray(np.random.default_rng(100 + rank).standard_normal((F, F + D)))
want = sum(np.random.default_rng(100 + r).standard_normal((F, F + D)) for r in range(world))
if comm is not None:
    got = comm.allreduce_gc(GC.copy(order="F"), D)
    err = max(np.max(np.abs(np.tril(got[:, :F]) - np.tril(want[:, :F]))), np.max(np.abs(got[:, F:] - want[:, F:])))
    report("rcb_allreduce_gc (lower triangle + C)", float(err), 1e-12)
    g = comm.allgather(np.full(5, rank, np.int32))
    report("rcb_comm_allgather", float(np.max(np.abs(g - np.arange(world, dtype=np.int32)[:, None]))), 0.0)
dist.destroy_process_group()
sys.exit(1 if fails else 0)
