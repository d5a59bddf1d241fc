"""World-size-2 gloo tests (CPU) of the multi-GPU HOST logic: sharding, the single all-reduce of the packed
[G | C], the replicated solve, and the time-chunk bookkeeping of BASELINE config 2.  The per-rank compute is injected
(the oracle stands in for rcb_train_partial / rcb_ridge_solve, which need a GPU); the GPU suite runs the same
train_sharded with the CUDA compute on one rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    from oracle import esn_oracle as O

    rng = np.random.default_rng(3)
    N, B, T, w = 40, 6, 60, 7
    W = O.rand_sparse_like(rng, N, radius=0.9, sparsity=6 / N)
    Win = O.scaled_rand_like(rng, N, 3)
    series = O.lorenz_series(T + B + 1)
    data = np.stack([series[:, b:b + T] for b in range(B)], axis=2).astype(np.float32)
    target = np.stack([series[:, b + 1:b + T + 1] for b in range(B)], axis=2).astype(np.float32)
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    return dict(N=N, B=B, T=T, w=w, W=W, Win=Win, data=data, target=target, h0=h0, lam=1e-3)


def _oracle_partial(P):
    from oracle import esn_oracle as O

    def partial(rc, ps, st, d3, y3, washout):
        layer = O.ESNCellParams(input_matrix=P["Win"], reservoir_matrix=P["W"], leak=0.6, act=O.ACT_TANH, modifiers=((O.MOD_NLAT2, 0.0),))
        F = P["N"]
        GC = np.zeros((F, F + y3.shape[0]), dtype=np.float64, order="F")
        hT = np.zeros((F, d3.shape[2]), dtype=d3.dtype, order="F")
        carry = st.reservoir.carry[0]
        for b in range(d3.shape[2]):
            s, hs = O.collectstates([layer], d3[:, :, b], [carry[:, b]])
            Z, Y = s[:, washout:].astype(np.float64), y3[:, washout:, b].astype(np.float64)
            GC[:, :F] += np.tril(Z @ Z.T)
            GC[:, F:] += Z @ Y.T
            hT[:, b] = hs[0]
        return GC, hT, [st.reservoir]

    return partial


def _numpy_solve(GC, d_out, reg):
    F = GC.shape[0]
    G = np.tril(GC[:, :F]) + np.tril(GC[:, :F], -1).T + reg * np.eye(F)
    return np.linalg.solve(G, GC[:, F:]).T


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import rcb200 as R

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = _problem()
        esn = R.ESN(3, P["N"], 3, init_reservoir=lambda rng, *d: P["W"], init_input=lambda rng, *d: P["Win"],
                    state_modifiers=R.NLAT2, leak_coefficient=0.6)
        ps, st = R.setup(np.random.default_rng(0), esn)
        lo, hi = R.shard_range(P["B"], rank, world)
        st = R.resetcarry(None, esn, st, init_carry=lambda rng, n: P["h0"][:, lo:hi])
        ps2, st2 = R.train_sharded(esn, np.asfortranarray(P["data"][:, :, lo:hi]), np.asfortranarray(P["target"][:, :, lo:hi]), ps, st,
                                   objective=R.RidgeRegression(P["lam"]), washout=P["w"], _partial=_oracle_partial(P), _solve=_numpy_solve)
        q.put((rank, lo, hi, np.asarray(ps2.readout.weight), np.asarray(st2.reservoir.carry[0])))
    finally:
        dist.destroy_process_group()


def test_shard_range_and_time_chunks():
    import rcb200 as R

    for n, w in [(1024, 8), (10, 3), (7, 8), (4096, 8)]:
        parts = [R.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1
    T, wo, ov = 1_000_000, 1000, 500
    own = []
    for r in range(8):
        start, lo, hi = R.time_chunks(T, wo, r, 8, ov)
        assert (start == 0) if r == 0 else (start == lo - ov)
        own.append((lo, hi))
    assert own[0][0] == wo and own[-1][1] == T and all(own[i][1] == own[i + 1][0] for i in range(7))


def test_chunk_windows_cover_the_samples_once():
    import rcb200 as R

    for T, w, n, ov in [(1_000_000, 1000, 1024, 500), (6000, 300, 7, 250), (100, 10, 9, 10)]:
        Tc, wins, tail = R.chunk_windows(T, w, n, ov)
        assert len(wins) == n and all(hi - lo == Tc and lo - s == ov and s >= 0 for s, lo, hi in wins)
        assert wins[0][1] == w and all(a[2] == b[1] for a, b in zip(wins, wins[1:]))
        end = wins[-1][2]
        if tail is not None:
            assert tail[1] == end and tail[2] == T and tail[1] - tail[0] == ov
            end = tail[2]
        assert end == T
    with pytest.raises(R.ArgumentError):
        R.chunk_windows(1000, 10, 4, 50)      # overlap > washout: windows would start before t = 0
    with pytest.raises(R.ArgumentError):
        R.chunk_windows(20, 10, 11, 5)


def test_train_sharded_two_ranks_gloo():
    import torch.multiprocessing as mp

    from oracle import esn_oracle as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P = _problem()
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 3, 3, 6)
    assert np.array_equal(res[0][3], res[1][3])          # every rank solves the same reduced system
    # single-process reference: pooled ridge over all sequences (Float64 augmented QR of the oracle)
    layer = O.ESNCellParams(input_matrix=P["Win"], reservoir_matrix=P["W"], leak=0.6, act=O.ACT_TANH, modifiers=((O.MOD_NLAT2, 0.0),))
    Zs, Ys = [], []
    for b in range(P["B"]):
        s, hs = O.collectstates([layer], P["data"][:, :, b], [P["h0"][:, b]])
        Zs.append(s[:, P["w"]:]); Ys.append(P["target"][:, P["w"]:, b])
        r = 0 if b < 3 else 1
        assert np.allclose(res[r][4][:, b - 3 * r], hs[0], atol=1e-6)
    Wref = O.train_ridge_qr(np.concatenate(Zs, axis=1), np.concatenate(Ys, axis=1), P["lam"])
    assert np.linalg.norm(res[0][3] - Wref) / np.linalg.norm(Wref) <= 1e-6
