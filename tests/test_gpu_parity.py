"""GPU parity tests proper: the CUDA path (through the C-ABI, via the host mirror rcb200) against the
CPU oracle on the same seeded inputs, against the reference's golden vectors, and — at full
BASELINE sizes — through size-independent properties.

Tolerances (BASELINE.json north_star): teacher-forced states ≤ 1e-5 relative (FP32), readout weights
≤ 1e-4 relative, sparsity pattern bit-exact."""
import os
import sys

import numpy as np
import pytest

from oracle import esn_oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "reservoircomputing.jl_b200"))

STATE_RTOL = 1e-5
WEIGHT_RTOL = 1e-4


@pytest.fixture(scope="module")
def R():
    import build as rcb_build

    rcb_build.build()
    import rcb200

    return rcb200


def _mods_to_oracle(mods):
    return tuple((m.kind, m.param) for m in mods)


def _act_to_oracle(R, f):
    return {R.identity: O.ACT_IDENTITY, R.tanh: O.ACT_TANH, R.tanh_fast: O.ACT_TANH_FAST, R.sigmoid: O.ACT_SIGMOID,
            R.relu: O.ACT_RELU}[f]


def oracle_layers(R, rc, ps):
    import scipy.sparse as sp

    cps = ps.cells if isinstance(rc, R.DeepESN) else (ps.reservoir,)
    out = []
    for c, p, m in zip(rc.cells, cps, rc.layer_modifiers):
        W = p.reservoir_matrix
        W = sp.csr_matrix(W) if (sp.issparse(W) or W.shape[0] > 1500) else np.asarray(W)
        extra = {}
        if isinstance(c, R.ES2NCell):
            extra = dict(cell="es2n", orthogonal_matrix=np.asarray(p.orthogonal_matrix), proximity=c.proximity)
        elif isinstance(c, R.ResESNCell):
            extra = dict(cell="resesn", orthogonal_matrix=np.asarray(p.orthogonal_matrix), alpha=c.alpha, beta=c.beta)
        elif isinstance(c, R.EuSNCell):
            extra = dict(cell="eusn", diffusion=c.diffusion)
            W = np.asarray(W.toarray() if sp.issparse(W) else W)
        out.append(O.ESNCellParams(input_matrix=np.asarray(p.input_matrix), reservoir_matrix=W,
                                   bias=np.asarray(p.bias) if c.use_bias else None, leak=c.leak_coefficient,
                                   act=_act_to_oracle(R, c.activation), modifiers=_mods_to_oracle(m), **extra))
    return out


def state_err(a, b):
    """max_t ||a(:,t) - b(:,t)||_inf / max(||b(:,t)||_inf, 1e-30)   (SURVEY.md §8d parity metric)"""
    num = np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)), axis=0)
    den = np.maximum(np.max(np.abs(b.astype(np.float64)), axis=0), 1e-30)
    return float(np.max(num / den))


def fixed_h0(R, st, rc, hs):
    """st with explicit carries (so that GPU and oracle start from the same hidden state)."""
    return R.resetcarry(None, rc, st, init_carry=[(lambda rng, n, h=h: h) for h in hs])


def const_init(M):
    return lambda rng, *dims: np.asarray(M, dtype=np.float32).reshape(dims)


# ---- golden vectors ------------------------------------------------------------------------------
@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_modifier_goldens(R, T):  # test/states_tests.jl:9-26
    x = np.arange(1, 10).astype(T)
    cases = [(R.NLAT1, [1, 2, 9, 4, 25, 6, 49, 8, 81]), (R.NLAT2, [1, 2, 2, 4, 12, 6, 30, 8, 56]),
             (R.NLAT3, [1, 2, 8, 4, 24, 6, 48, 8, 9]), (R.PartialSquare(0.6), [1, 4, 9, 16, 25, 6, 7, 8, 9]),
             (R.ExtendedSquare, [1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 4, 9, 16, 25, 36, 49, 64, 81]),
             (R.Pad(10.0), [1, 2, 3, 4, 5, 6, 7, 8, 9, 10])]
    for mod, expected in cases:
        out = mod(x)
        assert out.dtype == T and np.array_equal(out, np.array(expected, T)), mod
    x0 = np.arange(0, 10).astype(T)  # doctests src/states.jl:186,257,329
    assert np.array_equal(R.NLAT2(x0), [0, 1, 0, 3, 6, 5, 20, 7, 42, 9])
    m = np.arange(21, dtype=T).reshape(3, 7).T
    assert np.array_equal(R.NLAT2(m), O.nlat2(m))  # matrix form: per column
    chain = (R.NLAT1, R.NLAT2, R.Pad(3.0), R.ExtendedSquare)  # generic (un-fused) chain
    assert np.array_equal(R.apply_modifiers(chain, np.asfortranarray(m)), O.apply_modifiers(_mods_to_oracle(chain), m))


def test_collectstates_golden_linear_cell(R):  # test/Interface/generic_contracts_tests.jl:296-300
    esn = R.ESN(2, 2, 2, R.identity, init_input=const_init(np.eye(2)), init_reservoir=const_init(np.eye(2)),
                init_state=R.zeros32)
    ps, st = R.setup(np.random.default_rng(0), esn)
    states, st2 = R.collectstates(esn, np.array([[3, 4], [4, 5]], np.float32), ps, st)
    assert states.dtype == np.float32
    assert np.array_equal(states, np.array([[3, 7], [4, 9]], np.float32))
    assert np.array_equal(st2.reservoir.carry[0], [7, 9])
    states3, _ = R.collectstates(esn, np.array([[1], [1]], np.float32), ps, st2)  # carry continues
    assert np.array_equal(states3[:, 0], [8, 10])


def test_cell_analytic_cases(R):  # test/Layers/esncell_tests.jl:131-263
    I3, Z3 = np.eye(3), np.zeros((3, 3))
    x = np.array([[10], [20], [30]], np.float32)
    kw = dict(init_state=R.zeros32)
    e = R.ESN(3, 3, 3, R.identity, init_input=const_init(I3), init_reservoir=const_init(Z3), **kw)
    ps, st = R.setup(np.random.default_rng(0), e)
    assert np.allclose(R.collectstates(e, x, ps, st)[0][:, 0], x[:, 0])  # identity + leak 1
    e0 = R.ESN(3, 3, 3, R.identity, init_input=const_init(I3), init_reservoir=const_init(I3), leak_coefficient=0.0, **kw)
    ps, st = R.setup(np.random.default_rng(0), e0)
    st = fixed_h0(R, st, e0, [np.array([4, 5, 6], np.float32)])
    assert np.allclose(R.collectstates(e0, x, ps, st)[0][:, 0], [4, 5, 6])  # leak 0 keeps h0
    eb = R.ESN(3, 3, 3, R.identity, init_input=const_init(I3), init_reservoir=const_init(Z3), use_bias=True,
               init_bias=const_init(np.ones(3)), **kw)
    ps, st = R.setup(np.random.default_rng(0), eb)
    assert np.allclose(R.collectstates(eb, x, ps, st)[0][:, 0], x[:, 0] + 1)  # bias adds
    # scalar vs vector leak (atol 1e-6), batch of 3
    rng = np.random.default_rng(42)
    xs = rng.random((4, 1, 3), dtype=np.float32)
    h0 = rng.random((6, 3), dtype=np.float32)
    outs = []
    for leak in (0.3, np.full(6, 0.3, np.float32)):
        ev = R.ESN(4, 6, 2, R.identity, init_input=const_init(np.eye(6, 4)), init_reservoir=const_init(np.eye(6)),
                   leak_coefficient=leak, **kw)
        ps, st = R.setup(np.random.default_rng(0), ev)
        st = fixed_h0(R, st, ev, [h0])
        outs.append(R.collectstates(ev, xs, ps, st)[0])
    assert np.allclose(outs[0], outs[1], atol=1e-6)
    assert np.allclose(outs[0][:, 0, :], 0.7 * h0 + 0.3 * (np.eye(6, 4) @ xs[:, 0, :] + h0), atol=1e-6)


# ---- C1: README Lorenz ESN -----------------------------------------------------------------------
@pytest.fixture(scope="module")
def c1(R):
    rng = np.random.default_rng(17)
    esn = R.ESN(3, 300, 3, init_reservoir=R.rand_sparse(radius=1.2, sparsity=6 / 300), state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, esn)
    series = O.lorenz_series(5000 + 1250)
    data = np.asfortranarray(series[:, :5000].astype(np.float32))
    target = np.asfortranarray(series[:, 1:5001].astype(np.float32))
    test = np.asfortranarray(series[:, 5000:6250].astype(np.float32))
    h0 = np.random.default_rng(1).standard_normal(300).astype(np.float32)
    st = fixed_h0(R, st, esn, [h0])
    return dict(esn=esn, ps=ps, st=st, data=data, target=target, test=test, h0=h0)


def test_c1_csr_pattern_bit_exact(R, c1):
    rp, ci, v = R.export_csr(c1["esn"], c1["ps"])
    rp0, ci0, v0 = O.dense_to_csr(c1["ps"].reservoir.reservoir_matrix)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and v.tobytes() == v0.tobytes()
    assert rp[-1] == 1800


def test_c1_states_parity(R, c1):
    states, st2 = R.collectstates(c1["esn"], c1["data"], c1["ps"], c1["st"])
    ref, hs = O.collectstates(oracle_layers(R, c1["esn"], c1["ps"]), c1["data"], [c1["h0"]])
    assert states.shape == (300, 5000) and states.dtype == np.float32
    assert state_err(states, ref) <= STATE_RTOL
    assert np.max(np.abs(st2.reservoir.carry[0] - hs[0])) <= STATE_RTOL * np.max(np.abs(hs[0]))


@pytest.mark.parametrize("lam", [1e-6, 1e-2])
def test_c1_train_weights_parity(R, c1, lam):
    (ps2, st2), sw = R.train(c1["esn"], c1["data"], c1["target"], c1["ps"], c1["st"],
                             objective=R.RidgeRegression(lam), washout=100, return_states=True)
    W = ps2.readout.weight
    assert W.shape == (3, 300) and W.dtype == np.float64  # Float64 λ promotes (train.jl:126-129)
    # solver parity on the SAME states (the reference's own test form, test_train_ridge.jl:93-141)
    Wref = O.train_ridge_qr(sw, c1["target"][:, 100:], lam)
    assert np.linalg.norm(W - Wref) / np.linalg.norm(Wref) <= WEIGHT_RTOL
    # end to end against the oracle's own states
    Wo, hs, so = O.train(oracle_layers(R, c1["esn"], c1["ps"]), c1["data"], c1["target"], [c1["h0"]], reg=lam, washout=100)
    assert state_err(sw, so) <= STATE_RTOL
    pred, pred_o = W @ so.astype(np.float64), Wo @ so.astype(np.float64)
    assert np.linalg.norm(pred - pred_o) / np.linalg.norm(pred_o) <= WEIGHT_RTOL
    if lam >= 1e-2:
        assert np.linalg.norm(W - Wo) / np.linalg.norm(Wo) <= WEIGHT_RTOL


def test_c1_predict_parity(R, c1):
    lam = 1e-6
    ps2, st2 = R.train(c1["esn"], c1["data"], c1["target"], c1["ps"], c1["st"], objective=R.RidgeRegression(lam))
    layers = oracle_layers(R, c1["esn"], c1["ps"])
    W = ps2.readout.weight
    hT = st2.reservoir.carry[0]
    # teacher-forced (src/predict.jl:163-177): Float32 state, Float64 readout
    Y, st3 = R.predict(c1["esn"], c1["test"][:, :200], ps2, st2)
    Yo, hso = O.predict_teacher(layers, W, c1["test"][:, :200], [hT])
    assert Y.dtype == np.float64 and Y.shape == (3, 200)
    assert state_err(Y, Yo) <= 1e-4
    # autoregressive (src/predict.jl:74-88): Float64 loop after the first feedback; chaotic, so short horizon
    Ya, st4 = R.predict(c1["esn"], 1250, ps2, st2, initialdata=c1["test"][:, 0])
    Yao, _ = O.predict_autoregressive(layers, W, 60, c1["test"][:, 0], [hT])
    assert Ya.shape == (3, 1250) and Ya.dtype == np.float64 and np.all(np.isfinite(Ya))
    assert state_err(Ya[:, :60], Yao) <= 1e-3
    assert st4.reservoir.carry[0].dtype == np.float64


# ---- batched recurrence: every sequences-per-CTA variant ------------------------------------------
@pytest.mark.parametrize("B", [1, 5, 150, 300, 500, 900])
def test_batched_matches_oracle_and_single(R, B):
    N, T, d = 96, 24, 3
    rng = np.random.default_rng(100 + B)
    esn = R.ESN(d, N, 2, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), state_modifiers=(R.NLAT2, R.Pad(1.0)),
                leak_coefficient=0.5, use_bias=True, init_bias=R.randn32)
    ps, st = R.setup(rng, esn)
    data = np.asfortranarray(rng.standard_normal((d, T, B)).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    st = fixed_h0(R, st, esn, [h0])
    states, st2 = R.collectstates(esn, data, ps, st)
    assert states.shape == (N + 1, T, B)
    layers = oracle_layers(R, esn, ps)
    for b in sorted(set([0, B // 2, B - 1])):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL, b
        assert np.allclose(st2.reservoir.carry[0][:, b], hs[0], rtol=1e-5, atol=1e-6)
    # a sequence computed inside a batch is bit-identical to the same sequence computed alone
    b = B - 1
    st1 = fixed_h0(R, st, esn, [h0[:, b:b + 1]])
    alone, _ = R.collectstates(esn, np.asfortranarray(data[:, :, b:b + 1]), ps, st1)
    assert np.array_equal(alone[:, :, 0], states[:, :, b])


@pytest.mark.parametrize("mods", ["none", "nlat1", "nlat3", "psq", "esq", "pad", "esq_pad", "nlat1_nlat2"])
def test_modifier_chains_in_recurrence(R, mods):
    chains = dict(none=(), nlat1=(R.NLAT1,), nlat3=(R.NLAT3,), psq=(R.PartialSquare(0.4),), esq=(R.ExtendedSquare,),
                  pad=(R.Pad(2.5),), esq_pad=(R.ExtendedSquare, R.Pad(1.0)), nlat1_nlat2=(R.NLAT1, R.NLAT2))
    N, T = 75, 17  # odd N: exercises the NLAT3 upper bound and the scalar store path
    rng = np.random.default_rng(7)
    esn = R.ESN(2, N, 2, init_reservoir=R.rand_sparse(radius=0.8, sparsity=0.1), state_modifiers=chains[mods])
    ps, st = R.setup(rng, esn)
    data = np.asfortranarray(rng.standard_normal((2, T)).astype(np.float32))
    h0 = rng.standard_normal(N).astype(np.float32)
    states, _ = R.collectstates(esn, data, ps, fixed_h0(R, st, esn, [h0]))
    ref, _ = O.collectstates(oracle_layers(R, esn, ps), data, [h0])
    assert states.shape == ref.shape
    assert state_err(states, ref) <= STATE_RTOL


def test_float64_data_and_sparse_handoff(R):
    N, T = 128, 40
    rng = np.random.default_rng(9)
    esn = R.ESN(1, N, 1, init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.05, return_sparse=True), state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, esn)  # SparseMatrixCSC hand-off (ext/RCSparseArraysExt.jl:5-7)
    data = np.asfortranarray(O.mackey_glass_series(T)[:, :T])
    assert data.dtype == np.float64
    h0 = rng.standard_normal(N)
    states, _ = R.collectstates(esn, data, ps, fixed_h0(R, st, esn, [h0]))
    ref, _ = O.collectstates(oracle_layers(R, esn, ps), data, [h0])
    assert states.dtype == np.float64  # eltype follows data (reservoircomputer.jl:123)
    assert state_err(states, ref) <= 1e-12
    rp, ci, v = R.export_csr(esn, ps)
    rp0, ci0, v0 = O.dense_to_csr(ps.reservoir.reservoir_matrix.toarray())
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and v.tobytes() == v0.tobytes()


def test_random_initial_state_when_no_carry(R):  # esn_cell.jl:169-173: h0 ~ init_state from st's rng
    esn = R.ESN(2, 40, 2)
    ps, st = R.setup(np.random.default_rng(3), esn)
    data = np.zeros((2, 3), np.float32)
    s1, st1 = R.collectstates(esn, data, ps, st)
    s2, _ = R.collectstates(esn, data, ps, st)
    assert np.array_equal(s1, s2) and np.any(s1 != 0)  # deterministic given st; non-zero start
    assert st1.reservoir.carry is not None
    st_reset = R.resetcarry(None, esn, st1)
    assert st_reset.reservoir.carry is None


# ---- DeepESN --------------------------------------------------------------------------------------
def test_deep_esn_parity(R):
    rng = np.random.default_rng(21)
    res = [64, 96, 80]
    desn = R.DeepESN(3, res, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.1), leak_coefficient=(1.0, 0.8, 0.6),
                     state_modifiers=(R.NLAT2, R.NLAT1, (R.NLAT2, R.Pad(1.0))))
    ps, st = R.setup(rng, desn)
    T, B = 50, 3
    data = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
    hs0 = [rng.standard_normal((n, B)).astype(np.float32) for n in res]
    st = fixed_h0(R, st, desn, hs0)
    states, st2 = R.collectstates(desn, data, ps, st)
    assert states.shape == (81, T, B)
    layers = oracle_layers(R, desn, ps)
    for b in range(B):
        ref, hs = O.collectstates(layers, data[:, :, b], [h[:, b] for h in hs0])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
        for l in range(3):
            assert np.allclose(st2.cells[l].carry[0][:, b], hs[l], rtol=1e-5, atol=1e-6)
    # train + autoregressive predict through all layers
    tgt = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
    ps2, st3 = R.train(desn, data, tgt, ps, st, objective=R.RidgeRegression(1e-3), washout=5)
    Y, _ = R.predict(desn, 20, ps2, st3, initialdata=data[:, 0, :])
    hT = [st3.cells[l].carry[0] for l in range(3)]
    for b in range(B):
        Yo, _ = O.predict_autoregressive(layers, ps2.readout.weight, 20, data[:, 0, b], [h[:, b] for h in hT])
        assert state_err(Y[:, :, b], Yo) <= 1e-4


# ---- ridge readout: the reference's known-answer tests (test/test_train_ridge.jl) -------------------
def _random_ridge_problem(rng, T, F, S, D):
    Z = rng.standard_normal((F, S)).astype(T)
    tw = rng.standard_normal((D, F)).astype(T)
    return Z, (tw @ Z + T(0.01) * rng.standard_normal((D, S)).astype(T)).astype(T), tw


def test_ridge_closed_form(R):  # :64-91
    Z = 2.0 * np.eye(4)
    tw = np.array([[1.0, 0.0, -1.0, 0.5], [0.0, 2.0, 0.0, -0.5]])
    W = R.fit_readout(R.RidgeRegression(0.25), Z, tw @ Z, solver=R.QRFactorization())
    assert W.shape == (2, 4) and np.allclose(W, tw * (4.0 / 4.25), rtol=1e-10)


@pytest.mark.parametrize("seed,F,S,D,lam", [(7, 5, 60, 2, 1e-2), (3, 8, 20, 3, 1e-4), (19, 4, 50, 2, 0.0),
                                            (41, 7, 45, 5, 5e-3), (43, 6, 25, 3, 1e-8), (5, 200, 1000, 3, 1e-3)])
def test_ridge_matches_qr_f64(R, seed, F, S, D, lam):  # :93-181,293-328
    Z, Y, _ = _random_ridge_problem(np.random.default_rng(seed), np.float64, F, S, D)
    W = R.fit_readout(R.RidgeRegression(lam), Z, Y)
    assert W.dtype == np.float64
    assert np.allclose(W, O.train_ridge_qr(Z, Y, lam), rtol=1e-9, atol=1e-12)
    assert np.allclose(W, O.reference_ridge(Z, Y, lam), rtol=1e-9, atol=1e-12)


def test_ridge_f32_and_promotion(R):  # :119-141, train.jl:126-129
    Z, Y, _ = _random_ridge_problem(np.random.default_rng(11), np.float32, 4, 50, 2)
    W = R.fit_readout(R.RidgeRegression(np.float32, 1e-3), Z, Y)
    assert W.dtype == np.float32 and np.allclose(W, O.reference_ridge(Z, Y, 1e-3), rtol=1e-4)
    assert R.fit_readout(R.RidgeRegression(1e-3), Z, Y).dtype == np.float64
    Zi = np.array([[1, 2, 3, 4], [2, 1, 0, 1]]); Yi = np.array([[1.0, 2.0, 3.0, 4.0]])
    assert np.allclose(R.fit_readout(R.RidgeRegression(1e-3), Zi, Yi), O.train_ridge_qr(Zi, Yi, 1e-3), rtol=1e-9)


def test_ridge_error_paths(R):  # :220-291
    with pytest.raises(R.DimensionMismatch, match="states has 10 samples, targets has 9"):
        R.fit_readout(R.RidgeRegression(1e-3), np.zeros((5, 10)), np.zeros((2, 9)))
    with pytest.raises(R.ArgumentError, match="at least one training sample"):
        R.fit_readout(R.RidgeRegression(1e-3), np.zeros((4, 0)), np.zeros((2, 0)))
    with pytest.raises(R.ArgumentError, match="must be ≥ 0"):
        R.fit_readout(R.RidgeRegression(-1e-3), np.ones((4, 20)), np.ones((2, 20)))
    with pytest.raises(R.ArgumentError, match="is not supported"):
        R.fit_readout(R.RidgeRegression(1e-3), np.ones((4, 20)), np.ones((2, 20)), solver="not_a_solver")
    with pytest.raises(R.ArgumentError, match="failed to solve"):  # singular Gram, λ = 0 (train.jl:194-196)
        R.fit_readout(R.RidgeRegression(0.0), np.ones((4, 20)), np.ones((2, 20)))


def test_train_and_predict_error_paths(R):
    esn = R.ESN(2, 30, 2)
    ps, st = R.setup(np.random.default_rng(0), esn)
    X = np.zeros((2, 10), np.float32)
    with pytest.raises(R.ArgumentError, match="at least one column"):
        R.collectstates(esn, np.zeros((2, 0), np.float32), ps, st)
    with pytest.raises(R.ArgumentError, match="washout must be ≥ 0"):
        R.train(esn, X, X, ps, st, washout=-1)
    with pytest.raises(R.ArgumentError, match="washout=10 is ≥ number of time steps=10"):
        R.train(esn, X, X, ps, st, washout=10)
    with pytest.raises(R.DimensionMismatch):
        R.train(esn, X, X[:, :9], ps, st)
    with pytest.raises(R.ArgumentError, match="steps must be ≥ 1"):
        R.predict(esn, 0, ps, st, initialdata=np.zeros(2, np.float32))
    e32 = R.ESN(2, 30, 3)
    ps3, st3 = R.setup(np.random.default_rng(0), e32)
    with pytest.raises(R.DimensionMismatch, match="autoregressive predict requires each output to have length 2"):
        R.predict(e32, 5, ps3, st3, initialdata=np.zeros(2, np.float32))  # predict.jl:62-72


def test_washout_and_autoregressive_identity(R):  # array_workflow_tests.jl:145-155, test_train_ridge.jl:330-342
    e = R.ESN(2, 2, 2, R.identity, init_input=const_init(np.eye(2)), init_reservoir=const_init(np.zeros((2, 2))),
              init_state=R.zeros32)
    ps, st = R.setup(np.random.default_rng(0), e)
    ps = R.addreadout(e, np.eye(2), ps, st)[0]
    Y, _ = R.predict(e, 3, ps, st, initialdata=np.array([1.0, 2.0], np.float32))
    assert np.array_equal(Y, [[1, 1, 1], [2, 2, 2]]) and Y.dtype == np.float64
    rng = np.random.default_rng(1)
    X = rng.standard_normal((2, 30)).astype(np.float32)
    A = rng.standard_normal((2, 2)).astype(np.float32)
    (ps2, _), sw = R.train(e, X, A @ X, ps, st, objective=R.RidgeRegression(1e-10), washout=4, return_states=True)
    assert sw.shape == (2, 26) and np.array_equal(sw, X[:, 4:])
    Yt, _ = R.predict(e, X, ps2, st)
    assert np.allclose(Yt, A @ X, atol=1e-4)  # train then teacher-forced predict recovers linear targets


# ---- full BASELINE sizes: size-independent properties ------------------------------------------------
def test_c3_shape_properties(R):
    """N=8192, many sequences: (i) spot-check sequences against the oracle, (ii) batch == alone,
    (iii) with identity activation the recurrence is linear in (h0, u)."""
    N, B, T = 8192, 300, 12
    rng = np.random.default_rng(17)
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=True),
                state_modifiers=R.NLAT2, leak_coefficient=0.5)
    ps, st = R.setup(rng, esn)
    lor = O.lorenz_series(T + B)
    data = np.asfortranarray(np.stack([lor[:, b:b + T] for b in range(B)], axis=2).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    states, st2 = R.collectstates(esn, data, ps, fixed_h0(R, st, esn, [h0]))
    assert states.shape == (N, T, B)
    layers = oracle_layers(R, esn, ps)
    for b in (0, 151, B - 1):
        ref, _ = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
    rp, ci, v = R.export_csr(esn, ps)
    assert rp[-1] == 6 * N and np.array_equal(np.bincount(ci, minlength=N), np.full(N, 6))
    lin = R.ESN(3, N, 3, R.identity, init_reservoir=const_init_sparse(ps.reservoir.reservoir_matrix),
                init_input=const_init(ps.reservoir.input_matrix), leak_coefficient=0.5)
    psl, stl = R.setup(np.random.default_rng(0), lin)
    a, b2 = data[:, :, :2], data[:, :, 2:4]
    sa, _ = R.collectstates(lin, np.asfortranarray(a), psl, fixed_h0(R, stl, lin, [h0[:, :2]]))
    sb, _ = R.collectstates(lin, np.asfortranarray(b2), psl, fixed_h0(R, stl, lin, [h0[:, 2:4]]))
    sab, _ = R.collectstates(lin, np.asfortranarray(a + 2 * b2), psl, fixed_h0(R, stl, lin, [h0[:, :2] + 2 * h0[:, 2:4]]))
    assert np.max(np.abs(sab - (sa + 2 * sb))) <= 1e-4 * np.max(np.abs(sab))


def const_init_sparse(M):
    return lambda rng, *dims: M


# ---- K2 on tensor cores (tcgen05 kind::i8 digit planes, exact INT32 accumulation, csrc/gram_int8.cu) ----------
@pytest.mark.parametrize("F,S,D", [(300, 5000, 3), (257, 1111, 1), (1024, 4096, 2), (640, 40000, 3), (384, 3000, 6), (512, 2000, 11)])
def test_gram_tensor_matches_exact(R, F, S, D):
    from rcb200 import _lib as L

    rng = np.random.default_rng(F + S)
    Z = np.asfortranarray((rng.standard_normal((F, S)) * rng.uniform(0.05, 1.0, (F, 1))).astype(np.float32))
    Y = np.asfortranarray(rng.standard_normal((D, S)).astype(np.float64 if D == 6 else np.float32))  # D = 6: Float64 targets
    g_t = R.gram(Z, Y, gram_mode=L.GRAM_TENSOR)   # D <= 8: cross term fused into the split pre-pass; D = 11: separate kernel
    g_e = R.gram(Z, Y, gram_mode=L.GRAM_FP64)
    Z64 = Z.astype(np.float64)
    G = np.tril(Z64 @ Z64.T)
    assert np.allclose(np.tril(g_e[:, :F]), G, rtol=1e-12, atol=1e-9)
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
    # fixed point: 23-bit quantisation per feature (unbiased) + exact integer accumulation
    assert np.max(np.abs(np.tril(g_t[:, :F]) - G) / scale) <= 1e-6
    assert np.linalg.norm(np.tril(g_t[:, :F]) - G) / np.linalg.norm(G) <= 2e-7
    assert np.allclose(g_t[:, F:], Z64 @ Y.astype(np.float64).T, rtol=1e-12, atol=1e-9)  # cross term: exact FP64 products


@pytest.mark.parametrize("env", [{"RCB_GRAM_PAIR": "1"}, {"RCB_GRAM_PAIR": "1", "RCB_GRAM_PHASES": "2"}, {"RCB_GRAM_PAIR": "0", "RCB_GRAM_PHASES": "2"},
                                 {"RCB_GRAM_PAIR": "0"}, {"RCB_GRAM_PAIR": "1", "RCB_GRAM_SYNC": "0"}])
def test_gram_tensor_pair_variant_and_phase_groups(R, monkeypatch, env):
    """Both forms of the kernel — one M = 256 MMA per CTA pair (cta_group::2, tensor-map copies completing on the leader's
    barrier; the default) and one M = 128 MMA per CTA with a multicast B tile — in the nine- and the eight-product mode,
    against the exact Gram: same digit products, same result."""
    from rcb200 import _lib as L

    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(5)
    F, S, D = 700, 70000, 3        # several super-tiles incl. a partial one; more than one 32768-sample accumulation range
    Z = np.asfortranarray((rng.standard_normal((F, S)) * rng.uniform(0.05, 1.0, (F, 1))).astype(np.float32))
    Y = np.asfortranarray(rng.standard_normal((D, S)).astype(np.float32))
    g_t = R.gram(Z, Y, gram_mode=L.GRAM_TENSOR)
    Z64 = Z.astype(np.float64)
    G = np.tril(Z64 @ Z64.T)
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
    assert np.max(np.abs(np.tril(g_t[:, :F]) - G) / scale) <= 1e-6
    assert np.linalg.norm(np.tril(g_t[:, :F]) - G) / np.linalg.norm(G) <= 2e-7
    assert np.allclose(g_t[:, F:], Z64 @ Y.astype(np.float64).T, rtol=1e-12, atol=1e-9)


def test_gram_tensor_sub_chunks(R):
    """Several sample sub-chunks per call (what F = 8192 x 256k samples needs): G and the fused cross term."""
    from rcb200 import _lib as L

    rng = np.random.default_rng(77)
    F, S, D = 384, 9000, 3
    Z = np.asfortranarray(rng.standard_normal((F, S)).astype(np.float32))
    Y = np.asfortranarray(rng.standard_normal((D, S)).astype(np.float32))
    os.environ["RCB_GRAM_SUB_SAMPLES"] = "2048"
    try:
        g_t = R.gram(Z, Y, gram_mode=L.GRAM_TENSOR)
    finally:
        os.environ.pop("RCB_GRAM_SUB_SAMPLES", None)
    Z64 = Z.astype(np.float64)
    G = np.tril(Z64 @ Z64.T)
    assert np.max(np.abs(np.tril(g_t[:, :F]) - G)) / np.max(np.diag(G)) <= 1e-6
    assert np.allclose(g_t[:, F:], Z64 @ Y.astype(np.float64).T, rtol=1e-12, atol=1e-9)


def test_ridge_tensor_gram_weights(R):  # the reference's own Float32 case (test_train_ridge.jl:119-141) at rtol 1e-4
    from rcb200 import _lib as L

    rng = np.random.default_rng(5)
    F, S, D, lam = 512, 20000, 3, 1e-3
    Z = np.asfortranarray(rng.standard_normal((F, S)).astype(np.float32))
    Wt = rng.standard_normal((D, F)).astype(np.float32)
    Y = np.asfortranarray((Wt @ Z + 0.01 * rng.standard_normal((D, S))).astype(np.float32))
    W = R.fit_readout(R.RidgeRegression(lam), Z, Y, gram_mode=L.GRAM_TENSOR)
    Wref = O.train_ridge_qr(Z, Y, lam)
    assert np.linalg.norm(W - Wref) / np.linalg.norm(Wref) <= WEIGHT_RTOL


def test_train_tensor_gram_washout_and_batch(R):
    """rcb_train with the tensor Gram over several sequences + washout against the exact-Gram run: same sample
    set (index arithmetic), predictions of the two readouts agree."""
    from rcb200 import _lib as L

    rng = np.random.default_rng(11)
    N, B, T, w = 512, 6, 700, 37
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), state_modifiers=(R.NLAT2, R.Pad(1.0)),
                leak_coefficient=0.7)
    ps, st = R.setup(rng, esn)
    series = O.lorenz_series(T + 1)
    data = np.asfortranarray(np.stack([series[:, :T] * (1 + 0.01 * b) for b in range(B)], axis=2).astype(np.float32))
    target = np.asfortranarray(np.stack([series[:, 1:T + 1] * (1 + 0.01 * b) for b in range(B)], axis=2).astype(np.float32))
    (ps_t, _), sw = R.train(esn, data, target, ps, st, objective=R.RidgeRegression(1e-3), washout=w, return_states=True,
                            gram_mode=L.GRAM_TENSOR)
    ps_e, _ = R.train(esn, data, target, ps, st, objective=R.RidgeRegression(1e-3), washout=w, gram_mode=L.GRAM_FP64)
    Zs = sw.reshape(sw.shape[0], -1, order="F").astype(np.float64)
    pt, pe = ps_t.readout.weight @ Zs, ps_e.readout.weight @ Zs
    assert np.linalg.norm(pt - pe) / np.linalg.norm(pe) <= 1e-5


# ---- K1 fast7 (csrc/recurrence_fast7.cu): the 7-sequences-per-CTA kernel the C3 bench line runs ----------
@pytest.mark.parametrize("N,B,T,mod,leak,d_in", [(4096, 1030, 5, "nlat2", 0.5, 3), (8192, 900, 3, "nlat2", 0.5, 3),
                                                  (4096, 1036, 4, "none", 1.0, 3), (4096, 1000, 6, "none", 0.7, 1)])
def test_fast7_kernel_parity(R, N, B, T, mod, leak, d_in):
    """B > 6 x #SMs selects k1_fast7; checked (i) against the oracle on spot sequences (incl. the ragged last CTA),
    (ii) bit-for-bit determinism, (iii) against the generic kernel on every sequence."""
    rng = np.random.default_rng(N + B)
    mods = R.NLAT2 if mod == "nlat2" else ()
    esn = R.ESN(d_in, N, d_in, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=True),
                state_modifiers=mods, leak_coefficient=leak)
    ps, st = R.setup(rng, esn)
    lor = O.lorenz_series(T + B)[:d_in]
    data = np.asfortranarray(np.stack([lor[:, b:b + T] for b in range(B)], axis=2).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    stf = fixed_h0(R, st, esn, [h0])
    os.environ.pop("RCB_K1_MODE", None)
    states, st2 = R.collectstates(esn, data, ps, stf)
    if R.default_device().sm_count() * 6 < B:
        assert R.default_device().last_kernel().startswith("k1_fast7"), R.default_device().last_kernel()
    layers = oracle_layers(R, esn, ps)
    for b in (0, 7 * 73 + 6, B - 1):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
        assert np.max(np.abs(st2.reservoir.carry[0][:, b] - hs[0])) <= STATE_RTOL * np.max(np.abs(hs[0]))
    again, _ = R.collectstates(esn, data, ps, stf)
    assert again.tobytes() == states.tobytes()
    if N == 8192 and d_in == 3:  # the experimental 512-thread schedule of the same kernel (weights are re-installed under the env switch)
        os.environ["RCB_K1_MODE"] = "fast7_512"
        try:
            esn2 = R.ESN(3, N, 3, init_reservoir=lambda rng, *d: ps.reservoir.reservoir_matrix,
                         init_input=lambda rng, *d: ps.reservoir.input_matrix, state_modifiers=mods, leak_coefficient=leak)
            ps2, st2b = R.setup(np.random.default_rng(0), esn2)
            half, _ = R.collectstates(esn2, data, ps2, fixed_h0(R, st2b, esn2, [h0]))
            assert "512thr" in R.default_device().last_kernel()
        finally:
            os.environ.pop("RCB_K1_MODE", None)
        e5 = np.max(np.abs(states - half).reshape(N, -1), axis=0) / np.maximum(np.max(np.abs(states).reshape(N, -1), axis=0), 1e-30)
        assert float(e5.max()) <= STATE_RTOL
    os.environ["RCB_K1_MODE"] = "v1"
    try:
        generic, _ = R.collectstates(esn, data, ps, stf)
        assert not R.default_device().last_kernel().startswith("k1_fast7")
    finally:
        os.environ.pop("RCB_K1_MODE", None)
    err = np.max(np.abs(states - generic).reshape(N, -1), axis=0) / np.maximum(np.max(np.abs(generic).reshape(N, -1), axis=0), 1e-30)
    assert float(err.max()) <= STATE_RTOL


@pytest.mark.parametrize("N,B,T,mod,bias,leakv,d_in", [(2048, 1030, 6, "nlat2", True, True, 3), (4096, 900, 4, "none", True, False, 1),
                                                        (2048, 1000, 5, "nlat2", False, False, 3), (8192, 890, 3, "nlat2", False, True, 3)])
def test_fast7_kernel_bias_leak_vector_and_n2048(R, N, B, T, mod, bias, leakv, d_in):
    """The 7-sequences-per-CTA kernel beyond the benchmark shape (r02): N = 2048, a cell bias, a per-neuron leak — against the
    oracle on spot sequences and against the generic kernel on every sequence."""
    rng = np.random.default_rng(N + B + 1)
    mods = R.NLAT2 if mod == "nlat2" else ()
    kw = dict(init_bias=lambda r, *d: (0.2 * r.standard_normal(d)).astype(np.float32)) if bias else {}
    esn = R.ESN(d_in, N, d_in, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=True), state_modifiers=mods,
                use_bias=bias, leak_coefficient=rng.uniform(0.2, 1.0, N).astype(np.float32) if leakv else 0.6, **kw)
    ps, st = R.setup(rng, esn)
    lor = O.lorenz_series(T + B)[:d_in]
    data = np.asfortranarray(np.stack([lor[:, b:b + T] for b in range(B)], axis=2).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    stf = fixed_h0(R, st, esn, [h0])
    os.environ.pop("RCB_K1_MODE", None)
    states, st2 = R.collectstates(esn, data, ps, stf)
    if R.default_device().sm_count() * 6 < B:
        name = R.default_device().last_kernel()
        assert name.startswith("k1_fast7") and (("bias/leak-vector" in name) == (bias or leakv)), name
    layers = oracle_layers(R, esn, ps)
    for b in (0, 7 * 61 + 3, B - 1):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
        assert np.max(np.abs(st2.reservoir.carry[0][:, b] - hs[0])) <= STATE_RTOL * np.max(np.abs(hs[0]))
    os.environ["RCB_K1_MODE"] = "v1"
    try:
        generic, _ = R.collectstates(esn, data, ps, stf)
    finally:
        os.environ.pop("RCB_K1_MODE", None)
    err = np.max(np.abs(states - generic).reshape(N, -1), axis=0) / np.maximum(np.max(np.abs(generic).reshape(N, -1), axis=0), 1e-30)
    assert float(err.max()) <= STATE_RTOL


def test_train_sharded_single_rank_matches_train(R):
    """train_sharded (rcb_train_partial -> [all-reduce] -> rcb_ridge_solve) == train (rcb_train) on one rank."""
    rng = np.random.default_rng(23)
    N, B, T, w = 256, 5, 300, 20
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), state_modifiers=R.NLAT2, leak_coefficient=0.8)
    ps, st = R.setup(rng, esn)
    series = O.lorenz_series(T + B + 1)
    data = np.asfortranarray(np.stack([series[:, b:b + T] for b in range(B)], axis=2).astype(np.float32))
    target = np.asfortranarray(np.stack([series[:, b + 1:b + T + 1] for b in range(B)], axis=2).astype(np.float32))
    st = fixed_h0(R, st, esn, [rng.standard_normal((N, B)).astype(np.float32)])
    ps_a, st_a = R.train(esn, data, target, ps, st, objective=R.RidgeRegression(1e-4), washout=w)
    ps_b, st_b = R.train_sharded(esn, data, target, ps, st, objective=R.RidgeRegression(1e-4), washout=w)
    # FP64 atomics make the Gram's summation order run-dependent (1e-16), the conditioning amplifies it
    assert np.linalg.norm(ps_a.readout.weight - ps_b.readout.weight) / np.linalg.norm(ps_a.readout.weight) <= 1e-5
    assert np.array_equal(st_a.reservoir.carry[0], st_b.reservoir.carry[0])


# ---- C5: ensemble / hyper-parameter sweep (rcb_esn_set_member_params + rcb_train_members) -----------------------
def test_ensemble_members_match_individual_esns(R):
    """Member b == the reference's train() on an ESN with reservoir_matrix * w_scale[b] and leak[b], for every lambda."""
    rng = np.random.default_rng(41)
    N, T, w = 96, 260, 30
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=1.0, sparsity=6 / N), state_modifiers=R.NLAT2, leak_coefficient=1.0)
    ps, st = R.setup(rng, esn)
    series = O.lorenz_series(T + 1)
    data = np.asfortranarray(series[:, :T].astype(np.float32))
    target = np.asfortranarray(series[:, 1:T + 1].astype(np.float32))
    scales = np.array([0.5, 0.9, 1.2, 0.9, 0.7], np.float32)
    leaks = np.array([1.0, 0.3, 0.6, 0.8, 0.45], np.float32)
    lams = np.array([1e-6, 1e-3, 1e-1])
    B = len(scales)
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    W, hT = R.train_ensemble(esn, data, target, ps, fixed_h0(R, st, esn, [h0]), w_scale=scales, leak=leaks, lambdas=lams, washout=w)
    assert W.shape == (B, len(lams), 3, N)
    for b in range(B):
        layer = O.ESNCellParams(input_matrix=ps.reservoir.input_matrix,
                                reservoir_matrix=(ps.reservoir.reservoir_matrix * scales[b]).astype(np.float32),
                                leak=float(leaks[b]), act=O.ACT_TANH, modifiers=((O.MOD_NLAT2, 0.0),))
        so, hs = O.collectstates([layer], data, [h0[:, b]])
        assert np.max(np.abs(hT[:, b] - hs[0])) <= 2e-5 * np.max(np.abs(hs[0]))
        for l, lam in enumerate(lams):
            Wo = O.train_ridge_qr(so[:, w:], target[:, w:], lam)
            pred, pred_o = W[b, l] @ so[:, w:].astype(np.float64), Wo @ so[:, w:].astype(np.float64)
            assert np.linalg.norm(pred - pred_o) / np.linalg.norm(pred_o) <= WEIGHT_RTOL
            if lam >= 1e-3:
                assert np.linalg.norm(W[b, l] - Wo) / np.linalg.norm(Wo) <= 2e-3


# ---- C2: one long sequence, time-chunked after the washout (an approximation: echo-state re-synchronisation) ----
def test_time_chunked_train_approaches_sequential(R):
    rng = np.random.default_rng(5)
    N, T, w = 200, 6000, 300
    esn = R.ESN(1, N, 1, init_reservoir=R.rand_sparse(radius=0.6, sparsity=6 / N), leak_coefficient=1.0)
    ps, st = R.setup(rng, esn)
    mg = O.mackey_glass_series(T + 1)
    data = np.asfortranarray(mg[:, :T].astype(np.float32))
    target = np.asfortranarray(mg[:, 1:T + 1].astype(np.float32))
    st0 = fixed_h0(R, st, esn, [np.zeros(N, np.float32)])
    ps_seq, _ = R.train(esn, data, target, ps, st0, objective=R.RidgeRegression(1e-4), washout=w)
    ps_chk, st_chk = R.train_time_chunked(esn, data, target, ps, st0, objective=R.RidgeRegression(1e-4), washout=w, n_chunks=7, overlap=250)
    Ws, Wc = ps_seq.readout.weight, ps_chk.readout.weight
    states, _ = R.collectstates(esn, data, ps, st0)
    ys, yc = Ws @ states[:, w:].astype(np.float64), Wc @ states[:, w:].astype(np.float64)
    assert np.linalg.norm(ys - yc) / np.linalg.norm(ys) <= 1e-4      # radius 0.6: 0.6^250 is far below FP32 resolution
    assert st_chk.reservoir.carry[0].shape == (N,)
    Tc, wins, tail = R.chunk_windows(T, w, 7, 250)
    assert wins[0][1] == w and (tail[2] if tail else wins[-1][2]) == T and all(a[2] == b[1] for a, b in zip(wins, wins[1:]))
    with pytest.raises(R.ArgumentError):
        R.chunk_windows(T, 100, 4, 250)


def test_gc_pack_roundtrip(R):
    from rcb200 import _lib as L

    rng = np.random.default_rng(3)
    F, D = 37, 3
    GC = np.asfortranarray(rng.standard_normal((F, F + D)))
    lib, dev = R.lib(), R.default_device()
    n = int(lib.rcb_gc_packed_len(F, D))
    assert n == F * (F + 1) // 2 + F * D
    packed = np.empty(n)
    L.check(lib.rcb_gc_pack(dev.handle, L.ptr(GC), F, D, L.ptr(packed)))
    ref = np.concatenate([GC[j:, j] for j in range(F)] + [GC[:, F + d] for d in range(D)])
    assert np.array_equal(packed, ref)
    out = np.zeros_like(GC, order="F")
    L.check(lib.rcb_gc_unpack(dev.handle, L.ptr(packed), F, D, L.ptr(out)))
    assert np.array_equal(np.tril(out[:, :F]), np.tril(GC[:, :F])) and np.array_equal(out[:, F:], GC[:, F:])
    assert np.count_nonzero(np.triu(out[:, :F], 1)) == 0


# ---- sibling cells: ES2N, ResESN, EuSN ------------------------------------------------
def _sibling_models(R, N, d_in, d_out, use_bias, mods):
    kw = dict(init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.1), use_bias=use_bias, init_bias=R.randn32, state_modifiers=mods)
    return {
        "es2n": R.ES2N(d_in, N, d_out, proximity=0.4, **kw),
        "resesn": R.ResESN(d_in, N, d_out, alpha=0.8, beta=0.6, **kw),
        "eusn": R.EuSN(d_in, N, d_out, leak_coefficient=0.6, diffusion=0.2, **kw),
    }


@pytest.mark.parametrize("name", ["es2n", "resesn", "eusn"])
@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("use_bias", [False, True])
def test_sibling_cells_collectstates(R, name, T, use_bias):  # the cells of test/Interface/generic_contracts_tests.jl:47-77
    N, steps, B = 96, 60, 3
    rng = np.random.default_rng(5)
    rc = _sibling_models(R, N, 3, 3, use_bias, (R.NLAT2, R.Pad(1.0)))[name]
    ps, st = R.setup(rng, rc)
    data = np.asfortranarray(rng.standard_normal((3, steps, B)).astype(T))
    h0 = rng.standard_normal((N, B)).astype(T)
    states, st2 = R.collectstates(rc, data, ps, fixed_h0(R, st, rc, [h0]))
    assert states.shape == (N + 1, steps, B) and states.dtype == T
    layers = oracle_layers(R, rc, ps)
    tol = STATE_RTOL if T == np.float32 else 1e-6  # Float64: gamma and (1-p) cross the ABI as doubles, O / W stay Float32
    for b in range(B):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= tol
        assert state_err(st2.reservoir.carry[0][:, b:b + 1], hs[0].reshape(-1, 1)) <= max(tol, STATE_RTOL)


def test_sibling_cells_degenerate_to_esn(R):  # the parameter choices of test/collectstates_regression_tests.jl:30-62
    N, steps = 64, 30
    rng = np.random.default_rng(11)
    W = R.rand_sparse(radius=0.9, sparsity=0.1)(rng, N, N)
    Win = R.scaled_rand(rng, N, 3)
    Om = R.orthogonal(rng, N, N)
    kw = dict(init_reservoir=const_init(W), init_input=const_init(Win))
    data = np.asfortranarray(rng.standard_normal((3, steps)).astype(np.float32))
    h0 = np.zeros(N, np.float32)

    def run(rc):
        ps, st = R.setup(np.random.default_rng(0), rc)
        return R.collectstates(rc, data, ps, fixed_h0(R, st, rc, [h0]))[0]

    esn = run(R.ESN(3, N, 2, leak_coefficient=1.0, **kw))
    es2n = run(R.ES2N(3, N, 2, proximity=1.0, init_orthogonal=const_init(Om), **kw))   # (1-1) O h + 1 cand
    res = run(R.ResESN(3, N, 2, alpha=0.0, beta=1.0, init_orthogonal=const_init(Om), **kw))
    assert state_err(es2n, esn) <= 1e-6 and state_err(res, esn) <= 1e-6
    Wsym = np.asfortranarray(W + W.T)  # W - W' = 0: EuSN with diffusion 0 is driven by the input alone
    eusn = run(R.EuSN(3, N, 2, leak_coefficient=1.0, diffusion=0.0, init_reservoir=const_init(Wsym), init_input=const_init(Win)))
    assert state_err(eusn, np.tanh(Win @ data)) <= 5e-6  # the kernels' tanh is 1-2 ulp intrinsics (~1e-7 absolute)


@pytest.mark.parametrize("name", ["es2n", "resesn", "eusn"])
def test_sibling_cells_train_predict(R, name):
    N, steps = 80, 300
    rng = np.random.default_rng(2)
    rc = _sibling_models(R, N, 3, 3, False, R.NLAT2)[name]
    ps, st = R.setup(rng, rc)
    series = O.lorenz_series(steps + 30).astype(np.float32) / 20
    data, tgt = np.asfortranarray(series[:, :steps]), np.asfortranarray(series[:, 1:steps + 1])
    h0 = rng.standard_normal(N).astype(np.float32)
    st = fixed_h0(R, st, rc, [h0])
    ps2, st2 = R.train(rc, data, tgt, ps, st, objective=R.RidgeRegression(1e-4), washout=20)
    layers = oracle_layers(R, rc, ps)
    Wo, hs, _ = O.train(layers, data, tgt, [h0], reg=1e-4, washout=20)
    assert np.linalg.norm(ps2.readout.weight - Wo) / np.linalg.norm(Wo) <= WEIGHT_RTOL
    test = np.asfortranarray(series[:, steps:steps + 25])
    Y, _ = R.predict(rc, test, ps2, st2)                      # teacher-forced
    Yo, _ = O.predict_teacher(layers, ps2.readout.weight, test, hs)
    assert state_err(Y, Yo) <= 1e-4
    Ya, _ = R.predict(rc, 10, ps2, st2, initialdata=test[:, 0])  # autoregressive
    Yao, _ = O.predict_autoregressive(layers, ps2.readout.weight, 10, test[:, 0], hs)
    assert state_err(Ya, Yao) <= 1e-3


def test_sibling_cell_error_paths(R):
    import ctypes as C
    from rcb200 import _lib as L

    dev = R.default_device()
    desc = L.EsnDesc(1, 2, 2, 0, 0, 0)
    layers = (L.LayerDesc * 1)()
    layers[0].res_dims, layers[0].activation, layers[0].leak_scalar = 8, 1, 1.0
    h = C.c_void_p()
    L.check(L.lib().rcb_esn_create(dev.handle, C.byref(desc), layers, C.byref(h)))
    try:
        with pytest.raises(R.ArgumentError):
            L.check(L.lib().rcb_esn_set_cell(h, 0, 9, 0.0, 0.0))
        Om = np.asfortranarray(np.eye(8, dtype=np.float32))
        with pytest.raises(R.ArgumentError):   # not an ES2N / ResESN layer yet
            L.check(L.lib().rcb_esn_set_orthogonal_dense(h, 0, L.ptr(Om), 8))
        L.check(L.lib().rcb_esn_set_cell(h, 0, L.CELL_ES2N, 0.5, 0.0))
        W = np.asfortranarray(np.eye(8, dtype=np.float32) * 0.5)
        Win = np.asfortranarray(np.ones((8, 2), np.float32))
        L.check(L.lib().rcb_esn_set_weights_dense(h, 0, L.ptr(W), 8, L.ptr(Win), 8, None))
        u = np.zeros((2, 4, 1), np.float32, order="F")
        out = np.zeros((8, 4, 1), np.float32, order="F")
        with pytest.raises(R.ArgumentError):   # orthogonal_matrix missing
            L.check(L.lib().rcb_collect(h, L.ptr(u), 4, 1, None, L.ptr(out), None))
        with pytest.raises(R.DimensionMismatch):
            L.check(L.lib().rcb_esn_set_orthogonal_dense(h, 0, L.ptr(Om), 4))
        L.check(L.lib().rcb_esn_set_orthogonal_dense(h, 0, L.ptr(Om), 8))
        L.check(L.lib().rcb_collect(h, L.ptr(u), 4, 1, None, L.ptr(out), None))
        assert np.all(np.isfinite(out))
    finally:
        L.lib().rcb_esn_destroy(h)


# ---- feature plumbing and conceptor blocks (8f-4) ---------------------------------------
def test_delay_layer_goldens_gpu(R):  # test/basic_tests.jl:204-246
    dl = R.DelayLayer(1, num_delays=2, stride=1)
    st = dl.initialstates(np.random.default_rng(0))
    x = np.arange(1, 7, dtype=np.float32)[None, :]
    out, st2 = dl.apply_sequence(x[:, :5], st)
    assert np.array_equal(out.T, [[1, 0, 0], [2, 1, 0], [3, 2, 1], [4, 3, 2], [5, 4, 3]])
    assert st2.clock == 5 and np.array_equal(st2.history, [[5, 4]])
    dl2 = R.DelayLayer(1, num_delays=2, stride=2)
    out, st2 = dl2.apply_sequence(x[:, :4], dl2.initialstates(np.random.default_rng(0)))
    assert np.array_equal(out.T, [[1, 0, 0], [2, 0, 0], [3, 2, 0], [4, 2, 0]])
    with pytest.raises(R.DimensionMismatch):
        dl.apply_sequence(np.zeros((2, 3), np.float32), st)  # :248-255


@pytest.mark.parametrize("D,stride,T1,T2", [(2, 1, 7, 5), (3, 2, 5, 8), (1, 3, 4, 9), (4, 5, 3, 30), (0, 1, 3, 2)])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_delay_layer_matches_oracle_and_resumes(R, D, stride, T1, T2, dt):
    n, B = 5, 3
    rng = np.random.default_rng(D * 10 + stride)
    x = np.asfortranarray(rng.standard_normal((n, T1 + T2, B)).astype(dt))
    inits = tuple((lambda r, d, b, k=k: np.full((d, b), k + 1.0, np.float32)) for k in range(D))
    dl = R.DelayLayer(n, num_delays=D, stride=stride, init_delay=inits)
    st = dl.initialstates(rng)
    o1, st1 = dl.apply_sequence(x[:, :T1], st)       # two calls: history and clock carry over exactly
    o2, st2 = dl.apply_sequence(x[:, T1:], st1)
    out = np.concatenate([o1, o2], axis=1)
    h0 = np.stack([np.full(n, k + 1.0) for k in range(D)], axis=1) if D else None
    for b in range(B):
        ref, hist, clock = O.delay_layer_sequence(x[:, :, b], D, stride, history=h0)
        assert np.array_equal(out[:, :, b], ref)      # pure data movement: bit-exact
        if D:
            assert np.array_equal(st2.history[:, :, b], hist)
    assert st2.clock == T1 + T2 and out.dtype == dt


def test_extend_features_gpu(R):
    rng = np.random.default_rng(1)
    u = np.asfortranarray(rng.standard_normal((3, 11, 4)).astype(np.float32))
    x = np.asfortranarray(rng.standard_normal((17, 11, 4)).astype(np.float32))
    assert np.array_equal(R.extend_features(u, x), np.concatenate([u, x], axis=0))
    assert np.array_equal(R.extend_features(u[:, :, 0].astype(np.float64), x[:, :, 0].astype(np.float64)),
                          O.extend(u[:, :, 0], x[:, :, 0]).astype(np.float64))


@pytest.mark.parametrize("kind", ["input", "state", "both"])
def test_delay_esn_models(R, kind):  # src/models/esn_delay.jl
    N, T = 60, 80
    rng = np.random.default_rng(8)
    kw = dict(init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.1), leak_coefficient=0.7)
    if kind == "input":
        rc = R.InputDelayESN(3, N, 3, num_delays=2, stride=2, state_modifiers=R.NLAT2, **kw)
        idl, sdl, post = (2, 2), None, ()
    elif kind == "state":
        rc = R.StateDelayESN(3, N, 3, num_delays=2, stride=1, state_modifiers=(R.NLAT2, R.Pad(1.0)), **kw)
        idl, sdl, post = None, (2, 1), ((O.MOD_NLAT2, 0.0), (O.MOD_PAD, 1.0))
    else:
        rc = R.DelayESN(3, N, 3, num_input_delays=1, input_stride=1, num_state_delays=2, state_stride=3, **kw)
        idl, sdl, post = (1, 1), (2, 3), ()
    ps, st = R.setup(rng, rc)
    assert ps.reservoir.input_matrix.shape == (N, 3 * ((idl[0] if idl else 0) + 1))
    series = O.lorenz_series(T + 20).astype(np.float32) / 20
    data, tgt = np.asfortranarray(series[:, :T]), np.asfortranarray(series[:, 1:T + 1])
    h0 = rng.standard_normal(N).astype(np.float32)
    st = fixed_h0(R, st, rc, [h0])
    states, st2 = R.collectstates(rc, data, ps, st)
    layers = oracle_layers(R, rc, ps)
    ref, hs = O.collectstates_delay(layers, data, [h0], idl, sdl, post)
    assert states.shape == ref.shape and state_err(states, ref) <= STATE_RTOL
    assert ps.readout.weight.shape[1] == N * ((sdl[0] if sdl else 0) + 1)  # esn_delay.jl:292,477 (Pad widens it at train time)
    # train (collect -> washout -> ridge) and teacher-forced predict continue the delay lines from st
    ps2, st3 = R.train(rc, data, tgt, ps, st, objective=R.RidgeRegression(1e-4), washout=10)
    Wo = O.train_ridge_qr(ref[:, 10:], tgt[:, 10:], 1e-4)
    assert np.linalg.norm(ps2.readout.weight - Wo) / np.linalg.norm(Wo) <= WEIGHT_RTOL
    test = np.asfortranarray(series[:, T:T + 15])
    Y, _ = R.predict(rc, test, ps2, st3)
    full, _ = O.collectstates_delay(layers, np.concatenate([data, test], axis=1), [h0], idl, sdl, post)
    Yo = O.linear_readout(ps2.readout.weight, full[:, T:].astype(np.float64))
    assert state_err(Y, Yo) <= 1e-4
    # autoregressive generation: the delay lines, clocks and the carry continue from st3 (history = the training tail)
    Ya, st4 = R.predict(rc, 12, ps2, st3, initialdata=test[:, 0])
    assert Ya.shape == (3, 12) and np.all(np.isfinite(Ya))
    # oracle: replay the training series so that its delay buffers and clocks reach the same content, then generate
    if idl is not None:
        _, ih, ic = O.delay_layer_sequence(data, *idl)
    zall, hs_all = O.collectstates_delay(layers, data, [h0], idl, None, ())
    if sdl is not None:
        _, sh, sc = O.delay_layer_sequence(zall, *sdl)
    cur, hcur, outs = test[:, 0].astype(np.float32), hs_all[0], []
    for _ in range(12):
        u = cur
        if idl is not None:
            u, ih, ic = O.delay_layer_step(cur.astype(np.float32), ih, ic, idl[1])
        hcur = O.esn_cell_step(layers[0], u.astype(np.float32), hcur.astype(np.float32))
        z = O.apply_modifiers(layers[0].modifiers, hcur)
        if sdl is not None:
            z, sh, sc = O.delay_layer_step(z, sh, sc, sdl[1])
            z = O.apply_modifiers(post, z)
        cur = O.linear_readout(ps2.readout.weight, z.astype(np.float64))
        outs.append(cur)
    assert state_err(Ya, np.stack(outs, axis=1)) <= 1e-3


@pytest.mark.parametrize("F,S,dt", [(20, 200, np.float64), (6, 30, np.float32), (300, 4000, np.float32)])
def test_correlation_matrix_and_ridge_map(R, F, S, dt):  # test/Models/conceptors_tests.jl:10-13,30-40
    rng = np.random.default_rng(2024)
    Z = np.asfortranarray(rng.standard_normal((F, S)).astype(dt))
    Rm = R.correlation_matrix(Z)
    ref = O.correlation_matrix(Z.astype(np.float64))
    assert Rm.dtype == dt and np.array_equal(Rm, Rm.T)
    assert np.max(np.abs(Rm - ref)) <= (1e-12 if dt == np.float64 else 1e-6)
    with pytest.raises(R.ArgumentError):
        R.correlation_matrix(np.zeros((3, 0)))
    Y = np.asfortranarray(rng.standard_normal((4, S)).astype(dt))
    M = R.ridge_map(Z, Y, 1e-2)                         # conceptors.jl:639-649
    Mo = O.train_ridge_qr(Z, Y, 1e-2, reg_dtype=dt)
    assert M.dtype == dt and np.linalg.norm(M - Mo) / np.linalg.norm(Mo) <= WEIGHT_RTOL


def test_correlation_matrix_tensor_path(R):
    rng = np.random.default_rng(5)
    Z = np.asfortranarray(rng.standard_normal((512, 20000)).astype(np.float32))
    Rm = R.correlation_matrix(Z, gram_mode=R.GRAM_TENSOR)
    ref = (Z.astype(np.float64) @ Z.astype(np.float64).T) / Z.shape[1]
    assert np.array_equal(Rm, Rm.T) and np.max(np.abs(Rm - ref)) <= 1e-6


# ---- host setup on the GPU: spectral radius for scale_radius! ------------------------------------------
@pytest.mark.parametrize("n,k,sparse", [(1, 1, False), (50, 6, False), (300, 6, False), (1024, 6, True), (1500, 6, True)])
def test_spectral_radius_matches_dense_eigvals(R, n, k, sparse):  # inits_components.jl:127; test/inits_tests.jl:9-15 tolerance 1e-5
    import scipy.sparse as sp

    rng = np.random.default_rng(n)
    W = O.rand_sparse_like(rng, n, radius=1.0, sparsity=k / n) * np.float32(1.7) if n > 1 else np.array([[-0.3]], np.float32)
    ref = float(np.max(np.abs(np.linalg.eigvals(W.astype(np.float64)))))
    rho, dim = R.spectral_radius(sp.csc_matrix(W) if sparse else W, return_dim=True)
    assert abs(rho - ref) <= 1e-7 * ref and 1 <= dim <= n
    Ws = R.scale_radius(sp.csc_matrix(W) if sparse else W, 0.9)
    Wd = Ws.toarray() if sparse else Ws
    assert abs(np.max(np.abs(np.linalg.eigvals(Wd.astype(np.float64)))) - 0.9) <= 1e-5


def test_spectral_radius_large_and_degenerate(R):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    n = 8192
    rng = np.random.default_rng(17)
    rows = np.stack([rng.choice(n, size=6, replace=False) for _ in range(n)], axis=1)
    W = sp.csc_matrix((rng.standard_normal(6 * n).astype(np.float32), rows.T.ravel(), np.arange(0, 6 * n + 1, 6)), shape=(n, n))
    rho, dim = R.spectral_radius(W, return_dim=True)
    # ARPACK with a handful of Ritz values and one start vector can miss the leading pair (it did in 1 of 8 tries at
    # N = 4096, where dense eigvals confirmed the GPU value): take the best of three starts with a wider wanted set
    ref = max(np.max(np.abs(spla.eigs(W.astype(np.float64), k=24, which="LM", return_eigenvectors=False, tol=1e-10, maxiter=20000,
                                      v0=np.random.default_rng(s).standard_normal(n)))) for s in range(3))
    assert abs(rho - ref) <= 1e-6 * ref and dim < 1500
    assert R.spectral_radius(np.zeros((64, 64), np.float32)) == 0.0          # zero matrix
    with pytest.raises(R.ArgumentError):                                        # scale_radius! of it -> Inf (inits_components.jl:129-136)
        R.scale_radius(np.zeros((64, 64), np.float32), 1.0)
    P = np.roll(np.eye(64, dtype=np.float32), 1, axis=0) * np.float32(0.5)      # all 64 eigenvalues on one circle
    assert abs(R.spectral_radius(P) - 0.5) <= 1e-9
    with pytest.raises(R.NumericError):
        R.spectral_radius(np.full((4, 4), np.nan, np.float32))


@pytest.mark.parametrize("res,T,B", [([30, 51, 40], 70, 2), ([256, 384], 300, 1), ([130, 129, 131], 129, 3)])
def test_deep_esn_input_projection_shapes(R, res, T, B):
    """DeepESN layers >= 2 consume a dense GEMM W_in * Z over all time steps (esn_deep.jl:270-280): tile-boundary and
    unaligned-K shapes of that GEMM against the step-by-step oracle."""
    rng = np.random.default_rng(sum(res))
    desn = R.DeepESN(2, res, 2, init_reservoir=R.rand_sparse(radius=0.8, sparsity=0.2), leak_coefficient=0.7, state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, desn)
    data = np.asfortranarray(rng.standard_normal((2, T, B)).astype(np.float32))
    hs0 = [rng.standard_normal((n, B)).astype(np.float32) for n in res]
    states, _ = R.collectstates(desn, data, ps, fixed_h0(R, st, desn, hs0))
    layers = oracle_layers(R, desn, ps)
    for b in range(B):
        ref, _ = O.collectstates(layers, data[:, :, b], [h[:, b] for h in hs0])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL


# ---- randomized configurations: every kernel variant picks itself from (N, B, dtype, leak kind, bias, activation, chain) ----
def _fuzz_cases():
    rng = np.random.default_rng(20240)
    acts = ["identity", "tanh", "tanh_fast", "sigmoid", "relu"]
    chains = ["none", "nlat1", "nlat2", "nlat3", "pad", "nlat2_pad", "esq", "psq", "nlat1_nlat3_pad"]
    cases = []
    for i in range(36):
        N = int(rng.choice([1, 2, 3, 7, 31, 32, 33, 64, 100, 257, 300, 513, 1024, 1100]))
        cases.append(dict(seed=i, N=N, d_in=int(rng.integers(1, 12)), B=int(rng.choice([1, 1, 2, 3, 5, 9, 40])),
                          T=int(rng.choice([1, 2, 5, 23])), f64=bool(rng.integers(0, 4) == 0), act=str(rng.choice(acts)),
                          chain=str(rng.choice(chains)), bias=bool(rng.integers(0, 2)), vleak=bool(rng.integers(0, 3) == 0),
                          sparse=bool(rng.integers(0, 2)), density=float(rng.choice([0.02, 0.1, 0.5, 1.0]))))
    return cases


@pytest.mark.parametrize("c", _fuzz_cases(), ids=lambda c: f"{c['seed']}-N{c['N']}-B{c['B']}-T{c['T']}-{c['act']}-{c['chain']}")
def test_randomized_configurations(R, c):
    import scipy.sparse as sp

    rng = np.random.default_rng(1000 + c["seed"])
    N, d_in, B, T = c["N"], c["d_in"], c["B"], c["T"]
    mods = dict(none=(), nlat1=(R.NLAT1,), nlat2=(R.NLAT2,), nlat3=(R.NLAT3,), pad=(R.Pad(-0.5),), nlat2_pad=(R.NLAT2, R.Pad(1.0)),
                esq=(R.ExtendedSquare,), psq=(R.PartialSquare(0.37),), nlat1_nlat3_pad=(R.NLAT1, R.NLAT3, R.Pad(2.0)))[c["chain"]]
    act = dict(identity=R.identity, tanh=R.tanh, tanh_fast=R.tanh_fast, sigmoid=R.sigmoid, relu=R.relu)[c["act"]]
    W = (rng.standard_normal((N, N)) * (rng.random((N, N)) < c["density"])).astype(np.float32)
    W *= np.float32(0.8 / max(1e-3, np.max(np.abs(np.linalg.eigvals(W.astype(np.float64)))))) if N <= 300 else np.float32(0.5 / np.sqrt(max(1.0, c["density"] * N)))
    Win = rng.uniform(-0.5, 0.5, (N, d_in)).astype(np.float32)
    leak = rng.uniform(0.1, 1.0, N).astype(np.float32) if c["vleak"] else float(rng.uniform(0.1, 1.0))
    Wm = sp.csc_matrix(W) if c["sparse"] else np.asfortranarray(W)
    esn = R.ESN(d_in, N, 2, act, init_reservoir=lambda r, *d: Wm, init_input=const_init(Win), use_bias=c["bias"],
                init_bias=R.randn32, leak_coefficient=leak, state_modifiers=mods)
    ps, st = R.setup(rng, esn)
    dt = np.float64 if c["f64"] else np.float32
    data = np.asfortranarray(rng.standard_normal((d_in, T, B)).astype(dt))
    h0 = rng.standard_normal((N, B)).astype(dt)
    states, st2 = R.collectstates(esn, data, ps, fixed_h0(R, st, esn, [h0]))
    layers = oracle_layers(R, esn, ps)
    tol = 1e-11 if c["f64"] else STATE_RTOL
    for b in range(B):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert states[:, :, b].shape == ref.shape and states.dtype == dt
        if N >= 8:
            assert state_err(states[:, :, b], ref) <= tol
        else:  # a column of 1-3 numbers can pass through zero: its own max-norm is no scale (FMA vs separate rounding); use the trajectory's
            assert np.max(np.abs(states[:, :, b].astype(np.float64) - ref)) <= tol * max(1.0, np.max(np.abs(ref)))
        assert state_err(st2.reservoir.carry[0][:, b:b + 1], hs[0].reshape(-1, 1)) <= max(tol, 1e-11) or N < 8
    rp, ci, v = R.export_csr(esn, ps)                      # pattern + values bit-exact for every hand-off form
    rp0, ci0, v0 = O.dense_to_csr(W)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and v.tobytes() == v0.astype(np.float32).tobytes()


def _fuzz_train_cases():
    rng = np.random.default_rng(777)
    cases = []
    for i in range(14):
        cases.append(dict(seed=i, N=int(rng.choice([5, 20, 64, 150, 300])), d=int(rng.integers(1, 4)), B=int(rng.choice([1, 1, 2, 4])),
                          T=int(rng.choice([60, 150, 400])), washout=int(rng.choice([0, 1, 17])), lam=float(rng.choice([1e-4, 1e-2, 1.0])),
                          chain=str(rng.choice(["none", "nlat2", "nlat2_pad", "esq"])), ro_act=str(rng.choice(["identity", "tanh"])),
                          f64=bool(rng.integers(0, 3) == 0)))
    return cases


@pytest.mark.parametrize("c", _fuzz_train_cases(), ids=lambda c: f"{c['seed']}-N{c['N']}-B{c['B']}-T{c['T']}-{c['chain']}-{c['ro_act']}")
def test_randomized_train_predict(R, c):
    """train (collect -> washout -> ridge) and both predict modes on random small configurations against the oracle."""
    rng = np.random.default_rng(5000 + c["seed"])
    N, d, B, T, w = c["N"], c["d"], c["B"], c["T"], c["washout"]
    mods = dict(none=(), nlat2=(R.NLAT2,), nlat2_pad=(R.NLAT2, R.Pad(1.0)), esq=(R.ExtendedSquare,))[c["chain"]]
    ro_act = dict(identity=R.identity, tanh=R.tanh)[c["ro_act"]]
    esn = R.ESN(d, N, d, init_reservoir=R.rand_sparse(radius=0.8, sparsity=min(1.0, 6 / N)), init_input=R.scaled_rand(scaling=0.5),
                leak_coefficient=float(rng.uniform(0.3, 1.0)), state_modifiers=mods, readout_activation=ro_act)
    ps, st = R.setup(rng, esn)
    dt = np.float64 if c["f64"] else np.float32
    tt = np.arange(T + 31)[None, :, None]   # sustained O(1) drive: the per-column relative metric needs columns of O(1) norm
    series = np.sin(tt * rng.uniform(0.05, 0.4, (d, 1, B)) + rng.uniform(0, 6.28, (d, 1, B))) + 0.3 * rng.standard_normal((d, T + 31, B))
    series = (series / np.max(np.abs(series))).astype(dt)
    data, tgt = np.asfortranarray(series[:, :T]), np.asfortranarray(0.5 * series[:, 1:T + 1])
    h0 = rng.standard_normal((N, B)).astype(dt)
    st = fixed_h0(R, st, esn, [h0])
    (ps2, st2), states_wo = R.train(esn, data, tgt, ps, st, objective=R.RidgeRegression(c["lam"]), washout=w, return_states=True)
    layers = oracle_layers(R, esn, ps)
    Zs, Ys, hT = [], [], []
    for b in range(B):
        z, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        stol = 1e-11 if c["f64"] else STATE_RTOL                                                # post-washout states (train.jl:240-243)
        if N >= 8:
            assert state_err(states_wo[:, :, b], z[:, w:]) <= stol
        else:  # a handful of numbers per column: judge against the trajectory's scale (see test_randomized_configurations)
            assert np.max(np.abs(states_wo[:, :, b].astype(np.float64) - z[:, w:])) <= stol * max(1.0, np.max(np.abs(z)))
        Zs.append(states_wo[:, :, b]); Ys.append(tgt[:, w:, b]); hT.append(hs[0])
    # the solver is judged on identical inputs: Householder QR of the augmented system built from the states the GPU
    # collected (a ridge problem with few samples amplifies the 1e-7 state rounding differences far beyond 1e-4)
    Wo = O.train_ridge_qr(np.concatenate(Zs, axis=1), np.concatenate(Ys, axis=1), c["lam"])   # pooled over the sequences
    assert ps2.readout.weight.dtype == np.float64
    assert np.linalg.norm(ps2.readout.weight - Wo) / np.linalg.norm(Wo) <= WEIGHT_RTOL
    test = np.asfortranarray(series[:, T:T + 30])
    Y, _ = R.predict(esn, test, ps2, st2)
    Ya, _ = R.predict(esn, 12, ps2, st2, initialdata=test[:, 0])
    for b in range(B):
        Yo, _ = O.predict_teacher(layers, ps2.readout.weight, test[:, :, b], [hT[b]], ro_act=_act_to_oracle(R, ro_act))
        assert state_err(Y[:, :, b], Yo) <= 1e-4
        Yao, _ = O.predict_autoregressive(layers, ps2.readout.weight, 12, test[:, 0, b], [hT[b]], ro_act=_act_to_oracle(R, ro_act))
        assert state_err(Ya[:, :, b], Yao) <= 1e-3


@pytest.mark.parametrize("N,d_in,mod,B,T", [(300, 3, "nlat2", 1, 60), (300, 1, "none", 2, 33), (1024, 3, "nlat2", 1, 40), (2048, 3, "none", 3, 25),
                                             (4096, 1, "nlat2", 1, 30), (4096, 3, "nlat2", 2, 12), (1100, 3, "nlat2", 1, 20), (65, 1, "nlat2", 5, 50),
                                             (300, 3, "nlat2_pad", 2, 21), (2048, 1, "pad", 1, 9)])
def test_single_sequence_kernel_parity(R, N, d_in, mod, B, T):
    """The one-sequence-per-CTA kernel (recurrence_single.cu: BASELINE configs 1, 2 sequential, 4) against the oracle, incl.
    N that is not a multiple of the CTA width, both input widths, the carry, and a second call continuing from it."""
    rng = np.random.default_rng(N + 7 * B)
    esn = R.ESN(d_in, N, d_in, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), leak_coefficient=0.6,
                state_modifiers=dict(nlat2=(R.NLAT2,), none=(), nlat2_pad=(R.NLAT2, R.Pad(1.0)), pad=(R.Pad(-2.5),))[mod])
    ps, st = R.setup(rng, esn)
    data = np.asfortranarray(rng.standard_normal((d_in, 2 * T, B)).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    dev = R.default_device()
    s1, st1 = R.collectstates(esn, np.asfortranarray(data[:, :T]), ps, fixed_h0(R, st, esn, [h0]))
    assert ("k1_cluster" if N >= 1024 else "k1_single") in dev.last_kernel()    # from N = 1024 on a sequence is spread over a cluster of 8 SMs
    s2, st2 = R.collectstates(esn, np.asfortranarray(data[:, T:]), ps, st1)      # continues from the stored carry
    states = np.concatenate([s1, s2], axis=1)
    layers = oracle_layers(R, esn, ps)
    for b in range(B):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
        assert state_err(st2.reservoir.carry[0][:, b:b + 1], hs[0].reshape(-1, 1)) <= STATE_RTOL


def test_single_sequence_kernel_in_deep_esn(R):
    rng = np.random.default_rng(4)
    res = [256, 512, 320]
    desn = R.DeepESN(3, res, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.05), leak_coefficient=(1.0, 0.7, 0.5), state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, desn)
    T, B = 64, 2
    data = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
    hs0 = [rng.standard_normal((n, B)).astype(np.float32) for n in res]
    states, st2 = R.collectstates(desn, data, ps, fixed_h0(R, st, desn, hs0))
    assert "k1_single<din=0" in R.default_device().last_kernel()   # the last layer consumed a precomputed projection
    layers = oracle_layers(R, desn, ps)
    for b in range(B):
        ref, hs = O.collectstates(layers, data[:, :, b], [h[:, b] for h in hs0])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
        for l in range(3):
            assert state_err(st2.cells[l].carry[0][:, b:b + 1], hs[l].reshape(-1, 1)) <= STATE_RTOL


@pytest.mark.parametrize("chunk", [16, 37])
def test_deep_esn_layer_pipeline(R, chunk, monkeypatch):
    """One-sequence DeepESN: the layers run as a pipeline over time chunks on separate streams.  Forced to
    small chunks here; the result must equal the layer-by-layer pass bit for bit (same kernels, same arithmetic) and
    match the oracle, including the carries handed from chunk to chunk."""
    rng = np.random.default_rng(12)
    res = [128, 192, 96, 160]
    desn = R.DeepESN(3, res, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.08), leak_coefficient=(1.0, 0.8, 0.6, 0.4),
                     state_modifiers=(R.NLAT2, R.NLAT2, R.NLAT1, (R.NLAT2, R.Pad(1.0))))
    ps, st = R.setup(rng, desn)
    T = 150
    data = np.asfortranarray(rng.standard_normal((3, T)).astype(np.float32))
    hs0 = [rng.standard_normal(n).astype(np.float32) for n in res]
    st0 = fixed_h0(R, st, desn, hs0)
    monkeypatch.setenv("RCB_DEEP_CHUNK", "0")
    plain, st_plain = R.collectstates(desn, data, ps, st0)
    assert "pipeline" not in R.default_device().last_kernel()
    monkeypatch.setenv("RCB_DEEP_CHUNK", str(chunk))
    piped, st_piped = R.collectstates(desn, data, ps, st0)
    assert "layer pipeline" in R.default_device().last_kernel()
    assert np.array_equal(plain, piped)
    for l in range(4):
        assert np.array_equal(st_plain.cells[l].carry[0], st_piped.cells[l].carry[0])
    ref, hs = O.collectstates(oracle_layers(R, desn, ps), data, hs0)
    assert state_err(piped, ref) <= STATE_RTOL
    # train through the pipelined collect
    tgt = np.asfortranarray(rng.standard_normal((3, T)).astype(np.float32))
    ps_a, _ = R.train(desn, data, tgt, ps, st0, objective=R.RidgeRegression(1e-2), washout=10)
    monkeypatch.setenv("RCB_DEEP_CHUNK", "0")
    ps_b, _ = R.train(desn, data, tgt, ps, st0, objective=R.RidgeRegression(1e-2), washout=10)
    assert np.allclose(ps_a.readout.weight, ps_b.readout.weight, rtol=1e-9, atol=1e-12)


def test_plumbing_entry_point_error_paths(R):
    """Argument errors of the SURVEY 8f entry points map onto the reference's exception types (ArgumentError /
    DimensionMismatch) through the same status codes as the core path."""
    import ctypes as C
    from rcb200 import _lib as L

    dev = R.default_device()
    lib = L.lib()
    x = np.zeros((4, 6, 1), np.float32, order="F")
    out = np.zeros((12, 6, 1), np.float32, order="F")
    with pytest.raises(R.ArgumentError, match="num_delays"):
        L.check(lib.rcb_delay_features(dev.handle, L.ptr(x), 0, 4, 6, 1, -1, 1, None, 0, L.ptr(out), None))
    with pytest.raises(R.ArgumentError, match="stride"):
        L.check(lib.rcb_delay_features(dev.handle, L.ptr(x), 0, 4, 6, 1, 2, 0, None, 0, L.ptr(out), None))
    with pytest.raises(R.ArgumentError):
        L.check(lib.rcb_delay_features(dev.handle, None, 0, 4, 6, 1, 2, 1, None, 0, L.ptr(out), None))
    with pytest.raises(R.DimensionMismatch):      # DelayLayer row check, basic.jl:427-431
        R.DelayLayer(3).apply_sequence(np.zeros((4, 5), np.float32), R.DelayLayer(3).initialstates(np.random.default_rng(0)))
    with pytest.raises(R.DimensionMismatch):      # init_delay tuple length, basic.jl:381-388
        R.DelayLayer(3, num_delays=2, init_delay=(R.zeros32,))
    with pytest.raises(R.DimensionMismatch):
        R.extend_features(np.zeros((2, 5), np.float32), np.zeros((3, 6), np.float32))
    with pytest.raises(R.ArgumentError, match="at least one state"):   # conceptors.jl:29
        L.check(lib.rcb_correlation_matrix(dev.handle, L.ptr(x), 0, 4, 0, 4, 0, L.ptr(out)))
    with pytest.raises(R.DimensionMismatch):
        L.check(lib.rcb_correlation_matrix(dev.handle, L.ptr(x), 0, 4, 6, 2, 0, L.ptr(out)))
    rho, dim = C.c_double(), C.c_int32()
    with pytest.raises(R.ArgumentError):
        L.check(lib.rcb_spectral_radius_dense(dev.handle, None, 4, 4, 0.0, 0, C.byref(rho), C.byref(dim)))
    W = np.asfortranarray(np.diag([0.5, -2.0, 1.0, 0.25]).astype(np.float32))
    L.check(lib.rcb_spectral_radius_dense(dev.handle, L.ptr(W), 4, 4, 0.0, 0, C.byref(rho), C.byref(dim)))
    assert abs(rho.value - 2.0) < 1e-12 and dim.value <= 4
    ro = R.LinearReadout(5, 2)
    with pytest.raises(R.DimensionMismatch):      # readout weight vs feature count
        R.readout_apply(ro, R.NT(weight=np.zeros((2, 5))), np.zeros((4, 3), np.float32))
    y = R.readout_apply(R.LinearReadout(4, 2, R.tanh, use_bias=True), R.NT(weight=np.ones((2, 4)), bias=np.array([0.5, -0.5])),
                        np.full((4, 3), 0.25, np.float32))
    assert np.allclose(y, np.tanh(np.array([[1.5], [0.5]])) * np.ones((1, 3)))   # psi(W z + b), basic.jl:171-182


@pytest.mark.parametrize("N,d,mods,f64", [(300, 3, "nlat2", True), (300, 3, "nlat2_pad", False), (1024, 1, "none", True), (65, 3, "pad", True)])
def test_autoregressive_kernel_matches_generic_loop_and_oracle(R, N, d, mods, f64, monkeypatch):
    """The on-chip autoregressive kernel (k4_auto) against the general fused loop (same library, RCB_K4_MODE=generic)
    and against the oracle's step-by-step loop (src/predict.jl:74-88), in the reference's Float64 promotion and in Float32."""
    rng = np.random.default_rng(N + d)
    chain = dict(nlat2=(R.NLAT2,), nlat2_pad=(R.NLAT2, R.Pad(1.0)), none=(), pad=(R.Pad(0.5),))[mods]
    esn = R.ESN(d, N, d, init_reservoir=R.rand_sparse(radius=0.9, sparsity=min(1.0, 6 / N)), leak_coefficient=0.7, state_modifiers=chain)
    ps, st = R.setup(rng, esn)
    T = 400
    tt = np.arange(T + 1)[None, :]
    series = (np.sin(0.07 * tt * (1 + np.arange(d))[:, None]) * 0.6).astype(np.float32)
    data, tgt = np.asfortranarray(series[:, :T]), np.asfortranarray(series[:, 1:])
    obj = R.RidgeRegression(1e-3) if f64 else R.RidgeRegression(np.float32, 1e-3)
    ps2, st2 = R.train(esn, data, tgt, ps, st, objective=obj, washout=50)
    u0 = tgt[:, -1]
    Y, st3 = R.predict(esn, 40, ps2, st2, initialdata=u0)
    assert ("k4_cluster" if N >= 1024 else "k4_auto") in R.default_device().last_kernel() and Y.dtype == (np.float64 if f64 else np.float32)
    monkeypatch.setenv("RCB_K4_MODE", "generic")
    Yg, st3g = R.predict(esn, 40, ps2, st2, initialdata=u0)
    assert "k4_auto" not in R.default_device().last_kernel() and "k4_cluster" not in R.default_device().last_kernel()
    assert state_err(Y, Yg) <= (1e-9 if f64 else 1e-4)
    assert state_err(st3.reservoir.carry[0].reshape(-1, 1), st3g.reservoir.carry[0].reshape(-1, 1)) <= (1e-9 if f64 else 1e-4)
    layers = oracle_layers(R, esn, ps)
    Yo, _ = O.predict_autoregressive(layers, ps2.readout.weight, 40, u0, [st2.reservoir.carry[0]])
    # the reference's very first step runs in eltype(initialdata) = Float32 and only the fed-back Float64 outputs promote
    # the loop (esn_cell.jl:176); the kernels run every step in the promoted type: a one-step Float32 rounding apart
    assert state_err(Y, Yo) <= (1e-5 if f64 else 1e-3)


def test_deep_esn_float64_data(R):
    """Float64 data through a DeepESN: eltype(states) follows the data (esn_deep.jl:288); the inter-layer projection runs
    in Float64 with the Float32 W_in promoted, like the reference's gemv."""
    rng = np.random.default_rng(31)
    res = [40, 56, 33]
    desn = R.DeepESN(2, res, 2, init_reservoir=R.rand_sparse(radius=0.8, sparsity=0.2), leak_coefficient=0.8,
                     state_modifiers=(R.NLAT2, R.NLAT1, (R.NLAT3, R.Pad(1.0))))
    ps, st = R.setup(rng, desn)
    T = 30
    data = np.asfortranarray(rng.standard_normal((2, T, 2)))
    W = rng.standard_normal((2, 34))
    ps2 = ps.merge(readout=ps.readout.merge(weight=W))
    hs0 = [rng.standard_normal((n, 2)) for n in res]
    st0 = fixed_h0(R, st, desn, hs0)
    states, st1 = R.collectstates(desn, data, ps2, st0)
    assert states.dtype == np.float64 and states.shape == (34, T, 2)
    Y, _ = R.predict(desn, data, ps2, st0)
    assert Y.dtype == np.float64
    layers = oracle_layers(R, desn, ps2)
    for b in range(2):
        ref, hs = O.collectstates(layers, data[:, :, b], [h[:, b] for h in hs0])
        assert state_err(states[:, :, b], ref) <= 1e-11
        for l in range(3):
            assert state_err(st1.cells[l].carry[0][:, b:b + 1], hs[l].reshape(-1, 1)) <= 1e-11
        Yo, _ = O.predict_teacher(layers, W, data[:, :, b], [h[:, b] for h in hs0])
        assert state_err(Y[:, :, b], Yo) <= 1e-10



@pytest.mark.parametrize("name,N,mods", [("es2n", 300, "nlat2_pad"), ("resesn", 1024, "esq"), ("es2n", 130, "nlat1_nlat2"), ("resesn", 64, "none")])
def test_dense_operator_cells_cooperative_kernel(R, name, N, mods):
    """ES2N / ResESN, one sequence: the cooperative multi-SM kernel (recurrence_coop.cu: rows of O and W split over a grid,
    one grid barrier per step) against the oracle and against the single-CTA kernel, incl. carry in / out and a second
    call, fused and non-fused modifier chains, bias."""
    import os

    rng = np.random.default_rng(N)
    chain = dict(nlat2_pad=(R.NLAT2, R.Pad(1.0)), esq=(R.ExtendedSquare,), nlat1_nlat2=(R.NLAT1, R.NLAT2), none=())[mods]
    kw = dict(init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), use_bias=True, init_bias=R.randn32, state_modifiers=chain)
    rc = R.ES2N(3, N, 3, proximity=0.3, **kw) if name == "es2n" else R.ResESN(3, N, 3, alpha=0.7, beta=0.5, **kw)
    ps, st = R.setup(rng, rc)
    T = 50
    data = np.asfortranarray(rng.standard_normal((3, 2 * T)).astype(np.float32))
    h0 = rng.standard_normal(N).astype(np.float32)
    dev = R.default_device()
    s1, st1 = R.collectstates(rc, np.asfortranarray(data[:, :T]), ps, fixed_h0(R, st, rc, [h0]))
    assert "k1_coop" in dev.last_kernel()
    s2, st2 = R.collectstates(rc, np.asfortranarray(data[:, T:]), ps, st1)
    states = np.concatenate([s1, s2], axis=1)
    ref, hs = O.collectstates(oracle_layers(R, rc, ps), data, [h0])
    assert states.shape == ref.shape and state_err(states, ref) <= STATE_RTOL
    assert state_err(st2.reservoir.carry[0].reshape(-1, 1), hs[0].reshape(-1, 1)) <= STATE_RTOL
    os.environ["RCB_K1_MODE"] = "v1"   # the single-CTA path
    try:
        sg, _ = R.collectstates(rc, data, ps, fixed_h0(R, st, rc, [h0]))
        assert "k1_coop" not in dev.last_kernel()
    finally:
        del os.environ["RCB_K1_MODE"]
    assert state_err(states, sg) <= STATE_RTOL


@pytest.mark.parametrize("name,N,B,mods", [("es2n", 300, 5, "nlat2_pad"), ("resesn", 1024, 37, "esq"), ("es2n", 130, 70, "nlat1_nlat2"),
                                           ("resesn", 520, 150, "none"), ("es2n", 257, 2, "nlat2_pad")])
def test_dense_operator_cells_batched_cooperative_kernel(R, name, N, B, mods):
    """ES2N / ResESN in batches: the 2-D cooperative kernel (rows x sequences, each CTA's slice of O resident in shared memory,
    the states exchanged through a global double buffer with one grid barrier per step) against the oracle and against the
    one-CTA-per-sequence-tile kernel, incl. carry in / out over two calls, modifier chains, bias."""
    import os

    rng = np.random.default_rng(N + B)
    chain = dict(nlat2_pad=(R.NLAT2, R.Pad(1.0)), esq=(R.ExtendedSquare,), nlat1_nlat2=(R.NLAT1, R.NLAT2), none=())[mods]
    kw = dict(init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), use_bias=True, init_bias=R.randn32, state_modifiers=chain)
    rc = R.ES2N(3, N, 3, proximity=0.3, **kw) if name == "es2n" else R.ResESN(3, N, 3, alpha=0.7, beta=0.5, **kw)
    ps, st = R.setup(rng, rc)
    T = 40
    data = np.asfortranarray(rng.standard_normal((3, 2 * T, B)).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    dev = R.default_device()
    os.environ["RCB_K1_MODE"] = "coop"   # (by default the kernel is only taken in the shape range where it was measured to win)
    try:
        s1, st1 = R.collectstates(rc, np.asfortranarray(data[:, :T]), ps, fixed_h0(R, st, rc, [h0]))
        assert "k1_coop_batched" in dev.last_kernel()
        s2, st2 = R.collectstates(rc, np.asfortranarray(data[:, T:]), ps, st1)
    finally:
        del os.environ["RCB_K1_MODE"]
    states = np.concatenate([s1, s2], axis=1)
    layers = oracle_layers(R, rc, ps)
    for b in sorted({0, 1, B // 2, B - 1}):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert states[:, :, b].shape == ref.shape and state_err(states[:, :, b], ref) <= STATE_RTOL, b
        assert state_err(st2.reservoir.carry[0][:, b:b + 1], hs[0].reshape(-1, 1)) <= STATE_RTOL
    os.environ["RCB_K1_MODE"] = "v1"   # one CTA per tile of sequences
    try:
        sg, _ = R.collectstates(rc, data, ps, fixed_h0(R, st, rc, [h0]))
        assert "k1_coop" not in dev.last_kernel()
    finally:
        del os.environ["RCB_K1_MODE"]
    assert state_err(states, sg) <= STATE_RTOL


# ---- ReservoirChain with one cell: Extend, Collect points, delay lines --------------------------------
def _chain_cell(R, N, d_in, rng):
    W = R.rand_sparse(radius=0.9, sparsity=6 / N)(rng, N, N)
    Win = R.scaled_rand(rng, N, d_in)
    cell = R.ESNCell(d_in, N, R.tanh, init_reservoir=const_init(W), init_input=const_init(Win), leak_coefficient=0.8)
    return cell, O.ESNCellParams(input_matrix=Win, reservoir_matrix=W, leak=0.8, act=O.ACT_TANH)


def test_reservoir_chain_extend_example(R):  # the docs example of Extend, src/states.jl:113-124
    rng = np.random.default_rng(3)
    N, T = 300, 80
    cell, ocell = _chain_cell(R, N, 3, rng)
    rc = R.ReservoirChain(R.Extend(R.StatefulLayer(cell)), R.NLAT2, R.LinearReadout(N + 3, 3))
    ps, st = R.setup(rng, rc)
    assert list(ps.keys()) == ["layer_1", "layer_2", "layer_3", "layer_4"] and "op" in ps.layer_1   # Collect() inserted before the readout
    series = O.lorenz_series(T + 20).astype(np.float32) / 20
    data, tgt = np.asfortranarray(series[:, :T]), np.asfortranarray(series[:, 1:T + 1])
    h0 = rng.standard_normal(N).astype(np.float32)
    st = st.merge(layer_1=NT_with_carry(R, st.layer_1, h0, extend=True))
    states, st2 = R.collectstates(rc, data, ps, st)
    raw, hs = O.collectstates([ocell], data, [h0])
    ref = O.nlat2(O.extend(data, raw))                        # the chain order: vcat(input, state), then NLAT2 on all N+3 rows
    assert states.shape == (N + 3, T) and state_err(states, ref) <= STATE_RTOL
    ps2, st3 = R.train(rc, data, tgt, ps, st, objective=R.RidgeRegression(1e-4), washout=10)
    Wo = O.train_ridge_qr(states[:, 10:], tgt[:, 10:], 1e-4)
    assert ps2.layer_4.weight.shape == (3, N + 3) and np.linalg.norm(ps2.layer_4.weight - Wo) / np.linalg.norm(Wo) <= WEIGHT_RTOL
    test = np.asfortranarray(series[:, T:T + 15])
    Y, _ = R.predict(rc, test, ps2, st3)
    rawt, _ = O.collectstates([ocell], test, hs)
    assert state_err(Y, O.linear_readout(ps2.layer_4.weight, O.nlat2(O.extend(test, rawt)).astype(np.float64))) <= 1e-4
    Ya, _ = R.predict(rc, 8, ps2, st3, initialdata=test[:, 0])
    cur, h, outs = test[:, 0], hs[0], []
    for _ in range(8):
        h = O.esn_cell_step(ocell, cur.astype(np.float32), h.astype(np.float32))
        cur = O.linear_readout(ps2.layer_4.weight, O.nlat2(O.extend(cur.astype(np.float32), h)).astype(np.float64))
        outs.append(cur)
    assert state_err(Ya, np.stack(outs, axis=1)) <= 1e-3


def NT_with_carry(R, cell_st, h0, extend=False):
    inner = cell_st.op if extend else cell_st
    inner = inner.merge(carry=(np.asarray(h0),))
    return R.NT(op=inner) if extend else inner


def test_reservoir_chain_collect_points_and_delay(R):  # multiple Collect layers are vcat-ed in chain order, basic.jl:226-235,283-308
    rng = np.random.default_rng(9)
    N, T = 64, 40
    cell, ocell = _chain_cell(R, N, 2, rng)
    data = np.asfortranarray(rng.standard_normal((2, T)).astype(np.float32))
    h0 = rng.standard_normal(N).astype(np.float32)
    raw, _ = O.collectstates([ocell], data, [h0])
    # (1) explicit Collect after the cell + the readout's implicit one after NLAT2: features = [raw; NLAT2(raw)]
    rc = R.ReservoirChain(R.StatefulLayer(cell), R.Collect(), R.NLAT2, R.LinearReadout(2 * N, 2))
    ps, st = R.setup(rng, rc)
    st = st.merge(layer_1=NT_with_carry(R, st.layer_1, h0))
    states, _ = R.collectstates(rc, data, ps, st)
    assert state_err(states, np.concatenate([raw, O.nlat2(raw)], axis=0)) <= STATE_RTOL
    # (2) a delay line on the input, one on the state, a modifier after it, explicit Collect only
    rc = R.ReservoirChain(R.DelayLayer(2, num_delays=1), R.StatefulLayer(R.ESNCell(4, N, R.tanh, leak_coefficient=0.8,
                          init_reservoir=const_init(ocell.reservoir_matrix), init_input=R.scaled_rand)),
                          R.DelayLayer(N, num_delays=2, stride=2), R.NLAT1, R.Collect(), R.LinearReadout(3 * N, 2, include_collect=False))
    ps, st = R.setup(rng, rc)
    st = st.merge(layer_2=NT_with_carry(R, st.layer_2, h0))
    states, st2 = R.collectstates(rc, data, ps, st)
    ocell2 = O.ESNCellParams(input_matrix=np.asarray(ps.layer_2.input_matrix), reservoir_matrix=ocell.reservoir_matrix, leak=0.8, act=O.ACT_TANH)
    ud, _, _ = O.delay_layer_sequence(data, 1, 1)
    raw2, _ = O.collectstates([ocell2], ud, [h0])
    zd, _, _ = O.delay_layer_sequence(raw2, 2, 2)
    assert states.shape == (3 * N, T) and state_err(states, O.nlat1(zd)) <= STATE_RTOL
    assert st2.layer_1.clock == T and st2.layer_3.clock == T
    with pytest.raises(R.UnsupportedError):
        R.ReservoirChain(R.StatefulLayer(cell), R.StatefulLayer(cell), R.LinearReadout(N, 2))


@pytest.mark.parametrize("budget_mb,B", [(32, 900), (2, 40)])
def test_train_chunked_paths_match_unchunked(R, budget_mb, B, monkeypatch):
    """rcb_train cuts work that exceeds its state-buffer budget: in time for batches far wider than the machine (all
    sequences stay together, the carry links the windows), by sequences otherwise.  Both must reproduce the one-piece fit,
    including the washout that falls inside the first window and the final carry."""
    rng = np.random.default_rng(B)
    N, T, w = 256, 100, 37
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N), leak_coefficient=0.6, state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, esn)
    data = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
    tgt = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    st0 = fixed_h0(R, st, esn, [h0])
    ps_a, st_a = R.train(esn, data, tgt, ps, st0, objective=R.RidgeRegression(1e-2), washout=w)
    monkeypatch.setenv("RCB_TRAIN_BUDGET_MB", str(budget_mb))
    ps_b, st_b = R.train(esn, data, tgt, ps, st0, objective=R.RidgeRegression(1e-2), washout=w)
    assert np.linalg.norm(ps_a.readout.weight - ps_b.readout.weight) <= 1e-9 * np.linalg.norm(ps_a.readout.weight)
    assert np.array_equal(st_a.reservoir.carry[0], st_b.reservoir.carry[0])


# ---- single model call (rc)(inp, ps, st): src/reservoircomputer.jl:81-94 -------------------------------------------------
@pytest.mark.parametrize("batch", [None, 5])
def test_single_step_call_matches_a_loop_of_reference_steps(R, batch):
    """`apply` = cell step + modifiers + readout on the GPU (rcb_step); a Python loop of such calls threads the carry through
    `st` exactly like the reference's collectstates / teacher-forced predict loops do (reservoircomputer.jl:113-133)."""
    rng = np.random.default_rng(90)
    N, T = 120, 12
    esn = R.ESN(3, N, 2, init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.1), leak_coefficient=0.7, use_bias=True,
                state_modifiers=(R.NLAT2, R.Pad(1.0)))
    ps, st = R.setup(rng, esn)
    Wro = rng.standard_normal((2, N + 1))
    ps, st = R.addreadout(esn, Wro, ps, st)
    B = batch or 1
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    st = fixed_h0(R, st, esn, [h0 if batch else h0[:, 0]])
    data = rng.standard_normal((3, T, B)).astype(np.float32)
    layers = oracle_layers(R, esn, ps)
    ys, zs = [], []
    s = st
    for t in range(T):
        x = data[:, t, :] if batch else data[:, t, 0]
        z, _ = R.partial_apply(esn, x, ps, s)            # does not advance `s` here: the next call re-runs the same step
        y, s = R.apply(esn, x, ps, s)
        ys.append(y); zs.append(z)
    for b in range(B):
        so, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        zb = np.stack([z[:, b] if batch else z for z in zs], axis=1)
        yb = np.stack([y[:, b] if batch else y for y in ys], axis=1)
        assert state_err(zb, so) <= 1e-5
        assert np.max(np.abs(yb - Wro @ so.astype(np.float64))) <= 1e-5 * np.max(np.abs(yb))
        carry = s.reservoir.carry[0]
        assert np.max(np.abs((carry[:, b] if batch else carry) - hs[0])) <= 1e-5 * np.max(np.abs(hs[0]))


# ---- K1 cluster: one sequence spread over a thread-block cluster, state exchanged through distributed shared memory ------
@pytest.mark.parametrize("N,d_in,mod,B,T,csize", [(64, 3, "nlat2", 1, 40, 2), (300, 1, "none", 3, 33, 4), (1100, 3, "nlat2_pad", 2, 25, 8), (2048, 3, "nlat3", 1, 20, 4),
                                                   (4096, 1, "extsq", 2, 15, 8), (8192, 3, "nlat2", 1, 12, 8), (517, 2, "psq", 4, 30, 2)])
def test_cluster_kernel_parity(R, monkeypatch, N, d_in, mod, B, T, csize):
    """recurrence_cluster.cu forced at every size (by default it takes N >= 4096): rows not a multiple of the cluster / warp
    size, every fused modifier, cell bias and per-neuron leak, several sequences (one cluster each), the carry and a second
    call that continues from it."""
    monkeypatch.setenv("RCB_K1_MODE", "cluster")
    monkeypatch.setenv("RCB_CLUSTER_SIZE", str(csize))
    rng = np.random.default_rng(N + 11 * B)
    mods = dict(nlat2=(R.NLAT2,), none=(), nlat2_pad=(R.NLAT2, R.Pad(1.0)), nlat3=(R.NLAT3,), extsq=(R.ExtendedSquare,), psq=(R.PartialSquare(0.4),))[mod]
    esn = R.ESN(d_in, N, d_in, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / N, return_sparse=N > 1500), use_bias=True,
                init_bias=lambda r, *d: (0.1 * r.standard_normal(d)).astype(np.float32),
                leak_coefficient=rng.uniform(0.3, 1.0, N).astype(np.float32) if N % 2 else 0.6, state_modifiers=mods)
    ps, st = R.setup(rng, esn)
    data = np.asfortranarray(rng.standard_normal((d_in, 2 * T, B)).astype(np.float32))
    h0 = rng.standard_normal((N, B)).astype(np.float32)
    dev = R.default_device()
    s1, st1 = R.collectstates(esn, np.asfortranarray(data[:, :T]), ps, fixed_h0(R, st, esn, [h0]))
    assert "k1_cluster" in dev.last_kernel()
    s2, st2 = R.collectstates(esn, np.asfortranarray(data[:, T:]), ps, st1)
    states = np.concatenate([s1, s2], axis=1)
    layers = oracle_layers(R, esn, ps)
    for b in range(B):
        ref, hs = O.collectstates(layers, data[:, :, b], [h0[:, b]])
        assert state_err(states[:, :, b], ref) <= STATE_RTOL
        assert np.max(np.abs(st2.reservoir.carry[0][:, b] - hs[0])) <= STATE_RTOL * np.max(np.abs(hs[0]))


def test_cluster_kernel_deep_layers(R, monkeypatch):
    """DeepESN through the cluster kernel: layers >= 2 read their precomputed input projection (stored in the position order
    of the layer's schedule) through the row -> position map."""
    monkeypatch.setenv("RCB_K1_MODE", "cluster")
    rng = np.random.default_rng(77)
    sizes = [160, 96, 224]
    esn = R.DeepESN(3, sizes, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=0.08), leak_coefficient=(1.0, 0.7, 0.5),
                    state_modifiers=(R.NLAT2, R.NLAT1, R.Pad(1.0)))
    ps, st = R.setup(rng, esn)
    T = 60
    data = np.asfortranarray(rng.standard_normal((3, T)).astype(np.float32))
    h0 = [rng.standard_normal(n).astype(np.float32) for n in sizes]
    states, st2 = R.collectstates(esn, data, ps, fixed_h0(R, st, esn, h0))
    ref, hs = O.collectstates(oracle_layers(R, esn, ps), data, h0)
    assert state_err(states, ref) <= STATE_RTOL
    for l in range(3):
        assert np.max(np.abs(st2.cells[l].carry[0] - hs[l])) <= STATE_RTOL * np.max(np.abs(hs[l]))


@pytest.mark.parametrize("N,d,mods,f64", [(64, 3, "nlat2", True), (300, 3, "nlat2_pad", False), (517, 2, "none", True), (1100, 1, "pad", True),
                                           (2048, 3, "nlat2", True), (4096, 3, "nlat2", False)])
def test_autoregressive_cluster_kernel(R, N, d, mods, f64, monkeypatch):
    """k4_cluster (recurrence_cluster.cu): the autoregressive loop with the sequence spread over 8 SMs — forced at every size —
    against the one-CTA / general loops of the same library and against the oracle's step-by-step loop (src/predict.jl:74-88)."""
    rng = np.random.default_rng(N + d + 5)
    chain = dict(nlat2=(R.NLAT2,), nlat2_pad=(R.NLAT2, R.Pad(1.0)), none=(), pad=(R.Pad(0.5),))[mods]
    esn = R.ESN(d, N, d, init_reservoir=R.rand_sparse(radius=0.9, sparsity=min(1.0, 6 / N), return_sparse=N > 1500), leak_coefficient=0.7,
                state_modifiers=chain)
    ps, st = R.setup(rng, esn)
    T = 300
    tt = np.arange(T + 1)[None, :]
    series = (np.sin(0.07 * tt * (1 + np.arange(d))[:, None]) * 0.6).astype(np.float32)
    data, tgt = np.asfortranarray(series[:, :T]), np.asfortranarray(series[:, 1:])
    obj = R.RidgeRegression(1e-3) if f64 else R.RidgeRegression(np.float32, 1e-3)
    ps2, st2 = R.train(esn, data, tgt, ps, st, objective=obj, washout=50)
    u0 = tgt[:, -1]
    monkeypatch.setenv("RCB_K4_MODE", "cluster")
    Y, st3 = R.predict(esn, 30, ps2, st2, initialdata=u0)
    assert "k4_cluster" in R.default_device().last_kernel() and Y.dtype == (np.float64 if f64 else np.float32)
    monkeypatch.setenv("RCB_K4_MODE", "generic")
    Yg, st3g = R.predict(esn, 30, ps2, st2, initialdata=u0)
    assert "k4_cluster" not in R.default_device().last_kernel()
    assert state_err(Y, Yg) <= (1e-9 if f64 else 1e-4)
    assert state_err(st3.reservoir.carry[0].reshape(-1, 1), st3g.reservoir.carry[0].reshape(-1, 1)) <= (1e-9 if f64 else 1e-4)
    if N <= 1100:
        Yo, _ = O.predict_autoregressive(oracle_layers(R, esn, ps), ps2.readout.weight, 30, u0, [st2.reservoir.carry[0]])
        assert state_err(Y, Yo) <= (1e-5 if f64 else 1e-3)
