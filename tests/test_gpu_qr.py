"""GPU parity of the QR-grade ridge path (csrc/qr_solve.cu) against the oracle's Householder QR of the augmented system
(src/train.jl:108-147,182) — the reference's default objective is RidgeRegression(0.0) (train.jl:234), where the
Cholesky-on-Gram path squares the condition number."""
import os

import numpy as np
import pytest

from oracle import esn_oracle as O
from test_gpu_parity import R, c1, fixed_h0, oracle_layers, state_err  # noqa: F401  (fixtures + helpers)

pytestmark = pytest.mark.gpu

WEIGHT_RTOL = 1e-4


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - b) / np.linalg.norm(b))


def _illcond(rng, F, S, cond, dtype=np.float64):
    """F x S states with singular values log-spaced over `cond`."""
    U, _ = np.linalg.qr(rng.standard_normal((F, F)))
    V, _ = np.linalg.qr(rng.standard_normal((S, F)))
    s = np.logspace(0, -np.log10(cond), F)
    return np.asfortranarray((U * s) @ V.T * np.sqrt(S)).astype(dtype)


@pytest.mark.parametrize("F,S,d_out", [(5, 40, 1), (33, 100, 2), (64, 1000, 3), (300, 5000, 3), (517, 1300, 2)])
@pytest.mark.parametrize("lam", [0.0, 1e-8, 1e-2])
def test_qr_fit_matches_oracle_qr(R, F, S, d_out, lam):
    rng = np.random.default_rng(F + S)
    Z = np.asfortranarray(rng.standard_normal((F, S)))
    Y = np.asfortranarray(rng.standard_normal((d_out, F)) @ Z + 0.01 * rng.standard_normal((d_out, S)))
    W = R.fit_readout(R.RidgeRegression(lam), Z, Y, solver=R.QRFactorization())
    Wref = O.train_ridge_qr(Z, Y, lam)
    assert W.shape == (d_out, F) and W.dtype == np.float64
    assert rel(W, Wref) <= 1e-10


@pytest.mark.parametrize("cond", [1e6, 1e9, 1e11])
def test_qr_keeps_cond_not_cond_squared(R, cond):
    """lambda = 0 on ill-conditioned features: the normal equations lose cond^2 * eps, Householder QR cond * eps.
    AUTO must take the QR path for lambda == 0 (and agree with the explicit solver)."""
    rng = np.random.default_rng(int(np.log10(cond)))
    F, S = 48, 600
    Z = _illcond(rng, F, S, cond)
    Wtrue = rng.standard_normal((2, F))
    Y = np.asfortranarray(Wtrue @ Z)
    Wref = O.train_ridge_qr(Z, Y, 0.0)
    Wq = R.fit_readout(R.RidgeRegression(0.0), Z, Y, solver=R.QRSolver())
    Wa = R.fit_readout(R.RidgeRegression(0.0), Z, Y)
    # both QRs carry an error ~ cond * eps of their own: compare the fits through the residual and the well-determined part
    tol = max(1e-10, cond * 1e-15)
    assert rel(Wq, Wref) <= tol
    assert np.array_equal(Wq, Wa)
    assert np.linalg.norm(Wq @ Z - Y) / np.linalg.norm(Y) <= 1e-9


def test_qr_auto_escalates_when_cholesky_fails(R):
    """A Gram matrix that is not numerically positive definite at a tiny lambda: AUTO retries by QR instead of raising;
    the pinned normal-equation solver raises the reference's ArgumentError class."""
    rng = np.random.default_rng(3)
    F, S = 40, 500
    Z = _illcond(rng, F, S, 1e12)  # cond(G) = 1e24: a dozen pivots sit below the rounding noise of the factorisation
    Y = np.asfortranarray(rng.standard_normal((1, F)) @ Z)
    lam = 1e-26
    with pytest.raises(R.ArgumentError):
        R.fit_readout(R.RidgeRegression(lam), Z, Y, solver=R.NormalCholeskyFactorization())
    W = R.fit_readout(R.RidgeRegression(lam), Z, Y)  # AUTO
    assert np.all(np.isfinite(W))
    assert np.linalg.norm(W @ Z - Y) / np.linalg.norm(Y) <= 1e-6


def test_qr_rank_deficient_raises(R):
    """Exactly rank-deficient features at lambda = 0: R gets a zero diagonal entry -> the solver-failure class
    (train.jl:194-196), not a silent garbage weight."""
    Z = np.zeros((4, 20))
    Z[0] = np.arange(20.0)
    Z[1] = 1.0
    Z = np.asfortranarray(Z)  # rows 2, 3 are identically zero
    Y = np.asfortranarray(Z[:1] * 2.0)
    with pytest.raises(R.ArgumentError):
        R.fit_readout(R.RidgeRegression(0.0), Z, Y, solver=R.QRFactorization())
    W = R.fit_readout(R.RidgeRegression(1e-3), Z, Y, solver=R.QRFactorization())  # the lambda rows restore full rank
    assert rel(W, O.train_ridge_qr(Z, Y, 1e-3)) <= 1e-10


def test_qr_chunked_equals_one_shot(R, monkeypatch):
    """Streaming over sample chunks (triangular-pentagonal steps) is a QR of the same matrix."""
    rng = np.random.default_rng(9)
    F, S = 96, 3000
    Z = np.asfortranarray(rng.standard_normal((F, S)).astype(np.float32))
    Y = np.asfortranarray(rng.standard_normal((3, S)).astype(np.float32))
    W1 = R.fit_readout(R.RidgeRegression(1e-6), Z, Y, solver=R.QRFactorization())
    monkeypatch.setenv("RCB_QR_CHUNK", "700")
    W2 = R.fit_readout(R.RidgeRegression(1e-6), Z, Y, solver=R.QRFactorization())
    Wref = O.train_ridge_qr(Z, Y, 1e-6)
    assert rel(W1, Wref) <= 1e-10 and rel(W2, Wref) <= 1e-10
    assert W1.dtype == np.float64  # Float32 data promote with the Float64 lambda (train.jl:126-129)


def test_qr_partial_factors_merge(R):
    """Two ranks' partial factors merged by rcb_qr_merge solve the pooled problem (the QR analogue of summing [G | C])."""
    rng = np.random.default_rng(4)
    F, d_out = 70, 2
    Za, Zb = (np.asfortranarray(rng.standard_normal((F, n))) for n in (400, 650))
    Ya, Yb = (np.asfortranarray(rng.standard_normal((d_out, n))) for n in (400, 650))
    RQa = R.gram(Za, Ya, gram_mode=R.FIT_QR)
    RQb = R.gram(Zb, Yb, gram_mode=R.FIT_QR)
    assert np.allclose(np.tril(RQa[:, :F], -1), 0.0)  # R is upper triangular
    R.qr_merge(RQa, RQb, d_out)
    for lam in (0.0, 1e-3):
        W = R.qr_solve(RQa, d_out, lam)
        Wref = O.train_ridge_qr(np.hstack([Za, Zb]), np.hstack([Ya, Yb]), lam)
        assert rel(W, Wref) <= 1e-10
    # R'R is the Gram matrix
    G = R.gram(np.asfortranarray(np.hstack([Za, Zb])), np.asfortranarray(np.hstack([Ya, Yb])), gram_mode=R.GRAM_FP64)
    Rf = np.triu(RQa[:, :F])
    assert rel(np.tril(Rf.T @ Rf), np.tril(G[:, :F])) <= 1e-12


# ---- BASELINE config 1 exactly as the README calls it: train(esn, X, Y, ps, st) -> lambda = 0 (README.md:114) ----------
def test_c1_readme_default_lambda_zero(R, c1):
    (ps2, st2), sw = R.train(c1["esn"], c1["data"], c1["target"], c1["ps"], c1["st"], return_states=True)
    W = ps2.readout.weight
    assert W.shape == (3, 300) and W.dtype == np.float64
    # states parity first (they feed the fit)
    so, hs = O.collectstates(oracle_layers(R, c1["esn"], c1["ps"]), c1["data"], [c1["h0"]])
    assert state_err(sw, so) <= 1e-5
    # solver parity on the same states: the reference's QR (the same algorithm class: error ~ cond * eps)
    Wref = O.train_ridge_qr(sw, c1["target"], 0.0)
    assert rel(W, Wref) <= WEIGHT_RTOL
    # the fit itself: the training residual of W equals that of the oracle's weights on the oracle's states
    r_gpu = np.linalg.norm(W @ so.astype(np.float64) - c1["target"]) / np.linalg.norm(c1["target"])
    Wo = O.train_ridge_qr(so, c1["target"], 0.0)
    r_ref = np.linalg.norm(Wo @ so.astype(np.float64) - c1["target"]) / np.linalg.norm(c1["target"])
    assert r_gpu <= 1.05 * r_ref + 1e-9
    # ... and the README's next line: predict(esn, 1250, ps, st; initialdata = test[:, 1])
    Y, _ = R.predict(c1["esn"], 40, ps2, st2, initialdata=c1["test"][:, 0])
    Yo, _ = O.predict_autoregressive(oracle_layers(R, c1["esn"], c1["ps"]), W, 40, c1["test"][:, 0], [st2.reservoir.carry[0]])
    assert state_err(Y, Yo) <= 1e-3


@pytest.mark.parametrize("lam", [1e-6, 1e-2])
def test_c1_weights_end_to_end_qr_grade(R, c1, lam):
    """VERDICT r01 'weak 2': W against the QR oracle at lambda = 1e-6 without the `if lam >= 1e-2` escape — on the same
    states for every solver route (AUTO = Cholesky, explicit QR)."""
    for solver in (None, R.QRFactorization()):
        (ps2, _), sw = R.train(c1["esn"], c1["data"], c1["target"], c1["ps"], c1["st"], objective=R.RidgeRegression(lam),
                               solver=solver, washout=100, return_states=True)
        Wref = O.train_ridge_qr(sw, c1["target"][:, 100:], lam)
        assert rel(ps2.readout.weight, Wref) <= WEIGHT_RTOL, (solver, lam)


def test_train_members_qr_fallback(R):
    """Ensemble sweep with lambda = 0 among the regularisations: those systems go through QR (per member one factor,
    every lambda absorbs its own rows), the others through the batched Cholesky — all against the QR oracle."""
    rng = np.random.default_rng(21)
    esn = R.ESN(3, 64, 3, init_reservoir=R.rand_sparse(radius=0.9, sparsity=6 / 64), state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, esn)
    T, B = 400, 3
    series = O.lorenz_series(T + 1)
    X = np.asfortranarray(series[:, :T].astype(np.float32))
    Y = np.asfortranarray(series[:, 1:T + 1].astype(np.float32))
    scale = np.array([0.8, 1.0, 1.1], np.float32)
    leak = np.array([1.0, 0.7, 0.4], np.float32)
    lambdas = np.array([0.0, 1e-6, 1e-2])
    h0 = rng.standard_normal((64, B)).astype(np.float32)
    st = fixed_h0(R, st, esn, [h0])
    for mode in (R.GRAM_AUTO, R.FIT_QR):
        W, hT = R.train_ensemble(esn, X, Y, ps, st, w_scale=scale, leak=leak, lambdas=lambdas, washout=50, gram_mode=mode)
        for b in range(B):
            layer = O.ESNCellParams(input_matrix=ps.reservoir.input_matrix, reservoir_matrix=ps.reservoir.reservoir_matrix * scale[b],
                                    leak=float(leak[b]), act=O.ACT_TANH, modifiers=((O.MOD_NLAT2, 0.0),))
            so, _ = O.collectstates([layer], X, [h0[:, b]])
            for li, lam in enumerate(lambdas):
                Wref = O.train_ridge_qr(so[:, 50:], Y[:, 50:], lam)
                pred, pred_o = W[b, li] @ so[:, 50:].astype(np.float64), Wref @ so[:, 50:].astype(np.float64)
                assert np.linalg.norm(pred - pred_o) / np.linalg.norm(pred_o) <= 1e-4, (mode, b, lam)


# ---- iterative refinement of the Cholesky-on-Gram solution (corrected semi-normal equations) ----------------
@pytest.fixture(scope="module")
def esn_states(R):
    """Real reservoir features: N = 512 NLAT2 states of a Lorenz-driven ESN (strongly correlated, cond(G) ~ 1e12)."""
    rng = np.random.default_rng(512)
    N, T, w = 512, 2100, 100
    esn = R.ESN(3, N, 3, init_reservoir=R.rand_sparse(radius=1.0, sparsity=6 / N), state_modifiers=R.NLAT2)
    ps, st = R.setup(rng, esn)
    series = O.lorenz_series(T + 1)
    data = np.asfortranarray(series[:, :T].astype(np.float32))
    target = np.asfortranarray(series[:, 1:T + 1].astype(np.float32))
    states, _ = R.collectstates(esn, data, ps, fixed_h0(R, st, esn, [np.zeros(N, np.float32)]))
    return dict(esn=esn, ps=ps, st=st, data=data, target=target, Z=np.asfortranarray(states[:, w:]), Y=np.asfortranarray(target[:, w:]), w=w)


@pytest.mark.parametrize("lam", [1e-2, 1e-4, 1e-6])
def test_auto_refines_tensor_gram_to_qr_grade_weights(R, esn_states, lam, monkeypatch):
    """The tensor-core Gram is the Gram of states rounded to 22-bit fixed point: its raw Cholesky solution is off by
    ~2^-23 / (lambda per sample) (8.8e-3 at lambda = 1e-2, S = 2000).  AUTO corrects it with exact residuals: the WEIGHTS then
    meet the 1e-4 tolerance against the oracle's Float64 QR on the same states."""
    Z, Y = esn_states["Z"], esn_states["Y"]
    Wref = O.train_ridge_qr(Z, Y, lam)
    monkeypatch.setenv("RCB_GRAM", "tensor")          # force AUTO onto the tensor path at this (small) size
    W = R.fit_readout(R.RidgeRegression(lam), Z, Y)
    steps, last = R.default_device().last_refine()
    assert steps >= 1 and last <= 1e-3
    assert rel(W, Wref) <= WEIGHT_RTOL, (lam, steps, last)
    monkeypatch.setenv("RCB_REFINE", "0")             # the raw solution, for the record: outside the tolerance
    W0 = R.fit_readout(R.RidgeRegression(lam), Z, Y)
    assert R.default_device().last_refine()[0] == 0
    assert rel(W0, Wref) > rel(W, Wref)
    if lam <= 1e-2:
        assert rel(W0, Wref) > WEIGHT_RTOL


@pytest.mark.parametrize("lam", [1e-6, 1e-8])
def test_auto_refines_fp64_cholesky_at_small_lambda(R, esn_states, lam, monkeypatch):
    """Cholesky on the exact Gram loses cond(G + lambda I) * eps; one or two corrections on the exact normal equations
    bring the weights to QR grade (error ~ cond(design) * eps)."""
    Z, Y = esn_states["Z"], esn_states["Y"]
    Wref = O.train_ridge_qr(Z, Y, lam)
    monkeypatch.setenv("RCB_GRAM", "fp64")
    W = R.fit_readout(R.RidgeRegression(lam), Z, Y)
    assert R.default_device().last_refine()[0] >= 1
    assert rel(W, Wref) <= WEIGHT_RTOL, lam


def test_train_refines_end_to_end(R, esn_states, monkeypatch):
    """train() (fused: collect -> Gram -> solve -> refinement passes that re-run the recurrence) against the oracle's QR on
    the oracle-equal states; windows forced so that the residual pass walks several chunks."""
    e = esn_states
    lam = 1e-3
    st0 = fixed_h0(R, e["st"], e["esn"], [np.zeros(512, np.float32)])
    Wref = O.train_ridge_qr(e["Z"], e["Y"], lam)
    monkeypatch.setenv("RCB_GRAM", "tensor")
    for budget in (None, "3"):                        # 3 MB: the time-chunked accumulation path (1.5k-step windows)
        if budget:
            monkeypatch.setenv("RCB_TRAIN_BUDGET_MB", budget)
        ps2, _ = R.train(e["esn"], e["data"], e["target"], e["ps"], st0, objective=R.RidgeRegression(lam), washout=e["w"])
        assert R.default_device().last_refine()[0] >= 1
        assert rel(ps2.readout.weight, Wref) <= WEIGHT_RTOL, budget


@pytest.mark.parametrize("F,S,d_out,dt", [(300, 1999, 3, np.float32), (256, 700, 1, np.float32), (1028, 4100, 5, np.float32),
                                          (384, 1500, 2, np.float64), (512, 33, 8, np.float32), (302, 900, 3, np.float32)])
def test_refinement_residual_ring_matches_register_kernel(R, F, S, d_out, dt, monkeypatch):
    """The exact residual E = Y - W Z has two kernels: sample tiles staged through a shared-memory ring by bulk copies (16-byte
    aligned slices: F * sizeof % 16 == 0) and the register version (everything else, here F = 302).  Same products, FP64
    accumulation: the refined weights agree to rounding, and both meet the QR oracle."""
    rng = np.random.default_rng(F + S)
    Z = np.asfortranarray(np.tanh(rng.standard_normal((F, 24)) @ rng.standard_normal((24, S)) * 0.4 + 0.05 * rng.standard_normal((F, S))).astype(dt))
    Y = np.asfortranarray((rng.standard_normal((d_out, F)) @ Z / np.sqrt(F)).astype(dt))
    lam = 1e-5 * S
    monkeypatch.setenv("RCB_GRAM", "fp64")
    out = {}
    for ring in ("1", "0"):
        monkeypatch.setenv("RCB_RESID_RING", ring)
        out[ring] = R.fit_readout(R.RidgeRegression(lam), Z, Y)
        assert R.default_device().last_refine()[0] >= 1
    assert rel(out["1"], out["0"]) <= 1e-9
    assert rel(out["1"], O.train_ridge_qr(Z, Y, lam)) <= WEIGHT_RTOL
