"""Pin the CPU oracle against the reference's own golden vectors / known-answer tests.

Every case cites the reference test it restates (paths relative to /root/reference)."""
import numpy as np
import pytest

from oracle import esn_oracle as O

DTYPES = [np.float64, np.float32, np.float16]


# ---- test/states_tests.jl:9-26 -----------------------------------------------------------------
@pytest.mark.parametrize("T", DTYPES)
def test_states_golden(T):
    x = np.arange(1, 10).astype(T)
    cases = [
        (O.nlat1, [1, 2, 9, 4, 25, 6, 49, 8, 81]),
        (O.nlat2, [1, 2, 2, 4, 12, 6, 30, 8, 56]),
        (O.nlat3, [1, 2, 8, 4, 24, 6, 48, 8, 9]),
        (lambda v: O.partial_square(v, 0.6), [1, 4, 9, 16, 25, 6, 7, 8, 9]),
        (O.extended_square, [1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 4, 9, 16, 25, 36, 49, 64, 81]),
        (lambda v: O.pad(v, 10.0), [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]),
    ]
    for fn, expected in cases:
        out = fn(x)
        assert out.dtype == T
        assert np.array_equal(out, np.array(expected, dtype=T))


# ---- doctests in src/states.jl:186,257,329 (0:9 inputs) ----------------------------------------
def test_states_doctests():
    x = np.arange(0, 10, dtype=np.float64)
    assert np.array_equal(O.nlat1(x), [0, 1, 4, 3, 16, 5, 36, 7, 64, 9])
    assert np.array_equal(O.nlat2(x), [0, 1, 0, 3, 6, 5, 20, 7, 42, 9])
    assert np.array_equal(O.nlat3(x), [0, 1, 3, 3, 15, 5, 35, 7, 63, 9])
    # matrix form applies per column (src/states.jl:1-15)
    m = np.arange(21, dtype=np.float64).reshape(3, 7).T  # 7x3, columns 0..6, 7..13, 14..20
    out = O.nlat2(m)
    for j in range(3):
        assert np.array_equal(out[:, j], O.nlat2(m[:, j]))
    assert np.array_equal(O.pad(m, 1.0)[-1], [1, 1, 1])
    mods = ((O.MOD_NLAT2, 0.0), (O.MOD_PAD, 1.0))
    assert O.feature_dim(7, mods) == 8
    assert O.apply_modifiers(mods, m[:, 0]).shape == (8,)


# ---- test/Layers/esncell_tests.jl:131-289 ------------------------------------------------------
def _cell(n_in, n, W_in, W, **kw):
    return O.ESNCellParams(input_matrix=W_in.astype(np.float32), reservoir_matrix=W.astype(np.float32), **kw)


def test_cell_identity_linear_map():  # :131-153
    p = _cell(3, 3, np.eye(3), np.zeros((3, 3)), act=O.ACT_IDENTITY, leak=1.0)
    x = np.array([1, 2, 3], np.float32)
    y = O.esn_cell_step(p, x, np.zeros(3, np.float32))
    assert np.allclose(y, x)


def test_cell_mix_extremes():  # :155-192
    x = np.array([10, 20, 30], np.float32)
    h0 = np.array([4, 5, 6], np.float32)
    p0 = _cell(3, 3, np.eye(3), np.eye(3), act=O.ACT_IDENTITY, leak=0.0)
    assert np.allclose(O.esn_cell_step(p0, x, h0), h0)
    p1 = _cell(3, 3, np.eye(3), np.zeros((3, 3)), act=O.ACT_IDENTITY, leak=1.0, bias=np.ones(3, np.float32))
    assert np.allclose(O.esn_cell_step(p1, x, np.zeros(3, np.float32)), x + 1.0)


def test_cell_matrix_batch():  # :194-215
    p = _cell(3, 3, np.eye(3), np.zeros((3, 3)), act=O.ACT_IDENTITY, leak=1.0)
    X = np.array([[1, 2], [3, 4], [5, 6]], np.float32)
    Y = O.esn_cell_step(p, X, np.zeros((3, 2), np.float32))
    assert Y.shape == (3, 2) and np.allclose(Y, X)


def test_cell_scalar_vs_vector_leak():  # :218-263
    rng = np.random.default_rng(42)
    n_in, n, B = 4, 6, 3
    W_in = np.eye(n, n_in); W = np.eye(n)
    x = rng.random((n_in, B), dtype=np.float32); h0 = rng.random((n, B), dtype=np.float32)
    ps = _cell(n_in, n, W_in, W, act=O.ACT_IDENTITY, leak=np.float32(0.3))
    pv = _cell(n_in, n, W_in, W, act=O.ACT_IDENTITY, leak=np.full(n, 0.3, np.float32))
    assert np.allclose(O.esn_cell_step(ps, x, h0), O.esn_cell_step(pv, x, h0), atol=1e-6)


# ---- test/Interface/generic_contracts_tests.jl:296-300 -----------------------------------------
def test_collectstates_golden_linear_cell():
    # linear cell h_t = x_t + h_{t-1}: ESNCell with identity matrices/activation, zero h0
    p = _cell(2, 2, np.eye(2), np.eye(2), act=O.ACT_IDENTITY, leak=1.0)
    data = np.array([[3, 4], [4, 5]], np.float32)
    states, hs = O.collectstates([p], data, [np.zeros(2, np.float32)])
    assert np.array_equal(states, np.array([[3, 7], [4, 9]], np.float32))
    assert states.dtype == np.float32
    assert np.array_equal(hs[0], [7, 9])


# ---- test/Models/esn_tests.jl:193-246 (modifiers applied in order, after the cell) -------------
def test_modifier_order_and_carry_unmodified():
    p = _cell(3, 3, np.eye(3), np.zeros((3, 3)), act=O.ACT_IDENTITY, leak=1.0,
              modifiers=((O.MOD_NLAT1, 0), (O.MOD_PAD, 1.0)))
    data = np.array([[2.0], [3.0], [4.0]], np.float32)
    states, hs = O.collectstates([p], data, [np.zeros(3, np.float32)])
    assert np.array_equal(states[:, 0], [4, 3, 16, 1])  # NLAT1 then Pad
    assert np.array_equal(hs[0], [2, 3, 4])  # carry is the un-modified state


# ---- test/Models/esn_deep_tests.jl:111-135,218-242 ---------------------------------------------
def test_deep_modifiers_compose_and_last_layer_collected():
    l1 = _cell(2, 2, np.eye(2), np.zeros((2, 2)), act=O.ACT_IDENTITY, leak=1.0, modifiers=((O.MOD_NLAT1, 0),))
    l2 = _cell(2, 2, np.eye(2), np.zeros((2, 2)), act=O.ACT_IDENTITY, leak=1.0, modifiers=((O.MOD_PAD, 5.0),))
    data = np.array([[3.0, 1.0], [2.0, 1.0]], np.float32)
    states, hs = O.collectstates([l1, l2], data, [np.zeros(2, np.float32)] * 2)
    # layer1: [3,2] -> NLAT1 -> [9,2]; layer2 sees [9,2] -> identity -> pad 5
    assert np.array_equal(states[:, 0], [9, 2, 5])
    assert np.array_equal(states[:, 1], [1, 1, 5])
    assert np.array_equal(hs[0], [1, 1]) and np.array_equal(hs[1], [1, 1])


# ---- test/test_train_ridge.jl ------------------------------------------------------------------
def _random_ridge_problem(rng, T, F, S, D):  # :29-40
    states = rng.standard_normal((F, S)).astype(T)
    tw = rng.standard_normal((D, F)).astype(T)
    targets = (tw @ states + T(0.01) * rng.standard_normal((D, S)).astype(T)).astype(T)
    return states, targets, tw


def test_ridge_closed_form_orthogonal():  # :64-91
    scale, lam = 2.0, 0.25
    states = scale * np.eye(4)
    tw = np.array([[1.0, 0.0, -1.0, 0.5], [0.0, 2.0, 0.0, -0.5]])
    W = O.train_ridge_qr(states, tw @ states, lam)
    assert W.shape == (2, 4)
    assert np.allclose(W, tw * (scale ** 2 / (scale ** 2 + lam)), rtol=1e-10)


@pytest.mark.parametrize("seed,F,S,D,lam,rtol", [
    (7, 5, 60, 2, 1e-2, 1e-10),   # :93-117
    (3, 8, 20, 3, 1e-4, 1e-9),    # :143-161
    (19, 4, 50, 2, 0.0, 1e-9),    # :163-181
    (41, 7, 45, 5, 5e-3, 1e-10),  # :293-311
    (43, 6, 25, 3, 1e-8, 1e-9),   # :313-328
])
def test_ridge_qr_equals_gram_form_f64(seed, F, S, D, lam, rtol):
    Z, Y, _ = _random_ridge_problem(np.random.default_rng(seed), np.float64, F, S, D)
    W = O.train_ridge_qr(Z, Y, lam)
    ref = O.reference_ridge(Z, Y, lam)
    assert W.shape == (D, F)
    assert np.allclose(W, ref, rtol=rtol, atol=0)


def test_ridge_f32_agreement():  # :119-141
    Z, Y, _ = _random_ridge_problem(np.random.default_rng(11), np.float32, 4, 50, 2)
    W = O.train_ridge_qr(Z, Y, np.float32(1e-3), reg_dtype=np.float32)
    assert W.dtype == np.float32
    assert np.allclose(W, O.reference_ridge(Z, Y, 1e-3), rtol=1e-4)


def test_ridge_promotion_rule():  # src/train.jl:126-129
    Z, Y, _ = _random_ridge_problem(np.random.default_rng(5), np.float32, 4, 30, 2)
    assert O.train_ridge_qr(Z, Y, 1e-3).dtype == np.float64  # Float64 λ promotes
    Zi = np.array([[1, 2, 3, 4], [2, 1, 0, 1]]); Yi = np.array([[1.0, 2.0, 3.0, 4.0]])
    assert O.train_ridge_qr(Zi, Yi, 1e-3).shape == (1, 2)  # :268-276


def test_ridge_multi_output_consistency():  # :197-218
    rng = np.random.default_rng(29)
    Z = rng.standard_normal((6, 40)); Y = rng.standard_normal((4, 40))
    W = O.train_ridge_qr(Z, Y, 1e-2)
    for i in range(4):
        assert np.allclose(W[i:i + 1], O.train_ridge_qr(Z, Y[i:i + 1], 1e-2), rtol=1e-10)


def test_ridge_error_paths():  # :220-291
    with pytest.raises(O.DimensionMismatch):
        O.train_ridge_qr(np.zeros((5, 10)), np.zeros((2, 9)), 1e-3)
    with pytest.raises(O.ArgumentError, match="at least one training sample"):
        O.train_ridge_qr(np.zeros((4, 0)), np.zeros((2, 0)), 1e-3)
    with pytest.raises(O.ArgumentError):
        O.train_ridge_qr(np.ones((4, 20)), np.ones((2, 20)), -1e-3)


def test_apply_washout():  # :330-342
    states = np.arange(1.0, 21.0).reshape(5, 4).T
    targets = np.arange(100.0, 110.0).reshape(5, 2).T
    s, t = O.apply_washout(states, targets, 2)
    assert s.shape == (4, 3) and t.shape == (2, 3)
    assert np.array_equal(s, states[:, 2:5]) and np.array_equal(t, targets[:, 2:5])
    with pytest.raises(O.ArgumentError):
        O.apply_washout(states, targets, -1)
    with pytest.raises(O.ArgumentError):
        O.apply_washout(states, targets, 5)


# ---- test/Interface/array_workflow_tests.jl:145-163 --------------------------------------------
def test_autoregressive_identity_constant_output():
    p = _cell(2, 2, np.eye(2), np.zeros((2, 2)), act=O.ACT_IDENTITY, leak=1.0)
    W = np.eye(2)  # Float64 readout -> outputs promote to Float64
    Y, hs = O.predict_autoregressive([p], W, 3, np.array([1.0, 2.0], np.float32), [np.zeros(2, np.float32)])
    assert np.array_equal(Y, [[1, 1, 1], [2, 2, 2]])
    assert Y.dtype == np.float64
    with pytest.raises(O.ArgumentError):
        O.predict_autoregressive([p], W, 0, np.array([1.0, 2.0]), [np.zeros(2)])
    with pytest.raises(O.DimensionMismatch):  # closed-loop dimension, src/predict.jl:62-72
        O.predict_autoregressive([p], np.ones((3, 2)), 2, np.array([1.0, 2.0]), [np.zeros(2)])


# ---- test/Interface/array_workflow_tests.jl:111-143 (train, then teacher-forced predict) -------
def test_train_then_teacher_forced_predict_recovers_linear_targets():
    rng = np.random.default_rng(0)
    n_in, n, T = 2, 5, 40
    p = _cell(n_in, n, rng.standard_normal((n, n_in)), np.zeros((n, n)), act=O.ACT_IDENTITY, leak=1.0)
    data = rng.standard_normal((n_in, T)).astype(np.float32)
    A = rng.standard_normal((3, n_in)).astype(np.float32)
    targets = A @ data
    h0 = [np.zeros(n, np.float32)]
    W, hs, _ = O.train([p], data, targets, h0, reg=1e-10)
    Y, _ = O.predict_teacher([p], W, data, h0)
    assert np.allclose(Y, targets, atol=1e-4)


# ---- structure of rand_sparse / scaled_rand (test/inits_tests.jl:54-88) ------------------------
@pytest.mark.parametrize("n", [300, 1024, 2048, 4096, 8192])
def test_nnz_per_column_is_six(n):
    assert O.nnz_per_column(n, 6 / n) == 6


def test_rand_sparse_like_structure():
    rng = np.random.default_rng(17)
    W = O.rand_sparse_like(rng, 300, radius=1.2, sparsity=6 / 300)
    assert W.dtype == np.float32 and W.shape == (300, 300)
    assert np.all((W != 0).sum(axis=0) == 6)
    assert abs(O.spectral_radius(W) - 1.2) < 1e-5 * 10  # test/inits_tests.jl:67-70 (atol 1e-5, f32 scale)
    Win = O.scaled_rand_like(rng, 300, 3)
    assert Win.dtype == np.float32 and np.all(np.abs(Win) <= 0.1)
    rp, ci, v = O.dense_to_csr(W)
    assert rp[-1] == 1800 and ci.dtype == np.int32


# ---- the plain-C port (cpu_baseline) must agree with the pinned NumPy oracle --------------------
@pytest.mark.parametrize("sparse_form", [True, False])
def test_c_port_matches_numpy_oracle(sparse_form):
    from oracle import c_oracle

    rng = np.random.default_rng(5)
    n, T, B = 120, 30, 3
    W = O.rand_sparse_like(rng, n, radius=0.9, sparsity=0.05, as_csc=sparse_form)
    Win = O.scaled_rand_like(rng, n, 3)
    u = np.asfortranarray(rng.standard_normal((3, T, B)).astype(np.float32))
    h0 = rng.standard_normal((n, B)).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    states, hT = c_oracle.collect(W, Win, u, h0=h0, bias=bias, leak=0.5, act=O.ACT_TANH, mod_kind=O.MOD_NLAT2, pad=1.0)
    p = O.ESNCellParams(input_matrix=Win, reservoir_matrix=W, bias=bias, leak=np.float32(0.5), act=O.ACT_TANH,
                        modifiers=((O.MOD_NLAT2, 0), (O.MOD_PAD, 1.0)))
    for b in range(B):
        ref, hs = O.collectstates([p], u[:, :, b], [h0[:, b]])
        assert np.allclose(states[:, :, b], ref, rtol=1e-5, atol=1e-6)
        assert np.allclose(hT[:, b], hs[0], rtol=1e-5, atol=1e-6)


# ---- sibling cells.  The reference holds no numeric goldens for ES2N / ResESN / EuSN (its tests check
# shapes, eltypes and dispatch equality, test/collectstates_regression_tests.jl:28-112); the oracle is pinned through
# the hand-evaluated formulas of src/layers/{es2n,resesn,eusn}_cell.jl and their degenerate ESN limits. ----------------
def test_eusn_asym_operator_by_hand():  # eusn_cell.jl:120-123
    W = np.array([[1, 2], [3, 4]], np.float32)
    A = O.compute_asym_recurrent(W, np.float32(0.5))
    assert A.dtype == np.float32 and np.array_equal(A, np.array([[-0.5, -1], [1, -0.5]], np.float32))
    assert O.compute_asym_recurrent(W, np.float64(0.5)).dtype == np.float64  # gamma = T(diffusion) promotes


def test_sibling_cells_by_hand():
    W = np.array([[0, 1], [2, 0]], np.float32)
    Win = np.eye(2, dtype=np.float32)
    Om = np.array([[0, 1], [-1, 0]], np.float32)
    h = np.array([1, 2], np.float32)
    u = np.array([10, 20], np.float32)
    cand = u + W @ h  # identity activation: [12, 22]
    p = O.ESNCellParams(Win, W, act=O.ACT_IDENTITY, cell="es2n", orthogonal_matrix=Om, proximity=0.25)
    assert np.allclose(O.esn_cell_step(p, u, h), 0.75 * (Om @ h) + 0.25 * cand)      # es2n_cell.jl:113-114
    p = O.ESNCellParams(Win, W, act=O.ACT_IDENTITY, cell="resesn", orthogonal_matrix=Om, alpha=0.5, beta=2.0)
    assert np.allclose(O.esn_cell_step(p, u, h), 0.5 * (Om @ h) + 2.0 * cand)         # resesn_cell.jl:121-122
    p = O.ESNCellParams(Win, W, act=O.ACT_IDENTITY, cell="eusn", leak=0.5, diffusion=1.0)
    A = np.array([[-1, -1], [1, -1]], np.float32)
    assert np.allclose(O.esn_cell_step(p, u, h), 0.5 * h + 0.5 * (u + A @ h))         # eusn_cell.jl:100-104
    # matrix input (batched columns) takes the same path column by column
    H = np.stack([h, 2 * h], axis=1)
    U = np.stack([u, u], axis=1)
    out = O.esn_cell_step(p, U, H)
    assert np.allclose(out[:, 0], O.esn_cell_step(p, u, h))


def test_sibling_cells_degenerate_limits():  # parameter choices of collectstates_regression_tests.jl:30-62
    rng = np.random.default_rng(0)
    W = (0.3 * rng.standard_normal((8, 8))).astype(np.float32)
    Win = (0.5 * rng.standard_normal((8, 3))).astype(np.float32)
    Om = (0.3 * rng.standard_normal((8, 8))).astype(np.float32)
    data = rng.standard_normal((3, 20)).astype(np.float32)
    h0 = [np.zeros(8, np.float32)]
    esn, _ = O.collectstates([O.ESNCellParams(Win, W, act=O.ACT_IDENTITY, leak=1.0)], data, h0)
    es2n, _ = O.collectstates([O.ESNCellParams(Win, W, act=O.ACT_IDENTITY, cell="es2n", orthogonal_matrix=Om, proximity=1.0)], data, h0)
    assert np.array_equal(es2n, esn)
    res, _ = O.collectstates([O.ESNCellParams(Win, W, act=O.ACT_IDENTITY, cell="resesn", orthogonal_matrix=Om, alpha=0.0, beta=1.0)], data, h0)
    assert np.array_equal(res, esn)


# ---- DelayLayer goldens, test/basic_tests.jl:204-246 ----------------------------------------------------------------
def test_delay_layer_goldens():
    x = np.arange(1, 7, dtype=np.float32)[None, :]
    out, hist, clock = O.delay_layer_sequence(x[:, :5], 2, 1)           # :204-222
    assert np.array_equal(out.T, [[1, 0, 0], [2, 1, 0], [3, 2, 1], [4, 3, 2], [5, 4, 3]])
    assert clock == 5 and np.array_equal(hist, [[5, 4]])
    out, hist, clock = O.delay_layer_sequence(x[:, :4], 2, 2)           # :225-246
    assert np.array_equal(out.T, [[1, 0, 0], [2, 0, 0], [3, 2, 0], [4, 2, 0]])
    assert np.array_equal(hist, [[4, 2]])
    # :187-202 — explicit init columns show up unshifted after the first call with stride 2
    h0 = np.array([[1, 2], [1, 2]], np.float32)
    o, h1, c1 = O.delay_layer_step(np.array([0.3, 0.7], np.float32), h0, 0, 2)
    assert c1 == 1 and np.array_equal(h1, h0) and o.shape == (6,)
    assert np.array_equal(o[2:], [1, 1, 2, 2])  # vec(history) is column-major


def test_extend_and_correlation():
    u = np.array([[1.0, 2.0]]); x = np.array([[3.0, 4.0], [5.0, 6.0]])
    assert np.array_equal(O.extend(u, x), [[1, 2], [3, 4], [5, 6]])      # src/states.jl:138-141: input first
    Z = np.array([[1, 2, 3], [0, 1, 0]], np.float32)
    R = O.correlation_matrix(Z)
    assert R.dtype == np.float32 and np.allclose(R, np.array([[14, 2], [2, 1]]) / 3)   # src/conceptors.jl:28-33
    with pytest.raises(O.ArgumentError):
        O.correlation_matrix(np.zeros((3, 0)))                            # test/Models/conceptors_tests.jl:30
