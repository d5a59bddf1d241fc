"""CPU oracle for the ESN hot path of SciML/ReservoirComputing.jl (v0.12.35).

TEST INFRASTRUCTURE ONLY.  This module is a NumPy restatement of the reference's
algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``reservoircomputing.jl_b200/``) never imports, calls or links anything in ``oracle/``.

Parity pinning:
  * PINNED against the reference's own golden vectors / known-answer tests
    (``tests/test_oracle_golden.py``): state modifiers (test/states_tests.jl:9-16 and the
    doctests in src/states.jl), the ESNCell analytic cases (test/Layers/esncell_tests.jl:131-289),
    the collectstates golden (test/Interface/generic_contracts_tests.jl:296-300), the ridge
    closed form and Gram-form equivalences (test/test_train_ridge.jl:64-181,293-342), and the
    workflow goldens (test/Interface/array_workflow_tests.jl:111-155).
  * PARITY UNPINNED (third-party, un-vendored, no Manifest.toml in the reference; Julia is not
    installed here, so streams cannot be replayed): the *values* drawn by
    WeightInitializers.sparse_init / rand32 / randn32 and Random's Xoshiro / MersenneTwister
    streams, and NNlib.tanh_fast.  ``rand_sparse_like`` / ``scaled_rand_like`` restate the
    *structure* of those initializers on NumPy's PCG64; they are workload generators, not
    bit-parity claims.  Weights always cross the product boundary as explicit arrays.

Every function cites the reference file:line it follows.  Arrays follow Julia's layout
convention: features x time, ``(rows = feature, cols = time)``; NumPy arrays here are indexed the
same way (``states[f, t]``) and are converted to column-major at the C-ABI boundary by the tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------------------------
# activations  (src/layers/esn_cell.jl:129 default tanh_fast; src/models/esn.jl:95 default tanh)
# ----------------------------------------------------------------------------------------------
ACT_IDENTITY, ACT_TANH, ACT_TANH_FAST, ACT_SIGMOID, ACT_RELU = 0, 1, 2, 3, 4


def activation(act: int) -> Callable[[np.ndarray], np.ndarray]:
    if act == ACT_IDENTITY:
        return lambda x: x
    if act in (ACT_TANH, ACT_TANH_FAST):  # tanh_fast is NNlib (unvendored): "tanh within a few ulp"
        return np.tanh
    if act == ACT_SIGMOID:
        return lambda x: (1 / (1 + np.exp(-x))).astype(x.dtype)
    if act == ACT_RELU:
        return lambda x: np.maximum(x, 0).astype(x.dtype)
    raise ValueError(f"unknown activation id {act}")


# ----------------------------------------------------------------------------------------------
# state modifiers (src/states.jl)
# ----------------------------------------------------------------------------------------------
def _apply_tomatrix(fn, states: np.ndarray) -> np.ndarray:
    """src/states.jl:1-15 — apply a vector modifier column by column."""
    cols = [fn(states[:, j]) for j in range(states.shape[1])]
    return np.stack(cols, axis=1)


def nlat1(x: np.ndarray) -> np.ndarray:
    """src/states.jl:205-217 — square the 1-based odd rows (0-based even)."""
    if x.ndim == 2:
        return _apply_tomatrix(nlat1, x)
    z = x.copy()
    for i in range(x.shape[0]):  # i is 0-based; Julia idx = i+1
        if (i + 1) % 2 == 1:
            z[i] = x[i] * x[i]
    return z


def nlat2(x: np.ndarray) -> np.ndarray:
    """src/states.jl:277-289 — for 1-based odd idx > 1: z[idx] = x[idx-1]*x[idx-2]."""
    if x.ndim == 2:
        return _apply_tomatrix(nlat2, x)
    z = x.copy()
    for i in range(x.shape[0]):
        idx = i + 1
        if 1 < idx and idx % 2 == 1:
            z[i] = x[i - 1] * x[i - 2]
    return z


def nlat3(x: np.ndarray) -> np.ndarray:
    """src/states.jl:349-361 — for 1-based odd 1 < idx < n: z[idx] = x[idx-1]*x[idx+1]."""
    if x.ndim == 2:
        return _apply_tomatrix(nlat3, x)
    z = x.copy()
    n = x.shape[0]
    for i in range(n):
        idx = i + 1
        if 1 < idx < n and idx % 2 == 1:
            z[i] = x[i - 1] * x[i + 1]
    return z


def pad(x: np.ndarray, padding: float = 1.0) -> np.ndarray:
    """src/states.jl:73-88 — append a constant row ``T(padding)``."""
    if x.ndim == 2:
        row = np.full((1, x.shape[1]), padding, dtype=x.dtype)
        return np.concatenate([x, row], axis=0)
    return np.concatenate([x, np.array([padding], dtype=x.dtype)])


def partial_square(x: np.ndarray, eta: float) -> np.ndarray:
    """src/states.jl:404-421 — square the first floor(eta*n) rows."""
    if x.ndim == 2:
        return _apply_tomatrix(lambda c: partial_square(c, eta), x)
    z = x.copy()
    thr = int(math.floor(eta * x.shape[0]))
    z[:thr] = x[:thr] ** 2
    return z


def extended_square(x: np.ndarray) -> np.ndarray:
    """src/states.jl:466-473 — vcat(x, x.^2)."""
    if x.ndim == 2:
        return _apply_tomatrix(extended_square, x)
    return np.concatenate([x, x * x])


MOD_NLAT1, MOD_NLAT2, MOD_NLAT3, MOD_PAD, MOD_PARTIAL_SQUARE, MOD_EXTENDED_SQUARE = 1, 2, 3, 4, 5, 6


def apply_modifiers(mods: Sequence[tuple], x: np.ndarray) -> np.ndarray:
    """src/reservoircomputer.jl:72-79 ``__apply_seq`` — modifiers applied in order.

    ``mods`` is a sequence of ``(kind, param)`` with kind one of the MOD_* ids.
    """
    for kind, param in mods:
        if kind == MOD_NLAT1:
            x = nlat1(x)
        elif kind == MOD_NLAT2:
            x = nlat2(x)
        elif kind == MOD_NLAT3:
            x = nlat3(x)
        elif kind == MOD_PAD:
            x = pad(x, param)
        elif kind == MOD_PARTIAL_SQUARE:
            x = partial_square(x, param)
        elif kind == MOD_EXTENDED_SQUARE:
            x = extended_square(x)
        else:
            raise ValueError(f"unknown modifier {kind}")
    return x


def feature_dim(n: int, mods: Sequence[tuple]) -> int:
    for kind, _ in mods:
        if kind == MOD_PAD:
            n += 1
        elif kind == MOD_EXTENDED_SQUARE:
            n *= 2
    return n


# ----------------------------------------------------------------------------------------------
# ESN cell (src/layers/esn_cell.jl:175-200, src/generics.jl:83-96)
# ----------------------------------------------------------------------------------------------
@dataclass
class ESNCellParams:
    """``ps`` of an ESNCell: src/layers/esn_cell.jl:150-159."""

    input_matrix: np.ndarray  # (N, d_in)
    reservoir_matrix: np.ndarray  # (N, N) dense (rand_sparse default) — zeros kept
    bias: Optional[np.ndarray] = None  # (N,) only when use_bias
    leak: object = 1.0  # scalar or (N,) vector, esn_cell.jl:131-142
    act: int = ACT_TANH
    modifiers: tuple = ()  # per-layer state modifiers ((kind, param), ...)
    # sibling cells: "esn" | "es2n" (proximity) | "resesn" (alpha, beta) | "eusn" (diffusion; leak above)
    cell: str = "esn"
    orthogonal_matrix: Optional[np.ndarray] = None  # (N, N), ES2N / ResESN
    proximity: float = 1.0
    alpha: float = 1.0
    beta: float = 1.0
    diffusion: float = 1.0

    @property
    def n(self) -> int:
        return self.reservoir_matrix.shape[0]


def compute_asym_recurrent(W: np.ndarray, gamma) -> np.ndarray:
    """src/layers/eusn_cell.jl:120-123 — W .- transpose(W) .- gamma .* I, one fused broadcast:
    (W_ij - W_ji) in eltype(W), then minus gamma*I_ij in the promoted type."""
    eye = np.eye(W.shape[0], dtype=W.dtype)
    return (W - W.T) - gamma * eye


def esn_cell_step(p: ESNCellParams, u: np.ndarray, h: np.ndarray) -> np.ndarray:
    """One leaky step, src/layers/esn_cell.jl:175-184.

    ``u``: (d_in,) or (d_in, B); ``h``: (N, 1)/(N,) or (N, B).  T = eltype(inp) (:176).
    """
    T = u.dtype
    win_inp = p.input_matrix @ u  # generics.jl:91-96 (no bias on the input projection)
    if p.cell == "eusn":  # eusn_cell.jl:95-106: the same leaky step on the anti-symmetric operator
        w_state = compute_asym_recurrent(p.reservoir_matrix, T.type(p.diffusion)) @ h
    else:
        w_state = p.reservoir_matrix @ h  # generics.jl:83-96
    if p.bias is not None:
        w_state = w_state + (p.bias if w_state.ndim == 1 else p.bias[:, None])
    if win_inp.ndim == 1 and w_state.ndim == 2:  # vector input, N x 1 hidden state (lux_layers.jl:212-216)
        win_inp = win_inp[:, None]
    cand = activation(p.act)(win_inp + w_state)  # :180
    if p.cell == "es2n":  # es2n_cell.jl:106-116: ((1-p) .* O) * h .+ p .* cand  (.* and * associate left)
        prox = T.type(p.proximity)
        return ((T.type(1) - prox) * p.orthogonal_matrix) @ h + prox * cand
    if p.cell == "resesn":  # resesn_cell.jl:112-125: alpha .* (O * h) .+ beta .* cand
        return T.type(p.alpha) * (p.orthogonal_matrix @ h) + T.type(p.beta) * cand
    if np.isscalar(p.leak):  # __format_leak :186-188
        lc = T.type(p.leak)
        oml = T.type(1) - lc  # __one_minus_leak :194-196
    else:
        lc = np.asarray(p.leak).astype(T).reshape(-1, 1)  # :190-192
        oml = T.type(1) - lc
        if h.ndim == 1:
            lc, oml = lc[:, 0], oml[:, 0]
    return oml * h + lc * cand  # :182


# ----------------------------------------------------------------------------------------------
# collectstates (src/reservoircomputer.jl:113-133) and the DeepESN variant (esn_deep.jl:261-291)
# ----------------------------------------------------------------------------------------------
def collectstates(layers: Sequence[ESNCellParams], data: np.ndarray, h0: Sequence[np.ndarray]):
    """Teacher-forced state collection for an ESN (one layer) or DeepESN (several).

    ``data``: (d_in, T); ``h0``: one (N_l,) initial hidden state per layer (the reference draws
    it from ``init_state`` on the first call, esn_cell.jl:169-173 — here it is an explicit input).
    Returns ``(states (F, T) with eltype(data), carries [h_T per layer])``.

    Layer l+1 receives the *modified* state of layer l at the same step; only the last layer's
    modified state is collected (esn_deep.jl:270-281).  The carry is the *unmodified* h.
    """
    if data.shape[1] < 1:  # reservoircomputer.jl:65-70
        raise ValueError("collectstates data must have at least one column, got 0.")
    hs = [np.asarray(h, dtype=data.dtype).copy() for h in h0]
    cols = []
    for t in range(data.shape[1]):
        inp = data[:, t]
        for l, p in enumerate(layers):
            hs[l] = esn_cell_step(p, inp, hs[l])
            inp = apply_modifiers(p.modifiers, hs[l])
        cols.append(inp.copy())
    states = np.stack(cols, axis=1).astype(data.dtype)  # eltype(data), reservoircomputer.jl:123 / esn_deep.jl:288
    return states, hs


# ----------------------------------------------------------------------------------------------
# feature plumbing: Extend, DelayLayer, the delay ESNs; conceptors' correlation (8f-4)
# ----------------------------------------------------------------------------------------------
def extend(u: np.ndarray, x: np.ndarray) -> np.ndarray:
    """src/states.jl:138-141 — Extend: vcat(inp, state)."""
    return np.concatenate([u, x], axis=0)


def delay_layer_step(x: np.ndarray, history: np.ndarray, clock: int, stride: int):
    """One call of DelayLayer, src/layers/basic.jl:426-441: output vcat(inp, vec(history)) with the history as it
    stood BEFORE the call; then clock += 1 and, when clock % stride == 0, shift the columns and store inp in column 1.
    ``history``: (n, num_delays), column j = j-th most recent stored input.  Returns (out, history', clock')."""
    out = np.concatenate([x, history.reshape(-1, order="F")])
    clock += 1
    if history.shape[1] > 0 and clock % stride == 0:
        history = history.copy()
        history[:, 1:] = history[:, :-1].copy()
        history[:, 0] = x
    return out, history, clock


def delay_layer_sequence(x: np.ndarray, num_delays: int, stride: int = 1, history=None, clock: int = 0):
    """DelayLayer over the columns of x (n, T).  history None -> zeros (init_delay = zeros32, basic.jl:395,417-424)."""
    n, T = x.shape
    hist = np.zeros((n, num_delays), x.dtype) if history is None else np.array(history, dtype=x.dtype).reshape(n, num_delays)
    cols = []
    for t in range(T):
        o, hist, clock = delay_layer_step(x[:, t], hist, clock, stride)
        cols.append(o)
    return np.stack(cols, axis=1), hist, clock


def collectstates_delay(layers, data, h0, input_delay=None, state_delay=None, post_mods=()):
    """InputDelayESN / StateDelayESN / DelayESN, src/models/esn_delay.jl:548-558 (+ the plain __partial_apply for
    StateDelayESN, whose DelayLayer is simply the first state modifier, :283-296): input delay -> cell -> state delay ->
    remaining modifiers.  ``input_delay`` / ``state_delay``: (num_delays, stride) or None; ``layers``: one ESNCellParams
    whose own ``modifiers`` must be empty when a state delay is present."""
    u = data
    if input_delay is not None:
        u, _, _ = delay_layer_sequence(data, *input_delay)
    z, hs = collectstates(layers, u, h0)
    if state_delay is not None:
        z, _, _ = delay_layer_sequence(z, *state_delay)
        z = apply_modifiers(post_mods, z)
    return z, hs


def correlation_matrix(states: np.ndarray) -> np.ndarray:
    """src/conceptors.jl:28-33 — (Z Z') / S in float(eltype(states))."""
    if states.size == 0:
        raise ArgumentError("states must contain at least one state")
    Z = states if np.issubdtype(states.dtype, np.floating) else states.astype(np.float64)
    return (Z @ Z.T) / Z.dtype.type(Z.shape[1])


# ----------------------------------------------------------------------------------------------
# ridge readout fit (src/train.jl)
# ----------------------------------------------------------------------------------------------
class ArgumentError(ValueError):
    """Julia ArgumentError."""


class DimensionMismatch(ValueError):
    """Julia DimensionMismatch."""


def apply_washout(states: np.ndarray, targets: np.ndarray, washout: int):
    """src/train.jl:36-47."""
    if washout < 0:
        raise ArgumentError(f"washout must be ≥ 0, got {washout}")
    n = states.shape[1]
    if not washout < n:
        raise ArgumentError(f"washout={washout} is ≥ number of time steps={n}")
    return states[:, washout:], targets[:, washout:]


def ridge_augmented_system(states: np.ndarray, targets: np.ndarray, reg, reg_dtype=np.float64):
    """src/train.jl:108-138 — design=[Z'; sqrt(λ) I], rhs=[Y'; 0] in the promoted eltype.

    ``reg_dtype`` models ``typeof(objective.reg)``: Float64 for ``RidgeRegression(1e-4)``,
    Float32 for ``RidgeRegression(Float32, λ)`` (train.jl:24-34).
    """
    if states.shape[1] != targets.shape[1]:
        raise DimensionMismatch(
            f"states has {states.shape[1]} samples, targets has {targets.shape[1]}")
    if states.shape[1] < 1:
        raise ArgumentError("ridge regression requires at least one training sample")
    T = np.promote_types(np.promote_types(states.dtype, targets.dtype), reg_dtype)
    if not np.issubdtype(T, np.floating):
        T = np.dtype(np.float64)
    Z = states.astype(T)
    Y = targets.astype(T)
    lam = T.type(reg)
    if not lam >= 0:
        raise ArgumentError(f"RidgeRegression regularization must be ≥ 0, got reg={reg}")
    F = Z.shape[0]
    design = np.concatenate([Z.T, np.sqrt(lam) * np.eye(F, dtype=T)], axis=0)
    rhs = np.concatenate([Y.T, np.zeros((F, Y.shape[0]), dtype=T)], axis=0)
    return design, rhs


def train_ridge_qr(states: np.ndarray, targets: np.ndarray, reg, reg_dtype=np.float64) -> np.ndarray:
    """src/train.jl:140-147 (``qr(design) \\ rhs``) ≡ LinearSolve QRFactorization (:149-198).

    Householder QR least squares (LAPACK geqrf/ormqr/trtrs).  Returns W (d_out, F).
    """
    import scipy.linalg as sla

    design, rhs = ridge_augmented_system(states, targets, reg, reg_dtype)
    q, r = sla.qr(design, mode="economic")
    wt = sla.solve_triangular(r, q.T @ rhs, lower=False)
    return np.ascontiguousarray(wt.T)


def reference_ridge(states: np.ndarray, targets: np.ndarray, reg) -> np.ndarray:
    """The Gram-form cross-check of the reference's own tests, test/test_train_ridge.jl:18-27."""
    F = states.shape[0]
    lam = states.dtype.type(reg)
    gram = states @ states.T + lam * np.eye(F, dtype=states.dtype)
    return np.linalg.solve(gram, states @ targets.T).T


# ----------------------------------------------------------------------------------------------
# readout + predict (src/layers/basic.jl:171-182, src/predict.jl)
# ----------------------------------------------------------------------------------------------
def linear_readout(weight: np.ndarray, z: np.ndarray, bias: Optional[np.ndarray] = None,
                   act: int = ACT_IDENTITY) -> np.ndarray:
    """src/layers/basic.jl:171-182 — y = ψ(W z (+ b))."""
    out = weight @ z
    if bias is not None:
        out = out + (bias if out.ndim == 1 else bias.reshape(-1, *([1] * (out.ndim - 1))))
    return activation(act)(out)


def _model_step(layers, weight, ro_bias, ro_act, inp, hs):
    """(rc)(inp, ps, st): reservoircomputer.jl:90-94 / esn_deep.jl:213-217."""
    for l, p in enumerate(layers):
        h = hs[l]
        if h.dtype != inp.dtype:  # Julia promotes h through W*h with eltype(inp) leak (esn_cell.jl:176-182)
            h = h.astype(np.promote_types(h.dtype, inp.dtype))
        hs[l] = esn_cell_step(p, inp, h)
        inp = apply_modifiers(p.modifiers, hs[l])
    y = linear_readout(weight, inp, ro_bias, ro_act)
    return y, hs


def predict_teacher(layers, weight, data, h, ro_bias=None, ro_act=ACT_IDENTITY):
    """src/predict.jl:163-177 — y_t = model(u_t) for every column."""
    if data.shape[1] < 1:
        raise ArgumentError("predict data must have at least one column, got 0.")
    hs = [np.array(x, copy=True) for x in h]
    outs = []
    for t in range(data.shape[1]):
        y, hs = _model_step(layers, weight, ro_bias, ro_act, data[:, t], hs)
        outs.append(y)
    return np.stack(outs, axis=1), hs


def predict_autoregressive(layers, weight, steps: int, initialdata, h, ro_bias=None,
                           ro_act=ACT_IDENTITY):
    """src/predict.jl:74-88 — y_1 = model(u0); y_{t+1} = model(y_t)."""
    if steps < 1:
        raise ArgumentError(f"steps must be ≥ 1, got {steps}")
    hs = [np.array(x, copy=True) for x in h]
    n_in = len(initialdata)
    cur = np.asarray(initialdata)
    outs = []
    for step in range(1, steps + 1):
        cur, hs = _model_step(layers, weight, ro_bias, ro_act, cur, hs)
        if cur.shape[0] != n_in:  # predict.jl:62-72
            raise DimensionMismatch(
                f"autoregressive predict requires each output to have length {n_in} so it can be "
                f"used as the next input; step {step} produced length {cur.shape[0]}")
        outs.append(cur)
    return np.stack(outs, axis=1), hs


def train(layers, data, targets, h0, reg=0.0, washout=0, reg_dtype=np.float64):
    """src/train.jl:232-251 — collectstates → washout → ridge → W_out; returns (W, carries, states_wo)."""
    states, hs = collectstates(layers, data, h0)
    if washout > 0:
        states_wo, targets_wo = apply_washout(states, targets, washout)
    else:
        states_wo, targets_wo = states, targets
    W = train_ridge_qr(states_wo, targets_wo, reg, reg_dtype)
    return W, hs, states_wo


# ----------------------------------------------------------------------------------------------
# weight producers: structure of rand_sparse / scaled_rand (values: parity unpinned, see header)
# ----------------------------------------------------------------------------------------------
def nnz_per_column(n: int, sparsity: float) -> int:
    """rand_sparse passes 1-sparsity to sparse_init (src/inits/inits_reservoir.jl:62-63), which
    zeroes ceil((1-s)*rows) entries per column: nnz/col = rows - ceil((1 - s) * rows).  With the
    README's ``sparsity = 6/300`` (a Float64) ``T(1) - sparsity`` promotes to Float64."""
    lcl = 1.0 - float(sparsity)
    return n - int(math.ceil(lcl * n))


def spectral_radius(W) -> float:
    """max |eigvals(W)| (src/inits/inits_components.jl:127).  Dense LAPACK below N=1500, ARPACK
    (sparse Arnoldi) above: setup only, SURVEY.md §7.3 item 10."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    n = W.shape[0]
    if n <= 1500:
        Wd = W.toarray() if sp.issparse(W) else np.asarray(W)
        return float(np.max(np.abs(np.linalg.eigvals(Wd.astype(np.float64)))))
    Ws = sp.csr_matrix(W).astype(np.float64)
    v0 = np.random.default_rng(0).standard_normal(n)
    vals = spla.eigs(Ws, k=6, which="LM", v0=v0, return_eigenvectors=False, tol=1e-10, maxiter=20000)
    return float(np.max(np.abs(vals)))


def rand_sparse_like(rng: np.random.Generator, n: int, radius: float = 1.0, sparsity: float = 0.1,
                     std: float = 1.0, as_csc: bool = False):
    """Structure of ``rand_sparse`` (src/inits/inits_reservoir.jl:55-67): exactly
    ``nnz_per_column`` N(0,std) non-zeros per column at uniformly shuffled rows, rescaled so the
    spectral radius equals ``radius`` (scale_radius!, src/inits/inits_components.jl:126-138).
    Returns dense float32 (N,N) (return_sparse=false default) or a scipy CSC (return_sparse=true,
    ext/RCSparseArraysExt.jl:5-7)."""
    import scipy.sparse as sp

    k = nnz_per_column(n, sparsity)
    rows = np.empty((n, k), dtype=np.int64)
    for j in range(n):
        rows[j] = np.sort(rng.choice(n, size=k, replace=False))
    vals = (rng.standard_normal((n, k)) * std).astype(np.float32)
    colptr = np.arange(0, n * k + 1, k, dtype=np.int64)
    W = sp.csc_matrix((vals.ravel(), rows.ravel(), colptr), shape=(n, n))
    rho = spectral_radius(W)
    W = W.multiply(np.float32(radius) / np.float32(rho)).astype(np.float32).tocsc()
    if not np.all(np.isfinite(W.data)):  # check_inf_nan, inits_components.jl:99-111
        raise ArgumentError("Created matrix contains invalid values")
    return W if as_csc else np.asarray(W.toarray(), dtype=np.float32)


def scaled_rand_like(rng: np.random.Generator, n: int, d_in: int, scaling: float = 0.1) -> np.ndarray:
    """``scaled_rand``: U[0,1) mapped by (x-0.5)*2*scaling (src/inits/inits_input.jl:80-88,
    src/inits/inits_components.jl:1-4)."""
    x = rng.random((n, d_in), dtype=np.float32)
    return ((x - np.float32(0.5)) * (np.float32(2) * np.float32(scaling))).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# sparse hand-off formats (what the product's CSR export must reproduce bit-exactly)
# ----------------------------------------------------------------------------------------------
def dense_to_csr(W: np.ndarray):
    """Row-major CSR of the non-zeros of dense W, columns ascending (0-based int32)."""
    n = W.shape[0]
    rowptr = np.zeros(n + 1, dtype=np.int64)
    colind, vals = [], []
    for i in range(n):
        nz = np.nonzero(W[i])[0]
        rowptr[i + 1] = rowptr[i] + nz.size
        colind.append(nz.astype(np.int32))
        vals.append(W[i, nz])
    return rowptr, (np.concatenate(colind) if colind else np.zeros(0, np.int32)), \
        (np.concatenate(vals) if vals else np.zeros(0, W.dtype))


# ----------------------------------------------------------------------------------------------
# synthetic series
# ----------------------------------------------------------------------------------------------
def lorenz_series(T: int, dt: float = 0.02, x0=(1.0, 0.0, 0.0), discard: int = 300) -> np.ndarray:
    """Lorenz-63 (σ=10, ρ=28, β=8/3) by RK4; returns float64 (3, T+1) after ``discard`` steps."""
    s, r, b = 10.0, 28.0, 8.0 / 3.0

    def f(v):
        return np.array([s * (v[1] - v[0]), v[0] * (r - v[2]) - v[1], v[0] * v[1] - b * v[2]])

    v = np.array(x0, dtype=np.float64)
    out = np.empty((3, T + 1 + discard))
    for i in range(T + 1 + discard):
        out[:, i] = v
        k1 = f(v); k2 = f(v + 0.5 * dt * k1); k3 = f(v + 0.5 * dt * k2); k4 = f(v + dt * k3)
        v = v + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    return out[:, discard:]


def mackey_glass_series(T: int, tau: int = 17, beta: float = 0.2, gamma: float = 0.1, n: int = 10,
                        transient: int = 1000, substeps: int = 10) -> np.ndarray:
    """Mackey-Glass (dt=1, RK4 sub-stepped with linear history interpolation), x(≤0)=1.2,
    affine-normalised to zero mean / unit max-abs; float64 (1, T+1)."""
    h = 1.0 / substeps
    hist = tau * substeps
    total = (T + 1 + transient) * substeps
    x = np.empty(total + hist + 1)
    x[: hist + 1] = 1.2

    def f(xt, xd):
        return beta * xd / (1.0 + xd ** n) - gamma * xt

    for i in range(hist, hist + total):
        xd0, xd1 = x[i - hist], x[i - hist + 1]
        xdm = 0.5 * (xd0 + xd1)
        k1 = f(x[i], xd0); k2 = f(x[i] + 0.5 * h * k1, xdm)
        k3 = f(x[i] + 0.5 * h * k2, xdm); k4 = f(x[i] + h * k3, xd1)
        x[i + 1] = x[i] + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    s = x[hist + transient * substeps:: substeps][: T + 1]
    s = s - s.mean()
    s = s / np.max(np.abs(s))
    return s[None, :]
