"""Committed golden fixtures (tests/golden/, generator: tests/golden/make_golden.py).

CPU (`-m "not gpu"`): the oracle reproduces the reference's known-answer vectors and the stored
oracle outputs (so fixture and oracle cannot drift apart silently).
GPU (`-m gpu`): the CUDA path, through the C-ABI, against the stored arrays — no oracle code runs."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "reservoircomputing.jl_b200"))


@pytest.fixture(scope="module")
def ref():
    with open(os.path.join(GOLD, "reference_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def small():
    z = np.load(os.path.join(GOLD, "esn_small.npz"))
    d = {k: z[k] for k in z.files}
    N = int(d["N"])
    W = np.zeros((N, N), np.float32)
    W[d["rows"], d["cols"]] = d["vals"]
    d["W"] = W
    return d


def _state_err(a, b):
    num = np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)), axis=0)
    den = np.maximum(np.max(np.abs(b.astype(np.float64)), axis=0), 1e-30)
    return float(np.max(num / den))


# ---- CPU: oracle vs fixtures -----------------------------------------------------------------------
def test_oracle_reproduces_reference_vectors(ref):
    from oracle import esn_oracle as O

    for key in ("modifiers_1to9", "modifiers_0to9"):
        g = ref[key]
        x = np.array(g["input"], np.float64)
        fns = {"NLAT1": O.nlat1, "NLAT2": O.nlat2, "NLAT3": O.nlat3, "ExtendedSquare": O.extended_square,
               "PartialSquare_0.6": lambda v: O.partial_square(v, 0.6), "Pad_10": lambda v: O.pad(v, 10.0)}
        for name, fn in fns.items():
            if name in g:
                assert np.array_equal(fn(x), np.array(g[name], np.float64)), (key, name)
    g = ref["collectstates_linear_cell"]
    p = O.ESNCellParams(input_matrix=np.eye(2, dtype=np.float32), reservoir_matrix=np.eye(2, dtype=np.float32), leak=1.0,
                        act=O.ACT_IDENTITY)
    s, _ = O.collectstates([p], np.array(g["data"], np.float32), [np.zeros(2, np.float32)])
    assert np.array_equal(s, np.array(g["states"], np.float32))
    g = ref["ridge_closed_form"]
    tw = np.array(g["W_true"])
    Z = g["s"] * np.eye(tw.shape[1])
    W = O.train_ridge_qr(Z, tw @ Z, g["lambda"])
    assert np.allclose(W, tw * g["s"] ** 2 / (g["s"] ** 2 + g["lambda"]), rtol=g["rtol"])
    for key in ("delay_layer_stride1", "delay_layer_stride2"):
        g = ref[key]
        x = np.array(g["inputs"], np.float32)[None, :]
        out, _, clock = O.delay_layer_sequence(x, g["num_delays"], g["stride"])
        assert np.array_equal(out.T, np.array(g["outputs"], np.float32)) and clock == len(g["inputs"]), key


def test_oracle_reproduces_esn_small(small):
    from oracle import esn_oracle as O

    layer = O.ESNCellParams(input_matrix=small["Win"], reservoir_matrix=small["W"], leak=1.0, act=O.ACT_TANH,
                            modifiers=((O.MOD_NLAT2, 0.0),))
    s, hs = O.collectstates([layer], small["data"], [small["h0"]])
    assert _state_err(s, small["states32"]) <= 1e-6
    # the float64 "truth" run bounds the float32 oracle's own rounding: both inside the 1e-5 budget
    assert _state_err(small["states32"], small["states64_as32"]) <= 1e-5
    w = int(small["washout"])
    W = O.train_ridge_qr(small["states32"][:, w:], small["target"][:, w:], float(small["lam"]))
    assert np.linalg.norm(W - small["Wout"]) / np.linalg.norm(small["Wout"]) <= 1e-8


# ---- GPU: CUDA path vs fixtures ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def R():
    import build as rcb_build

    rcb_build.build()
    import rcb200

    return rcb200


@pytest.mark.gpu
def test_cuda_reference_vectors(R, ref):
    for key in ("modifiers_1to9", "modifiers_0to9"):
        g = ref[key]
        x = np.array(g["input"], np.float64)
        mods = {"NLAT1": R.NLAT1, "NLAT2": R.NLAT2, "NLAT3": R.NLAT3, "ExtendedSquare": R.ExtendedSquare,
                "PartialSquare_0.6": R.PartialSquare(0.6), "Pad_10": R.Pad(10.0)}
        for name, m in mods.items():
            if name in g:
                for T in (np.float64, np.float32):
                    out = R.apply_modifiers(m, x.astype(T))
                    assert out.dtype == T and np.array_equal(out, np.array(g[name], T)), (key, name)
    g = ref["ridge_closed_form"]
    tw = np.array(g["W_true"])
    Z = np.asfortranarray(g["s"] * np.eye(tw.shape[1]))
    W = R.fit_readout(R.RidgeRegression(g["lambda"]), Z, np.asfortranarray(tw @ Z))
    assert np.allclose(W, tw * g["s"] ** 2 / (g["s"] ** 2 + g["lambda"]), rtol=g["rtol"])
    for key in ("delay_layer_stride1", "delay_layer_stride2"):
        g = ref[key]
        dl = R.DelayLayer(1, num_delays=g["num_delays"], stride=g["stride"])
        x = np.array(g["inputs"], np.float32)[None, :]
        out, st = dl.apply_sequence(x, dl.initialstates(np.random.default_rng(0)))
        assert np.array_equal(out.T, np.array(g["outputs"], np.float32)) and st.clock == len(g["inputs"]), key
        # call by call, like the reference's loop: same outputs, the buffer and the clock carried in st
        st = dl.initialstates(np.random.default_rng(0))
        for i, row in enumerate(g["outputs"]):
            o, st = dl.apply_sequence(x[:, i:i + 1], st)
            assert np.array_equal(o[:, 0], np.array(row, np.float32)), (key, i)


@pytest.mark.gpu
@pytest.mark.parametrize("handoff", ["dense", "csc"])
def test_cuda_esn_small(R, small, handoff):
    import scipy.sparse as sp

    W = small["W"] if handoff == "dense" else sp.csc_matrix(small["W"])
    esn = R.ESN(3, int(small["N"]), 3, init_reservoir=lambda rng, *d: W, init_input=lambda rng, *d: small["Win"],
                state_modifiers=R.NLAT2)
    ps, st = R.setup(np.random.default_rng(0), esn)
    st = R.resetcarry(None, esn, st, init_carry=lambda rng, n: small["h0"])
    rp, ci, v = R.export_csr(esn, ps)
    order = np.lexsort((small["cols"], small["rows"]))
    assert np.array_equal(ci, small["cols"][order]) and v.tobytes() == small["vals"][order].tobytes()
    assert np.array_equal(np.diff(rp), np.bincount(small["rows"], minlength=int(small["N"])))
    data, target, w = np.asfortranarray(small["data"]), np.asfortranarray(small["target"]), int(small["washout"])
    states, st2 = R.collectstates(esn, data, ps, st)
    assert _state_err(states, small["states32"]) <= 1e-5
    assert _state_err(states, small["states64_as32"]) <= 1e-5
    assert np.max(np.abs(st2.reservoir.carry[0] - small["hT"])) <= 1e-5 * np.max(np.abs(small["hT"]))
    ps2, _ = R.train(esn, data, target, ps, st, objective=R.RidgeRegression(float(small["lam"])), washout=w)
    # same-states solver parity: predictions of the fitted readout on the stored states
    pred = ps2.readout.weight @ small["states32"][:, w:].astype(np.float64)
    pred_ref = small["Wout"] @ small["states32"][:, w:].astype(np.float64)
    assert np.linalg.norm(pred - pred_ref) / np.linalg.norm(pred_ref) <= 1e-4
