"""Generator of the committed golden fixtures (run from the repo root: python tests/golden/make_golden.py).

Two kinds of fixture:
  reference_vectors.json  known-answer vectors transcribed from the REFERENCE's own tests and doctests
                          (SciML/ReservoirComputing.jl v0.12.35; each entry cites file:line).  Julia is
                          not installed in the build container, so the reference cannot be executed to
                          regenerate them; they are restated literally.
  esn_small.npz           seeded inputs + outputs of the pinned CPU oracle (oracle/esn_oracle.py) for a
                          README-shaped ESN (3 -> 300 -> 3, 6 nnz/col, NLAT2): weights as COO triplets,
                          W_in, h0, data, targets, FP32 states, FP64 states, Float64 QR readout weights.
                          The GPU parity tests compare the CUDA path to these arrays without executing
                          the oracle; a CPU test checks that the oracle still reproduces them.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import esn_oracle as O  # noqa: E402


def reference_vectors():
    return {
        "source": "SciML/ReservoirComputing.jl v0.12.35",
        "modifiers_1to9": {  # test/states_tests.jl:9-16
            "cite": "test/states_tests.jl:9-16",
            "input": list(range(1, 10)),
            "NLAT1": [1, 2, 9, 4, 25, 6, 49, 8, 81],
            "NLAT2": [1, 2, 2, 4, 12, 6, 30, 8, 56],
            "NLAT3": [1, 2, 8, 4, 24, 6, 48, 8, 9],
            "PartialSquare_0.6": [1, 4, 9, 16, 25, 6, 7, 8, 9],
            "ExtendedSquare": [1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 4, 9, 16, 25, 36, 49, 64, 81],
            "Pad_10": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10],
        },
        "modifiers_0to9": {  # doctests src/states.jl:186,257,329
            "cite": "src/states.jl:186,257,329",
            "input": list(range(0, 10)),
            "NLAT1": [0, 1, 4, 3, 16, 5, 36, 7, 64, 9],
            "NLAT2": [0, 1, 0, 3, 6, 5, 20, 7, 42, 9],
            "NLAT3": [0, 1, 3, 3, 15, 5, 35, 7, 63, 9],
        },
        "collectstates_linear_cell": {  # h_t = x_t + h_{t-1}, h_0 = 0
            "cite": "test/Interface/generic_contracts_tests.jl:296-300",
            "data": [[3, 4], [4, 5]],
            "states": [[3, 7], [4, 9]],
        },
        "ridge_closed_form": {  # W = W_true * s^2 / (s^2 + lambda) for states = s*I
            "cite": "test/test_train_ridge.jl:64-91",
            "s": 2.0, "lambda": 0.5, "W_true": [[1.0, -2.0, 0.5], [0.0, 3.0, -1.0]], "rtol": 1e-10,
        },
        "delay_layer_stride1": {  # DelayLayer(1; num_delays=2, stride=1, init_delay=zeros32) fed 1..5
            "cite": "test/basic_tests.jl:204-222",
            "num_delays": 2, "stride": 1, "inputs": [1, 2, 3, 4, 5],
            "outputs": [[1, 0, 0], [2, 1, 0], [3, 2, 1], [4, 3, 2], [5, 4, 3]],
        },
        "delay_layer_stride2": {  # the same layer with stride=2 fed 1..6: the buffer shifts on every second call
            "cite": "test/basic_tests.jl:225-246",
            "num_delays": 2, "stride": 2, "inputs": [1, 2, 3, 4],
            "outputs": [[1, 0, 0], [2, 0, 0], [3, 2, 0], [4, 2, 0]],
        },
        "autoregressive_identity": {  # identity readout on a 2-dim identity reservoir feeds [1,2] back
            "cite": "test/Interface/array_workflow_tests.jl:145-155",
            "initialdata": [1.0, 2.0], "steps": 3, "output": [[1, 1, 1], [2, 2, 2]], "eltype": "Float64",
        },
    }


def esn_small(seed=17, N=300, k=6, T=100, washout=20, lam=1e-6):
    rng = np.random.default_rng(seed)
    W = O.rand_sparse_like(rng, N, radius=1.2, sparsity=k / N)
    Win = O.scaled_rand_like(rng, N, 3)
    h0 = rng.standard_normal(N).astype(np.float32)
    series = O.lorenz_series(T + 1)
    data = series[:, :T].astype(np.float32)
    target = series[:, 1:T + 1].astype(np.float32)
    layer = O.ESNCellParams(input_matrix=Win, reservoir_matrix=W, leak=1.0, act=O.ACT_TANH, modifiers=((O.MOD_NLAT2, 0.0),))
    states32, hs = O.collectstates([layer], data, [h0])
    states64, _ = O.collectstates([layer], data.astype(np.float64), [h0.astype(np.float64)])
    Wout = O.train_ridge_qr(states32[:, washout:], target[:, washout:], lam)
    r, c = np.nonzero(W)
    return dict(N=N, k=k, T=T, washout=washout, lam=lam, rows=r.astype(np.int32), cols=c.astype(np.int32),
                vals=W[r, c].astype(np.float32), Win=Win.astype(np.float32), h0=h0, data=data, target=target,
                states32=states32.astype(np.float32), states64_as32=states64.astype(np.float32),
                hT=hs[0].astype(np.float32), Wout=Wout.astype(np.float64))


def main():
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(reference_vectors(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "esn_small.npz"), **esn_small())
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
