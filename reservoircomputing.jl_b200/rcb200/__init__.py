"""rcb200 — host-side mirror of ReservoirComputing.jl's ESN hot-path API over librcb200.so (B200, sm_100a).

There is no CPU fallback: every call ends in a CUDA kernel of the in-tree library, and importing
this package without the built library raises on first use."""
from ._lib import (ArgumentError, CudaError, DimensionMismatch, NumericError, UnsupportedError, LIB_PATH,
                   GRAM_AUTO, GRAM_FP64, GRAM_TENSOR, FIT_QR, TIMER_RECURRENCE, TIMER_GRAM, TIMER_SOLVE, TIMER_EXCHANGE, TIMER_REFINE, lib)
from .api import (NT, Device, Comm, comm_for, ReservoirChain, StatefulLayer, Extend, Collect, DelayLayer, InputDelayESN, StateDelayESN, DelayESN, extend_features, readout_apply,
                  correlation_matrix, ridge_map, DeepESN, ESN, ESNCell, ES2N, ES2NCell, ResESN, ResESNCell, EuSN, EuSNCell, orthogonal, ExtendedSquare, LinearReadout, NLAT1, NLAT2, NLAT3, Pad,
                  PartialSquare, QRFactorization, QRSolver, NormalCholeskyFactorization, RidgeRegression, StandardRidge, addreadout,
                  apply, partial_apply, apply_modifiers, collectstates, csr_from_matrix, eusn_operator_csr, spectral_radius, scale_radius, hessenberg_eigvals, default_device, device_count, export_csr, fit_readout, gram,
                  ridge_solve, qr_merge, qr_solve, invalidate, identity,
                  predict, rand32, rand_sparse, randn32, relu, resetcarry, scaled_rand, setup, sigmoid, tanh,
                  tanh_fast, train, train_ensemble, train_sharded, train_time_chunked, chunk_windows, shard_range, time_chunks, zeros32)
