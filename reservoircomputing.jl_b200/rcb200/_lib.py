"""ctypes binding of librcb200.so (include/rcb200.h).  Fails loudly when the CUDA library is
missing: there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RCB200_LIB") or os.path.join(os.path.dirname(_HERE), "librcb200.so")

RCB_OK, RCB_ERR_ARGUMENT, RCB_ERR_DIMENSION, RCB_ERR_CUDA, RCB_ERR_NUMERIC, RCB_ERR_UNSUPPORTED, RCB_ERR_ALLOC = 0, -1, -2, -3, -4, -5, -6
RCB_F32, RCB_F64 = 0, 1
ACT_IDENTITY, ACT_TANH, ACT_TANH_FAST, ACT_SIGMOID, ACT_RELU = 0, 1, 2, 3, 4
MOD_NLAT1, MOD_NLAT2, MOD_NLAT3, MOD_PAD, MOD_PARTIAL_SQUARE, MOD_EXTENDED_SQUARE = 1, 2, 3, 4, 5, 6
GRAM_AUTO, GRAM_FP64, GRAM_TENSOR, FIT_QR = 0, 1, 2, 3
TIMER_RECURRENCE, TIMER_GRAM, TIMER_SOLVE, TIMER_EXCHANGE, TIMER_REFINE = 0, 1, 2, 3, 4
COMM_ID_BYTES = 128
MAX_MODIFIERS, MAX_LAYERS = 8, 8


class ArgumentError(ValueError):
    """Julia ArgumentError (RCB_ERR_ARGUMENT)."""


class DimensionMismatch(ValueError):
    """Julia DimensionMismatch (RCB_ERR_DIMENSION)."""


class CudaError(RuntimeError):
    """RCB_ERR_CUDA / RCB_ERR_ALLOC: no usable B200 or a CUDA runtime failure."""


class NumericError(ArgumentError):
    """RCB_ERR_NUMERIC: the reference raises ArgumentError for a failed solve (train.jl:194-196)."""


class UnsupportedError(NotImplementedError):
    """RCB_ERR_UNSUPPORTED."""


class LayerDesc(C.Structure):
    _fields_ = [("res_dims", C.c_int32), ("activation", C.c_int32), ("use_bias", C.c_int32),
                ("leak_is_vector", C.c_int32), ("leak_scalar", C.c_double), ("n_modifiers", C.c_int32),
                ("modifier_kind", C.c_int32 * MAX_MODIFIERS), ("modifier_param", C.c_double * MAX_MODIFIERS)]


class EsnDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("in_dims", C.c_int32), ("out_dims", C.c_int32),
                ("readout_activation", C.c_int32), ("readout_use_bias", C.c_int32), ("data_dtype", C.c_int32)]


_P = C.c_void_p
_I32, _I64, _F64 = C.c_int32, C.c_int64, C.c_double

# name -> (restype, argtypes); mirrors include/rcb200.h one to one
SIGNATURES = {
    "rcb_version": (_I32, []),
    "rcb_last_error": (C.c_char_p, []),
    "rcb_device_count": (_I32, [C.POINTER(_I32)]),
    "rcb_ctx_create": (_I32, [_I32, _P, C.POINTER(_P)]),
    "rcb_ctx_destroy": (_I32, [_P]),
    "rcb_ctx_synchronize": (_I32, [_P]),
    "rcb_ctx_sm_count": (_I32, [_P, _P]),
    "rcb_ctx_launch_count": (_I32, [_P, C.POINTER(_I64)]),
    "rcb_ctx_last_kernel": (_I32, [_P, C.c_char_p, _I32]),
    "rcb_ctx_last_kernel_ms": (_I32, [_P, _I32, C.POINTER(C.c_float)]),
    "rcb_ctx_last_refine": (_I32, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "rcb_ctx_last_nonfinite": (_I32, [_P, C.POINTER(C.c_int32)]),
    "rcb_esn_create": (_I32, [_P, C.POINTER(EsnDesc), C.POINTER(LayerDesc), C.POINTER(_P)]),
    "rcb_esn_destroy": (_I32, [_P]),
    "rcb_esn_feature_dims": (_I32, [_P, C.POINTER(_I32)]),
    "rcb_esn_carry_dims": (_I32, [_P, C.POINTER(_I32)]),
    "rcb_esn_set_weights_dense": (_I32, [_P, _I32, _P, _I64, _P, _I64, _P]),
    "rcb_esn_set_weights_csc": (_I32, [_P, _I32, _P, _P, _P, _P, _I64, _P]),
    "rcb_esn_set_leak_vector": (_I32, [_P, _I32, _P]),
    "rcb_esn_set_cell": (_I32, [_P, _I32, _I32, _F64, _F64]),
    "rcb_esn_set_orthogonal_dense": (_I32, [_P, _I32, _P, _I64]),
    "rcb_esn_csr_nnz": (_I32, [_P, _I32, C.POINTER(_I64)]),
    "rcb_esn_export_csr": (_I32, [_P, _I32, _P, _P, _P]),
    "rcb_csr_from_dense": (_I32, [_P, _I64, _I32, C.POINTER(_I64), _P, _P, _P]),
    "rcb_csr_from_csc": (_I32, [_P, _P, _P, _I32, C.POINTER(_I64), _P, _P, _P]),
    "rcb_eusn_operator_from_dense": (_I32, [_P, _I64, _I32, _F64, C.POINTER(_I64), _P, _P, _P]),
    "rcb_sell_stats_from_dense": (_I32, [_P, _I64, _I32, _I32, _I32, _P]),
    "rcb_esn_set_member_params": (_I32, [_P, _I64, _P, _P]),
    "rcb_esn_set_readout": (_I32, [_P, _P, _P, _I64]),
    "rcb_collect": (_I32, [_P, _P, _I64, _I64, _P, _P, _P]),
    "rcb_fit_readout": (_I32, [_P, _P, _I32, _I64, _I64, _I64, _P, _I32, _I64, _I64, _I64, _F64, _I32, _P]),
    "rcb_gram_accumulate": (_I32, [_P, _P, _I32, _I64, _I64, _I64, _P, _I32, _I64, _I64, _I32, _I32, _P]),
    "rcb_gc_packed_len": (_I64, [_I64, _I64]),
    "rcb_gc_pack": (_I32, [_P, _P, _I64, _I64, _P]),
    "rcb_gc_unpack": (_I32, [_P, _P, _I64, _I64, _P]),
    "rcb_ridge_solve": (_I32, [_P, _P, _I64, _I64, _F64, _P]),
    "rcb_qr_merge": (_I32, [_P, _P, _P, _I64, _I64]),
    "rcb_qr_solve": (_I32, [_P, _P, _I64, _I64, _F64, _P]),
    "rcb_comm_unique_id": (_I32, [_P]),
    "rcb_comm_create": (_I32, [_P, _P, _I32, _I32, C.POINTER(_P)]),
    "rcb_comm_create_local": (_I32, [_P, _I32, _P]),
    "rcb_comm_destroy": (_I32, [_P]),
    "rcb_comm_rank": (_I32, [_P, C.POINTER(_I32), C.POINTER(_I32)]),
    "rcb_allreduce_gc": (_I32, [_P, _P, _I64, _I64]),
    "rcb_allreduce_gc_local": (_I32, [_P, _P, _I32, _I64, _I64]),
    "rcb_comm_allgather": (_I32, [_P, _P, _P, _I64]),
    "rcb_train_sharded": (_I32, [_P, _P, _P, _P, _I32, _I64, _I64, _I64, _F64, _I32, _P, _P, _P]),
    "rcb_chunk_windows": (_I32, [_I64, _I64, _I64, _I64, C.POINTER(_I64), C.POINTER(_I64)]),
    "rcb_train_time_chunked": (_I32, [_P, _P, _P, _P, _I32, _I64, _I64, _I64, _I64, _F64, _I32, _P, _P]),
    "rcb_train": (_I32, [_P, _P, _P, _I32, _I64, _I64, _I64, _F64, _I32, _P, _P, _P, _P]),
    "rcb_train_members": (_I32, [_P, _P, _P, _I32, _I64, _I64, _I64, _I32, _P, _I32, _P, _P, _P]),
    "rcb_train_partial": (_I32, [_P, _P, _P, _I32, _I64, _I64, _I64, _I32, _P, _P, _P]),
    "rcb_predict_teacher": (_I32, [_P, _P, _I64, _I64, _P, _P]),
    "rcb_step": (_I32, [_P, _P, _I64, _P, _P, _P]),
    "rcb_predict_autoregressive": (_I32, [_P, _P, _I64, _I64, _I32, _P, _P]),
    "rcb_extend_features": (_I32, [_P, _P, _I64, _P, _I64, _I64, _I32, _P]),
    "rcb_delay_features": (_I32, [_P, _P, _I32, _I64, _I64, _I64, _I32, _I32, _P, _I64, _P, _P]),
    "rcb_readout_apply": (_I32, [_P, _P, _P, _I32, _P, _I32, _I64, _I64, _I64, _P]),
    "rcb_spectral_radius_dense": (_I32, [_P, _P, _I64, _I32, _F64, _I32, C.POINTER(_F64), C.POINTER(_I32)]),
    "rcb_spectral_radius_csc": (_I32, [_P, _P, _P, _P, _I32, _F64, _I32, C.POINTER(_F64), C.POINTER(_I32)]),
    "rcb_hessenberg_eigvals": (_I32, [_P, _I32, _P, _P]),
    "rcb_correlation_matrix": (_I32, [_P, _P, _I32, _I64, _I64, _I64, _I32, _P]),
    "rcb_apply_modifiers": (_I32, [_P, _I32, _P, _P, _P, _I32, _I64, _I64, _P]),
}

CELL_ESN, CELL_ES2N, CELL_RESESN, CELL_EUSN = 0, 1, 2, 3

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CudaError(f"{LIB_PATH} is missing: build it with `python reservoircomputing.jl_b200/build.py` "
                            "(nvcc, sm_100a).  This package has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(status: int):
    if status == RCB_OK:
        return
    msg = lib().rcb_last_error().decode("utf-8", "replace")
    exc = {RCB_ERR_ARGUMENT: ArgumentError, RCB_ERR_DIMENSION: DimensionMismatch, RCB_ERR_CUDA: CudaError,
           RCB_ERR_NUMERIC: NumericError, RCB_ERR_UNSUPPORTED: UnsupportedError, RCB_ERR_ALLOC: CudaError}.get(status, RuntimeError)
    raise exc(msg)


def ptr(a):
    """void* of a numpy array (host), a torch tensor (host or CUDA), an int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(f"cannot take a pointer of {type(a)}")


def dtype_id(dt) -> int:
    dt = np.dtype(dt)
    if dt == np.float32:
        return RCB_F32
    if dt == np.float64:
        return RCB_F64
    raise ArgumentError(f"unsupported element type {dt}; the B200 path handles Float32 and Float64 data")
