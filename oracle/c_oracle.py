"""ctypes wrapper of the plain-C oracle port (oracle/esn_oracle.c).  TEST / BASELINE INFRASTRUCTURE
ONLY — same rules as oracle/esn_oracle.py."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    src = os.path.join(_HERE, "esn_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_collect.restype = C.c_int
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return lib().oracle_max_threads()


def collect(W, Win, u, h0=None, bias=None, leak=1.0, act=1, mod_kind=0, pad=None, want_states=True, nthreads=0):
    """W: scipy CSC/CSR/dense (n,n) float32; Win (n,d_in); u (d_in,T,B) F-order float32; h0 (n,B).
    Returns (states (F,T,B) or None, hT (n,B))."""
    import scipy.sparse as sp

    n = W.shape[0]
    u = np.asfortranarray(u, dtype=np.float32)
    d_in, T, B = u.shape
    Win = np.asfortranarray(Win, dtype=np.float32)
    if sp.issparse(W):
        Wc = sp.csc_matrix(W)
        colptr = Wc.indptr.astype(np.int64); rowval = Wc.indices.astype(np.int32)
        nz = np.ascontiguousarray(Wc.data, dtype=np.float32)
        cp, rv = colptr.ctypes.data_as(C.c_void_p), rowval.ctypes.data_as(C.c_void_p)
    else:
        nz = np.asfortranarray(W, dtype=np.float32)
        cp, rv = None, None
    F = n + (1 if pad is not None else 0)
    states = np.empty((F, T, B), np.float32, order="F") if want_states else None
    hT = np.empty((n, B), np.float32, order="F")
    h0a = None if h0 is None else np.asfortranarray(np.asarray(h0, np.float32).reshape(n, B))
    lv = None if np.isscalar(leak) else np.ascontiguousarray(leak, dtype=np.float32)
    ba = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    rc = lib().oracle_collect(C.c_int(n), C.c_int(d_in), cp, rv, p(nz), p(Win), p(ba), p(lv),
                              C.c_float(0.0 if lv is not None else float(leak)), C.c_int(act), C.c_int(mod_kind),
                              C.c_int(0 if pad is None else 1), C.c_float(0.0 if pad is None else float(pad)),
                              C.c_int64(T), C.c_int64(B), p(u), p(h0a), p(states), p(hT), C.c_int(nthreads))
    if rc != 0:
        raise MemoryError("oracle_collect: allocation failed")
    return states, hT
