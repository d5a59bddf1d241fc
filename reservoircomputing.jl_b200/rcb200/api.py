"""Host-side mirror of the reference's operator interface for the ESN hot path.

Julia is not available in the build image, so the host side above the C-ABI is written in Python
with the same names, argument meaning and error behaviour as the reference:

    ESN / DeepESN / ESNCell / LinearReadout        src/models/esn.jl:93-106, esn_deep.jl:102-147,
                                                   src/layers/esn_cell.jl:127-148, basic.jl:139-154
    NLAT1/NLAT2/NLAT3/Pad/PartialSquare/ExtendedSquare   src/states.jl
    setup(rng, model) -> (ps, st)                  LuxCore.setup; reservoircomputer.jl:51-63
    collectstates(rc, data, ps, st)                src/reservoircomputer.jl:96-133
    train(rc, X, Y, ps, st; objective, washout, return_states)   src/train.jl:232-251
    predict(rc, steps|data, ps, st; initialdata)   src/predict.jl:55-60,90-103,118-129
    fit_readout(objective, states, targets)        __fit_readout, src/train.jl:100-106
    addreadout / resetcarry                        reservoircomputer.jl:137-144,207-232

All numerics run in librcb200.so on the B200; this module only marshals arrays (column-major,
features x time) and threads the functional ``ps`` / ``st`` values.  Extension over the reference:
``data`` may be 3-D ``(d_in, T, B)`` for B independent sequences (the reference loops over them).
"""
from __future__ import annotations

import copy
import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from ._lib import ArgumentError, DimensionMismatch, check, lib, ptr


# ---- tiny immutable NamedTuple stand-in ---------------------------------------------------------
class NT:
    """Immutable record with ``merge`` (Julia ``merge(nt, (; k = v))``)."""

    def __init__(self, **kw):
        object.__setattr__(self, "_d", dict(kw))

    def __getattr__(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        raise AttributeError("ps/st records are immutable; use merge()")

    def merge(self, **kw):
        d = dict(self._d)
        d.update(kw)
        return NT(**d)

    def keys(self):
        return self._d.keys()

    def __contains__(self, k):
        return k in self._d

    def __repr__(self):
        return "(" + ", ".join(f"{k} = {type(v).__name__ if isinstance(v, np.ndarray) else v!r}" for k, v in self._d.items()) + ")"

    def __eq__(self, o):
        if not isinstance(o, NT) or self._d.keys() != o._d.keys():
            return False
        for k in self._d:
            a, b = self._d[k], o._d[k]
            if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
                if not np.array_equal(a, b):
                    return False
            elif isinstance(a, np.random.Generator):
                continue
            elif a != b:
                return False
        return True


# ---- activations / modifiers --------------------------------------------------------------------
def identity(x):
    return x


def tanh(x):
    return np.tanh(x)


def tanh_fast(x):
    return np.tanh(x)


def sigmoid(x):
    return 1 / (1 + np.exp(-x))


def relu(x):
    return np.maximum(x, 0)


_ACT_IDS = {identity: L.ACT_IDENTITY, tanh: L.ACT_TANH, np.tanh: L.ACT_TANH, tanh_fast: L.ACT_TANH_FAST,
            sigmoid: L.ACT_SIGMOID, relu: L.ACT_RELU}


def _act_id(f) -> int:
    if f in _ACT_IDS:
        return _ACT_IDS[f]
    raise L.UnsupportedError(f"activation {f!r} is not implemented by the B200 recurrence "
                             "(identity, tanh, tanh_fast, sigmoid, relu)")


class _Modifier:
    kind = 0
    param = 0.0

    def __call__(self, x):  # modifiers are stateless functions in the reference
        x = np.asarray(x)
        vec = x.ndim == 1
        X = np.asfortranarray(x.reshape(x.shape[0], -1))
        out = apply_modifiers((self,), X)
        return out[:, 0] if vec else out

    def __repr__(self):
        return type(self).__name__


class _NLAT1(_Modifier):
    kind = L.MOD_NLAT1


class _NLAT2(_Modifier):
    kind = L.MOD_NLAT2


class _NLAT3(_Modifier):
    kind = L.MOD_NLAT3


class _ExtendedSquare(_Modifier):
    kind = L.MOD_EXTENDED_SQUARE


NLAT1, NLAT2, NLAT3, ExtendedSquare = _NLAT1(), _NLAT2(), _NLAT3(), _ExtendedSquare()


class Pad(_Modifier):  # src/states.jl:73-77
    kind = L.MOD_PAD

    def __init__(self, padding=1.0):
        self.param = float(padding)

    def __repr__(self):
        return f"Pad({self.param})"


class PartialSquare(_Modifier):  # src/states.jl:404-406
    kind = L.MOD_PARTIAL_SQUARE

    def __init__(self, eta):
        self.param = float(eta)


def _as_mods(m) -> tuple:
    if m is None:
        return ()
    if isinstance(m, (tuple, list)):
        return tuple(m)
    return (m,)


def _feature_dims(n: int, mods: Sequence[_Modifier]) -> int:
    for m in mods:
        if m.kind == L.MOD_PAD:
            n += 1
        elif m.kind == L.MOD_EXTENDED_SQUARE:
            n *= 2
    return n


# ---- device context -----------------------------------------------------------------------------
class Device:
    """One GPU + one stream (rcb_ctx).  ``stream`` may be a raw cudaStream_t (e.g. torch's)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        # stream: None -> the library creates its own (non-blocking) stream; a cudaStream_t handle -> that stream.
        # torch reports its default stream as handle 0, which CUDA spells cudaStreamLegacy (0x1) when passed explicitly.
        h = C.c_void_p()
        arg = None if stream is None else C.c_void_p(int(stream) if int(stream) != 0 else 1)
        check(lib().rcb_ctx_create(int(device), arg, C.byref(h)))
        self.handle = h
        self.device = int(device)

    def synchronize(self):
        check(lib().rcb_ctx_synchronize(self.handle))

    def launch_count(self) -> int:
        n = C.c_int64()
        check(lib().rcb_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def kernel_ms(self, which: int) -> float:
        ms = C.c_float()
        check(lib().rcb_ctx_last_kernel_ms(self.handle, which, C.byref(ms)))
        return ms.value

    def sm_count(self) -> int:
        n = C.c_int32()
        check(lib().rcb_ctx_sm_count(self.handle, C.byref(n)))
        return n.value

    def last_refine(self):
        """(refinement passes, |delta W| / |W| of the last one) of the most recent AUTO ridge fit on this context."""
        n, r = C.c_int32(0), C.c_double(0.0)
        check(lib().rcb_ctx_last_refine(self.handle, C.byref(n), C.byref(r)))
        return int(n.value), float(r.value)

    def last_nonfinite(self) -> bool:
        """True if the carry of a recurrence of the most recent call held a NaN / Inf (a diverged reservoir)."""
        f = C.c_int32(0)
        check(lib().rcb_ctx_last_nonfinite(self.handle, C.byref(f)))
        return bool(f.value)

    def last_kernel(self) -> str:
        buf = C.create_string_buffer(128)
        check(lib().rcb_ctx_last_kernel(self.handle, buf, 128))
        return buf.value.decode()

    def close(self):
        if self.handle:
            lib().rcb_ctx_destroy(self.handle)
            self.handle = None


_default = {}


def default_device(device: Optional[int] = None) -> Device:
    import os

    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _default:
        _default[device] = Device(device)
    return _default[device]


def device_count() -> int:
    n = C.c_int32()
    check(lib().rcb_device_count(C.byref(n)))
    return n.value


class Comm:
    """Library-owned NCCL communicator of one rank (rcb_comm).  The 128-byte NCCL id travels over whatever process group the
    host already has (torch.distributed here: plumbing only; MPI.jl / Distributed for a Julia host) — every collective on the
    data path is then issued by the library on the context's stream."""

    def __init__(self, dev: Device, ident, rank: int, world: int):
        h = C.c_void_p()
        check(lib().rcb_comm_create(dev.handle, ptr(ident), int(rank), int(world), C.byref(h)))
        self.handle, self.dev, self.rank, self.world = h, dev, int(rank), int(world)

    def allreduce_gc(self, GC, d_out: int):
        F = GC.shape[0] if isinstance(GC, np.ndarray) else GC.shape[1]
        check(lib().rcb_allreduce_gc(self.handle, ptr(GC), F, d_out))
        return GC

    def allgather(self, send: np.ndarray):
        send = np.ascontiguousarray(send)
        recv = np.empty((self.world,) + send.shape, dtype=send.dtype)
        check(lib().rcb_comm_allgather(self.handle, ptr(send), ptr(recv), send.nbytes))
        return recv

    def close(self):
        if self.handle:
            lib().rcb_comm_destroy(self.handle)
            self.handle = None


_comms = {}


def comm_for(dev: Optional[Device] = None, group=None) -> Optional[Comm]:
    """The rcb_comm that mirrors a torch.distributed process group on this rank's device (None for a single rank).
    Collective: every rank of the group must call it the first time."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    dev = dev or default_device()
    key = (dev.device, id(group) if group is not None else None)
    if key not in _comms:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = np.zeros(L.COMM_ID_BYTES, dtype=np.uint8)
        if rank == 0:
            check(lib().rcb_comm_unique_id(ptr(ident)))
        src = dist.get_global_rank(group, 0) if group is not None else 0
        t = torch.from_numpy(ident)
        if dist.get_backend(group) == "nccl":
            t = t.cuda(dev.device)
            dist.broadcast(t, src=src, group=group)
            ident = t.cpu().numpy()
        else:
            dist.broadcast(t, src=src, group=group)
        _comms[key] = Comm(dev, ident, rank, world)
    return _comms[key]


# ---- initializers (host-side producers) ------------------------------------
class _Init:
    """Curried initializer: ``rand_sparse(radius=1.2, sparsity=6/300)`` returns a callable
    ``(rng, *dims) -> array`` (src/initializer_utils.jl:39-73).  Values come from NumPy streams:
    parity with Julia's RNG streams is unpinned upstream."""

    def __init__(self, fn, **kw):
        self.fn, self.kw = fn, kw

    def __call__(self, *args, **kw):
        if args and isinstance(args[0], np.random.Generator):
            k = dict(self.kw)
            k.update(kw)
            return self.fn(*args, **k)
        k = dict(self.kw)
        k.update(kw)
        return _Init(self.fn, **k)


def _spectral_radius(W) -> float:
    import scipy.sparse as sp

    n = W.shape[0]
    if n <= 1500:  # dense eigvals, src/inits/inits_components.jl:127
        Wd = W.toarray() if sp.issparse(W) else np.asarray(W)
        return float(np.max(np.abs(np.linalg.eigvals(Wd.astype(np.float64)))))
    return spectral_radius(W)  # above that: Arnoldi on the GPU (rcb_spectral_radius_*)


def _rand_sparse(rng, *dims, radius=1.0, sparsity=0.1, std=1.0, return_sparse=False):
    """src/inits/inits_reservoir.jl:55-67: sparse_init(sparsity = 1 - s) then scale_radius!."""
    import scipy.sparse as sp

    n, m = dims
    if n != m:
        raise DimensionMismatch(f"reservoir matrix must be square, got {dims}")
    k = n - int(math.ceil((1.0 - float(sparsity)) * n))
    rows = np.argsort(rng.random((n, n)), axis=0)[:k, :] if n <= 2048 else np.stack(
        [rng.choice(n, size=k, replace=False) for _ in range(n)], axis=1)
    rows = np.sort(rows, axis=0)
    vals = (rng.standard_normal((k, n)) * std).astype(np.float32)
    W = sp.csc_matrix((vals.T.ravel(), rows.T.ravel(), np.arange(0, n * k + 1, k)), shape=(n, n))
    rho = _spectral_radius(W)
    W = (W * (np.float32(radius) / np.float32(rho))).astype(np.float32).tocsc()
    if not np.all(np.isfinite(W.data)):
        raise ArgumentError("Created matrix contains invalid values (NaN/Inf)")
    return W if return_sparse else np.asfortranarray(W.toarray().astype(np.float32))


def _scaled_rand(rng, *dims, scaling=0.1):
    """src/inits/inits_input.jl:80-88 + apply_scale! (inits_components.jl:1-4)."""
    x = rng.random(dims, dtype=np.float32)
    return np.asfortranarray((x - np.float32(0.5)) * (np.float32(2) * np.float32(scaling)))


rand_sparse = _Init(_rand_sparse)
scaled_rand = _Init(_scaled_rand)
randn32 = _Init(lambda rng, *dims: rng.standard_normal(dims).astype(np.float32))
rand32 = _Init(lambda rng, *dims: rng.random(dims, dtype=np.float32))
zeros32 = _Init(lambda rng, *dims: np.zeros(dims, np.float32))


# ---- layers / models ------------------------------------------------------------------------------
class ESNCell:
    """src/layers/esn_cell.jl:114-148."""

    def __init__(self, in_dims, out_dims, activation=tanh_fast, *, use_bias=False, init_bias=zeros32,
                 init_reservoir=rand_sparse, init_input=scaled_rand, init_state=randn32, leak_coefficient=1.0):
        if isinstance(leak_coefficient, (list, tuple, np.ndarray)):
            leak_coefficient = np.asarray(leak_coefficient)
            if leak_coefficient.shape[0] != out_dims:
                raise DimensionMismatch(f"leak_coefficient must have length out_dims={out_dims}, got {leak_coefficient.shape[0]}.")
        self.in_dims, self.out_dims, self.activation = int(in_dims), int(out_dims), activation
        self.use_bias, self.init_bias, self.init_reservoir = bool(use_bias), init_bias, init_reservoir
        self.init_input, self.init_state, self.leak_coefficient = init_input, init_state, leak_coefficient

    def initialparameters(self, rng):  # esn_cell.jl:150-159 — draw order: input, reservoir, bias
        ps = dict(input_matrix=self.init_input(rng, self.out_dims, self.in_dims),
                  reservoir_matrix=self.init_reservoir(rng, self.out_dims, self.out_dims))
        if self.use_bias:
            ps["bias"] = self.init_bias(rng, self.out_dims)
        return NT(**ps)

    def initialstates(self, rng):  # esn_cell.jl:161-163, lux_layers.jl:207-210
        rng.random()
        return NT(cell=NT(rng=copy.deepcopy(rng)), carry=None)


def _orthogonal(rng, *dims, gain=1.0):
    """WeightInitializers.orthogonal (un-vendored; values unpinned upstream): Q of a Gaussian matrix,
    columns sign-fixed by diag(R)."""
    n, m = dims
    q, r = np.linalg.qr(rng.standard_normal((max(n, m), min(n, m))))
    q = q * np.sign(np.diag(r))
    q = q if n >= m else q.T
    return np.asfortranarray((gain * q[:n, :m]).astype(np.float32))


orthogonal = _Init(_orthogonal)


class _SkipCell(ESNCell):
    """Cells whose skip path is a dense orthogonal operator: ps gains ``orthogonal_matrix``, drawn after
    ``reservoir_matrix`` (src/layers/es2n_cell.jl:93-104, src/layers/resesn_cell.jl:100-110)."""

    def __init__(self, in_dims, out_dims, activation=tanh_fast, *, init_orthogonal=orthogonal, **kw):
        kw.pop("leak_coefficient", None)
        super().__init__(in_dims, out_dims, activation, **kw)
        self.init_orthogonal = init_orthogonal

    def initialparameters(self, rng):
        ps = dict(input_matrix=self.init_input(rng, self.out_dims, self.in_dims),
                  reservoir_matrix=self.init_reservoir(rng, self.out_dims, self.out_dims),
                  orthogonal_matrix=self.init_orthogonal(rng, self.out_dims, self.out_dims))
        if self.use_bias:
            ps["bias"] = self.init_bias(rng, self.out_dims)
        return NT(**ps)


class ES2NCell(_SkipCell):
    """src/layers/es2n_cell.jl:66-91: h' = (1-proximity) O h + proximity phi(W_in u + W h + b)."""

    cell_kind = L.CELL_ES2N

    def __init__(self, in_dims, out_dims, activation=tanh_fast, *, proximity=1.0, **kw):
        super().__init__(in_dims, out_dims, activation, **kw)
        self.proximity = float(proximity)

    @property
    def cell_params(self):
        return (self.proximity, 0.0)


class ResESNCell(_SkipCell):
    """src/layers/resesn_cell.jl:70-98: h' = alpha O h + beta phi(W_in u + W h + b)."""

    cell_kind = L.CELL_RESESN

    def __init__(self, in_dims, out_dims, activation=tanh_fast, *, alpha=1.0, beta=1.0, **kw):
        super().__init__(in_dims, out_dims, activation, **kw)
        self.alpha, self.beta = float(alpha), float(beta)

    @property
    def cell_params(self):
        return (self.alpha, self.beta)


class EuSNCell(ESNCell):
    """src/layers/eusn_cell.jl:68-93: the leaky ESN step on W - W' - diffusion I (:95-106,120-123)."""

    cell_kind = L.CELL_EUSN

    def __init__(self, in_dims, out_dims, activation=tanh_fast, *, leak_coefficient=1.0, diffusion=1.0, **kw):
        super().__init__(in_dims, out_dims, activation, leak_coefficient=float(leak_coefficient), **kw)
        self.diffusion = float(diffusion)

    @property
    def cell_params(self):
        return (self.diffusion, 0.0)


class LinearReadout:
    """src/layers/basic.jl:129-154."""

    def __init__(self, in_dims, out_dims, activation=identity, *, init_weight=rand32, init_bias=rand32, use_bias=False,
                 include_collect=True):
        self.in_dims, self.out_dims, self.activation = int(in_dims), int(out_dims), activation
        self.init_weight, self.init_bias, self.use_bias = init_weight, init_bias, bool(use_bias)
        self.include_collect = bool(include_collect)

    def initialparameters(self, rng):  # basic.jl:156-164
        ps = dict(weight=self.init_weight(rng, self.out_dims, self.in_dims))
        if self.use_bias:
            ps["bias"] = self.init_bias(rng, self.out_dims)
        return NT(**ps)


class ESN:
    """ESN(in_dims, res_dims, out_dims, activation=tanh; readout_activation, state_modifiers, cell kwargs...)
    src/models/esn.jl:93-106."""

    def __init__(self, in_dims, res_dims, out_dims, activation=tanh, *, readout_activation=identity,
                 state_modifiers=(), **cell_kw):
        self.reservoir = ESNCell(in_dims, res_dims, activation, **cell_kw)
        self.state_modifiers = _as_mods(state_modifiers)
        self.readout = LinearReadout(res_dims, out_dims, readout_activation)
        self._cache = {}

    @property
    def cells(self):
        return (self.reservoir,)

    @property
    def layer_modifiers(self):
        return (self.state_modifiers,)


class _SingleCellModel(ESN):
    cell_type = ESNCell

    def __init__(self, in_dims, res_dims, out_dims, activation=tanh, *, readout_activation=identity,
                 state_modifiers=(), **cell_kw):
        self.reservoir = self.cell_type(in_dims, res_dims, activation, **cell_kw)
        self.state_modifiers = _as_mods(state_modifiers)
        self.readout = LinearReadout(res_dims, out_dims, readout_activation)
        self._cache = {}


class ES2N(_SingleCellModel):
    """src/models/es2n.jl:86-99."""
    cell_type = ES2NCell


class ResESN(_SingleCellModel):
    """src/models/resesn.jl:89-102."""
    cell_type = ResESNCell


class EuSN(_SingleCellModel):
    """src/models/eusn.jl:84-97."""
    cell_type = EuSNCell


class DeepESN:
    """DeepESN(in_dims, res_dims::Vector, out_dims, activation=tanh; ...) src/models/esn_deep.jl:102-154.
    Per-layer kwargs broadcast like __asvec (src/models/esn_generics.jl:51-79)."""

    def __init__(self, in_dims, res_dims, out_dims, activation=tanh, *, depth=2, leak_coefficient=1.0,
                 init_reservoir=rand_sparse, init_input=scaled_rand, init_bias=zeros32, init_state=randn32,
                 use_bias=False, state_modifiers=(), readout_activation=identity):
        if isinstance(res_dims, int):
            res_dims = [res_dims] * depth
        n = len(res_dims)

        def asvec(x, allow_seq=True):
            if allow_seq and isinstance(x, (tuple, list)):
                if len(x) == n:
                    return list(x)
                if len(x) == 1:
                    return [x[0]] * n
                if len(x) == 0:
                    return [None] * n
                raise DimensionMismatch(f"Expected length {n} or 1, got {len(x)}")
            return [x] * n

        acts, leaks, ires, iinp = asvec(activation), asvec(leak_coefficient), asvec(init_reservoir), asvec(init_input)
        ibias, istate, ub, mods = asvec(init_bias), asvec(init_state), asvec(use_bias), asvec(state_modifiers)
        cells, lmods, prev = [], [], int(in_dims)
        for i in range(n):
            cells.append(ESNCell(prev, res_dims[i], acts[i], use_bias=ub[i], init_bias=ibias[i], init_reservoir=ires[i],
                                 init_input=iinp[i], init_state=istate[i], leak_coefficient=leaks[i]))
            lm = _as_mods(mods[i])
            lmods.append(lm)
            prev = _feature_dims(res_dims[i], lm)  # the next cell's in_dims is declared as res_dims[i] upstream
            if prev != res_dims[i] and i + 1 < n:
                raise DimensionMismatch("size-changing modifiers between DeepESN layers change the next layer's input size")
        self._cells, self._mods = tuple(cells), tuple(lmods)
        self.readout = LinearReadout(res_dims[-1], out_dims, readout_activation)
        self._cache = {}

    @property
    def cells(self):
        return self._cells

    @property
    def layer_modifiers(self):
        return self._mods


def _is_deep(rc) -> bool:
    return isinstance(rc, DeepESN)


def setup(rng: np.random.Generator, rc):
    """LuxCore.setup(rng, model) -> (ps, st).  src/reservoircomputer.jl:51-63, esn_deep.jl:156-188."""
    if isinstance(rc, _DelayModel):
        return _setup_delay(rng, rc)
    if isinstance(rc, ReservoirChain):
        return _setup_chain(rng, rc)
    if _is_deep(rc):
        ps_cells = tuple(c.initialparameters(rng) for c in rc.cells)
        ps = NT(cells=ps_cells, state_modifiers=tuple(() for _ in rc.cells), readout=rc.readout.initialparameters(rng))
        st = NT(cells=tuple(c.initialstates(rng) for c in rc.cells), state_modifiers=tuple(() for _ in rc.cells), readout=NT())
    else:
        ps = NT(reservoir=rc.reservoir.initialparameters(rng), state_modifiers=tuple(NT() for _ in rc.state_modifiers),
                readout=rc.readout.initialparameters(rng))
        st = NT(reservoir=rc.reservoir.initialstates(rng), state_modifiers=tuple(NT() for _ in rc.state_modifiers), readout=NT())
    return ps, st


def _cell_ps(rc, ps):
    return ps.cells if _is_deep(rc) else (ps.reservoir,)


def _cell_st(rc, st):
    return st.cells if _is_deep(rc) else (st.reservoir,)


def _with_cell_st(rc, st, new):
    return st.merge(cells=tuple(new)) if _is_deep(rc) else st.merge(reservoir=new[0])


# ---- device handle --------------------------------------------------------------------------------
class _Handle:
    def __init__(self, rc, ps, dtype, dev: Device):
        import scipy.sparse as sp

        cells, lmods = rc.cells, rc.layer_modifiers
        nl = len(cells)
        if nl > L.MAX_LAYERS:
            raise L.UnsupportedError(f"at most {L.MAX_LAYERS} layers")
        layers = (L.LayerDesc * nl)()
        for i, c in enumerate(cells):
            d = layers[i]
            d.res_dims, d.activation, d.use_bias = c.out_dims, _act_id(c.activation), int(c.use_bias)
            vec = isinstance(c.leak_coefficient, np.ndarray)
            d.leak_is_vector = int(vec)
            d.leak_scalar = 0.0 if vec else float(c.leak_coefficient)
            if len(lmods[i]) > L.MAX_MODIFIERS:
                raise L.UnsupportedError(f"at most {L.MAX_MODIFIERS} state modifiers per layer")
            d.n_modifiers = len(lmods[i])
            for j, m in enumerate(lmods[i]):
                if not isinstance(m, _Modifier):
                    raise L.UnsupportedError(f"state modifier {m!r} has no B200 implementation")
                d.modifier_kind[j], d.modifier_param[j] = m.kind, float(m.param)
        desc = L.EsnDesc(nl, cells[0].in_dims, rc.readout.out_dims, _act_id(rc.readout.activation),
                         int(rc.readout.use_bias), L.dtype_id(dtype))
        h = C.c_void_p()
        check(lib().rcb_esn_create(dev.handle, C.byref(desc), layers, C.byref(h)))
        self.handle, self.dev, self.keep = h, dev, []
        for i, (c, p) in enumerate(zip(cells, _cell_ps(rc, ps))):
            kind = getattr(c, "cell_kind", L.CELL_ESN)
            if kind != L.CELL_ESN:
                check(lib().rcb_esn_set_cell(h, i, kind, *c.cell_params))
            Win = np.asfortranarray(p.input_matrix, dtype=np.float32)
            bias = np.ascontiguousarray(p.bias, dtype=np.float32) if c.use_bias else None
            W = p.reservoir_matrix
            if sp.issparse(W):  # SparseMatrixCSC hand-off (ext/RCSparseArraysExt.jl:5-7): 1-based Int64
                Wc = sp.csc_matrix(W)
                colptr = (Wc.indptr.astype(np.int64) + 1)
                rowval = (Wc.indices.astype(np.int64) + 1)
                nz = np.ascontiguousarray(Wc.data, dtype=np.float32)
                check(lib().rcb_esn_set_weights_csc(h, i, ptr(colptr), ptr(rowval), ptr(nz), ptr(Win), Win.shape[0], ptr(bias)))
            else:
                Wd = np.asfortranarray(W, dtype=np.float32)
                check(lib().rcb_esn_set_weights_dense(h, i, ptr(Wd), Wd.shape[0], ptr(Win), Win.shape[0], ptr(bias)))
            if isinstance(c.leak_coefficient, np.ndarray):
                lv = np.ascontiguousarray(c.leak_coefficient, dtype=np.float32)
                check(lib().rcb_esn_set_leak_vector(h, i, ptr(lv)))
            if kind in (L.CELL_ES2N, L.CELL_RESESN):
                O = np.asfortranarray(p.orthogonal_matrix, dtype=np.float32)
                check(lib().rcb_esn_set_orthogonal_dense(h, i, ptr(O), O.shape[0]))
        F, sN = C.c_int32(), C.c_int32()
        check(lib().rcb_esn_feature_dims(h, C.byref(F)))
        check(lib().rcb_esn_carry_dims(h, C.byref(sN)))
        self.F, self.sumN = F.value, sN.value
        self.readout_key = None

    def set_readout(self, ps_ro, use_bias):
        key = (_fingerprint(ps_ro.weight), _fingerprint(ps_ro.bias) if use_bias else None)
        if key == self.readout_key:
            return
        W = np.asfortranarray(ps_ro.weight, dtype=np.float64)
        if W.shape[1] != self.F:
            raise DimensionMismatch(f"readout weight has {W.shape[1]} columns but the model produces {self.F} features")
        b = np.ascontiguousarray(ps_ro.bias, dtype=np.float64) if use_bias else None
        check(lib().rcb_esn_set_readout(self.handle, ptr(W), ptr(b), 1))
        self.readout_key = key
        self.keep = [ps_ro.weight, ps_ro.bias if use_bias else None]

    def __del__(self):
        try:
            if self.handle:
                lib().rcb_esn_destroy(self.handle)
        except Exception:
            pass


def _fingerprint(a):
    """Identity + a cheap content fingerprint of a parameter array: the device handle is rebuilt when a weight is
    replaced (``ps.reservoir.merge(bias=new)``) AND when one is edited in place (``W *= 0.9``) — the reference reads ``ps``
    afresh on every call.  A strided sample of <= 4096 entries plus the end points is hashed: O(1) per call, and an edit
    that scales or shifts the array always shows (a single-entry edit between the sample points may not — call
    ``rcb200.invalidate(rc)`` after surgical edits)."""
    import scipy.sparse as sp

    if a is None:
        return None
    if sp.issparse(a):
        return (id(a), a.shape, a.nnz) + _fingerprint(a.data)[1:] + _fingerprint(a.indices)[1:]
    a = np.asarray(a)
    flat = a.ravel(order="K")  # a view for contiguous arrays
    step = max(1, flat.size // 4096)
    sample = np.ascontiguousarray(flat[::step])
    return (id(a), a.shape, a.dtype.str, hash(sample.tobytes()), float(flat[-1]) if flat.size else 0.0)


def invalidate(rc):
    """Drop the cached device handle of `rc` (weights are uploaded again on the next call)."""
    rc._cache.clear()


def _handle(rc, ps, dtype, dev: Optional[Device] = None) -> _Handle:
    dev = dev or default_device()
    cps = _cell_ps(rc, ps)
    key = (np.dtype(dtype).str, dev.device) + tuple(
        _fingerprint(getattr(p, k)) for p in cps for k in ("input_matrix", "reservoir_matrix", "orthogonal_matrix", "bias") if k in p)
    key += tuple((getattr(c, "cell_kind", L.CELL_ESN), getattr(c, "cell_params", None),
                  _fingerprint(c.leak_coefficient) if isinstance(c.leak_coefficient, np.ndarray) else c.leak_coefficient,
                  _act_id(c.activation), c.use_bias) for c in rc.cells)
    ent = rc._cache.get(key)
    if ent is None:
        ent = (_Handle(rc, ps, dtype, dev), cps)  # keep ps alive so ids are not recycled
        rc._cache.clear()
        rc._cache[key] = ent
    return ent[0]


def _initial_carry(rc, st, dtype, B):
    """h0 for every layer: the stored carry, or a fresh init_state draw from the replicated rng
    (src/layers/esn_cell.jl:169-173, src/layers/lux_layers.jl:35-46,212-216).  Returns
    (packed (sumN, B) array, new per-layer cell states with advanced rngs)."""
    parts, new_st = [], []
    for c, s in zip(rc.cells, _cell_st(rc, st)):
        if s.carry is None:
            rng = copy.deepcopy(s.cell.rng)
            h = np.asarray(c.init_state(rng, c.out_dims, B), dtype=dtype).reshape(c.out_dims, B)
            new_st.append(s.merge(cell=NT(rng=rng)))
        else:
            h = np.asarray(s.carry[0], dtype=dtype).reshape(c.out_dims, -1)
            if h.shape[1] != B:
                raise DimensionMismatch(f"carry holds {h.shape[1]} sequences, data has {B}")
            new_st.append(s)
        parts.append(h)
    return np.asfortranarray(np.concatenate(parts, axis=0)), new_st


def _store_carry(rc, st, cell_sts, hT, vector: bool):
    out, off = [], 0
    for c, s in zip(rc.cells, cell_sts):
        h = hT[off:off + c.out_dims]
        off += c.out_dims
        out.append(s.merge(carry=(np.ascontiguousarray(h[:, 0]) if vector else np.asfortranarray(h),)))
    return _with_cell_st(rc, st, out)


def _as3d(data):
    data = np.asarray(data)
    if data.ndim == 1:
        data = data[:, None]
    vec = data.ndim == 2
    d3 = data[:, :, None] if vec else data
    return np.asfortranarray(d3), vec


# ---- collectstates / train / predict ---------------------------------------------------------------
def collectstates(rc, data, ps, st, device: Optional[Device] = None):
    """states (F, T[, B]) with eltype(data), and the new st (final carry stored).
    src/reservoircomputer.jl:96-133, src/models/esn_deep.jl:261-291."""
    if isinstance(rc, _DelayModel):
        return _collect_delay(rc, data, ps, st, device)
    if isinstance(rc, ReservoirChain):
        return _collect_chain(rc, data, ps, st, device)
    d3, vec = _as3d(data)
    if d3.shape[1] < 1:
        raise ArgumentError(f"collectstates data must have at least one column, got {d3.shape[1]}.")
    return collectstates_core(rc, data, ps, st, device)


class RidgeRegression:
    """src/train.jl:24-34.  ``RidgeRegression(np.float32, lam)`` keeps a Float32 λ."""

    def __init__(self, *args):
        if len(args) == 0:
            self.reg, self.reg_dtype = 0.0, np.dtype(np.float64)
        elif len(args) == 1:
            self.reg = args[0]
            self.reg_dtype = np.dtype(type(args[0])) if isinstance(args[0], np.floating) else np.dtype(np.float64)
        else:
            self.reg_dtype = np.dtype(args[0])
            self.reg = self.reg_dtype.type(args[1])


StandardRidge = RidgeRegression  # src/deprecated.jl:4


class QRFactorization:
    """Solver facades (src/train.jl:85-98).  ``solver=None`` is the library's RCB_GRAM_AUTO: Cholesky on the Gram matrix,
    escalating by itself to the Householder QR of the augmented system (λ == 0, or a failed factorisation).  Passing
    QRFactorization() / QRSolver() explicitly asks for that QR outright (RCB_FIT_QR, the reference's own algorithm);
    NormalCholeskyFactorization() pins the normal equations on the exact FP64 Gram."""


class QRSolver:
    pass


class NormalCholeskyFactorization:
    """LinearSolve.NormalCholeskyFactorization, one of the algorithms train.jl:187-189 names."""


def _promote(objective, *arrays):
    T = objective.reg_dtype
    for a in arrays:
        T = np.promote_types(T, a.dtype)
    if not np.issubdtype(T, np.floating):
        T = np.dtype(np.float64)
    return T


def _check_solver(solver, gram_mode=L.GRAM_AUTO):
    """Validates `solver` (train.jl:158-178) and returns the library's fit mode for it."""
    if solver is None:
        return gram_mode
    if isinstance(solver, (QRFactorization, QRSolver)):
        return L.FIT_QR
    if isinstance(solver, NormalCholeskyFactorization):
        return L.GRAM_FP64
    raise ArgumentError(f"solver {type(solver).__name__} is not supported. Pass QRFactorization(), QRSolver(), "
                        "or a documented LinearSolve.jl algorithm instead.")


def fit_readout(objective: RidgeRegression, states, targets, solver=None, device: Optional[Device] = None, gram_mode=L.GRAM_AUTO):
    """__fit_readout(::RidgeRegression, states (F, S), targets (d_out, S)) -> W (d_out, F).
    src/train.jl:100-198."""
    gram_mode = _check_solver(solver, gram_mode)
    Z, Y = np.asarray(states), np.asarray(targets)
    if Z.dtype not in (np.float32, np.float64):
        Z = Z.astype(np.float64)
    if Y.dtype not in (np.float32, np.float64):
        Y = Y.astype(np.float64)
    Z, Y = np.asfortranarray(Z), np.asfortranarray(Y)
    dev = device or default_device()
    F, d_out = Z.shape[0], Y.shape[0]
    W = np.empty((d_out, F), dtype=np.float64, order="F")
    check(lib().rcb_fit_readout(dev.handle, ptr(Z), L.dtype_id(Z.dtype), F, Z.shape[1], max(F, 1), ptr(Y), L.dtype_id(Y.dtype),
                                d_out, Y.shape[1], max(d_out, 1), float(objective.reg), gram_mode, ptr(W)))
    return W.astype(_promote(objective, Z, Y))


def gram(states, targets, device: Optional[Device] = None, gram_mode=L.GRAM_AUTO, into=None):
    """[G | C] = [Z Z' | Z Y'] as an (F, F + d_out) Float64 array (lower triangle of G), the accumulator that
    ranks all-reduce before ridge_solve (rcb_gram_accumulate).  `into` adds onto an existing accumulator."""
    Z, Y = np.asfortranarray(states), np.asfortranarray(targets)
    dev = device or default_device()
    F, d_out = Z.shape[0], Y.shape[0]
    GC = np.zeros((F, F + d_out), dtype=np.float64, order="F") if into is None else into
    check(lib().rcb_gram_accumulate(dev.handle, ptr(Z), L.dtype_id(Z.dtype), F, Z.shape[1], max(F, 1), ptr(Y),
                                    L.dtype_id(Y.dtype), d_out, max(d_out, 1), gram_mode, 1 if into is None else 0, ptr(GC)))
    return GC


def ridge_solve(GC, d_out: int, reg: float, device: Optional[Device] = None):
    """(G + reg I) W' = C from a packed [G | C] (rcb_ridge_solve) -> W (d_out, F) Float64."""
    dev = device or default_device()
    GC = np.asfortranarray(GC, dtype=np.float64)
    F = GC.shape[0]
    W = np.empty((d_out, F), dtype=np.float64, order="F")
    check(lib().rcb_ridge_solve(dev.handle, ptr(GC), F, d_out, float(reg), ptr(W)))
    return W


def qr_merge(RQ, other, d_out: int, device: Optional[Device] = None):
    """RQ <- QR of [RQ; other] (rcb_qr_merge): merges another rank's partial factor [R | Q'b] into this one, in place."""
    dev = device or default_device()
    F = RQ.shape[0] if isinstance(RQ, np.ndarray) else RQ.shape[1]
    check(lib().rcb_qr_merge(dev.handle, ptr(RQ), ptr(other), F, d_out))
    return RQ


def qr_solve(RQ, d_out: int, reg: float, device: Optional[Device] = None):
    """W (d_out, F) from a factor [R | Q'b] accumulated with gram_mode = FIT_QR (rcb_qr_solve)."""
    dev = device or default_device()
    F = RQ.shape[0] if isinstance(RQ, np.ndarray) else RQ.shape[1]
    W = np.empty((d_out, F), dtype=np.float64, order="F")
    check(lib().rcb_qr_solve(dev.handle, ptr(RQ), F, d_out, float(reg), ptr(W)))
    return W


def addreadout(rc, output_matrix, ps, st):
    """addreadout!(rc, W, ps, st): functional update of ps.readout.weight (reservoircomputer.jl:137-144; the chain
    variant replaces the weight of its readout layer, train.jl:253-305)."""
    if isinstance(rc, ReservoirChain):
        name = rc.names()[-1]
        return ps.merge(**{name: getattr(ps, name).merge(weight=output_matrix)}), st
    return ps.merge(readout=ps.readout.merge(weight=output_matrix)), st


def train(rc, train_data, target_data, ps, st, *, objective: RidgeRegression = None, solver=None, washout: int = 0,
          return_states: bool = False, device: Optional[Device] = None, gram_mode=L.GRAM_AUTO):
    """train(rc, X, Y, ps, st; objective = RidgeRegression(0.0), solver, washout, return_states)
    -> (ps, st) or ((ps, st), states_wo).  src/train.jl:232-251."""
    objective = objective or RidgeRegression(0.0)
    gram_mode = _check_solver(solver, gram_mode)
    if isinstance(rc, (_DelayModel, ReservoirChain)):  # train.jl:240-250 step by step: the features are assembled outside the fused rcb_train
        states, st2 = collectstates(rc, train_data, ps, st, device)
        tgt = np.asarray(target_data)
        if washout < 0:
            raise ArgumentError(f"washout must be ≥ 0, got {washout}")
        if washout >= states.shape[1]:
            raise ArgumentError(f"washout={washout} is ≥ number of time steps={states.shape[1]}")
        sw, tw = states[:, washout:], tgt[:, washout:]
        if sw.ndim == 3:  # pooled over the sequences
            sw, tw = sw.reshape(sw.shape[0], -1, order="F"), tw.reshape(tw.shape[0], -1, order="F")
        W = fit_readout(objective, sw, tw, device=device, gram_mode=gram_mode)
        ps2, st2 = addreadout(rc, W, ps, st2)
        return ((ps2, st2), states[:, washout:]) if return_states else (ps2, st2)
    d3, vec = _as3d(train_data)
    y3, _ = _as3d(target_data)
    if y3.dtype not in (np.float32, np.float64):
        y3 = np.asfortranarray(y3.astype(np.float64))
    dtype = d3.dtype
    _, T, B = d3.shape
    if T < 1:
        raise ArgumentError(f"collectstates data must have at least one column, got {T}.")
    if washout < 0:
        raise ArgumentError(f"washout must be ≥ 0, got {washout}")
    if washout >= T:
        raise ArgumentError(f"washout={washout} is ≥ number of time steps={T}")
    if y3.shape[1] != T or y3.shape[2] != B:
        raise DimensionMismatch(f"states has {T - washout} samples, targets has {y3.shape[1] - washout}")
    h = _handle(rc, ps, dtype, device)
    h0, cell_sts = _initial_carry(rc, st, dtype, B)
    d_out = y3.shape[0]
    if d_out != rc.readout.out_dims:
        # the declared readout size is ignored by train (basic.jl:121-124); the kernel handle is sized by out_dims
        raise DimensionMismatch(f"targets have {d_out} rows but the readout was declared with out_dims={rc.readout.out_dims}")
    W = np.empty((d_out, h.F), dtype=np.float64, order="F")
    hT = np.empty((h.sumN, B), dtype=dtype, order="F")
    states = np.empty((h.F, T, B), dtype=dtype, order="F") if return_states else None
    check(lib().rcb_train(h.handle, ptr(d3), ptr(y3), L.dtype_id(y3.dtype), T, B, int(washout), float(objective.reg), gram_mode,
                          ptr(h0), ptr(W), ptr(states), ptr(hT)))
    W = W.astype(_promote(objective, d3, y3))
    ps2, st2 = addreadout(rc, W, ps, _store_carry(rc, st, cell_sts, hT, vec))
    if return_states:
        sw = states[:, washout:, :]
        return (ps2, st2), (sw[:, :, 0] if vec else sw)
    return ps2, st2


# ---- multi-GPU: one process per GPU, partial [G | C] summed by ONE all-reduce -----------------
def shard_range(n_items: int, rank: int, world: int):
    """Contiguous balanced shard [lo, hi) of n_items independent units (sequences / ensemble members)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def time_chunks(T: int, washout: int, rank: int, world: int, overlap: int):
    """BASELINE config 2: one long sequence, time-chunked after the washout.  Rank r owns the samples
    [own_lo, own_hi) of [washout, T); it runs the recurrence from start = max(0, own_lo - overlap) with a zero (rank > 0)
    or the given (rank 0) initial state and discards the first own_lo - start columns as its local washout — the
    echo-state property re-synchronises the state over the overlap.  Returns (start, own_lo, own_hi)."""
    lo, hi = shard_range(T - washout, rank, world)
    own_lo, own_hi = washout + lo, washout + hi
    start = 0 if rank == 0 else max(0, own_lo - overlap)
    return start, own_lo, own_hi


def _allreduce_gc(GC, d_out, group, device, use_lib=True):
    """Sum the packed [G | C] over the ranks: only the lower triangle of G (+ C) travels — half the bytes of the buffer
    (rcb_gc_pack / rcb_gc_unpack; the NumPy packing is the CPU-test stand-in with the same layout)."""
    import torch
    import torch.distributed as dist

    F = GC.shape[0]
    n = F * (F + 1) // 2 + F * d_out
    if use_lib:
        dev = device or default_device()
        packed = np.empty(n, dtype=np.float64)
        check(lib().rcb_gc_pack(dev.handle, ptr(GC), F, d_out, ptr(packed)))
    else:
        packed = np.concatenate([GC[j:, j] for j in range(F)] + [GC[:, F + d] for d in range(d_out)])
    t = torch.from_numpy(packed)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    packed = t.cpu().numpy()
    out = np.zeros_like(GC, order="F")
    if use_lib:
        check(lib().rcb_gc_unpack(dev.handle, ptr(packed), F, d_out, ptr(out)))
    else:
        o = 0
        for j in range(F):
            out[j:, j] = packed[o:o + F - j]
            o += F - j
        for d in range(d_out):
            out[:, F + d] = packed[o:o + F]
            o += F
    return out


def _partial_gram_cuda(rc, ps, st, d3, y3, washout, device, gram_mode):
    h = _handle(rc, ps, d3.dtype, device)
    _, T, B = d3.shape
    h0, cell_sts = _initial_carry(rc, st, d3.dtype, B)
    GC = np.empty((h.F, h.F + y3.shape[0]), dtype=np.float64, order="F")
    hT = np.empty((h.sumN, B), dtype=d3.dtype, order="F")
    check(lib().rcb_train_partial(h.handle, ptr(d3), ptr(y3), L.dtype_id(y3.dtype), T, B, int(washout), gram_mode, ptr(h0), ptr(GC),
                                  ptr(hT)))
    return GC, hT, cell_sts


def train_sharded(rc, train_data, target_data, ps, st, *, objective: RidgeRegression = None, solver=None, washout: int = 0, group=None,
                  device: Optional[Device] = None, gram_mode=L.GRAM_AUTO, _partial=None, _solve=None):
    """train() for a pooled readout over sequences sharded across ranks: every rank passes ITS shard (d_in, T, B_r),
    accumulates its partial [G | C] on its GPU (rcb_train_partial), the partials are summed by one all-reduce over the
    process group (NCCL over NVLink; gloo in the CPU tests) and every rank solves the same FP64 system
    (rcb_ridge_solve), so all ranks end with identical readout weights.  The recurrence itself needs no collective.
    `_partial` / `_solve` are test seams for the CPU-only gloo tests of this host logic."""
    import torch
    import torch.distributed as dist

    objective = objective or RidgeRegression(0.0)
    d3, vec = _as3d(train_data)
    y3, _ = _as3d(target_data)
    if y3.dtype not in (np.float32, np.float64):
        y3 = np.asfortranarray(y3.astype(np.float64))
    _, T, B = d3.shape
    if washout < 0:
        raise ArgumentError(f"washout must be ≥ 0, got {washout}")
    if washout >= T:
        raise ArgumentError(f"washout={washout} is ≥ number of time steps={T}")
    if y3.shape[1] != T or y3.shape[2] != B:
        raise DimensionMismatch(f"states has {T - washout} samples, targets has {y3.shape[1] - washout}")
    d_out = y3.shape[0]
    if _partial is None and _solve is None:
        # the CUDA path: partial accumulation, the NCCL exchange and the replicated solve are ONE library call
        # (rcb_train_sharded); torch.distributed only carried the communicator id (comm_for)
        dev = device or default_device()
        h = _handle(rc, ps, d3.dtype, dev)
        F = h.F
        h0, cell_sts = _initial_carry(rc, st, d3.dtype, B)
        hT = np.empty((h.sumN, B), dtype=d3.dtype, order="F")
        comm = comm_for(dev, group)
        Wd = np.empty((d_out, F), dtype=np.float64, order="F")
        check(lib().rcb_train_sharded(h.handle, comm.handle if comm else None, ptr(d3), ptr(y3), L.dtype_id(y3.dtype), T, B,
                                      int(washout), float(objective.reg), _check_solver(solver, gram_mode), ptr(h0), ptr(Wd), ptr(hT)))
        W = Wd.astype(_promote(objective, d3, y3))
        return addreadout(rc, W, ps, _store_carry(rc, st, cell_sts, hT, vec))
    partial = _partial or (lambda: _partial_gram_cuda(rc, ps, st, d3, y3, washout, device, gram_mode))
    GC, hT, cell_sts = partial() if _partial is None else _partial(rc, ps, st, d3, y3, washout)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        GC = _allreduce_gc(GC, d_out, group, device, use_lib=_partial is None)
    W = (_solve or (lambda g, d, r: ridge_solve(g, d, r, device)))(GC, d_out, float(objective.reg))
    W = np.asarray(W).astype(_promote(objective, d3, y3))
    return addreadout(rc, W, ps, _store_carry(rc, st, cell_sts, hT, vec))


def chunk_windows(T: int, washout: int, n_chunks: int, overlap: int):
    """BASELINE config 2: cut the samples [washout, T) of ONE long sequence into n_chunks equal chunks (+ a ragged tail).
    Chunk c owns [lo, hi) and is computed from a window that starts `overlap` steps earlier from a zero state; those
    steps are its local washout (the echo-state property re-synchronises the state).  Needs overlap <= washout so that
    every window has the same length.  Returns (Tc, [(start, lo, hi), ...], tail or None)."""
    if not 0 <= overlap <= washout:
        raise ArgumentError(f"time chunking needs 0 ≤ overlap ≤ washout, got overlap={overlap}, washout={washout}")
    n = T - washout
    if n_chunks < 1 or n < n_chunks:
        raise ArgumentError(f"cannot cut {n} samples into {n_chunks} chunks")
    Tc = n // n_chunks
    wins = [(washout + c * Tc - overlap, washout + c * Tc, washout + (c + 1) * Tc) for c in range(n_chunks)]
    rem = n - Tc * n_chunks
    tail = (T - rem - overlap, T - rem, T) if rem else None
    return Tc, wins, tail


def train_time_chunked(rc, train_data, target_data, ps, st, *, objective: RidgeRegression = None, solver=None, washout: int,
                       n_chunks: int, overlap: int, group=None, device: Optional[Device] = None, gram_mode=L.GRAM_AUTO):
    """train() for one long sequence (d_in, T), time-chunked after the washout: the chunks
    become a BATCH for the batched recurrence kernel, ranks take contiguous slices of the chunk list, and the partial
    [G | C] are summed by one all-reduce before the replicated FP64 solve — all inside rcb_train_time_chunked.  This is an
    approximation of the sequential semantics (exact only in the limit of a long overlap); the sequential single-GPU
    `train` stays the parity reference.  The carry stored in the returned st is the state at the end of this rank's last
    chunk."""
    objective = objective or RidgeRegression(0.0)
    mode = _check_solver(solver, gram_mode)
    X, Y = np.asarray(train_data), np.asarray(target_data)
    if X.ndim != 2 or Y.ndim != 2:
        raise DimensionMismatch("train_time_chunked takes one sequence: data (d_in, T), targets (d_out, T)")
    T = X.shape[1]
    if Y.shape[1] != T:
        raise DimensionMismatch(f"states has {T - washout} samples, targets has {Y.shape[1] - washout}")
    if washout >= T:
        raise ArgumentError(f"washout={washout} is ≥ number of time steps={T}")
    if Y.dtype not in (np.float32, np.float64):
        Y = Y.astype(np.float64)
    chunk_windows(T, washout, n_chunks, overlap)  # argument checks (the library repeats them)
    X, Y = np.asfortranarray(X), np.asfortranarray(Y)
    dev = device or default_device()
    h = _handle(rc, ps, X.dtype, dev)
    F, d_out = h.F, Y.shape[0]
    comm = comm_for(dev, group)
    Wd = np.empty((d_out, F), dtype=np.float64, order="F")
    hT = np.zeros((h.sumN, 1), dtype=X.dtype, order="F")
    check(lib().rcb_train_time_chunked(h.handle, comm.handle if comm else None, ptr(X), ptr(Y), L.dtype_id(Y.dtype), T, int(washout),
                                       int(n_chunks), int(overlap), float(objective.reg), mode, ptr(Wd), ptr(hT)))
    W = Wd.astype(_promote(objective, X, Y))
    cell_sts = list(_cell_st(rc, st))
    return addreadout(rc, W, ps, _store_carry(rc, st, cell_sts, hT, True))


def _tile_members(a3, B):
    """A shared (d, T, 1) series as one (d, T) block per member in the library's column-major layout — one broadcast write
    (np.repeat + asfortranarray took 23 ms per 15 MB array in config 5, this takes 1 ms)."""
    out = np.empty((a3.shape[0], a3.shape[1], B), dtype=a3.dtype, order="F")
    out[...] = a3
    return out


def train_ensemble(rc, train_data, target_data, ps, st, *, w_scale, leak, lambdas, washout: int = 0,
                   device: Optional[Device] = None, gram_mode=L.GRAM_AUTO):
    """Hyper-parameter / ensemble sweep (BASELINE config 5): member b runs the SAME reservoir with W * w_scale[b] and leak
    leak[b] on train_data[:, :, b] (pass (d_in, T) data to share it between all members) and gets one readout per
    regularisation in `lambdas`.  Each member is exactly `train(ESN(... radius*w_scale[b], leak[b]), ...; RidgeRegression(λ))`
    of the reference (src/train.jl:232-251); the recurrence and the Gram run once per member, the solve once per
    (member, λ).  Returns W (B, n_lambda, d_out, F) Float64 and the final carries (N, B).  Members are independent: shard
    them over ranks with shard_range — no collective."""
    w_scale = np.ascontiguousarray(w_scale, dtype=np.float32)
    leak = np.ascontiguousarray(leak, dtype=np.float32)
    lambdas = np.ascontiguousarray(lambdas, dtype=np.float64)
    B = w_scale.shape[0]
    if leak.shape[0] != B:
        raise DimensionMismatch(f"w_scale has {B} members, leak has {leak.shape[0]}")
    d3, _ = _as3d(train_data)
    y3, _ = _as3d(target_data)
    if y3.dtype not in (np.float32, np.float64):
        y3 = np.asfortranarray(y3.astype(np.float64))
    if d3.shape[2] == 1 and B > 1:
        d3 = _tile_members(d3, B)
    if y3.shape[2] == 1 and B > 1:
        y3 = _tile_members(y3, B)
    _, T, Bd = d3.shape
    if Bd != B or y3.shape[2] != B:
        raise DimensionMismatch(f"{B} members but data has {Bd} and targets {y3.shape[2]} sequences")
    if washout < 0:
        raise ArgumentError(f"washout must be ≥ 0, got {washout}")
    if washout >= T:
        raise ArgumentError(f"washout={washout} is ≥ number of time steps={T}")
    if y3.shape[1] != T:
        raise DimensionMismatch(f"states has {T - washout} samples, targets has {y3.shape[1] - washout}")
    h = _handle(rc, ps, d3.dtype, device)
    h0, _ = _initial_carry(rc, st, d3.dtype, B)
    W = np.empty((y3.shape[0], h.F, lambdas.shape[0], B), dtype=np.float64, order="F")
    hT = np.empty((h.sumN, B), dtype=d3.dtype, order="F")
    check(lib().rcb_esn_set_member_params(h.handle, B, ptr(w_scale), ptr(leak)))
    try:
        check(lib().rcb_train_members(h.handle, ptr(d3), ptr(y3), L.dtype_id(y3.dtype), T, B, int(washout), int(lambdas.shape[0]),
                                      ptr(lambdas), gram_mode, ptr(h0), ptr(W), ptr(hT)))
    finally:
        check(lib().rcb_esn_set_member_params(h.handle, 0, None, None))
    return W.transpose(3, 2, 0, 1), hT   # a view of the (d_out, F, n_lambda, B) result: W[b, l] is the (d_out, F) readout, column-major


def predict(rc, steps_or_data, ps, st, *, initialdata=None, device: Optional[Device] = None):
    """predict(rc, steps::Integer, ps, st; initialdata) — autoregressive (src/predict.jl:55-60,74-88);
    predict(rc, data::AbstractMatrix, ps, st) — teacher-forced (:90-103,163-177).
    Returns (outputs (d_out, T[, B]), st)."""
    if isinstance(rc, ReservoirChain):
        ro_ps = getattr(ps, rc.names()[-1])
        if isinstance(steps_or_data, (int, np.integer)):
            return _predict_delay_autoregressive(rc, int(steps_or_data), ps, st, initialdata, device, ro_ps=ro_ps)
        if np.asarray(steps_or_data).shape[1] < 1:
            raise ArgumentError("predict data must have at least one column, got 0.")
        z, st2 = collectstates(rc, steps_or_data, ps, st, device)
        return readout_apply(rc.readout, ro_ps, z, device), st2
    Wro = np.asarray(ps.readout.weight)
    if isinstance(rc, _DelayModel):
        if isinstance(steps_or_data, (int, np.integer)):
            return _predict_delay_autoregressive(rc, int(steps_or_data), ps, st, initialdata, device)
        if np.asarray(steps_or_data).shape[1] < 1:
            raise ArgumentError("predict data must have at least one column, got 0.")
        z, st2 = collectstates(rc, steps_or_data, ps, st, device)  # y_t = readout(features_t), predict.jl:163-177
        return readout_apply(rc.readout, ps.readout, z, device), st2
    if isinstance(steps_or_data, (int, np.integer)):
        steps = int(steps_or_data)
        if initialdata is None:
            raise ArgumentError("autoregressive predict needs initialdata")
        if steps < 1:
            raise ArgumentError(f"steps must be ≥ 1, got {steps}")
        u0 = np.asarray(initialdata)
        vec = u0.ndim == 1
        u0 = np.asfortranarray(u0.reshape(u0.shape[0], -1))
        data_dtype = u0.dtype if u0.dtype in (np.float32, np.float64) else np.dtype(np.float64)
        u0 = u0.astype(data_dtype)
        B = u0.shape[1]
        # Float64 W_out (default train) feeds Float64 outputs back: the loop runs in Float64 (esn_cell.jl:176)
        out_dtype = np.promote_types(Wro.dtype, data_dtype)
        h = _handle(rc, ps, data_dtype, device)
        h.set_readout(ps.readout, rc.readout.use_bias)
        h0, cell_sts = _initial_carry(rc, st, out_dtype, B)
        Y = np.empty((rc.readout.out_dims, steps, B), dtype=np.float64, order="F")
        check(lib().rcb_predict_autoregressive(h.handle, ptr(u0), steps, B, L.dtype_id(out_dtype), ptr(h0), ptr(Y)))
        st2 = _store_carry(rc, st, cell_sts, h0, vec)
        Y = Y.astype(out_dtype)
        return (Y[:, :, 0] if vec else Y), st2
    d3, vec = _as3d(steps_or_data)
    if d3.shape[1] < 1:
        raise ArgumentError(f"predict data must have at least one column, got {d3.shape[1]}.")
    dtype = d3.dtype
    _, T, B = d3.shape
    h = _handle(rc, ps, dtype, device)
    h.set_readout(ps.readout, rc.readout.use_bias)
    h0, cell_sts = _initial_carry(rc, st, dtype, B)
    Y = np.empty((rc.readout.out_dims, T, B), dtype=np.float64, order="F")
    check(lib().rcb_predict_teacher(h.handle, ptr(d3), T, B, ptr(h0), ptr(Y)))
    st2 = _store_carry(rc, st, cell_sts, h0, vec)
    Y = Y.astype(np.promote_types(Wro.dtype, dtype))
    return (Y[:, :, 0] if vec else Y), st2


def apply(rc, inp, ps, st, device: Optional[Device] = None, partial: bool = False):
    """One call of the model, ``(rc)(inp, ps, st)`` (src/reservoircomputer.jl:90-94): cell step + state modifiers
    (``__partial_apply``, :81-88) + readout, on the GPU (rcb_step).  ``inp`` is a vector (d_in,) or one column per sequence
    (d_in, B).  Returns (y, st') — or (features, st') with ``partial=True`` — with the carry threaded through ``st`` exactly as
    a loop of such calls does in the reference (that loop is ``collectstates`` / ``predict``)."""
    x = np.asarray(inp)
    vec = x.ndim == 1
    x = np.asfortranarray(x.reshape(x.shape[0], -1))
    dtype = x.dtype if x.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    x = x.astype(dtype)
    B = x.shape[1]
    h = _handle(rc, ps, dtype, device)
    h0, cell_sts = _initial_carry(rc, st, dtype, B)
    feats = np.empty((h.F, B), dtype=dtype, order="F")
    y = None
    if not partial:
        h.set_readout(ps.readout, rc.readout.use_bias)
        y = np.empty((rc.readout.out_dims, B), dtype=np.float64, order="F")
    check(lib().rcb_step(h.handle, ptr(x), B, ptr(h0), ptr(feats), ptr(y)))
    st2 = _store_carry(rc, st, cell_sts, h0, vec)
    if partial:
        return (feats[:, 0] if vec else feats), st2
    y = y.astype(np.promote_types(np.asarray(ps.readout.weight).dtype, dtype))
    return (y[:, 0] if vec else y), st2


def partial_apply(rc, inp, ps, st, device: Optional[Device] = None):
    """``__partial_apply(rc, inp, ps, st)``: cell step + state modifiers, no readout (src/reservoircomputer.jl:81-88)."""
    return apply(rc, inp, ps, st, device, partial=True)


def resetcarry(rng, rc, st, init_carry=None):
    """resetcarry!(rng, rc, st; init_carry) — src/reservoircomputer.jl:207-232, esn_deep.jl:219-259."""
    new = []
    for i, (c, s) in enumerate(zip(rc.cells, _cell_st(rc, st))):
        f = init_carry[i] if isinstance(init_carry, (tuple, list)) else init_carry
        new.append(s.merge(carry=None if f is None else (np.asarray(f(rng, c.out_dims)),)))
    return _with_cell_st(rc, st, new)


def apply_modifiers(mods, x, device: Optional[Device] = None):
    """State modifiers as plain functions (src/states.jl): a vector (n,) -> (F,) like NLAT2(::AbstractVector), or
    the matrix form (__apply_tomatrix, :1-15): x (n, cols) -> (F, cols)."""
    mods = _as_mods(mods)
    X = np.asarray(x)
    if X.ndim == 1:
        return apply_modifiers(mods, X.reshape(-1, 1), device)[:, 0]
    X = np.asfortranarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = np.asfortranarray(X.astype(np.float64))
    n, cols = X.shape
    out = np.empty((_feature_dims(n, mods), cols), dtype=X.dtype, order="F")
    kinds = (C.c_int32 * max(len(mods), 1))(*[m.kind for m in mods])
    params = (C.c_double * max(len(mods), 1))(*[float(m.param) for m in mods])
    dev = device or default_device()
    check(lib().rcb_apply_modifiers(dev.handle, len(mods), kinds, params, ptr(X), L.dtype_id(X.dtype), n, cols, ptr(out)))
    return out


def export_csr(rc, ps, layer: int = 0, device: Optional[Device] = None):
    """Canonical CSR (rowptr int64, colind int32, vals float32) the kernels were built from."""
    h = _handle(rc, ps, np.float32, device)
    nnz = C.c_int64()
    check(lib().rcb_esn_csr_nnz(h.handle, layer, C.byref(nnz)))
    n = rc.cells[layer].out_dims
    rowptr, colind, vals = np.empty(n + 1, np.int64), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float32)
    check(lib().rcb_esn_export_csr(h.handle, layer, ptr(rowptr), ptr(colind), ptr(vals)))
    return rowptr, colind, vals


def csr_from_matrix(W):
    """Host-only (no GPU): canonical CSR of a dense (N, N) float32 matrix or a scipy sparse matrix
    handed over as Julia's SparseMatrixCSC — the routine the device handle itself uses."""
    import scipy.sparse as sp

    nnz = C.c_int64()
    if sp.issparse(W):
        Wc = sp.csc_matrix(W)
        n = Wc.shape[0]
        colptr, rowval = Wc.indptr.astype(np.int64) + 1, Wc.indices.astype(np.int64) + 1
        nz = np.ascontiguousarray(Wc.data, dtype=np.float32)
        check(lib().rcb_csr_from_csc(ptr(colptr), ptr(rowval), ptr(nz), n, C.byref(nnz), None, None, None))
        rowptr, colind, vals = np.empty(n + 1, np.int64), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float32)
        check(lib().rcb_csr_from_csc(ptr(colptr), ptr(rowval), ptr(nz), n, C.byref(nnz), ptr(rowptr), ptr(colind), ptr(vals)))
        return rowptr, colind, vals
    Wd = np.asfortranarray(W, dtype=np.float32)
    n = Wd.shape[0]
    check(lib().rcb_csr_from_dense(ptr(Wd), n, n, C.byref(nnz), None, None, None))
    rowptr, colind, vals = np.empty(n + 1, np.int64), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float32)
    check(lib().rcb_csr_from_dense(ptr(Wd), n, n, C.byref(nnz), ptr(rowptr), ptr(colind), ptr(vals)))
    return rowptr, colind, vals


def eusn_operator_csr(W, diffusion: float):
    """Host-only (no GPU): canonical CSR of the EuSN operator ``W - W' - diffusion*I`` the kernels are built
    from (src/layers/eusn_cell.jl:120-123)."""
    Wd = np.asfortranarray(W.toarray() if hasattr(W, "toarray") else W, dtype=np.float32)
    n = Wd.shape[0]
    nnz = C.c_int64()
    check(lib().rcb_eusn_operator_from_dense(ptr(Wd), n, n, float(diffusion), C.byref(nnz), None, None, None))
    rowptr, colind, vals = np.empty(n + 1, np.int64), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float32)
    check(lib().rcb_eusn_operator_from_dense(ptr(Wd), n, n, float(diffusion), C.byref(nnz), ptr(rowptr), ptr(colind), ptr(vals)))
    return rowptr, colind, vals


# ---- feature plumbing around the recurrence: Extend, DelayLayer, the delay ESNs ------------------------
def extend_features(u, x, device: Optional[Device] = None):
    """``vcat(u, x)`` column by column on the GPU — Extend, src/states.jl:131-145.  u (d_in, T[, B]), x (n, T[, B])."""
    dev = device or default_device()
    u3, vec = _as3d(u)
    x3, _ = _as3d(x)
    if u3.shape[1:] != x3.shape[1:]:
        raise DimensionMismatch(f"Extend: input has {u3.shape[1:]} columns, state has {x3.shape[1:]}")
    dt = np.promote_types(u3.dtype, x3.dtype)
    u3, x3 = np.asfortranarray(u3, dtype=dt), np.asfortranarray(x3, dtype=dt)
    out = np.empty((u3.shape[0] + x3.shape[0],) + x3.shape[1:], dtype=dt, order="F")
    check(lib().rcb_extend_features(dev.handle, ptr(u3), u3.shape[0], ptr(x3), x3.shape[0], x3.shape[1] * x3.shape[2],
                                    L.dtype_id(dt), ptr(out)))
    return out[:, :, 0] if vec else out


class DelayLayer:
    """DelayLayer(in_dims; num_delays = 2, stride = 1, init_delay = zeros32) — src/layers/basic.jl:372-404."""

    def __init__(self, in_dims, *, num_delays=2, stride=1, init_delay=zeros32):
        if isinstance(init_delay, tuple):
            if len(init_delay) != num_delays:
                raise DimensionMismatch(f"init_delay tuple must have length num_delays={num_delays}, got {len(init_delay)}.")
        else:
            init_delay = tuple(init_delay for _ in range(num_delays))
        self.in_dims, self.num_delays, self.stride, self.init_delay = int(in_dims), int(num_delays), int(stride), init_delay

    def initialparameters(self, rng):
        return NT()

    def initialstates(self, rng):  # basic.jl:410-415
        return NT(history=None, clock=0, rng=copy.deepcopy(rng))

    def apply_sequence(self, x, st, device: Optional[Device] = None):
        """All T calls of the layer at once (rcb_delay_features): x (in_dims, T[, B]) -> ((num_delays+1) in_dims, T[, B]), st'."""
        dev = device or default_device()
        x3, vec = _as3d(x)
        n, T, B = x3.shape
        if n != self.in_dims:
            raise DimensionMismatch(f"DelayLayer expected input with {self.in_dims} rows, got {n}.")
        D = self.num_delays
        hist = st.history
        if hist is None and D > 0:  # init_delay_history, basic.jl:417-424: one draw per delay column
            rng = st.rng
            cols = [np.asarray(f(rng, n, 1), dtype=x3.dtype).reshape(n) for f in self.init_delay]
            hist = np.repeat(np.stack(cols, axis=1)[:, :, None], B, axis=2)
        hist = None if D == 0 else np.asfortranarray(np.asarray(hist, dtype=x3.dtype).reshape(n, D, -1))
        out = np.empty(((D + 1) * n, T, B), dtype=x3.dtype, order="F")
        hout = None if D == 0 else np.empty((n, D, B), dtype=x3.dtype, order="F")
        check(lib().rcb_delay_features(dev.handle, ptr(x3), L.dtype_id(x3.dtype), n, T, B, D, self.stride, ptr(hist),
                                       int(st.clock), ptr(out), ptr(hout)))
        new_hist = None if D == 0 else (hout[:, :, 0] if vec else hout)
        return (out[:, :, 0] if vec else out), st.merge(history=new_hist, clock=int(st.clock) + T)


class _DelayModel:
    """InputDelayESN / StateDelayESN / DelayESN (src/models/esn_delay.jl:107-130,276-300,456-490): an ESNCell with a
    DelayLayer on its input and / or as the first state modifier.  The recurrence runs through rcb_collect, the delay
    lines through rcb_delay_features, the remaining modifiers through rcb_apply_modifiers."""

    def __init__(self, in_dims, res_dims, out_dims, activation, n_in, s_in, n_st, s_st, readout_activation, state_modifiers,
                 cell_kw):
        self.input_delay = DelayLayer(in_dims, num_delays=n_in, stride=s_in) if n_in is not None else None
        aug = in_dims * ((n_in or 0) + 1)
        self.reservoir = ESNCell(aug, res_dims, activation, **cell_kw)
        self.state_delay = DelayLayer(res_dims, num_delays=n_st, stride=s_st) if n_st is not None else None
        self.state_modifiers = _as_mods(state_modifiers)  # the modifiers AFTER the state delay
        fdim = res_dims * ((n_st or 0) + 1)
        self.readout = LinearReadout(fdim, out_dims, readout_activation)
        self._cache = {}

    # the recurrence core: the same cell without modifiers (they act after the delay line)
    @property
    def cells(self):
        return (self.reservoir,)

    @property
    def layer_modifiers(self):
        return ((),) if self.state_delay is not None else (self.state_modifiers,)


class InputDelayESN(_DelayModel):
    def __init__(self, in_dims, res_dims, out_dims, activation=tanh, *, num_delays=2, stride=1, readout_activation=identity,
                 state_modifiers=(), **kw):
        super().__init__(in_dims, res_dims, out_dims, activation, num_delays, stride, None, None, readout_activation,
                         state_modifiers, kw)


class StateDelayESN(_DelayModel):
    def __init__(self, in_dims, res_dims, out_dims, activation=tanh, *, num_delays=2, stride=1, readout_activation=identity,
                 state_modifiers=(), **kw):
        super().__init__(in_dims, res_dims, out_dims, activation, None, None, num_delays, stride, readout_activation,
                         state_modifiers, kw)


class DelayESN(_DelayModel):
    def __init__(self, in_dims, res_dims, out_dims, activation=tanh, *, num_input_delays=1, input_stride=1,
                 num_state_delays=1, state_stride=1, readout_activation=identity, state_modifiers=(), **kw):
        super().__init__(in_dims, res_dims, out_dims, activation, num_input_delays, input_stride, num_state_delays,
                         state_stride, readout_activation, state_modifiers, kw)


def _setup_delay(rng, rc):  # esn_delay.jl:529-546; StateDelayESN keeps the plain layout with the delay as modifier 1
    ps = dict(reservoir=rc.reservoir.initialparameters(rng))
    st = dict(reservoir=rc.reservoir.initialstates(rng))
    if rc.input_delay is not None:
        ps = dict(input_delay=NT(), **ps)
        st = dict(input_delay=rc.input_delay.initialstates(rng), **st)
    nmods = len(rc.state_modifiers) + (1 if rc.state_delay is not None else 0)
    ps["state_modifiers"] = tuple(NT() for _ in range(nmods))
    mst = [NT() for _ in range(nmods)]
    if rc.state_delay is not None:
        mst[0] = rc.state_delay.initialstates(rng)
    st["state_modifiers"] = tuple(mst)
    ps["readout"] = rc.readout.initialparameters(rng)
    st["readout"] = NT()
    return NT(**ps), NT(**st)


def _collect_delay(rc, data, ps, st, device):
    dev = device or default_device()
    u = np.asarray(data)
    if u.shape[1] < 1:
        raise ArgumentError(f"collectstates data must have at least one column, got {u.shape[1]}.")
    st2 = st
    if rc.input_delay is not None:
        u, dst = rc.input_delay.apply_sequence(u, st.input_delay, dev)
        st2 = st2.merge(input_delay=dst)
    z, st_core = collectstates_core(rc, u, ps, st2, dev)
    st2 = st_core
    if rc.state_delay is not None:
        z, dst = rc.state_delay.apply_sequence(z, st.state_modifiers[0], dev)
        st2 = st2.merge(state_modifiers=(dst,) + tuple(st.state_modifiers[1:]))
        if rc.state_modifiers:
            z = apply_modifiers(rc.state_modifiers, z, dev)
    return z, st2


def collectstates_core(rc, data, ps, st, device=None):
    """rcb_collect for the model's cell (+ the modifiers K1 fuses) — shared by the plain and the delay models."""
    d3, vec = _as3d(data)
    dtype = d3.dtype
    h = _handle(rc, ps, dtype, device)
    _, T, B = d3.shape
    h0, cell_sts = _initial_carry(rc, st, dtype, B)
    states = np.empty((h.F, T, B), dtype=dtype, order="F")
    hT = np.empty((h.sumN, B), dtype=dtype, order="F")
    check(lib().rcb_collect(h.handle, ptr(d3), T, B, ptr(h0), ptr(states), ptr(hT)))
    return (states[:, :, 0] if vec else states), _store_carry(rc, st, cell_sts, hT, vec)


def readout_apply(ro: LinearReadout, ps_ro, features, device: Optional[Device] = None):
    """y = psi(W z + b) for every column (src/layers/basic.jl:171-182) on the GPU; Float64 result cast to the promoted eltype."""
    dev = device or default_device()
    z3, vec = _as3d(features)
    F, T, B = z3.shape
    W = np.asfortranarray(ps_ro.weight, dtype=np.float64)
    if W.shape[1] != F:
        raise DimensionMismatch(f"readout weight has {W.shape[1]} columns but the model produces {F} features")
    b = np.ascontiguousarray(ps_ro.bias, dtype=np.float64) if ro.use_bias else None
    y = np.empty((W.shape[0], T, B), dtype=np.float64, order="F")
    check(lib().rcb_readout_apply(dev.handle, ptr(W), ptr(b), _act_id(ro.activation), ptr(z3), L.dtype_id(z3.dtype), F, T * B,
                                  W.shape[0], ptr(y)))
    y = y.astype(np.promote_types(z3.dtype, np.asarray(ps_ro.weight).dtype))
    return y[:, :, 0] if vec else y


# ---- conceptors' building blocks ------------------------------------------------------------------------
def correlation_matrix(states, device: Optional[Device] = None, gram_mode=L.GRAM_AUTO):
    """correlation_matrix(states) = (Z Z') / S in float(eltype(states)) — src/conceptors.jl:28-33, on the K2 Gram kernels."""
    Z = np.asarray(states)
    if Z.size == 0:
        raise ArgumentError("states must contain at least one state")
    if Z.dtype not in (np.float32, np.float64):
        Z = Z.astype(np.float64)
    Z = np.asfortranarray(Z)
    dev = device or default_device()
    F, S = Z.shape
    R = np.empty((F, F), dtype=Z.dtype, order="F")
    check(lib().rcb_correlation_matrix(dev.handle, ptr(Z), L.dtype_id(Z.dtype), F, S, F, gram_mode, ptr(R)))
    return R


def ridge_map(features, targets, reg, device: Optional[Device] = None):
    """ridge_map (src/conceptors.jl:639-649): __fit_readout(RidgeRegression(element_type, reg), features, targets)."""
    f, t = np.asarray(features), np.asarray(targets)
    et = np.promote_types(f.dtype, t.dtype)
    if not np.issubdtype(et, np.floating):
        et = np.dtype(np.float64)
    return fit_readout(RidgeRegression(et, reg), f.astype(et), t.astype(et), device=device)


# ---- host setup on the GPU ------------------------------------------------------------------------------
def spectral_radius(W, device: Optional[Device] = None, tol: float = 0.0, max_dim: int = 0, return_dim: bool = False):
    """max |eigvals(W)| (what scale_radius! needs, src/inits/inits_components.jl:127) by Arnoldi on the GPU instead of a
    dense eigen-decomposition.  W: dense (N, N) float32 or a scipy sparse matrix (handed over as SparseMatrixCSC)."""
    import scipy.sparse as sp

    dev = device or default_device()
    rho, dim = C.c_double(), C.c_int32()
    if sp.issparse(W):
        Wc = sp.csc_matrix(W)
        Wc.sort_indices()
        colptr, rowval = Wc.indptr.astype(np.int64) + 1, Wc.indices.astype(np.int64) + 1
        nz = np.ascontiguousarray(Wc.data, dtype=np.float32)
        check(lib().rcb_spectral_radius_csc(dev.handle, ptr(colptr), ptr(rowval), ptr(nz), Wc.shape[0], float(tol), int(max_dim),
                                            C.byref(rho), C.byref(dim)))
    else:
        Wd = np.asfortranarray(W, dtype=np.float32)
        check(lib().rcb_spectral_radius_dense(dev.handle, ptr(Wd), Wd.shape[0], Wd.shape[0], float(tol), int(max_dim),
                                              C.byref(rho), C.byref(dim)))
    return (rho.value, dim.value) if return_dim else rho.value


def scale_radius(W, radius, device: Optional[Device] = None):
    """scale_radius!(W, radius): W .*= radius / rho(W) in W's eltype (src/inits/inits_components.jl:126-138); returns a
    new array (dense in, dense out; sparse in, sparse out)."""
    import scipy.sparse as sp

    rho = np.float32(spectral_radius(W, device))
    if not rho > 0:  # radius / 0 = Inf in the reference, caught by its Inf check (inits_components.jl:129-136)
        raise ArgumentError("Sparsity too low for size of the matrix. Increase res_size or increase sparsity.")
    f = np.float32(radius) / rho
    out = (W.multiply(f).astype(np.float32).tocsc()) if sp.issparse(W) else np.asfortranarray(np.asarray(W, np.float32) * f)
    vals = out.data if sp.issparse(out) else out
    if not np.all(np.isfinite(vals)):
        raise ArgumentError("Sparsity too low for size of the matrix. Increase res_size or increase sparsity.")
    return out


def hessenberg_eigvals(H):
    """Host-only: eigenvalues of an upper Hessenberg matrix by the library's shifted complex QR iteration."""
    Hf = np.asfortranarray(H, dtype=np.float64)
    m = Hf.shape[0]
    re, im = np.empty(m), np.empty(m)
    check(lib().rcb_hessenberg_eigvals(ptr(Hf), m, ptr(re), ptr(im)))
    return re + 1j * im


def _predict_delay_autoregressive(rc, steps, ps, st, initialdata, device, ro_ps=None):
    """y_1 = model(u0), y_{t+1} = model(y_t) (src/predict.jl:74-88) for the delay ESNs: the delay lines sit outside the
    fused predict kernel, so every step is the same four GPU calls as the teacher-forced path on ONE column (delay on
    the input, cell, delay on the state + modifiers, readout) with the delay buffers, clocks and the carry threaded
    through ``st`` — launch-latency bound (a few hundred microseconds per step), but every operation is a library kernel."""
    if initialdata is None:
        raise ArgumentError("autoregressive predict needs initialdata")
    if steps < 1:
        raise ArgumentError(f"steps must be ≥ 1, got {steps}")
    cur = np.asarray(initialdata)
    if cur.ndim != 1:
        raise ArgumentError("autoregressive predict of the delay ESNs takes one sequence (a vector initialdata)")
    n_in = cur.shape[0]
    outs = []
    for step in range(1, steps + 1):
        z, st = collectstates(rc, cur[:, None], ps, st, device)
        y = readout_apply(rc.readout, ro_ps if ro_ps is not None else ps.readout, z, device)[:, 0]
        if y.shape[0] != n_in:  # predict.jl:62-72
            raise DimensionMismatch(f"autoregressive predict requires each output to have length {n_in} so it can be used as "
                                    f"the next input; step {step} produced length {y.shape[0]}")
        outs.append(y)
        cur = y
    return np.stack(outs, axis=1), st


# ---- ReservoirChain: the one-cell chains of the reference's docs (Extend, Collect, delay lines, modifiers, readout) ----
class StatefulLayer:
    """StatefulLayer(cell) — src/layers/lux_layers.jl:27-46."""

    def __init__(self, cell):
        self.cell = cell


class Extend:
    """Extend(op): vcat(input, op(input)) — src/states.jl:131-145."""

    def __init__(self, op):
        if not isinstance(op, StatefulLayer):
            raise L.UnsupportedError("Extend wraps a StatefulLayer(cell) on the B200 path")
        self.op = op


class Collect:
    """Collect(): a collection point of features — src/layers/basic.jl:246-252."""


class _ChainCore:
    """What the device handle needs to know about the chain's cell: it runs without fused modifiers; the chain applies its
    own post-processing to the collected raw states."""

    def __init__(self, cell, readout):
        self.reservoir, self.readout, self._cache = cell, readout, {}
        self.cells, self.layer_modifiers = (cell,), ((),)


class ReservoirChain:
    """ReservoirChain(layers...) restricted to one recurrent cell (src/layers/lux_layers.jl:125-135):
    ``[DelayLayer...] , StatefulLayer(cell) | Extend(StatefulLayer(cell)) , [modifier | DelayLayer | Collect ...] , LinearReadout``.
    A readout built with include_collect=true (the default) contributes an implicit Collect right before it (:156-163);
    with several collection points the features are vcat-ed in chain order (src/layers/basic.jl:283-308)."""

    def __init__(self, *layers):
        flat = []
        for l in layers:
            if isinstance(l, LinearReadout):
                if getattr(l, "include_collect", True):
                    flat.append(Collect())
                flat.append(l)
            else:
                flat.append(l)
        self.layers = tuple(flat)
        kinds = [isinstance(l, (StatefulLayer, Extend)) for l in self.layers]
        if sum(kinds) != 1:
            raise L.UnsupportedError("the B200 path handles chains with exactly one recurrent cell")
        self._icell = kinds.index(True)
        if not isinstance(self.layers[-1], LinearReadout):
            raise L.UnsupportedError("the chain must end in a LinearReadout")
        for l in self.layers[:self._icell]:
            if not isinstance(l, DelayLayer):
                raise L.UnsupportedError("only DelayLayer may precede the cell on the B200 path")
        for l in self.layers[self._icell + 1:-1]:
            if not isinstance(l, (_Modifier, DelayLayer, Collect)):
                raise L.UnsupportedError(f"layer {l!r} has no B200 implementation")
        c = self.layers[self._icell]
        self.extend = isinstance(c, Extend)
        self.cell = (c.op if self.extend else c).cell
        self.readout = self.layers[-1]
        self._core = _ChainCore(self.cell, self.readout)

    def names(self):
        return [f"layer_{i + 1}" for i in range(len(self.layers))]   # named_tuple_layers, lux_layers.jl


def _setup_chain(rng, rc):
    ps, st = {}, {}
    for name, l in zip(rc.names(), rc.layers):
        if isinstance(l, StatefulLayer):
            ps[name], st[name] = l.cell.initialparameters(rng), l.cell.initialstates(rng)
        elif isinstance(l, Extend):
            ps[name], st[name] = NT(op=l.op.cell.initialparameters(rng)), NT(op=l.op.cell.initialstates(rng))
        elif isinstance(l, DelayLayer):
            ps[name], st[name] = NT(), l.initialstates(rng)
        elif isinstance(l, LinearReadout):
            ps[name], st[name] = l.initialparameters(rng), NT()
        else:
            ps[name], st[name] = NT(), NT()
    return NT(**ps), NT(**st)


def _collect_chain(rc, data, ps, st, device):
    dev = device or default_device()
    names = rc.names()
    u = np.asarray(data)
    if u.shape[1] < 1:
        raise ArgumentError(f"collectstates data must have at least one column, got {u.shape[1]}.")
    new_st = {}
    for i in range(rc._icell):  # delay lines on the input
        u, new_st[names[i]] = rc.layers[i].apply_sequence(u, getattr(st, names[i]), dev)
    cname = names[rc._icell]
    cps, cst = getattr(ps, cname), getattr(st, cname)
    if rc.extend:
        cps, cst = cps.op, cst.op
    raw, core_st = collectstates_core(rc._core, u, NT(reservoir=cps), NT(reservoir=cst), dev)
    new_st[cname] = NT(op=core_st.reservoir) if rc.extend else core_st.reservoir
    val = extend_features(u, raw, dev) if rc.extend else raw
    collected, pending = [], []

    def flush(v):
        nonlocal pending
        if pending:
            v = apply_modifiers(tuple(pending), v, dev)
            pending = []
        return v

    for i in range(rc._icell + 1, len(rc.layers) - 1):
        l = rc.layers[i]
        if isinstance(l, _Modifier):
            pending.append(l)
        elif isinstance(l, DelayLayer):
            val = flush(val)
            val, new_st[names[i]] = l.apply_sequence(val, getattr(st, names[i]), dev)
        else:  # Collect
            val = flush(val)
            collected.append(val)
    val = flush(val)
    feats = val if not collected else (collected[0] if len(collected) == 1 else np.asfortranarray(np.concatenate(collected, axis=0)))
    return feats, st.merge(**new_st)
