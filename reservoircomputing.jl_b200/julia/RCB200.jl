# RCB200.jl — reference-side binding of librcb200.so (include/rcb200.h) for SciML/ReservoirComputing.jl
# v0.12.35.  Source only: Julia is not installed in the build container, so this file is shipped as
# the binding a maintainer would add (INTEGRATION.md walks through it); the in-tree tests exercise
# the very same C entry points through ctypes (reservoircomputing.jl_b200/rcb200/_lib.py).
#
# The reference has no FFI for this path — its extension points are multiple-dispatch hooks
# (SURVEY.md §8b).  A reservoir wrapped in `B200Reservoir` selects the methods below:
#   __collectstates(reservoir, rc, data, ps, st)            src/reservoircomputer.jl:96-133
#   __predict(reservoir, rc, steps::Integer, ps, st; ...)   src/predict.jl:118-124,74-88
#   __predict(reservoir, rc, data::AbstractMatrix, ps, st)  src/predict.jl:126-129,163-177
#   __train_ridge(::B200Cholesky / ::B200QR, ::RidgeRegression, states, targets)   src/train.jl:140-198
#   train(rc::ReservoirComputer{<:B200Reservoir}, X, Y, ps, st; ...)               src/train.jl:232-251 (fused: rcb_train)
# `ESN`, `DeepESN`, `ReservoirComputer`, `LuxCore.setup`, `predict`, `collectstates`, `RidgeRegression`,
# `addreadout!` and `resetcarry!` are untouched: `ps`/`st` keep their layout (Appendix A of SURVEY.md),
# the final carry is copied back into `st.reservoir.carry`, and errors are re-raised as the
# exception types the reference throws (ArgumentError / DimensionMismatch).
module RCB200

using ReservoirComputing
using ReservoirComputing: AbstractReservoirComputer, ReservoirComputer, StatefulLayer, ESNCell, ES2NCell, ResESNCell,
                          EuSNCell, RidgeRegression, NLAT1, NLAT2, NLAT3, Pad, PartialSquare, ExtendedSquare
using ReservoirComputing: AbstractReservoirComputingSolver
using ReservoirComputing.LuxCore: AbstractLuxWrapperLayer
import ReservoirComputing: __collectstates, __predict, __fit_readout, __train_ridge, collectstates, train
using SparseArrays: SparseMatrixCSC

const LIB = get(ENV, "RCB200_LIB", joinpath(@__DIR__, "..", "librcb200.so"))

# ---- status -> exception (include/rcb200.h: rcb_status) ------------------------------------------
const RCB_OK = Int32(0)
lasterror() = unsafe_string(ccall((:rcb_last_error, LIB), Cstring, ()))
function check(status::Int32)
    status == RCB_OK && return nothing
    msg = lasterror()
    status == -1 && throw(ArgumentError(msg))
    status == -2 && throw(DimensionMismatch(msg))
    status == -4 && throw(ArgumentError(msg))   # solver failure: train.jl:194-196 raises ArgumentError
    error("librcb200: status $status: $msg")     # CUDA / allocation / unsupported: there is no CPU fallback
end

# ---- C structs (field order and sizes as in include/rcb200.h) ------------------------------------
const MAXMOD = 8
struct LayerDesc
    res_dims::Int32
    activation::Int32
    use_bias::Int32
    leak_is_vector::Int32
    leak_scalar::Float64
    n_modifiers::Int32
    modifier_kind::NTuple{MAXMOD, Int32}
    modifier_param::NTuple{MAXMOD, Float64}
end
struct ESNDesc
    n_layers::Int32
    in_dims::Int32
    out_dims::Int32
    readout_activation::Int32
    readout_use_bias::Int32
    data_dtype::Int32
end

activation_id(::typeof(identity)) = Int32(0)
activation_id(::typeof(tanh)) = Int32(1)
activation_id(f) = nameof(f) === :tanh_fast ? Int32(2) :
                   nameof(f) === :sigmoid ? Int32(3) : nameof(f) === :relu ? Int32(4) :
                   throw(ArgumentError("activation $(f) is outside the B200 hot path"))
dtype_id(::Type{Float32}) = Int32(0)
dtype_id(::Type{Float64}) = Int32(1)

modifier_entry(::typeof(NLAT1)) = (Int32(1), 0.0)
modifier_entry(::typeof(NLAT2)) = (Int32(2), 0.0)
modifier_entry(::typeof(NLAT3)) = (Int32(3), 0.0)
modifier_entry(p::Pad) = (Int32(4), Float64(p.padding))
modifier_entry(p::PartialSquare) = (Int32(5), Float64(p.eta))
modifier_entry(::typeof(ExtendedSquare)) = (Int32(6), 0.0)
modifier_entry(w) = hasproperty(w, :func) ? modifier_entry(w.func) :   # WrappedFunction, lux_layers.jl:2-6
                    throw(ArgumentError("state modifier $(w) is outside the B200 hot path"))

# ---- context + handle ------------------------------------------------------------------------------
mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(device::Integer = 0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rcb_ctx_create, LIB), Int32, (Int32, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), device, C_NULL, out))
        c = new(out[])
        finalizer(x -> ccall((:rcb_ctx_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), c)
    end
end
const DEFAULT_CTX = Ref{Union{Nothing, Context}}(nothing)
ctx() = something(DEFAULT_CTX[], (DEFAULT_CTX[] = Context(0)))

"""
    diverged(c::Context = ctx()) -> Bool

`true` if the carry of a recurrence of the most recent call held a NaN / Inf (`rcb_ctx_last_nonfinite`): a diverged
reservoir, which the reference path only shows as NaN weights later on.
"""
function diverged(c::Context = ctx())
    f = Ref{Int32}(0)
    check(ccall((:rcb_ctx_last_nonfinite, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), c.ptr, f))
    return f[] != 0
end

"""
    B200Reservoir(cell)

Marker wrapper selecting the CUDA path: `ESN(...; )` is built as usual and its `reservoir`
(`StatefulLayer(ESNCell)`) is wrapped — `ReservoirComputer(B200Reservoir(rc.reservoir), rc.state_modifiers, rc.readout)`.
Parameter and state trees are those of the wrapped layer (LuxCore.initialparameters/initialstates forward).
"""
struct B200Reservoir{L} <: AbstractLuxWrapperLayer{:layer}   # like StatefulLayer (lux_layers.jl:27): ps / st are those of `layer`
    layer::L
end
(r::B200Reservoir)(x, ps, st) = r.layer(x, ps, st)   # single steps stay on the reference path (reservoircomputer.jl:90-94)

mutable struct Handle
    ptr::Ptr{Cvoid}
    F::Int
    N::Int
end

# sibling cells (include/rcb200.h: rcb_cell_kind): (kind, p0, p1, leak)
cell_variant(c::ESNCell) = (Int32(0), 0.0, 0.0, c.leak_coefficient)
cell_variant(c::ES2NCell) = (Int32(1), Float64(c.proximity), 0.0, 1.0)               # es2n_cell.jl:106-116
cell_variant(c::ResESNCell) = (Int32(2), Float64(c.alpha), Float64(c.beta), 1.0)     # resesn_cell.jl:112-125
cell_variant(c::EuSNCell) = (Int32(3), Float64(c.diffusion), 0.0, c.leak_coefficient) # eusn_cell.jl:95-106
cell_variant(c) = throw(ArgumentError("cell $(typeof(c)) is outside the B200 hot path"))

function layer_desc(cell, layer_mods)
    kind, p0, p1, α = cell_variant(cell)
    mods = collect(modifier_entry(m) for m in layer_mods)
    length(mods) <= MAXMOD || throw(ArgumentError("at most $MAXMOD state modifiers per layer"))
    kinds = ntuple(i -> i <= length(mods) ? mods[i][1] : Int32(0), MAXMOD)
    params = ntuple(i -> i <= length(mods) ? mods[i][2] : 0.0, MAXMOD)
    return LayerDesc(cell.out_dims, activation_id(cell.activation), ReservoirComputing.known(cell.use_bias) ? 1 : 0,
                     α isa AbstractVector ? 1 : 0, α isa AbstractVector ? 0.0 : Float64(α), length(mods), kinds, params)
end

# One handle for `cells` (ESNCell-family cells, layer order), their per-layer modifier tuples and parameter NamedTuples.
function build_handle(cells, mods_per_layer, ps_cells, readout, ::Type{T}) where {T}
    nl = length(cells)
    lds = [layer_desc(cells[l], mods_per_layer[l]) for l in 1:nl]
    d = ESNDesc(nl, cells[1].in_dims, readout.out_dims, activation_id(readout.activation),
                ReservoirComputing.known(readout.use_bias) ? 1 : 0, dtype_id(T))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rcb_esn_create, LIB), Int32, (Ptr{Cvoid}, Ref{ESNDesc}, Ptr{LayerDesc}, Ref{Ptr{Cvoid}}),
                ctx().ptr, Ref(d), lds, out))
    h = Handle(out[], 0, sum(c.out_dims for c in cells))
    finalizer(x -> ccall((:rcb_esn_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
    for l in 1:nl
        cell, p, li = cells[l], ps_cells[l], Int32(l - 1)
        kind, p0, p1, α = cell_variant(cell)
        kind != 0 && check(ccall((:rcb_esn_set_cell, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64, Float64), h.ptr, li, kind, p0, p1))
        W, Win = p.reservoir_matrix, Matrix{Float32}(p.input_matrix)
        bias = haskey(p, :bias) ? Vector{Float32}(p.bias) : C_NULL
        if W isa SparseMatrixCSC                      # return_sparse=true, ext/RCSparseArraysExt.jl:5-7
            check(ccall((:rcb_esn_set_weights_csc, LIB), Int32,
                        (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}, Ptr{Float32}, Int64, Ptr{Float32}),
                        h.ptr, li, W.colptr, W.rowval, Vector{Float32}(W.nzval), Win, size(Win, 1), bias))
        else                                          # dense Matrix with explicit zeros (rand_sparse default)
            Wd = Matrix{Float32}(W)
            check(ccall((:rcb_esn_set_weights_dense, LIB), Int32,
                        (Ptr{Cvoid}, Int32, Ptr{Float32}, Int64, Ptr{Float32}, Int64, Ptr{Float32}),
                        h.ptr, li, Wd, size(Wd, 1), Win, size(Win, 1), bias))
        end
        α isa AbstractVector && check(ccall((:rcb_esn_set_leak_vector, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}),
                                            h.ptr, li, Vector{Float32}(α)))
        if kind == 1 || kind == 2                     # ps.reservoir.orthogonal_matrix, es2n_cell.jl:93-104
            Om = Matrix{Float32}(p.orthogonal_matrix)
            check(ccall((:rcb_esn_set_orthogonal_dense, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float32}, Int64), h.ptr, li, Om, size(Om, 1)))
        end
    end
    F = Ref{Int32}(0)
    check(ccall((:rcb_esn_feature_dims, LIB), Int32, (Ptr{Cvoid}, Ref{Int32}), h.ptr, F))
    h.F = F[]
    return h
end

Handle(rc::AbstractReservoirComputer, ps, ::Type{T}) where {T} =
    build_handle((rc.reservoir.layer.cell,), (rc.state_modifiers,), (ps.reservoir,), rc.readout, T)

# h0: the carry in `st`, or a fresh draw from the replicated rng exactly like the first reference step
# (esn_cell.jl:169-173, lux_layers.jl:212-216)
function initial_carry(rc, st, ::Type{T}) where {T}
    cell = rc.reservoir.layer.cell
    if st.reservoir.carry === nothing
        rng = ReservoirComputing.LuxCore.replicate(st.reservoir.cell.rng)
        return Vector{T}(vec(cell.init_state(rng, cell.out_dims, 1)))
    end
    return Vector{T}(vec(first(st.reservoir.carry)))
end
store_carry(st, h) = merge(st, (reservoir = merge(st.reservoir, (carry = (reshape(h, :, 1),),)),))

# ---- collectstates: src/reservoircomputer.jl:113-133 --------------------------------------------------
function __collectstates(res::B200Reservoir, rc::AbstractReservoirComputer, data::AbstractMatrix{T}, ps, st) where {T}
    size(data, 2) >= 1 || throw(ArgumentError("collectstates data must have at least one column, got 0."))
    h = Handle(rc, ps, T)
    Tn = size(data, 2)
    states = Matrix{T}(undef, h.F, Tn)                      # similar(data, F, T), :123
    h0 = initial_carry(rc, st, T)
    hT = similar(h0)
    check(ccall((:rcb_collect, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                h.ptr, Matrix{T}(data), Tn, 1, h0, states, hT))
    return states, store_carry(st, hT)
end

# ---- single model call: src/reservoircomputer.jl:81-94 -------------------------------------------------------------
# `(rc)(inp, ps, st)` keeps running on the reference path (B200Reservoir forwards the cell call: one step is launch-latency
# bound).  `step_b200` is the same call on the GPU — cell step + modifiers + readout in one rcb_step — for hosts that keep
# everything else there; `partial = true` stops before the readout (`__partial_apply`).
function step_b200(rc::ReservoirComputer{<:B200Reservoir}, inp::AbstractVector{T}, ps, st; partial::Bool = false) where {T <: Union{Float32, Float64}}
    h = Handle(rc, ps, T)
    partial || set_readout!(h, ps)
    carry = initial_carry(rc, st, T)
    feats = Vector{T}(undef, h.F)
    y = partial ? nothing : Vector{Float64}(undef, rc.readout.out_dims)
    check(ccall((:rcb_step, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                h.ptr, Vector{T}(inp), 1, carry, feats, partial ? C_NULL : y))
    out = partial ? feats : Vector{promote_type(T, eltype(ps.readout.weight))}(y)
    return out, store_carry(st, carry)
end

# ---- ridge: src/train.jl:100-198 ---------------------------------------------------------------------
# The objective stays the reference's `RidgeRegression(λ)`; the GPU path is selected the way the reference selects any
# other algorithm — through `solver` (train.jl:100-106, 149-198):
#     train(rc, X, Y, ps, st; objective = RidgeRegression(1e-6), solver = B200Cholesky())
#   B200Cholesky(gram = :auto | :fp64 | :tensor)  Gram form (K2) + FP64 Cholesky (K3); `:auto` escalates by itself: λ == 0
#                                                 -> QR, a failed pivot -> exact FP64 Gram -> QR (include/rcb200.h)
#   B200QR()                                      FP64 Householder QR of [Z'; √λ I] on the GPU — the reference's own algorithm
const GRAM_AUTO, GRAM_FP64, GRAM_TENSOR, FIT_QR = Int32(0), Int32(1), Int32(2), Int32(3)
struct B200Cholesky <: AbstractReservoirComputingSolver
    gram::Symbol
end
B200Cholesky(; gram::Symbol = :auto) = B200Cholesky(gram)
struct B200QR <: AbstractReservoirComputingSolver end
gram_mode(s::B200Cholesky) = s.gram === :auto ? GRAM_AUTO : s.gram === :fp64 ? GRAM_FP64 : s.gram === :tensor ? GRAM_TENSOR :
                             throw(ArgumentError("B200Cholesky: gram must be :auto, :fp64 or :tensor, got :$(s.gram)"))
gram_mode(::B200QR) = FIT_QR
const B200Solver = Union{B200Cholesky, B200QR}

dense32or64(x::Matrix{Float32}) = x
dense32or64(x::Matrix{Float64}) = x
dense32or64(x::AbstractMatrix{Float32}) = Matrix{Float32}(x)
dense32or64(x::AbstractMatrix) = Matrix{Float64}(x)

function __train_ridge(solver::B200Solver, objective::RidgeRegression, states::AbstractMatrix, targets::AbstractMatrix; kwargs...)
    size(states, 2) == size(targets, 2) ||
        throw(DimensionMismatch("states has $(size(states, 2)) samples, targets has $(size(targets, 2))"))
    F, S = size(states); D = size(targets, 1)
    Z, Y = dense32or64(states), dense32or64(targets)
    W = Matrix{Float64}(undef, D, F)
    check(ccall((:rcb_fit_readout, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Float64, Int32, Ptr{Float64}),
                ctx().ptr, Z, dtype_id(eltype(Z)), F, S, stride(Z, 2), Y, dtype_id(eltype(Y)), D, size(Y, 2), stride(Y, 2),
                Float64(objective.reg), gram_mode(solver), W))
    Tp = promote_type(eltype(states), eltype(targets), typeof(objective.reg))    # train.jl:126-129
    return Matrix{Tp}(W)
end

"""`B200Ridge(λ)` — shorthand objective: `RidgeRegression(λ)` solved with `B200Cholesky()`."""
struct B200Ridge{R <: Number}
    reg::R
end
__fit_readout(obj::B200Ridge, states::AbstractMatrix, targets::AbstractMatrix; solver = nothing, kwargs...) =
    __train_ridge(solver isa B200Solver ? solver : B200Cholesky(), RidgeRegression(obj.reg), states, targets; kwargs...)

# ---- train: src/train.jl:232-251, fused ---------------------------------------------------------------------------
# The generic `train` materialises the F x T state matrix on the host between `collectstates` and `__fit_readout`.  For a
# CUDA-backed reservoir with a GPU solver the whole chain collect -> washout -> Gram -> solve -> addreadout! is ONE call
# (rcb_train): the states never leave the device (the library cuts long inputs into windows that fit HBM) unless
# `return_states = true` asks for them.  Every other (objective, solver) combination falls through to the generic method.
function train(rc::ReservoirComputer{<:B200Reservoir}, train_data::AbstractMatrix{T}, target_data::AbstractMatrix, ps, st;
               objective = RidgeRegression(0.0), solver = nothing, washout::Integer = 0, return_states::Bool = false,
               kwargs...) where {T <: Union{Float32, Float64}}
    if objective isa B200Ridge
        objective, solver = RidgeRegression(objective.reg), solver isa B200Solver ? solver : B200Cholesky()
    end
    if !(objective isa RidgeRegression && solver isa B200Solver)
        return invoke(train, Tuple{Any, Any, Any, Any, Any}, rc, train_data, target_data, ps, st;
                      objective, solver, washout, return_states, kwargs...)
    end
    Tn = size(train_data, 2)
    Tn >= 1 || throw(ArgumentError("collectstates data must have at least one column, got 0."))
    washout >= 0 || throw(ArgumentError("washout must be ≥ 0, got $washout"))                       # train.jl:37
    washout < Tn || throw(ArgumentError("washout=$washout is ≥ number of time steps=$Tn"))          # train.jl:39-43
    size(target_data, 2) == Tn ||
        throw(DimensionMismatch("states has $(Tn - washout) samples, targets has $(size(target_data, 2) - washout)"))
    h = Handle(rc, ps, T)
    X, Y = Matrix{T}(train_data), dense32or64(target_data)
    D = size(Y, 1)
    W = Matrix{Float64}(undef, D, h.F)
    states = return_states ? Matrix{T}(undef, h.F, Tn) : nothing
    h0 = initial_carry(rc, st, T); hT = similar(h0)
    check(ccall((:rcb_train, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Float64, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}),
                h.ptr, X, Y, dtype_id(eltype(Y)), Tn, 1, washout, Float64(objective.reg), gram_mode(solver), h0, W,
                states === nothing ? C_NULL : states, hT))
    Tp = promote_type(T, eltype(target_data), typeof(objective.reg))                                 # train.jl:126-129
    ps2, st2 = ReservoirComputing.addreadout!(rc, Matrix{Tp}(W), ps, store_carry(st, hT))
    return return_states ? ((ps2, st2), states[:, (washout + 1):end]) : (ps2, st2)
end

# ---- streaming / sharded fit (SURVEY 8e): the pieces a multi-process Julia host (MPI.jl, Distributed) composes ----------
"""
    gram_accumulate!(GC, states, targets; gram = :auto, zero_first = false) -> GC

`[G | C] += [Z Z' | Z Y']` (F x (F + d_out) Float64, lower triangle of G) — one chunk of a streamed ridge fit.
`ridge_solve(GC, d_out, λ)` finishes it; ranks sum their `GC` in between (`allreduce_gc!`).
"""
function gram_accumulate!(GC::Matrix{Float64}, states::AbstractMatrix, targets::AbstractMatrix; gram::Symbol = :auto, zero_first::Bool = false)
    Z, Y = dense32or64(states), dense32or64(targets)
    F, S = size(Z); D = size(Y, 1)
    size(GC) == (F, F + D) || throw(DimensionMismatch("GC must be $F x $(F + D), got $(size(GC))"))
    check(ccall((:rcb_gram_accumulate, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Ptr{Cvoid}, Int32, Int64, Int64, Int32, Int32, Ptr{Float64}),
                ctx().ptr, Z, dtype_id(eltype(Z)), F, S, stride(Z, 2), Y, dtype_id(eltype(Y)), D, stride(Y, 2),
                gram_mode(B200Cholesky(gram)), zero_first ? 1 : 0, GC))
    return GC
end
function ridge_solve(GC::Matrix{Float64}, d_out::Integer, λ::Real)
    F = size(GC, 1)
    W = Matrix{Float64}(undef, d_out, F)
    check(ccall((:rcb_ridge_solve, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Float64, Ptr{Float64}), ctx().ptr, GC, F, d_out, Float64(λ), W))
    return W
end

"""
    Comm(rank, world, id) / comm_unique_id()

Library-owned NCCL communicator, one process per GPU: rank 0 draws `id = comm_unique_id()` (128 bytes), ships it to every
rank by the host's own means (`MPI.Bcast!`, a `RemoteChannel`, a file), then every rank constructs `Comm(rank, world, id)`.
"""
mutable struct Comm
    ptr::Ptr{Cvoid}
    function Comm(rank::Integer, world::Integer, id::Vector{UInt8})
        length(id) == 128 || throw(ArgumentError("the communicator id has 128 bytes, got $(length(id))"))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rcb_comm_create, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Ref{Ptr{Cvoid}}), ctx().ptr, id, rank, world, out))
        c = new(out[])
        finalizer(x -> ccall((:rcb_comm_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), c)
    end
end
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:rcb_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    return id
end
allreduce_gc!(comm::Comm, GC::Matrix{Float64}, d_out::Integer) =
    (check(ccall((:rcb_allreduce_gc, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64), comm.ptr, GC, size(GC, 1), d_out)); GC)

"""
    train_sharded(rc, comm, data, targets, ps, st; objective, solver = B200Cholesky(), washout = 0)

`train` with ONE pooled readout over sequences sharded across ranks (BASELINE config 3): this rank passes its own
`d_in x T x B_local` inputs and `d_out x T x B_local` targets; the partial `[G | C]` are summed by one NCCL all-reduce inside
the library and every rank gets the same readout.  `comm = nothing`: a single rank.
"""
function train_sharded(rc::ReservoirComputer{<:B200Reservoir}, comm::Union{Comm, Nothing}, data::Array{T, 3}, targets::Array{<:Union{Float32, Float64}, 3},
                       ps, st; objective::RidgeRegression = RidgeRegression(0.0), solver::B200Solver = B200Cholesky(), washout::Integer = 0) where {T <: Union{Float32, Float64}}
    _, Tn, B = size(data)
    (size(targets, 2), size(targets, 3)) == (Tn, B) || throw(DimensionMismatch("targets must be d_out x $Tn x $B, got $(size(targets))"))
    h = Handle(rc, ps, T)
    W = Matrix{Float64}(undef, size(targets, 1), h.F)
    h0 = repeat(initial_carry(rc, st, T), 1, B); hT = similar(h0)
    check(ccall((:rcb_train_sharded, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Float64, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                h.ptr, comm === nothing ? C_NULL : comm.ptr, data, targets, dtype_id(eltype(targets)), Tn, B, washout,
                Float64(objective.reg), gram_mode(solver), h0, W, hT))
    ps2, _ = ReservoirComputing.addreadout!(rc, W, ps, st)
    return ps2, hT
end

"""
    train_time_chunked(rc, comm, data, targets, ps, st; objective, solver, washout, n_chunks, overlap)

`train` of ONE long series (BASELINE config 2) cut into `n_chunks` windows after the washout; every window re-synchronises
over `overlap` steps from a zero state (echo-state property — an approximation; plain `train` is the sequential reference).
"""
function train_time_chunked(rc::ReservoirComputer{<:B200Reservoir}, comm::Union{Comm, Nothing}, data::Matrix{T}, targets::Matrix{<:Union{Float32, Float64}},
                            ps, st; objective::RidgeRegression = RidgeRegression(0.0), solver::B200Solver = B200Cholesky(), washout::Integer,
                            n_chunks::Integer, overlap::Integer = washout) where {T <: Union{Float32, Float64}}
    h = Handle(rc, ps, T)
    W = Matrix{Float64}(undef, size(targets, 1), h.F)
    hT = Vector{T}(undef, h.N)
    check(ccall((:rcb_train_time_chunked, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Int64, Float64, Int32, Ptr{Float64}, Ptr{Cvoid}),
                h.ptr, comm === nothing ? C_NULL : comm.ptr, data, targets, dtype_id(eltype(targets)), size(data, 2), washout, n_chunks,
                overlap, Float64(objective.reg), gram_mode(solver), W, hT))
    ps2, _ = ReservoirComputing.addreadout!(rc, W, ps, st)
    return ps2, store_carry(st, hT)
end

"""
    train_ensemble(rc, data, targets, ps, st; w_scale, leak, lambdas, washout = 0, solver = B200Cholesky()) -> W, hT

Hyper-parameter sweep (BASELINE config 5): member `b` runs the same reservoir with `W * w_scale[b]` and leak `leak[b]` and gets
one readout per `λ ∈ lambdas`; `W` is `d_out x F x length(lambdas) x B`.  Members are independent: shard them over ranks.
"""
function train_ensemble(rc::ReservoirComputer{<:B200Reservoir}, data::Array{T, 3}, targets::Array{<:Union{Float32, Float64}, 3}, ps, st;
                        w_scale::Vector{Float32}, leak::Vector{Float32}, lambdas::Vector{Float64}, washout::Integer = 0,
                        solver::B200Solver = B200Cholesky()) where {T <: Union{Float32, Float64}}
    _, Tn, B = size(data)
    length(w_scale) == B == length(leak) || throw(DimensionMismatch("w_scale / leak need one entry per member ($B)"))
    h = Handle(rc, ps, T)
    W = Array{Float64, 4}(undef, size(targets, 1), h.F, length(lambdas), B)
    h0 = repeat(initial_carry(rc, st, T), 1, B); hT = similar(h0)
    check(ccall((:rcb_esn_set_member_params, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Float32}, Ptr{Float32}), h.ptr, B, w_scale, leak))
    check(ccall((:rcb_train_members, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Int32, Ptr{Float64}, Int32, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}),
                h.ptr, data, targets, dtype_id(eltype(targets)), Tn, B, washout, length(lambdas), lambdas, gram_mode(solver), h0, W, hT))
    return W, hT
end

# ---- predict: src/predict.jl:74-88 (autoregressive) and :163-177 (teacher-forced) -----------------------
function set_readout!(h::Handle, ps)
    W = Matrix{Float64}(ps.readout.weight)
    b = haskey(ps.readout, :bias) ? Vector{Float64}(ps.readout.bias) : C_NULL
    check(ccall((:rcb_esn_set_readout, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), h.ptr, W, b, 1))
end

function __predict(res::B200Reservoir, rc::AbstractReservoirComputer, steps::Integer, ps, st; initialdata)
    steps >= 1 || throw(ArgumentError("steps must be ≥ 1, got $steps"))
    T = eltype(initialdata)
    h = Handle(rc, ps, T); set_readout!(h, ps)
    # the reference's loop silently runs in the eltype of the fed-back output (esn_cell.jl:176): Float64
    # after the default (Float64) ridge fit, Float32 after RidgeRegression(Float32, λ)
    Tc = promote_type(T, eltype(ps.readout.weight))
    carry = Vector{Tc}(initial_carry(rc, st, Tc))
    Y = Matrix{Float64}(undef, rc.readout.out_dims, steps)
    check(ccall((:rcb_predict_autoregressive, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Cvoid}, Ptr{Float64}),
                h.ptr, Vector{T}(initialdata), steps, 1, dtype_id(Tc), carry, Y))
    return Matrix{Tc}(Y), store_carry(st, carry)
end

function __predict(res::B200Reservoir, rc::AbstractReservoirComputer, data::AbstractMatrix{T}, ps, st) where {T}
    h = Handle(rc, ps, T); set_readout!(h, ps)
    carry = initial_carry(rc, st, T)
    Y = Matrix{Float64}(undef, rc.readout.out_dims, size(data, 2))
    check(ccall((:rcb_predict_teacher, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Float64}),
                h.ptr, Matrix{T}(data), size(data, 2), 1, carry, Y))
    return Matrix{promote_type(T, eltype(ps.readout.weight))}(Y), store_carry(st, carry)
end

# ---- DeepESN: src/models/esn_deep.jl:190-291 ------------------------------------------------------------------------
# DeepESN has its own `collectstates` method and no `:reservoir` field (SURVEY 8b item 6), so the CUDA path is selected
# by wrapping the whole model; `ps` / `st` keep DeepESN's layout (cells, state_modifiers, readout).
"""`B200DeepESN(desn)` / `b200(desn)` — a `DeepESN` whose `collectstates` and `predict` run on the GPU."""
struct B200DeepESN{M <: ReservoirComputing.DeepESN} <:
       ReservoirComputing.AbstractEchoStateNetwork{(:cells, :state_modifiers, :readout)}
    cells
    state_modifiers
    readout
    model::M
end
B200DeepESN(m::ReservoirComputing.DeepESN) = B200DeepESN(m.cells, m.state_modifiers, m.readout, m)
ReservoirComputing.LuxCore.initialparameters(rng, d::B200DeepESN) = ReservoirComputing.LuxCore.initialparameters(rng, d.model)
ReservoirComputing.LuxCore.initialstates(rng, d::B200DeepESN) = ReservoirComputing.LuxCore.initialstates(rng, d.model)
(d::B200DeepESN)(inp, ps, st) = d.model(inp, ps, st)     # single steps stay on the reference path
ReservoirComputing.resetcarry!(rng, d::B200DeepESN, st; kwargs...) = ReservoirComputing.resetcarry!(rng, d.model, st; kwargs...)
ReservoirComputing.addreadout!(d::B200DeepESN, W::AbstractMatrix, ps, st) = ReservoirComputing.addreadout!(d.model, W, ps, st)

deep_handle(d::B200DeepESN, ps, ::Type{T}) where {T} =
    build_handle(map(c -> c.cell, d.cells), map(m -> m === nothing ? () : m, d.state_modifiers), ps.cells, d.readout, T)

# packed carry (sum of N_l rows): stored carries, or fresh init_state draws like the first reference step of each cell
function deep_initial_carry(d::B200DeepESN, st, ::Type{T}) where {T}
    parts = map(zip(d.cells, st.cells)) do (c, s)
        if s.carry === nothing
            rng = ReservoirComputing.LuxCore.replicate(s.cell.rng)
            Vector{T}(vec(c.cell.init_state(rng, c.cell.out_dims, 1)))
        else
            Vector{T}(vec(first(s.carry)))
        end
    end
    return reduce(vcat, parts)
end
function deep_store_carry(d::B200DeepESN, st, h)
    off = 0
    new_cells = map(zip(d.cells, st.cells)) do (c, s)
        n = c.cell.out_dims
        part = reshape(h[(off + 1):(off + n)], :, 1)
        off += n
        merge(s, (carry = (part,),))
    end
    return merge(st, (cells = Tuple(new_cells),))
end

function collectstates(d::B200DeepESN, data::AbstractMatrix{T}, ps, st::NamedTuple) where {T}
    size(data, 2) >= 1 || throw(ArgumentError("collectstates data must have at least one column, got 0."))
    h = deep_handle(d, ps, T)
    states = Matrix{T}(undef, h.F, size(data, 2))            # eltype(data), esn_deep.jl:288
    h0 = deep_initial_carry(d, st, T); hT = similar(h0)
    check(ccall((:rcb_collect, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                h.ptr, Matrix{T}(data), size(data, 2), 1, h0, states, hT))
    return states, deep_store_carry(d, st, hT)
end

# predict(rc, ...) passes `nothing` as the reservoir for models without a :reservoir field (src/predict.jl:118-129)
function __predict(::Nothing, d::B200DeepESN, steps::Integer, ps, st; initialdata)
    steps >= 1 || throw(ArgumentError("steps must be ≥ 1, got $steps"))
    T = eltype(initialdata)
    h = deep_handle(d, ps, T); set_readout!(h, ps)
    Tc = promote_type(T, eltype(ps.readout.weight))
    carry = deep_initial_carry(d, st, Tc)
    Y = Matrix{Float64}(undef, d.readout.out_dims, steps)
    check(ccall((:rcb_predict_autoregressive, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Cvoid}, Ptr{Float64}),
                h.ptr, Vector{T}(initialdata), steps, 1, dtype_id(Tc), carry, Y))
    return Matrix{Tc}(Y), deep_store_carry(d, st, carry)
end
function __predict(::Nothing, d::B200DeepESN, data::AbstractMatrix{T}, ps, st) where {T}
    h = deep_handle(d, ps, T); set_readout!(h, ps)
    carry = deep_initial_carry(d, st, T)
    Y = Matrix{Float64}(undef, d.readout.out_dims, size(data, 2))
    check(ccall((:rcb_predict_teacher, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Ptr{Float64}),
                h.ptr, Matrix{T}(data), size(data, 2), 1, carry, Y))
    return Matrix{promote_type(T, eltype(ps.readout.weight))}(Y), deep_store_carry(d, st, carry)
end

# ---- feature plumbing + conceptor blocks (SURVEY 8f-2 / 8f-4) ------------------------------------------------
# The delay ESNs (src/models/esn_delay.jl) compose these around rcb_collect exactly like rcb200/api.py:_collect_delay:
#   input DelayLayer -> rcb_collect (cell) -> state DelayLayer -> remaining modifiers (rcb_apply_modifiers).
"""
    delay_features(x, num_delays, stride; history = nothing, clock = 0) -> (out, history′, clock′)

All `size(x, 2)` calls of a `DelayLayer` (src/layers/basic.jl:426-441) in one pass on the GPU.
"""
function delay_features(x::Matrix{T}, num_delays::Integer, stride::Integer; history = nothing, clock::Integer = 0) where {T}
    n, Tn = size(x)
    out = Matrix{T}(undef, n * (num_delays + 1), Tn)
    hout = Matrix{T}(undef, n, num_delays)
    hin = history === nothing ? C_NULL : Matrix{T}(history)
    check(ccall((:rcb_delay_features, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx().ptr, x, dtype_id(T), n, Tn, 1, num_delays, stride, hin, clock, out, num_delays > 0 ? hout : C_NULL))
    return out, hout, clock + Tn
end

"""`extend_features(u, x)` = `vcat(u, x)` column by column — `Extend`, src/states.jl:138-141."""
function extend_features(u::Matrix{T}, x::Matrix{T}) where {T}
    size(u, 2) == size(x, 2) || throw(DimensionMismatch("Extend: $(size(u, 2)) input columns, $(size(x, 2)) state columns"))
    out = Matrix{T}(undef, size(u, 1) + size(x, 1), size(x, 2))
    check(ccall((:rcb_extend_features, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Int32, Ptr{Cvoid}),
                ctx().ptr, u, size(u, 1), x, size(x, 1), size(x, 2), dtype_id(T), out))
    return out
end

"""`correlation_matrix_b200(states)` = `(states * states') / size(states, 2)` — src/conceptors.jl:28-33 on the K2 Gram kernels."""
function correlation_matrix_b200(states::AbstractMatrix{<:Real})
    isempty(states) && throw(ArgumentError("states must contain at least one state"))
    T = float(eltype(states)) === Float32 ? Float32 : Float64
    Z = Matrix{T}(states)
    R = Matrix{T}(undef, size(Z, 1), size(Z, 1))
    check(ccall((:rcb_correlation_matrix, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int64, Int64, Int32, Ptr{Cvoid}),
                ctx().ptr, Z, dtype_id(T), size(Z, 1), size(Z, 2), size(Z, 1), 0, R))
    return R
end

"""
    spectral_radius_b200(W) -> Float64

`maximum(abs, eigvals(W))` without the dense eigen-decomposition of `scale_radius!` (src/inits/inits_components.jl:126-138):
Arnoldi on the GPU (`rcb_spectral_radius_dense` / `_csc`).  `scale_radius_b200!(W, radius)` applies the same in-place scaling.
"""
function spectral_radius_b200(W::AbstractMatrix{<:Real})
    rho, dim = Ref{Float64}(0.0), Ref{Int32}(0)
    if W isa SparseMatrixCSC
        check(ccall((:rcb_spectral_radius_csc, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Float32}, Int32, Float64, Int32, Ref{Float64}, Ref{Int32}),
                    ctx().ptr, W.colptr, W.rowval, Vector{Float32}(W.nzval), size(W, 1), 0.0, 0, rho, dim))
    else
        Wd = Matrix{Float32}(W)
        check(ccall((:rcb_spectral_radius_dense, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Float32}, Int64, Int32, Float64, Int32, Ref{Float64}, Ref{Int32}),
                    ctx().ptr, Wd, size(Wd, 1), size(Wd, 1), 0.0, 0, rho, dim))
    end
    return rho[]
end
function scale_radius_b200!(W::AbstractMatrix, radius::AbstractFloat)
    rho = spectral_radius_b200(W)
    W .*= eltype(W)(radius) / eltype(W)(rho)
    any(!isfinite, W) && error("Sparsity too low for size of the matrix. Increase res_size or increase sparsity.")
    return W
end

"""
    b200(rc::AbstractReservoirComputer) -> ReservoirComputer

Same model (`ESN`, `ES2N`, `ResESN`, `EuSN` or a hand-built `ReservoirComputer` around one of their cells),
CUDA-backed reservoir: `train(b200(esn), X, Y, ps, st; objective = RidgeRegression(λ), solver = B200Cholesky())` runs
collect -> washout -> Gram -> solve as one library call.  `ps` / `st` keep the layout of the wrapped model (same field
names `reservoir`, `state_modifiers`, `readout`).
"""
b200(rc::AbstractReservoirComputer) = ReservoirComputer(B200Reservoir(rc.reservoir), rc.state_modifiers, rc.readout)
b200(rc::ReservoirComputing.DeepESN) = B200DeepESN(rc)

export B200Reservoir, B200DeepESN, B200Ridge, B200Cholesky, B200QR, b200, step_b200, gram_accumulate!, ridge_solve, Comm, comm_unique_id, allreduce_gc!,
       train_sharded, train_time_chunked, train_ensemble, spectral_radius_b200, scale_radius_b200!, delay_features, extend_features, correlation_matrix_b200

end # module
