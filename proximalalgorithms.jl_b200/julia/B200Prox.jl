# B200Prox.jl -- reference-side binding of libproxb200.so (include/proxb200.h) for ProximalAlgorithms.jl v0.7.
#
# STATUS: UNVERIFIED IN THIS REPOSITORY'S CI.  Julia is not installed in the build image nor on the GPU box, so this file
# has never been executed; it is the `ccall` shim a ProximalAlgorithms.jl maintainer would add (INTEGRATION.md walks
# through it).  Everything it calls is exercised, with the same arguments and in the same order, by the Python host
# (proximalalgorithms.jl_b200/algorithms.py), which IS tested against the oracle on B200 hardware.
#
# What it does: ProximalAlgorithms' structs are parametric in the array type `Tx`
# (src/algorithms/fast_forward_backward.jl:44,60), and `Base.iterate(iter, state)` is an ordinary method.  We add a
# device-vector type `B200Vector{T}` and specialise the two iterators on it, so that
#
#     x0 = B200Vector(ctx, zeros(Float32, n))
#     f  = B200LeastSquares(ctx, A, b);  g = ProximalOperators.NormL1(lam)
#     ProximalAlgorithms.FastForwardBackward(tol = 1f-6)(x0 = x0, f = f, g = g, Lf = Lf)
#
# runs the UNCHANGED driver loop (src/ProximalAlgorithms.jl:114-123) with one fused kernel per iteration.

module B200Prox

using LinearAlgebra
using ProximalCore
using ProximalOperators: NormL1, IndBox, IndBallL2, NormL21
import ProximalAlgorithms
import ProximalAlgorithms: FastForwardBackwardIteration, ForwardBackwardIteration, AdaptiveNesterovSequence

const LIB = get(ENV, "PROXB200_LIB", joinpath(@__DIR__, "..", "lib", "libproxb200.so"))

const PB_F32, PB_F64 = Cint(0), Cint(1)
const PB_PROX_ZERO, PB_PROX_L1, PB_PROX_BOX, PB_PROX_SCALE, PB_PROX_L21 = Cint(0), Cint(1), Cint(2), Cint(3), Cint(4)
const S_GSUM, S_RESSQ, S_GDR, S_RESINF, S_AUX = 1, 3, 5, 7, 9        # 1-based indices into the scalar block
const NSCALARS = 16

struct PbProx               # mirrors `struct pb_prox` (include/proxb200.h)
    kind::Cint
    group::Cint
    p0::Cdouble
    p1::Cdouble
    v0::Ptr{Cvoid}
    v1::Ptr{Cvoid}
end

pbdtype(::Type{Float32}) = PB_F32
pbdtype(::Type{Float64}) = PB_F64

function check(rc::Cint)
    rc == 0 || error("libproxb200: ", unsafe_string(ccall((:pb_last_error, LIB), Cstring, ())))
    return nothing
end

# ---- context ------------------------------------------------------------------------------------------------------
mutable struct B200Context
    h::Ptr{Cvoid}
    scalars::Vector{Float64}
    function B200Context(device::Integer = 0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pb_ctx_create, LIB), Cint, (Cint, Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), device, C_NULL, 0, h))
        ctx = new(h[], zeros(Float64, NSCALARS))
        finalizer(c -> ccall((:pb_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), ctx)
        return ctx
    end
end

"Copy the device scalar block to the host (the one synchronisation per iteration) and return it."
function read_scalars!(ctx::B200Context)
    check(ccall((:pb_read_scalars, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.h, ctx.scalars))
    return ctx.scalars
end
pairsum(s, i) = s[i] + s[i+1]          # rounded value of a double-double pair

# ---- device vector ------------------------------------------------------------------------------------------------
mutable struct B200Vector{T<:Union{Float32,Float64}} <: AbstractVector{T}
    ctx::B200Context
    ptr::Ptr{Cvoid}
    n::Int
    function B200Vector{T}(ctx::B200Context, n::Integer) where {T}
        p = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:pb_malloc, LIB), Cint, (Ptr{Cvoid}, Csize_t, Ref{Ptr{Cvoid}}), ctx.h, n * sizeof(T), p))
        v = new{T}(ctx, p[], n)
        finalizer(w -> ccall((:pb_free, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), w.ctx.h, w.ptr), v)
        return v
    end
end
function B200Vector(ctx::B200Context, host::Vector{T}) where {T}
    v = B200Vector{T}(ctx, length(host))
    check(ccall((:pb_upload, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), ctx.h, v.ptr, host, sizeof(host)))
    check(ccall((:pb_ctx_sync, LIB), Cint, (Ptr{Cvoid},), ctx.h))      # `host` may be collected after return
    return v
end
Base.size(v::B200Vector) = (v.n,)
Base.similar(v::B200Vector{T}) where {T} = B200Vector{T}(v.ctx, v.n)
function Base.copy(v::B200Vector{T}) where {T}                      # fast_forward_backward.jl:74, forward_backward.jl:66
    w = similar(v)
    check(ccall((:pb_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), v.ctx.h, w.ptr, v.ptr, v.n * sizeof(T)))
    return w
end
function Base.zero(v::B200Vector{T}) where {T}
    w = similar(v)
    check(ccall((:pb_memset_zero, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), v.ctx.h, w.ptr, v.n * sizeof(T)))
    return w
end
function Base.Array(v::B200Vector{T}) where {T}
    host = Vector{T}(undef, v.n)
    check(ccall((:pb_download, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), v.ctx.h, host, v.ptr, sizeof(host)))
    return host
end
Base.getindex(v::B200Vector, i::Int) = Array(v)[i]      # debugging only: every call is a full download

function LinearAlgebra.norm(v::B200Vector{T}, p::Real = 2) where {T}
    check(ccall((:pb_nrm2sq, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}), v.ctx.h, pbdtype(T), v.n, v.ptr))
    s = read_scalars!(v.ctx)
    p == Inf && return T(s[11])
    p == 2 || error("B200Vector: only norm(., 2) and norm(., Inf) are accelerated")
    return T(sqrt(pairsum(s, S_AUX)))
end
function LinearAlgebra.dot(a::B200Vector{T}, b::B200Vector{T}) where {T}
    check(ccall((:pb_dot, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}), a.ctx.h, pbdtype(T), a.n, a.ptr, b.ptr))
    return T(pairsum(read_scalars!(a.ctx), S_AUX))
end

# ---- proximable terms: descriptors for ProximalOperators' own types ---------------------------------------------------
descriptor(::ProximalCore.Zero, ::Type) = PbProx(PB_PROX_ZERO, 0, 0.0, 0.0, C_NULL, C_NULL)
descriptor(g::NormL1{<:Real}, ::Type{T}) where {T} = PbProx(PB_PROX_L1, 0, Float64(T(g.lambda)), 0.0, C_NULL, C_NULL)
descriptor(g::IndBox{<:Real,<:Real}, ::Type{T}) where {T} = PbProx(PB_PROX_BOX, 0, Float64(T(g.lb)), Float64(T(g.ub)), C_NULL, C_NULL)
descriptor(g::NormL21, ::Type{T}, group::Integer) where {T} = PbProx(PB_PROX_L21, Cint(group), Float64(T(g.lambda)), 0.0, C_NULL, C_NULL)
value_from(g::NormL1, ::Type{T}, s) where {T} = T(g.lambda) * T(pairsum(s, S_GSUM))
value_from(g, ::Type{T}, s) where {T} = zero(T)

"prox!(z, g, y, gamma) -> g(z)   (ProximalCore contract; call sites fast_forward_backward.jl:80,141)"
function ProximalCore.prox!(z::B200Vector{T}, g::Union{NormL1,IndBox,ProximalCore.Zero}, y::B200Vector{T}, gamma) where {T}
    d = Ref(descriptor(g, T))
    check(ccall((:pb_prox_apply, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cdouble, Ref{PbProx}, Ptr{Cvoid}),
                y.ctx.h, pbdtype(T), y.n, y.ptr, Float64(gamma), d, z.ptr))
    return value_from(g, T, read_scalars!(y.ctx))
end

# ---- smooth term: dense least squares with the benchmark's explicit gradient (benchmark/benchmarks.jl:11-17) -----------
struct B200LeastSquares{T}
    ctx::B200Context
    A::B200Vector{T}          # column-major m x n, as Julia stores it
    b::B200Vector{T}
    r::B200Vector{T}
    m::Int
    n::Int
end
function B200LeastSquares(ctx::B200Context, A::Matrix{T}, b::Vector{T}) where {T}
    m, n = size(A)
    return B200LeastSquares{T}(ctx, B200Vector(ctx, vec(A)), B200Vector(ctx, b), B200Vector{T}(ctx, m), m, n)
end

"Enqueue r = A x - b (AUX = ||r||^2) and grad = A' r; the value is read with the next scalar read-back."
function value_and_gradient_into!(grad::B200Vector{T}, f::B200LeastSquares{T}, x::B200Vector{T}) where {T}
    check(ccall((:pb_lsq_dense_residual, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                f.ctx.h, pbdtype(T), f.m, f.n, f.A.ptr, f.m, x.ptr, f.b.ptr, f.r.ptr))
    check(ccall((:pb_lsq_dense_gradient, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                f.ctx.h, pbdtype(T), f.m, f.n, f.A.ptr, f.m, f.r.ptr, grad.ptr))
    return nothing
end
lsq_value(::Type{T}, s) where {T} = (nr = T(sqrt(pairsum(s, S_AUX))); nr^2 / 2)      # norm(res)^2/2, sqrt-then-square

# ---- column shard of a dense A on a rank of a B200World (SURVEY.md section 8e): x is row-sharded, so A is column-sharded; the chunk
# partials of A x are all-gathered and folded in global chunk order INSIDE the combine kernel (pb_lsq_dense_residual_sharded), which
# makes r, f and the whole solve bit-identical to one GPU.  Shard boundaries: multiples of dense_chunk_cols(T, m, n).
struct B200LeastSquaresShard{T}
    ctx::B200Context
    A::B200Vector{T}             # this rank's columns, column-major (m x n_local)
    b::B200Vector{T}             # replicated
    r::B200Vector{T}             # replicated on return
    m::Int
    n_local::Int
    n_global::Int
    col_offset::Int
end
dense_chunk_cols(::Type{T}, m, n) where {T} = Int(ccall((:pb_lsq_dense_chunk_cols, LIB), Int64, (Cint, Int64, Int64), pbdtype(T), m, n))
function dense_shard_bounds(::Type{T}, m, n, P) where {T}       # 0-based half-open column ranges, one per rank
    cc = dense_chunk_cols(T, m, n)
    nch = cld(n, cc)
    base, rem = divrem(nch, P)
    out = Tuple{Int,Int}[]
    c = 0
    for r in 0:P-1
        k = base + (r < rem ? 1 : 0)
        push!(out, (min(n, c * cc), min(n, (c + k) * cc)))
        c += k
    end
    return out
end
function B200LeastSquaresShard(ctx::B200Context, A::Matrix{T}, b::Vector{T}, lo::Int, hi::Int) where {T}
    m, n = size(A)
    return B200LeastSquaresShard{T}(ctx, B200Vector(ctx, vec(A[:, lo+1:hi])), B200Vector(ctx, b), B200Vector{T}(ctx, m), m, hi - lo, n, lo)
end
function value_and_gradient_into!(grad::B200Vector{T}, f::B200LeastSquaresShard{T}, x::B200Vector{T}) where {T}
    check(ccall((:pb_lsq_dense_residual_sharded, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Cint),
                f.ctx.h, pbdtype(T), f.m, f.n_local, f.A.ptr, f.m, x.ptr, f.b.ptr, f.r.ptr, f.n_global, f.col_offset, 0))
    check(ccall((:pb_lsq_dense_gradient, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}),
                f.ctx.h, pbdtype(T), f.m, f.n_local, f.A.ptr, f.m, f.r.ptr, grad.ptr))
    return nothing
end

function ProximalAlgorithms.value_and_gradient(f::B200LeastSquares{T}, x::B200Vector{T}) where {T}
    grad = similar(x)
    value_and_gradient_into!(grad, f, x)
    return lsq_value(T, read_scalars!(f.ctx)), grad
end

# ---- FastForwardBackward specialised on B200Vector ----------------------------------------------------------------------
# Own state type: same field names as FastForwardBackwardState (fast_forward_backward.jl:60-71); `y` and `res` are
# materialised on demand (they cost 2 extra vector writes per iteration if always stored).
mutable struct B200FFBState{R,T}
    x::B200Vector{T}
    f_x::R
    grad_f_x::B200Vector{T}
    gamma::R
    z::B200Vector{T}
    g_z::R
    z_prev::B200Vector{T}
    extrapolation_sequence::Any
    x_next::B200Vector{T}
    beta_next::R
    res_inf::R
end

function fused_step!(st::B200FFBState{R,T}, g, beta) where {R,T}
    d = Ref(descriptor(g, T))
    check(ccall((:pb_ffb_step, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ref{PbProx}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                st.x.ctx.h, pbdtype(T), st.x.n, st.x.ptr, st.grad_f_x.ptr, st.z_prev.ptr, Float64(st.gamma), Float64(beta), d,
                C_NULL, st.z.ptr, C_NULL, st.x_next.ptr))
    s = read_scalars!(st.x.ctx)                       # the single host sync of the iteration
    st.g_z = value_from(g, T, s)
    st.res_inf = R(s[S_RESINF])
    return s
end

next_beta!(st::B200FFBState) = st.extrapolation_sequence isa AdaptiveNesterovSequence ?
    ProximalAlgorithms.next!(st.extrapolation_sequence, st.gamma) : first(st.extrapolation_sequence)

# init: fast_forward_backward.jl:73-97 (fixed stepsize: `gamma` or `Lf` given)
function Base.iterate(iter::FastForwardBackwardIteration{R,<:B200Vector{T}}) where {R,T}
    iter.gamma === nothing && error("B200Prox: this iterator specialisation covers fixed stepsizes; use b200_solve (pb_solve) for the adaptive line search")
    x = copy(iter.x0)
    grad = similar(x)
    value_and_gradient_into!(grad, iter.f, x)
    seq = iter.extrapolation_sequence !== nothing ? Iterators.Stateful(iter.extrapolation_sequence) : AdaptiveNesterovSequence(iter.mf)
    st = B200FFBState{R,T}(x, zero(R), grad, R(iter.gamma), similar(x), zero(R), copy(x), seq, similar(x), zero(R), zero(R))
    st.beta_next = next_beta!(st)
    s = fused_step!(st, iter.g, st.beta_next)
    st.f_x = lsq_value(R, s)
    return st, st
end

# step: fast_forward_backward.jl:106-145; the extrapolation of :135 was produced by the previous fused pass
function Base.iterate(iter::FastForwardBackwardIteration{R,<:B200Vector{T}}, st::B200FFBState{R,T}) where {R,T}
    st.x, st.x_next = st.x_next, st.x                 # :135
    st.z_prev, st.z = st.z, st.z_prev                 # :136
    value_and_gradient_into!(st.grad_f_x, iter.f, st.x)      # :138-139
    st.beta_next = next_beta!(st)
    s = fused_step!(st, iter.g, st.beta_next)         # :140-142 (+ next :135)
    st.f_x = lsq_value(R, s)
    return st, st
end

ProximalAlgorithms.default_stopping_criterion(tol, ::FastForwardBackwardIteration, st::B200FFBState) = st.res_inf / st.gamma <= tol
ProximalAlgorithms.default_solution(::FastForwardBackwardIteration, st::B200FFBState) = st.z

# ======================================================================================================================
# "Next" rows (SURVEY.md section 8f): L-BFGS operator and Douglas-Rachford on B200Vector.  Same status as the rest of this file:
# unexecuted here; the Python hosts accel.py / douglas_rachford.py issue exactly these calls and are tested on B200.
# ======================================================================================================================
const S_AUX2, S_AUX3 = 13, 15                              # <s,y>, <y,y> of pb_lbfgs_update (1-based hi words)
const PB_PROX_SQRL2 = Cint(5)

pair(s::Vector{Float64}, i::Int) = s[i] + s[i + 1]          # a double-double sum of the scalar block, rounded once

# ---- LBFGSOperator{M} on device vectors: src/accel/lbfgs.jl:5-28 (storage), :102-104 (initialize) ---------------------
mutable struct B200LBFGSOperator{R,T}
    h::Ptr{Cvoid}
    ctx::B200Context
end

function ProximalAlgorithms.initialize(::ProximalAlgorithms.LBFGS{M}, x::B200Vector{T}) where {M,T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pb_lbfgs_create, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Cint, Ref{Ptr{Cvoid}}), x.ctx.h, pbdtype(T), length(x), M, h))
    op = B200LBFGSOperator{real(T),T}(h[], x.ctx)
    finalizer(o -> ccall((:pb_lbfgs_destroy, LIB), Cint, (Ptr{Cvoid},), o.h), op)
    return op
end

ProximalAlgorithms.acceleration_style(::Type{<:B200LBFGSOperator}) = ProximalAlgorithms.QuasiNewtonStyle()

# reset!, lbfgs.jl:53-56
ProximalAlgorithms.reset!(L::B200LBFGSOperator) = (check(ccall((:pb_lbfgs_reset, LIB), Cint, (Ptr{Cvoid},), L.h)); L)

# update!(L, s, y), lbfgs.jl:30-51: one fused pass (ring slot + <s,y> + <y,y>), one scalar read-back, host-side `if ys > 0`
function ProximalAlgorithms.update!(L::B200LBFGSOperator, s::B200Vector, y::B200Vector)
    check(ccall((:pb_lbfgs_update, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                L.ctx.h, L.h, s.ptr, C_NULL, y.ptr, C_NULL))
    sc = read_scalars!(L.ctx)
    acc = Ref{Cint}(0)
    check(ccall((:pb_lbfgs_commit, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ref{Cint}), L.h, pair(sc, S_AUX2), pair(sc, S_AUX3), acc))
    return L
end

# mul!(d, L, v), lbfgs.jl:66-95: the two-loop recursion as 2*currmem + 2 launches with device-resident coefficients
function LinearAlgebra.mul!(d::B200Vector{T}, L::B200LBFGSOperator, v::B200Vector{T}) where {T}
    check(ccall((:pb_lbfgs_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                L.ctx.h, L.h, v.ptr, 1.0, d.ptr, C_NULL, C_NULL))
    return d
end
Base.:*(L::B200LBFGSOperator, v::B200Vector) = mul!(similar(v), L, v)

# PANOC's fused form (panoc.jl:114-117 + :183): d = -(H * res) and x_d = x + d in the last launch of the chain
function panoc_direction!(d::B200Vector{T}, x_d::B200Vector{T}, L::B200LBFGSOperator, res::B200Vector{T}, x::B200Vector{T}) where {T}
    check(ccall((:pb_lbfgs_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                L.ctx.h, L.h, res.ptr, -1.0, d.ptr, x.ptr, x_d.ptr))
    return d
end

# `tau .* x_d .+ (1 - tau) .* z_curr` and friends (panoc.jl:214-215, :234-237), products and sum rounded separately
function lincomb2!(out::B200Vector{T}, a, x::B200Vector{T}, b, y::B200Vector{T}) where {T}
    check(ccall((:pb_lincomb2, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}),
                out.ctx.h, pbdtype(T), length(out), Float64(a), x.ptr, Float64(b), y.ptr, out.ptr))
    return out
end

# ---- DouglasRachford for element-wise f, g: src/algorithms/douglas_rachford.jl:54-63 as ONE kernel per iteration ----------
mutable struct B200DRState{R,T}
    x::B200Vector{T}            # current x
    x_in::B200Vector{T}         # x before the last update (y, r, z, res are recomputed from it on demand)
    res_inf::R
end

function dr_pass!(iter, st::B200DRState{R,T}, y, r, z, res) where {R,T}
    ctx = st.x.ctx
    fd, gd = Ref(descriptor(iter.f, T)), Ref(descriptor(iter.g, T))
    check(ccall((:pb_dr_step, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cdouble, Ref{PbProx}, Ref{PbProx}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx.h, pbdtype(T), length(st.x), st.x_in.ptr, Float64(iter.gamma), fd, gd, st.x.ptr, y, r, z, res))
    st.res_inf = R(read_scalars!(ctx)[S_RESINF])
end

function Base.iterate(iter::ProximalAlgorithms.DouglasRachfordIteration{R,C,<:B200Vector{T}},
                      st::B200DRState{R,T} = B200DRState{R,T}(copy(iter.x0), similar(iter.x0), zero(R))) where {R,C,T}
    st.x_in, st.x = st.x, st.x_in
    dr_pass!(iter, st, C_NULL, C_NULL, C_NULL, C_NULL)
    return st, st
end

ProximalAlgorithms.default_stopping_criterion(tol, iter::ProximalAlgorithms.DouglasRachfordIteration, st::B200DRState) =
    st.res_inf / iter.gamma <= tol
function ProximalAlgorithms.default_solution(iter::ProximalAlgorithms.DouglasRachfordIteration, st::B200DRState{R,T}) where {R,T}
    y = similar(st.x)                                        # state.y = prox_f(x_in): re-run the pass with y requested
    dr_pass!(iter, st, y.ptr, C_NULL, C_NULL, C_NULL)
    return y
end

# ======================================================================================================================
# Whole solves inside the library: pb_solve (csrc/solve.cu, csrc/persist.cu, csrc/step_multi.cu).
#
# The iterator specialisations above keep ProximalAlgorithms' own driver loop (one `ccall` + one scalar read-back per iteration).
# For the built-in terms the library can also run the loop itself -- same kernels, same scalar arithmetic in R, identical
# iterates and iteration counts (tests/test_gpu_solvers.py::test_native_driver_equals_python_host) -- and then picks the fastest
# form on its own:
#   * dense least squares that stays cache resident (the benchmark suite's 5x10 .. 500x1000 problems): ONE persistent cooperative
#     kernel for the whole solve, line search and Nesterov sequence on the device (`persistent_ctas > 0` in the result);
#   * fixed stepsize + element-wise gradient source (SquaredDistance): ONE persistent kernel looping over the iterations
#     (`multi_iter_kernel = 1`), with the per-iteration scalar exchange between GPUs done inside the kernel;
#   * otherwise one fused kernel per iteration, pipelined one iteration ahead of the stop test.
# A maintainer wires it in by adding a method for the solver call, e.g.
#
#     function (alg::ProximalAlgorithms.IterativeAlgorithm{<:FastForwardBackwardIteration})(; x0::B200Vector, f::B200LeastSquares, g, kwargs...)
#         alg.stop === default_stop && !alg.verbose || return invoke(...)        # custom stop / display: keep Julia's loop
#         return b200_solve(f, g, x0; fast = true, maxit = alg.maxit, tol = <tol of the default stopping criterion>, kwargs...)
#     end
# ======================================================================================================================
const PB_F_LSQ_DENSE, PB_F_LSQ_BLOCKDIAG, PB_F_SQDIST, PB_F_LINEAR = Cint(0), Cint(1), Cint(2), Cint(3)
const PB_ALG_FB, PB_ALG_FFB = Cint(0), Cint(1)
const PB_SEQ_ADAPTIVE = Cint(0)
const PB_OPT_FUSED_EXCHANGE, PB_OPT_PERSISTENT, PB_OPT_MULTI_ITER = Cint(4), Cint(5), Cint(6)

struct PbSmooth              # mirrors `struct pb_smooth`
    kind::Cint
    pad::Cint
    m::Int64
    n::Int64
    lda::Int64
    nblk::Int64
    mb::Int64
    nb::Int64
    A::Ptr{Cvoid}
    b::Ptr{Cvoid}
    r::Ptr{Cvoid}
end

struct PbSolveOpts           # mirrors `struct pb_solve_opts`
    algorithm::Cint
    adaptive::Cint
    sequence::Cint
    profile::Cint
    maxit::Int64
    n_global::Int64
    tol::Cdouble
    gamma::Cdouble
    mf::Cdouble
    constant_beta::Cdouble
    minimum_gamma::Cdouble
    reduce_gamma::Cdouble
    increase_gamma::Cdouble
    spare_x::Ptr{Cvoid}
    spare_z::Ptr{Cvoid}
    spare_grad::Ptr{Cvoid}
end

mutable struct PbSolveResult  # mirrors `struct pb_solve_result`
    iterations::Int64
    backtracks::Int64
    gamma::Cdouble
    f_x::Cdouble
    g_z::Cdouble
    res_inf::Cdouble
    warned_small_gamma::Cint
    persistent_ctas::Cint
    x::Ptr{Cvoid}
    grad::Ptr{Cvoid}
    z::Ptr{Cvoid}
    z_prev::Ptr{Cvoid}
    loop_ms::Cdouble
    step_kernel_ms::Cdouble
    step_kernel_launches::Int64
    res_sq::Cdouble
    gdr::Cdouble
    gsum::Cdouble
    multi_iter_kernel::Cint
    pad::Cint
    PbSolveResult() = new()
end

smooth_descriptor(f::B200LeastSquares) = PbSmooth(PB_F_LSQ_DENSE, 0, f.m, f.n, f.m, 0, 0, 0, f.A.ptr, f.b.ptr, f.r.ptr)
smooth_descriptor(f::B200LeastSquaresShard) =      # pb_smooth of a column shard: nblk = col_offset, nb = n_global (proxb200.h)
    PbSmooth(PB_F_LSQ_DENSE, 0, f.m, f.n_local, f.m, f.col_offset, 0, f.n_global, f.A.ptr, f.b.ptr, f.r.ptr)

"SquaredDistance (benchmark/benchmarks.jl:19-28) on a device vector: f(x) = norm(x - b)^2 / 2."
struct B200SquaredDistance{T}
    b::B200Vector{T}
end
smooth_descriptor(f::B200SquaredDistance) = PbSmooth(PB_F_SQDIST, 0, 0, length(f.b), 0, 0, 0, 0, C_NULL, f.b.ptr, C_NULL)

"""
    b200_solve(f, g, x0; fast = true, maxit = 10_000, tol = 1e-8, Lf = nothing, gamma = nothing, mf = 0, n_global = 0, ...)

ForwardBackward (`fast = false`) / FastForwardBackward (`fast = true`) with the reference's keyword arguments and defaults
(forward_backward.jl:38-48, fast_forward_backward.jl:44-56), run by `pb_solve`.  Returns `(z, iterations, result)`; `x0` is not
mutated (upload copies, as every reference test asserts).
"""
function b200_solve(f, g, x0::B200Vector{T}; fast::Bool = true, maxit::Integer = 10_000, tol::Real = 1e-8, Lf = nothing, gamma = nothing,
                    adaptive = nothing, mf::Real = 0, minimum_gamma::Real = 1e-7, reduce_gamma::Real = 0.5, increase_gamma::Real = 1.0,
                    n_global::Integer = 0) where {T}
    R = real(T)
    gam = gamma === nothing ? (Lf === nothing ? nothing : 1 / Lf) : gamma          # fast_forward_backward.jl:49
    adapt = adaptive === nothing ? gam === nothing : adaptive                       # :50
    x = copy(x0)
    vecs = [similar(x) for _ in 1:8]                                                # grad, z, z_prev, x_next, grad_z, scratch, spare_x, spare_z
    grad, z, z_prev, x_next, grad_z, scratch, spare_x, spare_z = vecs
    fd, gd = Ref(smooth_descriptor(f)), Ref(descriptor(g, T))
    pipelined = fast && !adapt
    opts = Ref(PbSolveOpts(fast ? PB_ALG_FFB : PB_ALG_FB, adapt ? 1 : 0, PB_SEQ_ADAPTIVE, 0, maxit, n_global, Float64(tol),
                           gam === nothing ? 0.0 : Float64(R(gam)), Float64(R(mf)), 0.0, Float64(R(minimum_gamma)), Float64(R(reduce_gamma)),
                           Float64(R(increase_gamma)), pipelined ? spare_x.ptr : C_NULL, pipelined ? spare_z.ptr : C_NULL,
                           pipelined ? scratch.ptr : C_NULL))
    res = PbSolveResult()
    GC.@preserve x vecs f g begin
        check(ccall((:pb_solve, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Ref{PbSmooth}, Ref{PbProx}, Ref{PbSolveOpts}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                     Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{PbSolveResult}),
                    x.ctx.h, pbdtype(T), length(x), fd, gd, opts, x.ptr, grad.ptr, z.ptr, z_prev.ptr, x_next.ptr, grad_z.ptr, scratch.ptr, res))
    end
    res.warned_small_gamma != 0 && @warn "stepsize `gamma` became too small ($(R(res.gamma)))"      # fb_tools.jl:59-61
    zsol = first(v for v in (x, vecs...) if v.ptr == res.z)                         # buffers are swapped by pointer inside the loop
    return zsol, Int(res.iterations), res
end

# ---- one Julia process, all GPUs of the node (SURVEY.md section 8b "Threading"): one context and one task per device ------------
"""
    B200World(devices) -> contexts connected for the in-kernel scalar exchange (pb_xchg_connect_local)

Row-shard every n-vector over the contexts (32-element aligned ranges), then run the SAME `b200_solve` on every shard from its own
task with `n_global = n`: the step kernels exchange their scalar blocks over NVLink inside the kernel, every rank takes identical
decisions, and the concatenated shards equal the single-GPU solution bit for bit (tests/test_gpu_local_world.py).
Rule (as for NCCL): no device-wide synchronisation (allocation, `cudaDeviceSynchronize`) on one rank while another is inside a solve.
"""
struct B200World
    ctxs::Vector{B200Context}
end
function B200World(devices::AbstractVector{<:Integer})
    ctxs = [B200Context(d) for d in devices]
    P = length(ctxs)
    for (r, c) in enumerate(ctxs)
        check(ccall((:pb_xchg_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), c.h, r - 1, P, C_NULL))
    end
    hs = [c.h for c in ctxs]
    check(ccall((:pb_xchg_connect_local, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), hs, P))
    for c in ctxs
        check(ccall((:pb_ctx_set_option, LIB), Cint, (Ptr{Cvoid}, Cint, Cint), c.h, PB_OPT_FUSED_EXCHANGE, 1))
    end
    return B200World(ctxs)
end

"Run `solve_shard(rank, ctx)` (which calls `b200_solve` on that rank's shard) on every context concurrently; one task per GPU."
function on_every_gpu(solve_shard, w::B200World)
    out = Vector{Any}(undef, length(w.ctxs))
    Threads.@sync for (r, c) in enumerate(w.ctxs)
        Threads.@spawn begin
            check(ccall((:pb_ctx_make_current, LIB), Cint, (Ptr{Cvoid},), c.h))
            out[r] = solve_shard(r - 1, c)
        end
    end
    return out
end

end # module
