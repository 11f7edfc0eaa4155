# KANODEsB200.jl — Julia host side of the B200-native KAN-ODE hot path.
#
# Keeps the reference's driver surface (Lotka-Volterra/LV_driver_KANODE.jl:130-203,284 and the PDE scripts):
#   KDense(in, out, G; use_base_act, basis_func, normalizer)   -- same kwargs/defaults as src/kdense.jl:20-68
#   Lux-style setup:  ps, st = setup(rng, chain);  flat p = vec of [C1; W1; C2; W2] (ComponentArray data order)
#   node = NeuralODE(chain, tspan, Tsit5(); saveat);  Array(node(u0, p, st)[1])
#   Zygote.gradient(loss, p)[1]   for the reference's UNCHANGED loss(p): ChainRulesCore.rrule on the node call and on
#                                 `predict`; the pullback is kanode_solve_adjoint (dense forward + adjoint with dL/dpred)
#   (l::KDense)(x, p, st) -> (y, st)   direct layer call (kdense.jl:109-130) through a KANODE_RHS_MAP handle
#   u0::Matrix [n, B]              batches: one trajectory per column (the C layout u0[batch][n])
#   NeuralODE(...; devices = 0:7)  one handle driving several GPUs of the box (kanode_create_multi)
# and calls libkanode_b200.so (include/kanode.h) through `ccall` on pointers resolved once with dlsym.  Float64 arrays bind the *_f64 entry points (the
# reference drivers run Float64), Float32 arrays the float ones.
#
# NOTE: Julia is not installed in the build image of this repo, so this file has not been executed there; the
# identical C ABI is exercised by the Python mirror (kan_odes_b200/) and its tests.  Run `julia julia/runtests.jl`
# on a machine with Julia >= 1.10, ChainRulesCore and a B200 to validate it.
module KANODEsB200

using Libdl
import ChainRulesCore
const CRC = ChainRulesCore

export KDense, Chain, NeuralODE, SourceODE, Tsit5, setup, flatten_params, predict, loss_and_grad, mse_loss, solve_adjoint,
       edge_activations, set_regularizer!, rbf, rswaf, iqf, tanh_fast, softsign, sigmoid_fast, swish

const LIB = Ref{String}(get(ENV, "KANODE_B200_LIB",
                            joinpath(@__DIR__, "..", "kan_odes_b200", "csrc", "libkanode_b200.so")))
# symbols are resolved once (dlopen + dlsym) and called through the pointer: `ccall((name, lib), ...)` needs constants
const DL = Ref{Ptr{Cvoid}}(C_NULL)
const SYMS = Dict{Symbol, Ptr{Cvoid}}()
function fptr(name::Symbol)
    get!(SYMS, name) do
        DL[] == C_NULL && (DL[] = Libdl.dlopen(LIB[]))
        Libdl.dlsym(DL[], name)
    end
end

# ---- enums of include/kanode.h -------------------------------------------------------------------------
@enum Normalizer::Int32 NORM_TANH=0 NORM_SOFTSIGN=1 NORM_SIGMOID=2
@enum Basis::Int32 BASIS_RBF=0 BASIS_RSWAF=1 BASIS_IQF=2
const RHS_CHAIN = Int32(0); const RHS_SOURCE_LAPLACIAN = Int32(1); const RHS_MAP = Int32(2)
const ERR_SOLVER = Cint(-6)

# names the reference scripts pass as kwargs (utils.jl:8-62, NNlib)
struct Named{T}; code::T; end
const rbf = Named(BASIS_RBF); const rswaf = Named(BASIS_RSWAF); const iqf = Named(BASIS_IQF)
const tanh_fast = Named(NORM_TANH); const softsign = Named(NORM_SOFTSIGN); const sigmoid_fast = Named(NORM_SIGMOID)
const swish = :swish
struct Tsit5 end

# ---- POD descriptors (must match include/kanode.h byte for byte) ---------------------------------------
struct LayerDesc
    in_dims::Int32; out_dims::Int32; grid_len::Int32
    normalizer::Int32; basis::Int32; use_base_act::Int32
    grid_lo::Float32; grid_hi::Float32; denominator::Float32
    kind::Int32; dense_act::Int32          # kind 1: Lux.Dense(in => out, act) of the MLP-NODE baseline (LV_driver_MLP.jl:61)
end
const MAX_LAYERS = 8
struct Desc
    n_layers::Int32
    layers::NTuple{MAX_LAYERS, LayerDesc}
    rhs_kind::Int32
    n_state::Int32
    lap_coef::Float64
    dx::Float64
end
struct Stats; naccept::Int32; nreject::Int32; nf::Int32; retcode::Int32; end

# ---- layer surface (kdense.jl:5-107) ---------------------------------------------------------------------
struct KDense
    in_dims::Int; out_dims::Int; grid_len::Int
    normalizer::Named{Normalizer}; basis_func::Named{Basis}; use_base_act::Bool
    grid_lims::NTuple{2, Float32}; denominator::Float32
end
function KDense(in_dims::Int, out_dims::Int, grid_len::Int; normalizer = tanh_fast,
                grid_lims = (-1.0f0, 1.0f0), denominator = Float32(2 / (grid_len - 1)),
                basis_func = rbf, base_act = swish, use_base_act = true)
    @assert grid_lims[2] > grid_lims[1]                                    # kdense.jl:50-51
    base_act === swish || error("only base_act = swish is on the hot path")
    KDense(in_dims, out_dims, grid_len, normalizer, basis_func, use_base_act, Float32.(grid_lims), Float32(denominator))
end
parameterlength(l::KDense) = l.in_dims * l.grid_len * l.out_dims + (l.use_base_act ? l.in_dims * l.out_dims : 0)
statelength(l::KDense) = l.grid_len
struct Chain; layers::Vector{KDense}; end
Chain(ls::KDense...) = Chain(collect(ls))
parameterlength(c::Chain) = sum(parameterlength, c.layers)

glorot_uniform(rng, dims...) = (rand(rng, Float32, dims...) .- 0.5f0) .* sqrt(24.0f0 / sum(dims[1:2]))
function setup(rng, c::Chain)
    ps = [(; C = glorot_uniform(rng, l.out_dims, l.grid_len * l.in_dims),
             W = l.use_base_act ? glorot_uniform(rng, l.out_dims, l.in_dims) : nothing) for l in c.layers]
    st = [(; grid = collect(LinRange(l.grid_lims..., l.grid_len))) for l in c.layers]
    ps, st
end
flatten_params(ps) = vcat((vcat(vec(p.C), p.W === nothing ? Float32[] : vec(p.W)) for p in ps)...)

function Desc(c::Chain; rhs_kind = RHS_CHAIN, n_state = c.layers[1].in_dims, lap_coef = 0.0, dx = 1.0)
    zero_l = LayerDesc(0, 0, 0, 0, 0, 0, 0f0, 0f0, 0f0, 0, 0)
    ls = ntuple(MAX_LAYERS) do i
        i > length(c.layers) && return zero_l
        l = c.layers[i]
        LayerDesc(l.in_dims, l.out_dims, l.grid_len, Int32(l.normalizer.code), Int32(l.basis_func.code),
                  l.use_base_act, l.grid_lims[1], l.grid_lims[2], l.denominator, 0, 0)
    end
    Desc(length(c.layers), ls, rhs_kind, n_state, lap_coef, dx)
end

# ---- handle ---------------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}; n::Int; n_out::Int; np::Int
end
lasterr(h) = unsafe_string(ccall(fptr(:kanode_last_error), Cstring, (Ptr{Cvoid},), h))
check(rc, h, what) = rc == 0 || error("$what failed ($rc): $(lasterr(h))")

"One handle per model; `devices` with more than one entry shards every batch over those GPUs (kanode_create_multi)."
function Handle(c::Chain; device = 0, devices = nothing, kw...)
    d = Ref(Desc(c; kw...))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    if devices === nothing || length(devices) <= 1
        dev = devices === nothing ? device : first(devices)
        rc = ccall(fptr(:kanode_create), Cint, (Ref{Desc}, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), d, dev, C_NULL, out)
    else
        devs = collect(Cint, devices)
        rc = GC.@preserve devs ccall(fptr(:kanode_create_multi), Cint, (Ref{Desc}, Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), d, devs, length(devs), out)
    end
    check(rc, C_NULL, "kanode_create")        # KANODE_ERR_NO_DEVICE when no B200: there is no CPU fallback
    n_out = d[].rhs_kind == RHS_MAP ? c.layers[end].out_dims : Int(d[].n_state)
    h = Handle(out[], Int(d[].n_state), n_out, parameterlength(c))
    finalizer(x -> ccall(fptr(:kanode_destroy), Cint, (Ptr{Cvoid},), x.ptr), h)
end

suffix(::Type{Float64}) = "_f64"; suffix(::Type{Float32}) = ""
sym(name, T) = fptr(Symbol(name * suffix(T)))
const Real32or64 = Union{Float32, Float64}

function set_params!(h::Handle, p::AbstractVector{T}) where {T <: Real32or64}
    length(p) == h.np || error("expected $(h.np) parameters")
    pv = convert(Vector{T}, p)
    rc = GC.@preserve pv ccall(sym("kanode_set_params", T), Cint, (Ptr{Cvoid}, Ptr{T}, Csize_t), h.ptr, pv, length(pv))
    check(rc, h.ptr, "kanode_set_params")
end

# ---- direct layer call (kdense.jl:109-130; Activation_getter.jl:39, Allen-Cahn_Source.jl:91) --------------------
const MAP_HANDLES = IdDict{KDense, Handle}()
layer_params(p::NamedTuple) = vcat(vec(p.C), (haskey(p, :W) && p.W !== nothing) ? vec(p.W) : eltype(p.C)[])
layer_params(p::AbstractVector) = p
"`(l::KDense)(x, p, st) -> (y, st)`: x is [in_dims] or [in_dims, K] (one sample per column)."
function (l::KDense)(x::AbstractVecOrMat, p, st = nothing)
    h = get!(() -> Handle(Chain(l); rhs_kind = RHS_MAP), MAP_HANDLES, l)
    T = eltype(x) === Float64 ? Float64 : Float32
    set_params!(h, convert(Vector{T}, layer_params(p)))
    xm = convert(Matrix{T}, reshape(x, l.in_dims, :)); K = size(xm, 2)
    y = Matrix{T}(undef, l.out_dims, K)
    rc = GC.@preserve xm y ccall(sym("kanode_rhs", T), Cint, (Ptr{Cvoid}, Ptr{T}, Ptr{T}, Int64), h.ptr, xm, y, K)
    check(rc, h.ptr, "kanode_rhs")
    (x isa AbstractVector ? vec(y) : y), st
end

"act[o, i, k] of layer `layer` (1-based) at its inputs x [I_l, K] (LV/Activation_getter.jl); sum over i = layer output."
function edge_activations(h::Handle, c::Chain, layer::Int, x::AbstractMatrix{T}) where {T <: Real32or64}
    l = c.layers[layer]; K = size(x, 2)
    xm = convert(Matrix{T}, x); act = Array{T}(undef, l.out_dims, l.in_dims, K)
    rc = GC.@preserve xm act ccall(sym("kanode_edge_activations", T), Cint, (Ptr{Cvoid}, Int32, Ptr{T}, Ptr{T}, Int64), h.ptr, layer - 1, xm, act, K)
    check(rc, h.ptr, "kanode_edge_activations")
    act
end

"reg_loss(p, act_reg, entropy_reg) (LV_driver_KANODE.jl:187-201) is added inside every loss_and_grad / mse_loss from now on."
set_regularizer!(h::Handle, act_reg, entropy_reg = 0.0) =
    check(ccall(fptr(:kanode_set_regularizer), Cint, (Ptr{Cvoid}, Float64, Float64), h.ptr, act_reg, entropy_reg), h.ptr, "kanode_set_regularizer")

# ---- ODE surface (LV_driver_KANODE.jl:180-184; Allen-Cahn_Source.jl:96-99) ------------------------------------
struct NeuralODE
    model::Chain; tspan::NTuple{2, Float64}; saveat::Vector{Float64}; abstol::Float64; reltol::Float64; h::Handle
end
NeuralODE(model::Chain, tspan, ::Tsit5 = Tsit5(); saveat = Float64[], abstol = 1e-6, reltol = 1e-3, device = 0, devices = nothing) =
    NeuralODE(model, Float64.(tspan), collect(Float64, saveat), abstol, reltol, Handle(model; device, devices))
SourceODE(model::Chain, n_state, lap_coef, dx, tspan, ::Tsit5 = Tsit5(); saveat = Float64[], abstol = 1e-6, reltol = 1e-3, device = 0, devices = nothing) =
    NeuralODE(model, Float64.(tspan), collect(Float64, saveat), abstol, reltol,
              Handle(model; device, devices, rhs_kind = RHS_SOURCE_LAPLACIAN, n_state, lap_coef, dx))

# u: [n, nsave] for one trajectory (== Array(sol) of the reference), [n, nsave, B] for a batch
struct ODESolution{T, A <: AbstractArray{T}}; t::Vector{Float64}; u::A; stats::Vector{Stats}; end
Base.Array(s::ODESolution) = s.u

tol(::Type{Float32}, x) = Float32(x); tol(::Type{Float64}, x) = Float64(x)
batch_of(u0::AbstractVector) = 1
batch_of(u0::AbstractMatrix) = size(u0, 2)
shape_out(u0::AbstractVector, a) = reshape(a, size(a, 1), size(a, 2))
shape_out(u0::AbstractMatrix, a) = a

function solve_raw(node::NeuralODE, u0, p::AbstractVector{T}) where {T <: Real32or64}
    set_params!(node.h, p)
    B = batch_of(u0); ns = length(node.saveat)
    u = convert(Array{T}, float.(u0))                            # [n] or [n, B]: column b is trajectory b == u0[b][n] in C
    out = Array{T}(undef, node.h.n, ns, B); stats = Vector{Stats}(undef, B)
    rc = GC.@preserve u out stats ccall(sym("kanode_solve", T), Cint,
        (Ptr{Cvoid}, Ptr{T}, Int64, Float64, Float64, Ptr{Float64}, Int32, T, T, Ptr{T}, Ptr{Stats}),
        node.h.ptr, u, B, node.tspan[1], node.tspan[2], node.saveat, ns, tol(T, node.abstol), tol(T, node.reltol), out, stats)
    check(rc, node.h.ptr, "kanode_solve")
    ODESolution(copy(node.saveat), shape_out(u0, out), stats)
end

"`node(u0, p, st)` -> `(sol, st)`; u0::Vector is one trajectory like the reference drivers, u0::Matrix [n, B] a batch."
(node::NeuralODE)(u0::AbstractVecOrMat, p::AbstractVector{T}, st = nothing) where {T <: Real32or64} = (solve_raw(node, u0, p), st)

predict(node::NeuralODE, u0, p) = Array(node(u0, p)[1])

"""Pullback of the solve: given dL/dpred (the shape of `predict`) returns (pred, dL/dp, dL/du0) — kanode_solve_adjoint.
This is what `Zygote.gradient(loss, p)` needs for ANY loss of the predictions (LV_driver_KANODE.jl:197-203,284;
Burgers_Surrogate.jl:105-107,191)."""
function solve_adjoint(node::NeuralODE, u0, p::AbstractVector{T}, dpred::AbstractArray) where {T <: Real32or64}
    set_params!(node.h, p)
    B = batch_of(u0); ns = length(node.saveat)
    u = convert(Array{T}, float.(u0)); cot = convert(Array{T}, dpred)
    length(cot) == node.h.n * ns * B || error("cotangent has the wrong size")
    out = Array{T}(undef, node.h.n, ns, B); grad = Vector{T}(undef, node.h.np); du0 = similar(u)
    rc = GC.@preserve u cot out grad du0 ccall(sym("kanode_solve_adjoint", T), Cint,
        (Ptr{Cvoid}, Ptr{T}, Int64, Float64, Float64, Ptr{Float64}, Int32, T, T, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{T}, Ptr{Stats}, Ptr{Stats}),
        node.h.ptr, u, B, node.tspan[1], node.tspan[2], node.saveat, ns, tol(T, node.abstol), tol(T, node.reltol), cot, out, grad, du0, C_NULL, C_NULL)
    check(rc, node.h.ptr, "kanode_solve_adjoint")
    shape_out(u0, out), grad, du0
end

# rrule on `predict` and on the node call: the reference's `loss(p) = mean(abs2, Xn .- predict(ComponentArray(p, pM_axis)))`
# (optionally + reg_loss(p, 5e-4, 0)) differentiates unchanged under Zygote
function CRC.rrule(::typeof(predict), node::NeuralODE, u0, p)
    pred = predict(node, u0, p)
    function predict_pullback(Δ)
        _, g, du0 = solve_adjoint(node, u0, p, CRC.unthunk(Δ))
        (CRC.NoTangent(), CRC.NoTangent(), du0, g)
    end
    pred, predict_pullback
end
solution_cotangent(Δ) = Δ
solution_cotangent(Δ::CRC.Tangent) = hasproperty(Δ, :u) ? Δ.u : Δ          # Tangent{ODESolution}(u = ...)
function CRC.rrule(node::NeuralODE, u0::AbstractVecOrMat, p::AbstractVector, st)
    sol = solve_raw(node, u0, p)
    function node_pullback(Δ)
        Δ = CRC.unthunk(Δ)
        Δsol = Δ isa CRC.Tangent || Δ isa Tuple ? Δ[1] : Δ                  # cotangent of (sol, st)
        Δu = solution_cotangent(CRC.unthunk(Δsol))
        (Δu isa CRC.AbstractZero) && return (CRC.NoTangent(), CRC.ZeroTangent(), CRC.ZeroTangent(), CRC.NoTangent())
        _, g, du0 = solve_adjoint(node, u0, p, Δu)
        (CRC.NoTangent(), du0, g, CRC.NoTangent())
    end
    (sol, st), node_pullback
end

"loss(p) = mean(abs2, X .- predict(p)) and its gradient in ONE fused call (LV_driver_KANODE.jl:197-203,284); X: [n, nsave(, B)]."
function loss_and_grad(node::NeuralODE, u0, p::AbstractVector{T}, X::AbstractArray) where {T <: Real32or64}
    set_params!(node.h, p)
    B = batch_of(u0); ns = length(node.saveat)
    u = convert(Array{T}, float.(u0)); tg = convert(Array{T}, X)   # [n, nsave, B] column-major == target[B][nsave][n]
    loss = Ref{T}(0); grad = Vector{T}(undef, node.h.np); du0 = similar(u)
    rc = GC.@preserve u tg grad du0 ccall(sym("kanode_loss_grad", T), Cint,
        (Ptr{Cvoid}, Ptr{T}, Int64, Float64, Float64, Ptr{Float64}, Int32, Ptr{T}, T, T, Ref{T}, Ptr{T}, Ptr{T}, Ptr{Stats}, Ptr{Stats}),
        node.h.ptr, u, B, node.tspan[1], node.tspan[2], node.saveat, ns, tg, tol(T, node.abstol), tol(T, node.reltol),
        loss, grad, du0, C_NULL, C_NULL)
    check(rc, node.h.ptr, "kanode_loss_grad")                     # KANODE_ERR_SOLVER: a solve failed (the reference's loss throws)
    loss[], grad, du0
end

# the fused fast path for the mean-squared loss: one kanode_loss_grad call instead of solve + pullback
mse_loss(node::NeuralODE, u0, p, X) = loss_and_grad(node, u0, p, X)[1]
function CRC.rrule(::typeof(mse_loss), node::NeuralODE, u0, p, X)
    l, g, du0 = loss_and_grad(node, u0, p, X)
    pullback(Δ) = (CRC.NoTangent(), CRC.NoTangent(), Δ .* du0, Δ .* g, CRC.NoTangent())
    l, pullback
end

end # module
