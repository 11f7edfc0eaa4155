# KANODEsB200.jl — Julia host side of the B200-native KAN-ODE hot path.
#
# Keeps the reference's driver surface (Lotka-Volterra/LV_driver_KANODE.jl:130-203,284 and the PDE scripts):
#   KDense(in, out, G; use_base_act, basis_func, normalizer)   -- same kwargs/defaults as src/kdense.jl:20-68
#   Lux-style setup:  ps, st = setup(rng, chain);  flat p = vec of [C1; W1; C2; W2] (ComponentArray data order)
#   node = NeuralODE(chain, tspan, Tsit5(); saveat);  Array(node(u0, p, st)[1])
#   Zygote.gradient(loss, p)[1]   via a ChainRulesCore.rrule on `predict`
# and calls libkanode_b200.so (include/kanode.h) through `ccall`.  Float64 arrays bind the *_f64 entry points (the
# reference drivers run Float64), Float32 arrays the float ones.
#
# NOTE: Julia is not installed in the build image of this repo, so this file has not been executed there; the
# identical C ABI is exercised by the Python mirror (kan_odes_b200/) and its tests.  Run `julia julia/runtests.jl`
# on a machine with Julia >= 1.10, ChainRulesCore and a B200 to validate it.
module KANODEsB200

using Libdl
import ChainRulesCore
const CRC = ChainRulesCore

export KDense, Chain, NeuralODE, SourceODE, Tsit5, setup, flatten_params, predict, loss_and_grad,
       rbf, rswaf, iqf, tanh_fast, softsign, sigmoid_fast, swish

const LIB = Ref{String}(get(ENV, "KANODE_B200_LIB",
                            joinpath(@__DIR__, "..", "kan_odes_b200", "csrc", "libkanode_b200.so")))

# ---- enums of include/kanode.h -------------------------------------------------------------------------
@enum Normalizer::Int32 NORM_TANH=0 NORM_SOFTSIGN=1 NORM_SIGMOID=2
@enum Basis::Int32 BASIS_RBF=0 BASIS_RSWAF=1 BASIS_IQF=2
const RHS_CHAIN = Int32(0); const RHS_SOURCE_LAPLACIAN = Int32(1)

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
    zero_l = LayerDesc(0, 0, 0, 0, 0, 0, 0f0, 0f0, 0f0)
    ls = ntuple(MAX_LAYERS) do i
        i > length(c.layers) && return zero_l
        l = c.layers[i]
        LayerDesc(l.in_dims, l.out_dims, l.grid_len, Int32(l.normalizer.code), Int32(l.basis_func.code),
                  l.use_base_act, l.grid_lims[1], l.grid_lims[2], l.denominator)
    end
    Desc(length(c.layers), ls, rhs_kind, n_state, lap_coef, dx)
end

# ---- handle ---------------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}; n::Int; np::Int
end
lasterr(h) = unsafe_string(ccall((:kanode_last_error, LIB[]), Cstring, (Ptr{Cvoid},), h))
check(rc, h, what) = rc == 0 || error("$what failed ($rc): $(lasterr(h))")

function Handle(c::Chain; device = 0, kw...)
    d = Ref(Desc(c; kw...))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:kanode_create, LIB[]), Cint, (Ref{Desc}, Cint, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), d, device, C_NULL, out)
    check(rc, C_NULL, "kanode_create")        # KANODE_ERR_NO_DEVICE when no B200: there is no CPU fallback
    h = Handle(out[], Int(d[].n_state), parameterlength(c))
    finalizer(x -> ccall((:kanode_destroy, LIB[]), Cint, (Ptr{Cvoid},), x.ptr), h)
end

suffix(::Type{Float64}) = "_f64"; suffix(::Type{Float32}) = ""
sym(name, T) = Symbol(name * suffix(T))

function set_params!(h::Handle, p::Vector{T}) where {T <: Union{Float32, Float64}}
    length(p) == h.np || error("expected $(h.np) parameters")
    rc = GC.@preserve p ccall((sym("kanode_set_params", T), LIB[]), Cint, (Ptr{Cvoid}, Ptr{T}, Csize_t), h.ptr, p, length(p))
    check(rc, h.ptr, "kanode_set_params")
end

# ---- ODE surface (LV_driver_KANODE.jl:180-184; Allen-Cahn_Source.jl:96-99) ------------------------------------
struct NeuralODE
    model::Chain; tspan::NTuple{2, Float64}; saveat::Vector{Float64}; abstol::Float64; reltol::Float64; h::Handle
end
NeuralODE(model::Chain, tspan, ::Tsit5 = Tsit5(); saveat = Float64[], abstol = 1e-6, reltol = 1e-3, device = 0) =
    NeuralODE(model, Float64.(tspan), collect(Float64, saveat), abstol, reltol, Handle(model; device))
SourceODE(model::Chain, n_state, lap_coef, dx, tspan, ::Tsit5 = Tsit5(); saveat = Float64[], abstol = 1e-6, reltol = 1e-3, device = 0) =
    NeuralODE(model, Float64.(tspan), collect(Float64, saveat), abstol, reltol,
              Handle(model; device, rhs_kind = RHS_SOURCE_LAPLACIAN, n_state, lap_coef, dx))

struct ODESolution{T}; t::Vector{Float64}; u::Matrix{T}; stats::Vector{Stats}; end      # u: [n, nsave] == Array(sol)
Base.Array(s::ODESolution) = s.u

tol(::Type{Float32}, x) = Float32(x); tol(::Type{Float64}, x) = Float64(x)

"`node(u0, p, st)` -> `(sol, st)`; one trajectory (u0::Vector) like the reference drivers."
function (node::NeuralODE)(u0::AbstractVector, p::Vector{T}, st = nothing) where {T <: Union{Float32, Float64}}
    set_params!(node.h, p)
    u = convert(Vector{T}, float.(u0)); ns = length(node.saveat)
    out = Matrix{T}(undef, node.h.n, ns); stats = Vector{Stats}(undef, 1)
    rc = GC.@preserve u out stats ccall((sym("kanode_solve", T), LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{T}, Int64, Float64, Float64, Ptr{Float64}, Int32, T, T, Ptr{T}, Ptr{Stats}),
        node.h.ptr, u, 1, node.tspan[1], node.tspan[2], node.saveat, ns, tol(T, node.abstol), tol(T, node.reltol), out, stats)
    check(rc, node.h.ptr, "kanode_solve")
    ODESolution{T}(copy(node.saveat), out, stats), st
end

predict(node::NeuralODE, u0, p) = Array(node(u0, p)[1])

"loss(p) = mean(abs2, X .- predict(p)) and its gradient in one call (LV_driver_KANODE.jl:197-203,284)."
function loss_and_grad(node::NeuralODE, u0::AbstractVector, p::Vector{T}, X::AbstractMatrix) where {T <: Union{Float32, Float64}}
    set_params!(node.h, p)
    u = convert(Vector{T}, float.(u0)); ns = length(node.saveat)
    tg = convert(Matrix{T}, X)                                  # [n, nsave] column-major == target[1][nsave][n]
    loss = Ref{T}(0); grad = Vector{T}(undef, node.h.np); du0 = Vector{T}(undef, node.h.n)
    rc = GC.@preserve u tg grad du0 ccall((sym("kanode_loss_grad", T), LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{T}, Int64, Float64, Float64, Ptr{Float64}, Int32, Ptr{T}, T, T, Ref{T}, Ptr{T}, Ptr{T}, Ptr{Stats}, Ptr{Stats}),
        node.h.ptr, u, 1, node.tspan[1], node.tspan[2], node.saveat, ns, tg, tol(T, node.abstol), tol(T, node.reltol),
        loss, grad, du0, C_NULL, C_NULL)
    check(rc, node.h.ptr, "kanode_loss_grad")
    loss[], grad, du0
end

# Zygote.gradient(p -> mean(abs2, X .- predict(node, u0, p)), p) keeps working: the pullback of `predict` needs
# dL/dpred for arbitrary losses, which the C ABI fuses only for the mean-squared loss.  For that loss (the only one
# the reference uses) differentiate `mse_loss` below; its rrule calls kanode_loss_grad once.
mse_loss(node::NeuralODE, u0, p, X) = loss_and_grad(node, u0, p, X)[1]
function CRC.rrule(::typeof(mse_loss), node::NeuralODE, u0, p, X)
    l, g, du0 = loss_and_grad(node, u0, p, X)
    pullback(Δ) = (CRC.NoTangent(), CRC.NoTangent(), Δ .* du0, Δ .* g, CRC.NoTangent())
    l, pullback
end

end # module
