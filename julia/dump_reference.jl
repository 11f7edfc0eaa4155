# dump_reference.jl — pins the repo's CPU oracle to the REAL reference (closes "parity unpinned", DESIGN.md §2).
#
# Runs the unmodified reference sources (Lotka-Volterra/src/*.jl through `include`) and the bodies of
#   Lotka-Volterra/LV_driver_KANODE.jl:111-203,284   (data, model, NeuralODE, predict, loss, Zygote gradient)
#   PDE examples/Burgers_Surrogate.jl:82-107         (surrogate model, NeuralODE, predict, loss) + the gradient call of :191
# at FIXED parameters / initial conditions read from tests/golden/julia_inputs.json (written by scripts/make_julia_inputs.py, so
# both sides use bit-identical inputs; the Julia Xoshiro stream is never needed), and writes what tests/test_reference_dump.py
# compares with the oracle:  sol.t, Array(sol), sol.stats.{naccept,nreject,nf}, the accepted-step times of the forward solve,
# loss(p) and Zgrad(loss, p)[1].
#
#   julia --project=<reference>/Lotka-Volterra julia/dump_reference.jl <reference root> [tests/golden]
#
# Needs Julia 1.11.1 with the packages pinned in <reference>/Lotka-Volterra/Manifest.toml (instantiate that project first).
# This file cannot run in the repo's build image (no Julia, no network); it is the one command a maintainer with Julia runs.
using Random, Lux, LinearAlgebra, Statistics
using NNlib, ConcreteStructs, WeightInitializers, ChainRulesCore
using ComponentArrays
using OrdinaryDiffEq, DiffEqFlux, SciMLSensitivity
using Zygote: gradient as Zgrad

refroot = length(ARGS) >= 1 ? ARGS[1] : error("usage: dump_reference.jl <reference root> [golden dir]")
golden  = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "..", "tests", "golden")

include(joinpath(refroot, "Lotka-Volterra", "src", "KolmogorovArnold.jl"))
using .KolmogorovArnold

# ---- minimal JSON (numbers / vectors only): no package outside the reference Manifest is needed ----------------------
function read_inputs(path)
    txt = read(path, String)
    d = Dict{String, Vector{Float64}}()
    for m in eachmatch(r"\"([A-Za-z0-9_]+)\"\s*:\s*\[([^\]]*)\]", txt)
        d[m.captures[1]] = isempty(strip(m.captures[2])) ? Float64[] : parse.(Float64, split(m.captures[2], ","))
    end
    d
end
jnum(x::Integer) = string(x)
jnum(x::AbstractFloat) = isfinite(x) ? repr(Float64(x)) : "null"
jvec(v) = "[" * join((jnum(x) for x in v), ",") * "]"
function write_json(path, pairs)
    open(path, "w") do io
        println(io, "{")
        for (i, (k, v)) in enumerate(pairs)
            print(io, "  \"", k, "\": ", v isa AbstractArray ? jvec(vec(v)) : jnum(v), i < length(pairs) ? ",\n" : "\n")
        end
        println(io, "}")
    end
end

inp = read_inputs(joinpath(golden, "julia_inputs.json"))

# ======================================================================================================================
# Lotka-Volterra (LV_driver_KANODE.jl).  Two parameter sets: p_init = glorot/1e5 (driver state at iteration 0, :175) and
# p_dyn = unscaled glorot (non-trivial field).  Checklist of SURVEY.md §8c verified by each dumped field:
#   sol_u        a4-a8 (KDense forward, rbf, tanh_fast(::Float64), swish), a10 (Tsit5 stages), a15 (dense output at saveat)
#   naccept/nreject/nf, step_t   a11 (error norm), a12 (PI controller + fastpower), a13 (initial dt), a14 (tstop clipping)
#   zgrad        a6 (rrule(_rbf)), a16 (InterpolatingAdjoint: callback/FSAL order at the save times, error norm over [lambda; g])
#   loss         a17
# ======================================================================================================================
function lotka!(du, u, p, t)                                   # LV_driver_KANODE.jl:46-50
    α, β, γ, δ = p
    du[1] = α * u[1] - β * u[2] * u[1]
    du[2] = γ * u[1] * u[2] - δ * u[2]
end
timestep = 0.1                                                 # :111-127
tspan = (0.0, 14)
tspan_train = (0.0, 3.5)
u0 = [1, 1]
p_ = Float32[1.5, 1, 1, 3]
prob = ODEProblem(lotka!, u0, tspan, p_)
solution = solve(prob, Tsit5(), abstol = 1e-12, reltol = 1e-12, saveat = timestep)
end_index = Int64(floor(length(solution.t) * tspan_train[2] / tspan[2]))
t = solution.t
t_train = t[1:end_index]
X = Array(solution)
Xn = deepcopy(X)

basis_func = rbf                                               # :130-143
normalizer = tanh_fast
layer_width = 10
grid_size = 5
kan1 = Lux.Chain(
    KDense( 2, layer_width, grid_size; use_base_act = true, basis_func, normalizer),
    KDense(layer_width,  2, grid_size; use_base_act = true, basis_func, normalizer),
)
rng = Random.default_rng(); Random.seed!(rng, 0)
pM, stM = Lux.setup(rng, kan1)
pM_axis = getaxes(ComponentArray(pM))                          # :173-174

train_node = NeuralODE(kan1, tspan_train, Tsit5(), saveat = t_train)      # :180
function predict(p)                                            # :182-184
    Array(train_node(u0, p, stM)[1])
end
function loss(p)                                               # :197-203 with sparse_on == 0
    mean(abs2, Xn[:, 1:end_index] .- predict(ComponentArray(p, pM_axis)))
end

for (tag, key) in (("lv_init", "lv_p_init"), ("lv_dyn", "lv_p_dyn"))
    p = inp[key]                                               # Float64 like (pM_data)./1e5 of :175
    @assert length(p) == 240
    sol = train_node(u0, ComponentArray(p, pM_axis), stM)[1]
    # the same problem without saveat: sol.t are the accepted-step times of the forward integrator
    dudt(u, p_, t_) = first(kan1(u, p_, stM))
    free = solve(ODEProblem(dudt, float.(u0), tspan_train, ComponentArray(p, pM_axis)), Tsit5())
    g = Zgrad(loss, p)[1]                                      # :284
    write_json(joinpath(golden, "julia_$(tag).json"), [
        "sol_t" => sol.t, "sol_u" => Array(sol), "naccept" => sol.stats.naccept, "nreject" => sol.stats.nreject, "nf" => sol.stats.nf,
        "step_t" => free.t, "free_naccept" => free.stats.naccept, "free_nreject" => free.stats.nreject, "free_nf" => free.stats.nf,
        "target" => Xn[:, 1:end_index], "loss" => loss(p), "zgrad" => g])
    println("wrote julia_$(tag).json: naccept=", sol.stats.naccept, " nf=", sol.stats.nf, " loss=", loss(p))
end

# ======================================================================================================================
# Burgers surrogate (Burgers_Surrogate.jl:82-107,191): [n,10,n] softsign model; u0 / targets / p from the input file
# (the script's MethodOfLines data generation is not part of the hot path).  n = 41 like the reference (xspan -1:0.05:1).
# ======================================================================================================================
let
    n = 41
    basis_func  = rbf                                          # Burgers_Surrogate.jl:82-88
    normalizer  = softsign
    KANgrid     = 5
    kanb = Lux.Chain(
        KDense(n, 10, KANgrid; use_base_act = true, basis_func, normalizer),
        KDense(10, n, KANgrid; use_base_act = true, basis_func, normalizer),
    )
    rngb = Random.default_rng(); Random.seed!(rngb, 0)
    pMb, stMb = Lux.setup(rngb, kanb)
    pMb_axis = getaxes(ComponentArray(pMb))
    tspanb = (0.0, 1.0)
    dt_train = [0.0, 0.1, 0.3, 0.5, 0.7, 0.9]                  # :68
    u0b = inp["burgers_u0"]
    Xb = reshape(inp["burgers_target"], n, length(dt_train))'   # Xₙ: [nsave, n] like :69-74
    nodeb = NeuralODE(kanb, tspanb, Tsit5(), saveat = dt_train)               # :97
    predictb(p) = Array(nodeb(u0b, p, stMb)[1])                               # :100-102
    lossb(p) = mean(abs2, Xb .- predictb(ComponentArray(p, pMb_axis))')       # :105-107 (note the transposed target)
    p = Float32.(inp["burgers_p"])                             # the PDE scripts keep Float32 parameters (:159)
    sol = nodeb(u0b, ComponentArray(p, pMb_axis), stMb)[1]
    g = Zgrad(lossb, p)[1]                                     # :191
    write_json(joinpath(golden, "julia_burgers.json"), [
        "sol_t" => sol.t, "sol_u" => Array(sol), "naccept" => sol.stats.naccept, "nreject" => sol.stats.nreject, "nf" => sol.stats.nf,
        "loss" => lossb(p), "zgrad" => g])
    println("wrote julia_burgers.json: naccept=", sol.stats.naccept, " loss=", lossb(p))
end
