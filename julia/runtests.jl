# Smoke test of the Julia binding (needs Julia >= 1.10, ChainRulesCore, Zygote, a B200 and the built libkanode_b200.so).
# Not runnable in the build image (no Julia): see DESIGN.md §1.  The same C ABI is exercised by tests/ through ctypes.
include("KANODEsB200.jl")
using .KANODEsB200, Random, Test, Statistics
import Zygote
rng = Random.default_rng(); Random.seed!(rng, 0)
kan1 = Chain(KDense(2, 10, 5; use_base_act = true, basis_func = rbf, normalizer = tanh_fast),
             KDense(10, 2, 5; use_base_act = true, basis_func = rbf, normalizer = tanh_fast))
pM, stM = setup(rng, kan1)
p = Float64.(flatten_params(pM)) ./ 1e5                      # LV_driver_KANODE.jl:175
@test length(p) == 240
t_train = collect(0.0:0.1:3.4)
node = NeuralODE(kan1, (0.0, 3.5), Tsit5(); saveat = t_train)
pred = predict(node, [1, 1], p)
@test size(pred) == (2, 35)
X = ones(2, 35)
l, g, _ = loss_and_grad(node, [1, 1], p, X)
@test isfinite(l) && length(g) == 240
# the reference's loss shape, unchanged, under Zygote (LV_driver_KANODE.jl:197-203,284): pullback = kanode_solve_adjoint
loss(p) = mean(abs2, X .- Array(node([1, 1], p, stM)[1]))
gz = Zygote.gradient(loss, p)[1]
@test maximum(abs.(gz .- g)) <= 1e-8 * maximum(abs.(g))
# direct layer call and a batch of initial conditions
y, _ = kan1.layers[1]([1.0, 1.0], pM[1], stM[1])
@test length(y) == 10
sol, _ = node([1.0 0.5; 1.0 2.0], p, stM)
@test size(Array(sol)) == (2, 35, 2)
