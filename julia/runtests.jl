# Smoke test of the Julia binding (needs Julia >= 1.10, ChainRulesCore, a B200 and the built libkanode_b200.so).
# Not runnable in the build image (no Julia): see DESIGN.md §1.
include("KANODEsB200.jl")
using .KANODEsB200, Random, Test
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
