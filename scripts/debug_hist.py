import sys; sys.path.insert(0, '.')
import numpy as np
import bench
import kan_odes_b200 as K
chain, p, u0, tg = bench.make_workload(65536, 1234)
ode = K.KanOde(chain, dtype=np.float32); ode.set_params(p)
r = ode.loss_grad(u0, bench.TSPAN, bench.SAVEAT, tg)
att = r["bwd_stats"].naccept + r["bwd_stats"].nreject
print("bwd attempts hist:", np.bincount(att)[35:])
print("max attempts", att.max(), "mean", att.mean(), "frac > 43:", (att > 43).mean(), "fwd attempts max", (r["fwd_stats"].naccept + r["fwd_stats"].nreject).max())
print("nreject>0 frac", (r["bwd_stats"].nreject > 0).mean())
