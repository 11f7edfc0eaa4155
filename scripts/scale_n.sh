#!/bin/bash
# One N of the weak-scaling sweep, the way the driver launches it:  gpurun --gpus N -- 'bash scripts/scale_n.sh r03 N [bench flags]'
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
    bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_n$n.json"))
w = d.get("workloads") or {}
print($n, round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["kernel_ms"], d["clocks"]["reasons"], d["step_ms_spread"],
      {k: (round(v["value"], 1), round(v["ms_per_step"], 3)) for k, v in w.items()})
PY
