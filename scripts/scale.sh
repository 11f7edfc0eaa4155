#!/bin/bash
# 1/2/4/8-GPU weak-scaling sweep of bench.py on one box, the way the driver launches it (results under gpurun_out/).
#   gpurun --gpus 8 -- 'bash scripts/scale.sh r02j'
tag=${1:-scale}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \
    bench.py --impl reference --gpus 8 --steps 3 --warmup 1 --lean > gpurun_out/${tag}_ref_n8.json 2> gpurun_out/${tag}_ref_n8.err
python - <<PY
import json
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/${tag}_n{n}.json"))
        w = d.get("workloads") or {}
        print(n, round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["kernel_ms"], d["clocks"]["reasons"],
              {k: round(v["value"], 1) for k, v in w.items()})
    except Exception as e:
        print(n, "failed", e)
try:
    d = json.load(open("gpurun_out/${tag}_ref_n8.json")); print("ref n8", d["value"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("ref failed", e)
PY
