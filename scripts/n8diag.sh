#!/bin/bash
# 8-GPU diagnostics of the LV bench step: where does the time beyond the per-rank kernels go?
#   gpurun --gpus 8 -- 'bash scripts/n8diag.sh r02n8'
tag=${1:-n8diag}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 3 --lean --no-cpu "${@:3}" > gpurun_out/${tag}_$2.json 2> gpurun_out/${tag}_$2.err; }
run 29701 default
run 29702 fixed --fixed-params
KANODE_BENCH_NO_ALLREDUCE=1 run 29703 noallreduce
NCCL_NVLS_ENABLE=0 NCCL_ALGO=Ring NCCL_PROTO=LL run 29704 ringll
python - <<PY
import json
for k in ("default", "fixed", "noallreduce", "ringll"):
    try:
        d = json.load(open(f"gpurun_out/${tag}_{k}.json")); print(k, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 4), d["kernel_ms"])
    except Exception as e:
        print(k, "failed", e)
PY
