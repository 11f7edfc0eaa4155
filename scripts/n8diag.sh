#!/bin/bash
# Diagnostic (8 GPUs): LV bench with and without the gradient all-reduce, to isolate the collective's share of the step at N=8.
for v in 1 ""; do
  KANODE_BENCH_NO_ALLREDUCE=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2958${#v} bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null > gpurun_out/n8diag_${#v}.json
  python -c "
import json;d=json.load(open('gpurun_out/n8diag_${#v}.json'));print('no_allreduce=[$v]',d['value'],d['ms_per_step'],d['kernel_ms'])"
done
