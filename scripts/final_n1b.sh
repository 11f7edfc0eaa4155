#!/bin/bash
# GPU tests + the bench line (no ncu).   gpurun -- 'bash scripts/final_n1b.sh r03b'
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_n1.json"))
print(round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["kernel_ms"], d["roofline"]["frac"], d["clocks"]["reasons"])
print({k: (round(v["value"], 1), round(v["ms_per_step"], 3)) for k, v in (d.get("workloads") or {}).items()})
print(d.get("parity")); print(d.get("no_history")); print(d.get("f64"))
PY
