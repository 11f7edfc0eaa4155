#!/bin/bash
# End-of-round single-GPU evidence: GPU tests, the bench line, the reference arm, the ncu launch list and one --set full capture
# of the dominant kernel (each ncu pass only after its command has run clean without ncu).   gpurun -- 'bash scripts/final_n1.sh r03'
tag=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest.log
python bench.py > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref_n1.json 2> gpurun_out/${tag}_ref_n1.err; echo "ref rc=$?"
python bench.py --lean --no-cpu --steps 2 --warmup 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --lean --no-cpu --steps 2 --warmup 3 > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:small_backward_lg_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_lg \
    python bench.py --lean --no-cpu --steps 2 --warmup 3 > gpurun_out/${tag}_ncu2.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_n1.json"))
print(round(d["value"]), round(d["ms_per_step"], 4), round(d["e2e"]["value"]), d["kernel_ms"], d["roofline"]["frac"], d["clocks"]["reasons"])
print({k: (round(v["value"], 1), round(v["ms_per_step"], 3)) for k, v in (d.get("workloads") or {}).items()})
r = json.load(open("gpurun_out/${tag}_ref_n1.json")); print("ref", r.get("value"), r.get("cpu_baseline", {}).get("cores"))
PY
