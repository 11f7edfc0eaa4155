"""Turns gpurun_out/*.ncu-rep / launches csv into the small tracked summaries under profiles/.
usage: python scripts/summarize_profile.py <tag> <ncu-rep> [launches.csv]"""
import collections
import csv
import itertools
import re
import subprocess
import sys
from pathlib import Path

tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 else None
out = Path("profiles"); out.mkdir(exist_ok=True)
KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor", "gpu__dram_throughput", "lts__t_bytes.sum ", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__shared_mem_per_block", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
lines = [f"# ncu --set full summary: {rep} ({tag})", ""]
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    lines.append(f"## {r[hdr.index('Kernel Name')][:110]}")
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(k) or k in h for k in KEYS) and "max" not in h and "min" not in h:
            lines.append(f"{h} [{u}] = {v}")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
sections, cur = [], None
for r in srows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections:
    h, data = sec["hdr"], sec["data"]
    ix = {k: i for i, k in enumerate(h)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0
    tot = sum(f(r, "Instructions Executed") for r in data) or 1.0
    ops = collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        ops[m.group(2).split(".")[0] if m else "?"] += f(r, "Instructions Executed")
    lines += [f"## SASS of {sec['name'][:90]}: {len(data)} instructions, {tot:.4g} warp-instructions executed",
              "opcode mix: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in ops.most_common(12))]
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    st = collections.Counter({s_: sum(f(r, s_) for r in data) for s_ in stalls})
    ts = sum(st.values()) or 1
    lines.append("stall samples: " + ", ".join(f"{k[6:]} {v / ts * 100:.1f}%" for k, v in st.most_common(8)))
    lines.append("")
(out / f"{tag}_ncu_summary.md").write_text("\n".join(lines) + "\n")
if launches:
    r = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
    hd = r[0]; ik, iv = hd.index("Kernel Name"), hd.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for row in r[1:]:
        if len(row) > iv:
            nm = row[ik].split("<")[0].split("(")[0]; tot[nm] += float(row[iv].replace(",", "")); cnt[nm] += 1
    s = sum(tot.values())
    txt = [f"# ncu launch list ({tag}): gpu__time_duration.sum per kernel, --clock-control none (cold-cache, serialised: compare shares)", ""]
    txt += [f"{k:55s} launches={cnt[k]:3d} total_ms={v / 1e6:9.3f} share={v / s * 100:5.1f}%" for k, v in tot.most_common()]
    (out / f"{tag}_launches.md").write_text("\n".join(txt) + "\n")
    Path(out / f"{tag}_launches.csv").write_text(open(launches).read())
print("\n".join(lines[:60]))
