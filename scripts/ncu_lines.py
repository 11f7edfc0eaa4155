"""Per-source-line hot spots of one ncu report: python scripts/ncu_lines.py <ncu-rep> [top N]
Aggregates the `--page source --print-source cuda,sass` view: instructions executed and stall samples per file:line."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        ix = {k: i for i, k in enumerate(hdr)}
        try:
            inst = float(r[ix["Instructions Executed"]]); samp = float(r[ix["# Samples"]])
        except ValueError:
            continue
        st = {k[6:]: float(r[i]) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k and r[i] not in ("", "-")}
        out.append((fname, int(r[0]), r[1].strip()[:90], inst, samp, st))
ti = sum(o[3] for o in out) or 1; ts = sum(o[4] for o in out) or 1
print(f"total warp-instructions {ti:.4g}, samples {ts:.0f}")
print("--- by instructions ---")
for f, ln, src, inst, samp, st in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{f}:{ln:5d} inst {inst / ti * 100:5.2f}% samp {samp / ts * 100:5.2f}% {src}")
print("--- by samples ---")
for f, ln, src, inst, samp, st in sorted(out, key=lambda o: -o[4])[:top]:
    s3 = ", ".join(f"{k} {v / max(samp, 1) * 100:.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{ln:5d} inst {inst / ti * 100:5.2f}% samp {samp / ts * 100:5.2f}% [{s3}] {src}")
