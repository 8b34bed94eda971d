"""Opcode histogram, stall totals and hottest source lines of one kernel from an .ncu-rep (source page).
usage: python profiles/sass_hist.py report.ncu-rep <warp_steps: warps x steps, to normalise> [kernel-regex]"""
import csv, subprocess, sys, re
from collections import Counter
rep, ws = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
op, stall, tot = Counter(), Counter(), 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    try: n = int(r[ix["Instructions Executed"]])
    except Exception: continue
    tot += n
    o = r[ix["Source"]].split()
    name = o[1] if o[0].startswith("@") else o[0]
    op[name.split(".")[0]] += n
    for c in stall_cols:
        try: stall[c] += int(r[ix[c]])
        except Exception: pass
print(f"instructions per warp-step: {tot / ws:.1f}")
print("  ".join(f"{k} {v / ws:.1f}" for k, v in op.most_common(24)))
ts = sum(stall.values())
print("stalls: " + "  ".join(f"{k[6:]} {100 * v / ts:.1f}%" for k, v in stall.most_common(10)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
lines, cur_file, hdr = [], "", None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 3 and r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iN = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > iN and r[0] not in ("", "Function Name"):
        try: lines.append((int(r[iS] or 0), int(r[iN] or 0), cur_file, r[0], r[1].strip()[:100]))
        except Exception: pass
tot_s = sum(x[0] for x in lines) or 1
print("hottest lines (samples%, inst/warp-step, file:line, source):")
for smp, n, f, ln, src in sorted(lines, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"  {100 * smp / tot_s:5.1f}%  {n / ws:7.1f}  {f}:{ln:>4s}  {src}")
