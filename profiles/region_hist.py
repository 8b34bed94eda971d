"""Split a kernel's SASS at barriers / mbarrier waits and report stall samples and executed instructions per region.
usage: python profiles/region_hist.py report.ncu-rep <warp_steps>"""
import csv, subprocess, sys
rep, ws = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try: data.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), {k: int(r[ix[k]] or 0) for k in keys}))
    except Exception: pass
tot = sum(d[1] for d in data)
reg, cur = [], [0, 0, 0, "start", {}]
for i, (src, smp, n, st) in enumerate(data):
    cur[0] += smp; cur[1] += n; cur[2] += 1
    for k, v in st.items(): cur[4][k] = cur[4].get(k, 0) + v
    if any(k in src for k in ("BAR.SYNC", "UCGABAR_WAIT", "SYNCS.PHASECHK")):
        cur[3] += f" -> #{i} {src[:34]}"; reg.append(tuple(cur)); cur = [0, 0, 0, f"#{i}", {}]
reg.append(tuple(cur))
for smp, n, cnt, name, st in reg:
    if smp * 200 > tot or n / ws > 5:
        top = ", ".join(f"{k[6:]} {100 * v / max(smp, 1):.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:4])
        print(f"{100 * smp / tot:5.1f}%  inst/warp-step {n / ws:6.1f}  static {cnt:4d}  {name:62s} {top}")
