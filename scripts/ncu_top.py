"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples, executed
count and shared-memory wavefronts. Usage: python scripts/ncu_top.py file.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# first row is kernel name; second header
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
  try: return float(r[ix[k]].replace(",", ""))
  except Exception: return 0.0
tot_s = sum(num(r, "# Samples") for r in data)
tot_i = sum(num(r, "Instructions Executed") for r in data)
tot_w = sum(num(r, "L1 Wavefronts Shared") for r in data)
print(f"kernel: {rows[0][1][:100]}")
print(f"total samples {tot_s:.0f}  inst {tot_i:.3g}  smem wavefronts {tot_w:.3g}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(num(r, s) for r in data) for s in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100*v/max(tot_s,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("--- top by samples")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top]:
  st = sorted(((num(r, s), s[6:]) for s in stalls), reverse=True)[:2]
  print(f"{100*num(r,'# Samples')/max(tot_s,1):5.1f}%  exec {num(r,'Instructions Executed'):.3g}  wf {num(r,'L1 Wavefronts Shared'):.3g}/{num(r,'L1 Wavefronts Shared Ideal'):.3g}  {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}  {r[ix['Source']][:90]}")
print("--- top by smem wavefronts")
for r in sorted(data, key=lambda r: -num(r, "L1 Wavefronts Shared"))[:8]:
  print(f"wf {num(r,'L1 Wavefronts Shared'):.3g} ideal {num(r,'L1 Wavefronts Shared Ideal'):.3g} exec {num(r,'Instructions Executed'):.3g}  {r[ix['Source']][:90]}")
