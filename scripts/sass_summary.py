"""Per-kernel SASS evidence for libb200nufft.so (sm_100a): counts of the instructions that prove the
Blackwell-native paths -- UTMALDG (TMA tile loads), UTMAREDG (TMA reduce-add tile flush), SYNCS
(mbarrier), FFMA2 (packed fp32 FMA), REDG (vector global reductions), ATOMS (shared-memory atomics:
expected 0 in the spreaders), LDGSTS (cp.async), and the absence of tensor-core instructions
(north_star: no stage is a dense contraction).  Usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tensorflow_nufft_b200", "libb200nufft.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTMALDG", "UTMAREDG", "SYNCS", "FFMA2", "FFMA", "REDG", "ATOMS", "ATOMG", "LDGSTS", "LDS", "STS", "UTCHMMA", "HMMA", "LDTM"]
per = collections.OrderedDict()
cur = None
arch = set()
for ln in txt.splitlines():
  m = re.search(r"Function : (\S+)", ln)
  if m:
    cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
    cur = re.sub(r"\(.*", "", cur).replace("b200::", "").replace("void ", "")
    per.setdefault(cur, collections.Counter())
    continue
  m = re.search(r"arch = (sm_\w+)", ln)
  if m: arch.add(m.group(1))
  if cur is None: continue
  m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
  if not m: continue
  op = m.group(2)
  for k in KEYS:
    if op == k or (k in ("REDG", "ATOMS", "ATOMG", "LDS", "STS", "LDGSTS") and op.startswith(k)):
      per[cur][k] += 1
print(f"# SASS summary of {os.path.basename(so)}; cubin architectures: {sorted(arch)}")
print("# kernel family (template instances merged) : instances, then instruction counts summed over instances")
fam = collections.OrderedDict()
for name, c in per.items():
  f = re.sub(r"<.*", "", name)
  e = fam.setdefault(f, [0, collections.Counter()])
  e[0] += 1
  e[1].update(c)
tot = collections.Counter()
for f, (n, c) in sorted(fam.items()):
  tot.update(c)
  print(f"{f:34s} n={n:3d}  " + "  ".join(f"{k}={c[k]}" for k in KEYS if c[k]))
print("TOTAL".ljust(34) + "        " + "  ".join(f"{k}={tot[k]}" for k in KEYS))
print(f"tensor-core instructions (UTC*MMA / HMMA / LDTM): {tot['UTCHMMA'] + tot['HMMA'] + tot['LDTM']} (none expected)")
