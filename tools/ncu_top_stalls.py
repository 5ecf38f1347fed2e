#!/usr/bin/env python
"""Top stall sites of every kernel in an .ncu-rep, from `ncu --page source --csv` (needs -lineinfo and
--import-source on).  Usage: ncu_top_stalls.py rep [topN].  Prints per kernel the SASS lines with the most
warp-stall samples, with their source line, so that a capture can be summarised on the GPU box (the .ncu-rep of
a --set full run is too large to carry back)."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = re.split(r"\n(?=\"Kernel Name\")", out)
cur = None
rows = list(csv.reader(io.StringIO(out)))
# the source page prints one table per kernel, each introduced by a header row starting with "#"/"Address"
kernel = "?"; hdr = None; data = []
def flush():
    if not hdr or not data: return
    def col(name):
        for i, h in enumerate(hdr):
            if h.strip().lower() == name: return i
        return None
    ci = col("warp stall sampling (all samples)") or col("warp stall sampling (all cycles)") or col("# samples")
    si = col("source"); ai = col("address")
    if ci is None: print("  (no stall-sampling column; columns:", hdr[:12], ")"); return
    tot = 0; items = []
    for r in data:
        try: v = float(r[ci].replace(",", "") or 0)
        except (ValueError, IndexError): continue
        tot += v; items.append((v, r))
    items.sort(key=lambda x: -x[0])
    print(f"== {kernel[:100]}  total samples {tot:.0f}")
    for v, r in items[:topn]:
        print(f"   {100*v/max(tot,1):5.1f}%  {r[si][:110] if si is not None else ''}")
for r in rows:
    if not r: continue
    if r[0].startswith("Kernel Name") or (len(r) == 2 and r[0] == "Kernel Name"):
        flush(); kernel = r[1] if len(r) > 1 else "?"; hdr = None; data = []; continue
    if r[0] in ("#", "Address") or (hdr is None and any(h.strip() == "Source" for h in r)):
        flush(); hdr = r; data = []; continue
    if hdr is not None: data.append(r)
flush()
