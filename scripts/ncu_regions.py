"""Split a kernel's SASS (ncu source page csv) at barriers / mbarrier waits and report instructions + stall samples per region."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
reg = []; cur = {"start": 0, "inst": 0, "samp": 0, "wf": 0, "n": 0, "first": ""}
tot_i = tot_s = 0
for k, d in enumerate(data):
    src = d[ix["Source"]]
    inst = int(d[ix["Instructions Executed"]] or 0); samp = int(d[ix["# Samples"]] or 0)
    try: wf = int(d[ix["L1 Wavefronts Shared"]] or 0)
    except: wf = 0
    cur["inst"] += inst; cur["samp"] += samp; cur["wf"] += wf; cur["n"] += 1
    tot_i += inst; tot_s += samp
    if "BAR.SYNC" in src or "SYNCS.PHASECHK" in src or "EXIT" in src or "BAR.ARV" in src:
        cur["end"] = k; cur["last"] = src.strip()[:40]
        reg.append(cur); cur = {"start": k + 1, "inst": 0, "samp": 0, "wf": 0, "n": 0}
reg.append(cur)
print(f"total inst {tot_i} samples {tot_s}")
for r in reg:
    if r["inst"] == 0 and r["samp"] == 0: continue
    print(f"sass[{r['start']:5d}..{r.get('end', 0):5d}] n={r['n']:4d} inst {100*r['inst']/tot_i:5.1f}%  samples {100*r['samp']/tot_s:5.1f}%  smem-wf {r['wf']/1e6:7.1f}M  ends: {r.get('last','')}")
