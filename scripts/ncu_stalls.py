"""Summarise the source page of an ncu report: stall reasons (sampled) in total and the hottest SASS lines.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python scripts/ncu_stalls.py src.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# the file may hold several kernels: split at "Kernel Name" lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks:
    hdr, data = b["rows"][0], b["rows"][1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for d in data:
        for s in stall_cols:
            try: tot[s] += int(d[ix[s]])
            except Exception: pass
    allsamp = sum(tot.values())
    print("==", b["name"][:100], "samples", allsamp)
    for s, v in tot.most_common(10):
        print(f"   {s:28s} {v:9d} {100.0*v/max(allsamp,1):5.1f}%")
    si = ix["# Samples"]
    hot = sorted(data, key=lambda d: -int(d[si] or 0))[:top]
    for d in hot:
        st = sorted(((int(d[ix[s]] or 0), s) for s in stall_cols), reverse=True)[:2]
        print(f"   {int(d[si]):8d} {100.0*int(d[si])/max(allsamp,1):5.1f}%  {d[ix['Source']][:70]:70s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
