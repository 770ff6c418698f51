"""per-warp-role stall summary of an `ncu --page source --csv --print-source sass` export of a warp-specialised kernel:
rows are grouped by how often they executed relative to the number of pipeline stages"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = []
for r in rows[2:]:
    if len(r) < len(hdr): break
    data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
S = float(sys.argv[2])           # pipeline stages executed by the whole grid
groups = [(float(x.split(":")[0]), x.split(":")[1]) for x in sys.argv[3].split(",")]   # "1:loader,4:transform,..."
role = collections.defaultdict(collections.Counter); instr = collections.Counter()
for r in data:
    e = int(r[ci["Instructions Executed"]]); q = e / S
    k = "rare" if q < 0.02 else "other"
    for m, name in groups:
        if abs(q - m) < 0.06 * m: k = name
    instr[k] += e
    for h in hdr:
        if h.startswith("stall_") and "Not" not in h and r[ci[h]] not in ("", "0"): role[k][h] += int(r[ci[h]])
for k in role:
    print("%-16s %8.1f instr/stage  %7d samples  %s" % (k, instr[k] / S, sum(role[k].values()), dict(role[k].most_common(7))))
