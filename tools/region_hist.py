"""instruction / stall-sample histogram along the SASS address axis of an `ncu --page source --csv --print-source sass`
export (first kernel in the file): usage region_hist.py file.csv STAGES [chunk]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = []
for r in rows[2:]:
    if len(r) < len(hdr): break
    data.append(r)
ci = {h: i for i, h in enumerate(hdr)}
S = float(sys.argv[2]); CH = int(sys.argv[3]) if len(sys.argv) > 3 else 250
tot = 0
for c in range(0, len(data), CH):
    ch = data[c:c + CH]
    ins = sum(int(r[ci["Instructions Executed"]]) for r in ch) / S
    smp = sum(int(r[ci["# Samples"]]) for r in ch)
    tot += smp
    if ins < 1 and smp < 200: continue
    st = collections.Counter()
    for r in ch:
        for h in hdr:
            if h.startswith("stall_") and "Not" not in h and r[ci[h]] not in ("", "0"): st[h[6:]] += int(r[ci[h]])
    mult = collections.Counter(round(int(r[ci["Instructions Executed"]]) / S, 1) for r in ch)
    print(c, "%.0f instr/stage" % ins, smp, dict(st.most_common(6)), mult.most_common(3))
print("total samples", tot)
