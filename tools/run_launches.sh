cd $GRAFT_REPO_ROOT
T=${1:-ll}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python tools/scan_probe.py 1.0 0 > gpurun_out/ncu_b.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${T}_launches.csv")))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[h]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki].split('(')[0],float(r[vi].replace(',',''))) for r in rows[h+1:] if len(r)>vi]
idx=[i for i,(k,v) in enumerate(seq) if 'pair_scan' in k]
s,e=idx[-2],idx[-1]
for k,v in seq[s:e]: print("%-40s %9.1f us"%(k[:40],v/1e3))
PY
