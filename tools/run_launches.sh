cd $GRAFT_REPO_ROOT
T=${1:-l}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches.csv python tools/scan_probe.py 1.0 0 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${T}_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:][-14:]:
    print(r[ki][:60].ljust(60), r[vi], r[ui])
PY
