cd $GRAFT_REPO_ROOT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b32.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_f32.csv') if l.startswith('"'))]
for r in rows[-14:]:
    print(r[4][:50], r[7], r[8], r[-1])
PY
