cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --fp32-pairs --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f32.json'));print('f32',d['value'],d['e2e']['value'],d['breakdown_ms'], d['e2e'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fp32-pairs > gpurun_out/ncu_b32.log 2>&1
grep -v "^==" gpurun_out/launches_f32.csv | tail -9 | cut -d, -f5,15
