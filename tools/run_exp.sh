cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f32.json'));print('f32 value',round(d['value']),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],2),d['breakdown_ms'],d['e2e'])"
