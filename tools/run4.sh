cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 5 --warmup 3 --fp32-pairs --no-cpu-baseline > gpurun_out/bench_f32b.json 2> gpurun_out/bench_f32b.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f32b.json'));print('f32',d['value'],d['e2e']['value'],d['breakdown_ms'])"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f64b.json 2> gpurun_out/bench_f64b.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f64b.json'));print('f64',d['value'],d['e2e']['value'],d['breakdown_ms'])"
tail -3 gpurun_out/bench_f32b.err gpurun_out/bench_f64b.err
