cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "records_match or golden" 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --fp32-pairs --no-cpu-baseline > gpurun_out/bench_f32e.json 2> gpurun_out/bench_f32e.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f32e.json'));print('f32',d['value'],d['e2e']['value'],d['breakdown_ms'])"
tail -n 3 gpurun_out/bench_f32e.err
