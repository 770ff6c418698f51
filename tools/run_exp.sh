cd $GRAFT_REPO_ROOT
for ch in 2 3; do
DMX_CHUNKS=$ch timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f32.json'));print('chunks $ch value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['e2e']['ms_per_step'],2),d['e2e']['kernel_ms'],d['e2e']['run_ms'])"
done
