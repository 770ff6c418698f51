cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for m in "--fp32-pairs" ""; do
timeout 300 python bench.py --steps 5 --warmup 3 $m --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
python -c "
import json;d=json.load(open('gpurun_out/bench_tmp.json'));print('$m',d['value'],d['e2e']['value'],d['breakdown_ms'])"
tail -n 3 gpurun_out/bench_tmp.err
done
