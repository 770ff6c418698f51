cd $GRAFT_REPO_ROOT
for d in 0 1 3; do
DMX_DEBUG=$d timeout 300 python bench.py --steps 3 --warmup 3 --fp32-pairs --no-cpu-baseline --scale 0.3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
python -c "
import json;d=json.load(open('gpurun_out/bench_tmp.json'));print('dbg $d',d['breakdown_ms']['pair_scan'])"
done
