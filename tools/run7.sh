cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "records_match or golden" 2>&1 | tail -4
for d in 11 9 8; do
DMX_TRACE=gpurun_out/trace_$d.txt DMX_DEBUG=$d timeout 300 python bench.py --steps 3 --warmup 3 --fp32-pairs --no-cpu-baseline --scale 0.1 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
tail -n 2 gpurun_out/bench_tmp.err
python -c "
import json;d=json.load(open('gpurun_out/bench_tmp.json'));print('dbg $d',d['breakdown_ms']['pair_scan'])"
done
