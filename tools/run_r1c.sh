cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --fp32-pairs --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f32.json'));print('f32',d['value'],d['e2e']['value'],d['breakdown_ms'])"
tail -n 3 gpurun_out/bench_f32.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
python -c "
import json;d=json.load(open('gpurun_out/bench_f64.json'));print('f64',d['value'],d['e2e']['value'],d['breakdown_ms'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fp32-pairs > gpurun_out/ncu_b32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pair_scan|rescore|entq" -s 6 -c 3 -o gpurun_out/r1c_f32 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --scale 0.3 --fp32-pairs > gpurun_out/ncu_full32.log 2>&1
ls -la gpurun_out
