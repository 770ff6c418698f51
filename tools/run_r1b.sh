# round-1 state capture: GPU tests, bench (f64 / f32 pair scan), reference arm, ncu launch list + full capture
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; lscpu | grep -E 'Model name|Socket|Core|Thread' >> gpurun_out/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
tail -c 1500 gpurun_out/bench_f64.json
timeout 600 python bench.py --steps 10 --warmup 3 --fp32-pairs --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
tail -c 1500 gpurun_out/bench_f32.json; tail -n 3 gpurun_out/bench_f32.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f64.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --fp32-pairs > gpurun_out/ncu_b32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan -s 2 -c 1 -o gpurun_out/pair_scan_f64 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --scale 0.3 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan -s 2 -c 1 -o gpurun_out/pair_scan_f32 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --scale 0.3 --fp32-pairs > gpurun_out/ncu_full32.log 2>&1
ls -la gpurun_out
