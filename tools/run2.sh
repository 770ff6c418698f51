cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "records_match" 2>&1 | tail -15
timeout 300 python bench.py --steps 5 --warmup 3 --fp32-pairs --no-cpu-baseline > gpurun_out/bench_f32a.json 2> gpurun_out/bench_f32a.err
tail -c 1500 gpurun_out/bench_f32a.json; tail -5 gpurun_out/bench_f32a.err
