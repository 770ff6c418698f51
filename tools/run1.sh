set -x
cd $GRAFT_REPO_ROOT
./tools/f32_bench > gpurun_out/f32_bench.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
tail -c 600 gpurun_out/bench_r1e.json
