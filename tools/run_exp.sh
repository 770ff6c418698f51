cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python tools/write_bench.py > gpurun_out/r1d_write_bench.json 2> gpurun_out/r1d_write_bench.err
cat gpurun_out/r1d_write_bench.json; tail -3 gpurun_out/r1d_write_bench.err
