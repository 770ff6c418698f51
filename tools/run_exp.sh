cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_field.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/field_bench.py > gpurun_out/r1d_field_bench.json 2> gpurun_out/r1d_field_bench.err
cat gpurun_out/r1d_field_bench.json; tail -3 gpurun_out/r1d_field_bench.err
