cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_field.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/field_bench.py > gpurun_out/r1d_field_bench.json 2> gpurun_out/r1d_field_bench.err
cat gpurun_out/r1d_field_bench.json; tail -3 gpurun_out/r1d_field_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:field_ -s 9 -c 3 -o gpurun_out/r1d_field -f python tools/field_bench.py 100000 > gpurun_out/ncu_field.log 2>&1
tail -2 gpurun_out/ncu_field.log
