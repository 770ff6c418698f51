cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ./tools/tf32_bench > gpurun_out/r1d_tf32_bench.jsonl 2>&1
cat gpurun_out/r1d_tf32_bench.jsonl
