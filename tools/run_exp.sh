cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 ./tools/umma_bench > gpurun_out/r1d_umma_bench.jsonl 2>&1
echo rc=$?
cat gpurun_out/r1d_umma_bench.jsonl
nvidia-smi --query-gpu=name,temperature.gpu --format=csv,noheader
