cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1d_bench_n2.json 2> gpurun_out/r1d_bench_n2.err
echo rc=$?
tail -c 600 gpurun_out/r1d_bench_n2.json; tail -3 gpurun_out/r1d_bench_n2.err
