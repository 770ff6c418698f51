# round-2 closing check of the host-side changes (BAM / VCF readers): GPU tests, smoke, one bench line
cd $GRAFT_REPO_ROOT
T=${1:-r5}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${T}_pytest.txt 2>&1
tail -4 gpurun_out/${T}_pytest.txt
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.txt
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
