# round 2, first capture: the new config tests and one bench line per BASELINE.json config (before kernel changes)
cd $GRAFT_REPO_ROOT
T=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -25 > gpurun_out/${T}_pytest.txt; tail -25 gpurun_out/${T}_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_C2.json 2> gpurun_out/${T}_bench_C2.err; tail -c 300 gpurun_out/${T}_bench_C2.json; tail -2 gpurun_out/${T}_bench_C2.err
timeout 900 python bench.py --workload C4 --write-pair --steps 3 --warmup 3 > gpurun_out/${T}_bench_C4.json 2> gpurun_out/${T}_bench_C4.err; tail -c 300 gpurun_out/${T}_bench_C4.json; tail -2 gpurun_out/${T}_bench_C4.err
timeout 900 python bench.py --workload C5 --scale 0.0005 --steps 2 --warmup 3 > gpurun_out/${T}_bench_C5.json 2> gpurun_out/${T}_bench_C5.err; tail -c 300 gpurun_out/${T}_bench_C5.json; tail -2 gpurun_out/${T}_bench_C5.err
for TOOL in racecheck synccheck; do
echo "== compute-sanitizer --tool $TOOL python tools/sanitize_scan.py" >> gpurun_out/${T}_scan_sanitizer.txt
timeout 900 compute-sanitizer --tool $TOOL --print-limit 8 python tools/sanitize_scan.py 2>&1 | tail -30 >> gpurun_out/${T}_scan_sanitizer.txt
done
tail -40 gpurun_out/${T}_scan_sanitizer.txt
