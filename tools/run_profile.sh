# profiles for profiles/: smoke, full bench line (cpu baseline), reference arm, launch lists, ncu --set full of the scan
cd $GRAFT_REPO_ROOT
T=${1:-r1b}
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --fp64-pairs --no-cpu-baseline > gpurun_out/${T}_bench_f64.json 2>> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan_f32 -s 3 -c 1 -o gpurun_out/${T}_pair_scan_f32 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -12
