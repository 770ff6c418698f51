cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_f32 -s 3 -c 1 -o gpurun_out/prof_f32e python bench.py --steps 1 --warmup 3 --fp32-pairs --no-cpu-baseline --scale 0.3 > gpurun_out/ncu_f32e.log 2>&1
tail -3 gpurun_out/ncu_f32e.log | cut -c1-300
