# pair-scan profile: debug-switch timings, then one ncu --set full capture with source counters
cd $GRAFT_REPO_ROOT
T=${1:-prof}
mkdir -p gpurun_out
python tools/scan_probe.py 1.0 0 1 2 3 4 2>&1 | tail -8 | tee gpurun_out/${T}_probe.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan_f32 -s 3 -c 1 -o gpurun_out/${T}_pair_scan_f32 -f python tools/scan_probe.py 1.0 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/${T}_pair_scan_f32.ncu-rep
