cd $GRAFT_REPO_ROOT
T=${1:-ncu}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan_f32 -s 3 -c 1 -o gpurun_out/${T}_pair_scan_f32 -f python tools/scan_probe.py 1.0 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
