# ncu --set full of the fp32 scan on the C3 workload at SCALE (tools/scan_probe.py): usage run_ncu_probe.sh TAG SCALE
cd $GRAFT_REPO_ROOT
T=${1:-np}; S=${2:-0.25}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_scan_v2 -s 2 -c 1 -o gpurun_out/${T}_scan -f python tools/scan_probe.py $S 0 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
python tools/ncu_summary.py gpurun_out/${T}_scan.ncu-rep gpurun_out/${T}_scan_ncu.csv
ncu -i gpurun_out/${T}_scan.ncu-rep --page source --csv > gpurun_out/${T}_scan_src.csv 2>/dev/null
rm -f gpurun_out/${T}_scan.ncu-rep
cat gpurun_out/${T}_scan_ncu.csv
