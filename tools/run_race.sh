# racecheck of the fp32 scan alone (second generation), product build and the build in which every lane arrives
cd $GRAFT_REPO_ROOT
T=${1:-race}
mkdir -p gpurun_out
for L in demuxlet_b200/libdmx.so tools/libdmx_arrall.so; do
echo "== $L: compute-sanitizer --tool racecheck python tools/sanitize_scan.py fp32" >> gpurun_out/${T}_race_fp32.txt
DMX_LIB=$PWD/$L timeout 900 compute-sanitizer --tool racecheck --print-limit 200 python tools/sanitize_scan.py fp32 > gpurun_out/race_full.txt 2>&1
grep -E "Error: Race|RACECHECK SUMMARY|checksum" gpurun_out/race_full.txt | cut -c1-220 | sort | uniq -c | sort -rn | head -30 >> gpurun_out/${T}_race_fp32.txt
grep -oE "(and Write|and Read|Read|Write) access at [^+]*" gpurun_out/race_full.txt | cut -c1-160 | sort | uniq -c | sort -rn | head -20 >> gpurun_out/${T}_race_fp32.txt
done
cat gpurun_out/${T}_race_fp32.txt
