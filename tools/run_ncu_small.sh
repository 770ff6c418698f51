# ncu --set full of the per-entry / per-cell kernels around the scan (one launch each), C3 at full size
cd $GRAFT_REPO_ROOT
T=${1:-ns}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"classify|entq_aff|entq1|cell_scalars|rescore_terms|rescore_sum|merge_kernel|finalise_kernel|dtab|gprow" -s 30 -c 14 -o gpurun_out/${T}_small -f python tools/scan_probe.py 1.0 0 > gpurun_out/${T}_small.log 2>&1
tail -2 gpurun_out/${T}_small.log
ncu -i gpurun_out/${T}_small.ncu-rep --page raw --csv > gpurun_out/${T}_small_raw.csv 2>/dev/null
rm -f gpurun_out/${T}_small.ncu-rep
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${T}_small_raw.csv")))
h=rows[0]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","lts__t_sector_hit_rate.pct","smsp__inst_executed.sum"]
idx=[h.index(w) for w in want if w in h]
print([h[i] for i in idx])
for r in rows[2:]:
    print([r[i][:28] for i in idx])
PY
