"""summary of an `ncu --set full` report for profiles/: usage ncu_summary.py in.ncu-rep out.csv [kernel workload traffic.json]
with the last three arguments also writes the measured DRAM bytes of that launch as profiles/traffic_<kernel>_<workload>.json,
the file bench.py reads for `roofline.traffic`
(first kernel of the report; the metrics DESIGN.md quotes: time, DRAM bytes, occupancy, issue / pipe utilisation,
shared-memory conflicts, stall reasons per issued instruction)"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|"
                  r"launch__(registers_per_thread|shared_mem_per_block_dynamic|occupancy_limit_\w+|waves_per_multiprocessor)|"
                  r"lts__t_sector_hit_rate\.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|"
                  r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.(per_cycle_active|pct_of_peak_sustained_active)|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                  r"sm__pipe_(fma|fp64|alu|xu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, first = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in sorted(zip(hdr, units, first)):
        if KEEP.match(h):
            w.writerow([h, u, v])
if len(sys.argv) > 5:
    import json
    d = dict(zip(hdr, first))
    u = dict(zip(hdr, units))

    def nbytes(k):
        v = float(d[k].replace(",", ""))
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
    tot = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
    json.dump({"kernel": sys.argv[3], "workload": sys.argv[4], "dram_bytes_per_launch": tot,
               "dram_bytes_read": nbytes("dram__bytes_read.sum"), "dram_bytes_write": nbytes("dram__bytes_write.sum"),
               "l2_hit_rate_pct": float(d.get("lts__t_sector_hit_rate.pct", "nan")),
               "gpu_time_ns": d.get("gpu__time_duration.sum"),
               "source": "ncu --set full --clock-control none, one launch of %s (%s) under bench.py; tools/run_profile.sh" %
                         (sys.argv[3], sys.argv[1])}, open(sys.argv[5], "w"), indent=1)
