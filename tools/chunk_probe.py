"""per-chunk overhead of the run path: dmx_compute (inputs resident) with the barcodes in K chunks
usage: DMX_CHUNKS=K python tools/chunk_probe.py [scale]"""
import os, sys, time, json
sys.path.insert(0, ".")
import numpy as np
from demuxlet_b200.api import Demuxlet
from bench import make_workload
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
plp, alphas = make_workload("C3", 0, scale)
d = Demuxlet(plp.nv, alphas, fp32_pairs=True, trust_inputs=True, host_threads=4)   # < 8 threads: resident runs are chunked too
d.set_pileup(plp)
d.upload()
ts = []
for _ in range(6):
    t0 = time.perf_counter(); d.compute(); ts.append((time.perf_counter() - t0) * 1e3)
st = d.stats()
print(json.dumps({"K": os.environ.get("DMX_CHUNKS"), "compute_ms": round(float(np.median(ts[2:])), 3), "prep": round(st["prep_ms"], 3),
                  "pair": round(st["pair_kernel_ms"], 3), "kernels": round(st["kernel_ms"], 3), "launches": st["gpu_launches"]}))
