"""timing probe of the fp32 pair scan at the C3 shape: pair-scan CUDA-event time under the DMX_DEBUG switches
(1: consumers skip their arithmetic, 2: transform warps skip theirs, 4: no per-sample extras, 16: no pairs with sample 0)
usage: scan_probe.py [scale] [dbg values ...]"""
import os, sys, json
sys.path.insert(0, ".")
import numpy as np
from demuxlet_b200.api import Demuxlet
from bench import make_workload

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
dbgs = [int(x) for x in sys.argv[2:]] or [0]
plp, alphas = make_workload("C3", 0, scale)
d = Demuxlet(plp.nv, alphas, fp32_pairs=True, trust_inputs=True)
d.set_pileup(plp)
d.upload()
for dbg in dbgs:
    os.environ["DMX_DEBUG"] = str(dbg)
    ts = []
    for _ in range(6):
        d.compute()
        ts.append(d.stats()["pair_kernel_ms"])
    st = d.stats()
    print(json.dumps({"dbg": dbg, "pair_ms": round(float(np.median(ts[2:])), 3), "prep_ms": round(st["prep_ms"], 3),
                      "kernel_ms": round(st["kernel_ms"], 3)}))
