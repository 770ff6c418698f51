"""timing experiments on the pair scan: one workload, several DMX_DEBUG settings in one process
usage: python tools/exp.py SCALE DBG[,DBG...] [f64]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from demuxlet_b200 import api
from demuxlet_b200.synth import make_config

scale = float(sys.argv[1])
dbgs = [int(x) for x in sys.argv[2].split(",")]
fp32 = not (len(sys.argv) > 3 and sys.argv[3] == "f64")
plp, alphas = make_config("C3", scale=scale)
d = api.Demuxlet(plp.nv, alphas, fp32_pairs=fp32, trust_inputs=True)
d.set_pileup(plp)
d.upload()
for dbg in dbgs:
    os.environ["DMX_DEBUG"] = str(dbg)
    os.environ["DMX_TRACE"] = "gpurun_out/trace_%d.txt" % dbg
    ts = []
    for i in range(4):
        d.compute()
        ts.append(d.stats()["pair_kernel_ms"])
    st = d.stats()
    print("dbg %d pair_scan ms %s  (entries %d, per 10k cells: %.2f ms) prep %.2f kernels %.2f" % (
        dbg, ["%.3f" % t for t in ts], st["n_entries"], min(ts[1:]) / scale, st["prep_ms"], st["kernel_ms"]), flush=True)
