"""end-to-end step time of dmx_run (host buffers in, records out) at the headline shape for several chunk schedules
(DMX_CHUNKS / DMX_CHUNK_F0 are read when the pileup is set).  usage: python tools/e2e_probe.py [scale]   one JSON line each"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demuxlet_b200 import api                    # noqa: E402
from bench import make_workload                  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
plp, alphas = make_workload("C3", 0, scale)


def pin(a):
    out = api.pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


gp = pin(plp.gp)
arrays = [pin(x) for x in (plp.cell_ptr, plp.ent_snp, plp.ent_ptr, plp.reads, plp.totl, plp.npass, plp.uniq)]
configs = [(None, None), ("2", "0.10"), ("2", "0.18"), ("2", "0.24"), ("3", None), ("4", None), ("1", None)]
ref = None
for rep in range(2):
    for K, F in configs:
        for k, v in (("DMX_CHUNKS", K), ("DMX_CHUNK_F0", F)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        # a fresh context: the schedule of a context is kept while the same droplets come again
        d = api.Demuxlet(plp.nv, alphas, device=0, borrow=True, trust_inputs=True, fp32_pairs=True, host_threads=os.cpu_count() or 1)
        ts = []
        for i in range(13):
            t0 = time.perf_counter()
            d._chk(d.lib.dmx_set_gp(d._h, gp.ctypes.data, gp.shape[0]), "dmx_set_gp")
            d.set_pileup_arrays(*arrays)
            d.run()
            ts.append((time.perf_counter() - t0) * 1e3)
        best = d.best()
        if ref is None:
            ref = best.copy()
        st = d.stats()
        print(json.dumps({"chunks": K, "f0": F, "rep": rep, "e2e_ms_median": round(float(np.median(ts[3:])), 3),
                          "e2e_ms_min": round(float(np.min(ts[3:])), 3), "kernel_ms": round(st["kernel_ms"], 3),
                          "launches": st["gpu_launches"], "same_records": bool(best.tobytes() == ref.tobytes())}), flush=True)
        d.close()
