"""dmx_write_outputs (.single / .sing2 / .best / .pair with the reference's printf formats,
cmd_cram_demuxlet.cpp:470, :506, :760-874) on the C3 pileup: rows formatted by all host threads against one thread
(one thread = the reference's own way: one hprintf per row), byte-identical files.
usage: python tools/write_bench.py [SCALE=1.0] [PAIR_SCALE=0.1]    prints one JSON line"""
import filecmp
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demuxlet_b200 import api                      # noqa: E402
from demuxlet_b200.synth import make_config        # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
pscale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
res = {"workload": "C3 x %g (.single/.sing2/.best), C3 x %g (--write-pair)" % (scale, pscale), "cores": os.cpu_count()}
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
for tag, sc, pair in (("tsv", scale, False), ("pair", pscale, True)):
    plp, alphas = make_config("C3", scale=sc)
    out = {}
    for nth in (1, 0):
        d = api.Demuxlet(plp.nv, alphas, fp32_pairs=True, host_threads=nth)
        d.set_pileup(plp)
        d.run()
        pre = os.path.join(tmp, "%s_%d" % (tag, nth))
        t0 = time.perf_counter()
        d.write_outputs(pre, write_pair=pair)
        dt = time.perf_counter() - t0
        d.close()
        exts = [".single", ".sing2", ".best"] + ([".pair"] if pair else [])
        nbytes = sum(os.path.getsize(pre + e) for e in exts)
        out[nth] = (dt, nbytes, pre, exts)
    (t1, nb, p1, exts), (tn, _, pn, _) = out[1], out[0]
    same = all(filecmp.cmp(p1 + e, pn + e, shallow=False) for e in exts)
    res[tag] = {"cells": plp.ncells, "bytes": nb, "one_thread_s": round(t1, 3), "all_threads_s": round(tn, 3),
                "mb_per_s_one": round(nb / t1 / 1e6, 1), "mb_per_s_all": round(nb / tn / 1e6, 1),
                "speedup": round(t1 / tn, 2), "byte_identical": same}
    for _, _, pre, ex in out.values():
        for e in ex:
            os.unlink(pre + e)
print(json.dumps(res))
