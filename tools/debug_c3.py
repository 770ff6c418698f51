"""debug: C3 subset, fp32 scan vs oracle: which pairs differ and by how much"""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from demuxlet_b200.api import Demuxlet
from demuxlet_b200.synth import make_config
from oracle import pyoracle as po

p, alphas = make_config("C3", scale=0.025)
orc = po.Oracle(p, alphas)
ref = orc.run(nthreads=os.cpu_count() or 1, want_pass1=False)["best"]
res = {}
for mode, kw in (("f32", dict(fp32_pairs=True)), ("f64", dict(fp32_pairs=False)), ("f32host", dict(fp32_pairs=True, host_finalise=True))):
    d = Demuxlet(p.nv, alphas, **kw)
    d.set_pileup(p)
    d.run()
    res[mode] = d.best().copy()
    st = d.stats()
    print(mode, "rescored", st["n_rescored"], "gpu", st["n_rescored_gpu"], "fullcell", st["n_fullcell"], "gpu_exact", st["gpu_exact"])
    if mode == "f32":
        dd = d
b = res["f32"]
bad = np.nonzero((ref["j_best"] != b["dbl1"]) | (ref["k_best"] != b["dbl2"]))[0]
print("bad cells", bad.tolist())
for c in bad[:8]:
    AB = dd.pair(int(c), int(c) + 1)[0]     # dense fp64 block
    oj, ok = int(ref["j_best"][c]), int(ref["k_best"][c])
    gj, gk = int(b["dbl1"][c]), int(b["dbl2"][c])
    flat = AB[:, :, 1].copy()
    np.fill_diagonal(flat, -np.inf)
    top = np.argsort(flat.ravel())[::-1][:6]
    print("cell", c, "entries", int(p.cell_ptr[c + 1] - p.cell_ptr[c]), "oracle", (oj, ok), "llk12 %.6f" % ref["pair_llk12"][c],
          "gpu", (gj, gk), "llk12 %.6f" % b["llk12"][c], "nresc", int(b["n_rescored"][c]), "mask", hex(int(b["exact_mask"][c])),
          "f64mode", (int(res["f64"]["dbl1"][c]), int(res["f64"]["dbl2"][c])), "f32host", (int(res["f32host"]["dbl1"][c]), int(res["f32host"]["dbl2"][c])))
    print("   dense fp64: oracle pair %.6f gpu pair %.6f; top:" % (AB[oj, ok, 1], AB[gj, gk, 1]),
          [(int(t // p.nv), int(t % p.nv), round(float(flat.ravel()[t]), 5)) for t in top])
    print("   exact rescore oracle pair %.6f gpu pair %.6f" % (dd.rescore(int(c), oj, ok, 1), dd.rescore(int(c), gj, gk, 1)))
