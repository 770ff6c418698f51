"""small pair-scan runs for compute-sanitizer (racecheck / synccheck / memcheck): both scan kernels at nv = 64
(triangular tasks only) and at nv = 70 with a three-point alpha grid (rectangular tasks, a full plane)"""
import sys

sys.path.insert(0, ".")
from demuxlet_b200.api import Demuxlet          # noqa: E402
from demuxlet_b200.synth import make_synth      # noqa: E402

only = sys.argv[1] if len(sys.argv) > 1 else "both"      # "fp32", "fp64" or both scans
for nv, alphas, ncells in ((64, (0.0, 0.5), 12), (70, (0.0, 0.3, 0.5), 6)):
    p = make_synth(ncells, nv, 1500, "GP", seed=nv)
    for fp32 in (True, False):
        if (only == "fp32" and not fp32) or (only == "fp64" and fp32):
            continue
        d = Demuxlet(nv, alphas, fp32_pairs=fp32)
        d.set_pileup(p)
        d.run()
        b = d.best()
        print("nv %d alphas %s %s: %d cells, %d launches, checksum %d" %
              (nv, alphas, "fp32" if fp32 else "fp64", p.ncells, d.stats()["gpu_launches"],
               int(b["dbl1"].sum() + b["dbl2"].sum() + b["sng1"].sum())))
        d.close()
