cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
cat > /tmp/small.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from demuxlet_b200 import api
from demuxlet_b200.synth import make_synth
for nv, alphas, fp32 in ((64, (0.0, 0.5), True), (70, (0.0, 0.3, 0.5), True), (12, (0.0, 0.5), False)):
    p = make_synth(40, nv, 3000, "GP", seed=3)
    d = api.Demuxlet(p.nv, alphas, fp32_pairs=fp32)
    d.set_pileup(p); d.run(); d.upload(); d.compute()
    print(nv, fp32, d.best()["sng1"][:5], d.stats()["gpu_launches"])
    d.close()
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/small.py 2>&1 | tail -12
