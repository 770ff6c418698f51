cd $GRAFT_REPO_ROOT
for T in 16 4 1; do DMX_FIN_DEBUG=1 python - <<PY 2>&1 | tail -3
import sys; sys.path.insert(0,'.')
from demuxlet_b200.api import Demuxlet
from bench import make_workload
plp, alphas = make_workload("C3", 0, 0.5)
d = Demuxlet(plp.nv, alphas, fp32_pairs=True, trust_inputs=True, host_threads=$T)
d.set_pileup(plp); d.upload()
for _ in range(3): d.compute()
print($T, d.stats()["finalise_ms"])
PY
done
