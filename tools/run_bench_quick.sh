# C3 bench line (no cpu baseline) + first-generation scan for comparison
cd $GRAFT_REPO_ROOT
T=${1:-bq}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value %.0f e2e %.0f ms/step %.2f e2e ms %.2f pair %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
print(d["breakdown_ms"])
PY
