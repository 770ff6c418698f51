# quick GPU check: parity tests (optionally a -k filter), then one short C3 bench line
cd $GRAFT_REPO_ROOT
T=${1:-quick}
K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q -k "$K" 2>&1 | tail -15
else
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench.json"))
print("value %.0f e2e %.0f ms/step %.2f pair %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
print(d["breakdown_ms"])
PY
