cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for KF in "$@"; do
K=${KF%%:*}; F=${KF##*:}
DMX_CHUNKS=$K DMX_CHUNK_F0=$F timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ck_bench.json 2> gpurun_out/ck_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/ck_bench.json"))
print("K=$K F0=$F value %.0f e2e %.0f ms/step %.2f e2e ms %.2f run_ms %.2f h2d %.2f kern %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["run_ms"], d["e2e"]["h2d_ms"], d["e2e"]["kernel_ms"]))
PY
done
