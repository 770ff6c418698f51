# profiles for profiles/: smoke, full bench line (cpu baseline), reference arm, launch lists, ncu --set full of the scan,
# the VCF-field kernels and the writers
cd $GRAFT_REPO_ROOT
T=${1:-r1b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 400 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --fp64-pairs --no-cpu-baseline > gpurun_out/${T}_bench_f64.json 2>> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan_f32 -s 3 -c 1 -o gpurun_out/${T}_pair_scan_f32 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 300 python tools/field_bench.py > gpurun_out/${T}_field_bench.json 2> gpurun_out/${T}_field_bench.err
cat gpurun_out/${T}_field_bench.json; tail -3 gpurun_out/${T}_field_bench.err
for K in gt gp pl; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:field_${K} -s 5 -c 1 -o gpurun_out/${T}_field_${K} -f python tools/field_bench.py > gpurun_out/ncu_field.log 2>&1
tail -1 gpurun_out/ncu_field.log
done
timeout 600 python tools/write_bench.py > gpurun_out/${T}_write_bench.json 2> gpurun_out/${T}_write_bench.err
cat gpurun_out/${T}_write_bench.json
ls -la gpurun_out | tail -12
cat > /tmp/small_field.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from demuxlet_b200 import api
from tests.test_field import make_field
for field in ("GT", "GP", "PL"):
    for nv in (64, 7, 131):
        a = make_field(field, 700, nv, 5)
        g, af = api.build_gp(field, a, 0.01, want_af=True)
        print(field, nv, float(g.sum()))
print(api.site_filter(make_field("GT", 700, 64, 5)).sum())
PY
for TOOL in memcheck racecheck; do
echo "== compute-sanitizer --tool $TOOL (field kernels)" >> gpurun_out/${T}_field_sanitizer.txt
timeout 600 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/small_field.py 2>&1 | tail -14 >> gpurun_out/${T}_field_sanitizer.txt
done
cat gpurun_out/${T}_field_sanitizer.txt
# microbenchmarks behind the round-2 plan
for B in tf32_bench tmem_ld_bench umma_bench umma_pipe_bench; do
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -diag-suppress 128 -o tools/$B tools/$B.cu && timeout 120 ./tools/$B > gpurun_out/${T}_$B.jsonl 2>&1
tail -3 gpurun_out/${T}_$B.jsonl
done
