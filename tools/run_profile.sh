# profiles for profiles/: GPU tests, smoke, bench lines (C3 default + fp64, C2, C4 --write-pair, C5 shape, reference arm),
# ncu launch list, ncu --set full of the pair scan (+ the measured DRAM traffic json bench.py reads), sanitizers
# usage (under gpurun): bash tools/run_profile.sh TAG [parts]   parts = letters: t tests, b bench, c configs, n ncu,
# s sanitizers, f field/writer benches, m microbenchmarks (default tbcns)
cd $GRAFT_REPO_ROOT
T=${1:-r2}
P=${2:-tbcns}
mkdir -p gpurun_out
has() { case "$P" in *$1*) return 0;; *) return 1;; esac; }
if has t; then
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${T}_pytest.txt 2>&1
tail -4 gpurun_out/${T}_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.txt
fi
if has b; then
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -c 600 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --fp64-pairs --no-cpu-baseline > gpurun_out/${T}_bench_f64.json 2>> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_ref.json
fi
if has n; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_scan_v2 -s 3 -c 1 -o gpurun_out/${T}_pair_scan_v2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python tools/ncu_summary.py gpurun_out/${T}_pair_scan_v2.ncu-rep gpurun_out/${T}_pair_scan_v2_ncu.csv pair_scan_v2_kernel C3 gpurun_out/traffic_pair_scan_v2_kernel_C3.json
cat gpurun_out/traffic_pair_scan_v2_kernel_C3.json
timeout 900 ncu --set full --clock-control none -k regex:"pair_scan_kernel" -s 1 -c 1 -o gpurun_out/${T}_pair_scan_f64 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --fp64-pairs > gpurun_out/ncu_full64.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_pair_scan_f64.ncu-rep gpurun_out/${T}_pair_scan_f64_ncu.csv pair_scan_kernel C3 gpurun_out/traffic_pair_scan_kernel_C3.json
rm -f gpurun_out/${T}_pair_scan_f64.ncu-rep
fi
if has c; then
timeout 600 python bench.py --workload C2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_C2.json 2> gpurun_out/${T}_bench_C2.err
tail -c 300 gpurun_out/${T}_bench_C2.json
timeout 1200 python bench.py --workload C4 --write-pair --steps 3 --warmup 3 > gpurun_out/${T}_bench_C4.json 2> gpurun_out/${T}_bench_C4.err
tail -c 300 gpurun_out/${T}_bench_C4.json
timeout 1200 python bench.py --workload C5 --scale 0.0005 --steps 2 --warmup 3 > gpurun_out/${T}_bench_C5.json 2> gpurun_out/${T}_bench_C5.err
tail -c 300 gpurun_out/${T}_bench_C5.json
fi
if has s; then
# racecheck also on the build in which every lane arrives on the ring barriers (tools/build_variant.sh arrall -DDMX_ARRIVE_ALL=1)
if [ -f tools/libdmx_arrall.so ]; then
echo "== DMX_ARRIVE_ALL build: compute-sanitizer --tool racecheck python tools/sanitize_scan.py" >> gpurun_out/${T}_scan_sanitizer.txt
DMX_LIB=$PWD/tools/libdmx_arrall.so timeout 900 compute-sanitizer --tool racecheck --print-limit 8 python tools/sanitize_scan.py 2>&1 | cut -c1-400 | tail -30 >> gpurun_out/${T}_scan_sanitizer.txt
fi
for TOOL in racecheck synccheck memcheck; do
echo "== compute-sanitizer --tool $TOOL python tools/sanitize_scan.py" >> gpurun_out/${T}_scan_sanitizer.txt
timeout 900 compute-sanitizer --tool $TOOL --print-limit 8 python tools/sanitize_scan.py 2>&1 | cut -c1-400 | tail -40 >> gpurun_out/${T}_scan_sanitizer.txt
done
tail -12 gpurun_out/${T}_scan_sanitizer.txt
fi
if has f; then
timeout 300 python tools/field_bench.py > gpurun_out/${T}_field_bench.json 2> gpurun_out/${T}_field_bench.err
cat gpurun_out/${T}_field_bench.json; tail -3 gpurun_out/${T}_field_bench.err
for K in gt gp pl; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:field_${K} -s 5 -c 1 -o gpurun_out/${T}_field_${K} -f python tools/field_bench.py > gpurun_out/ncu_field.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_field_${K}.ncu-rep gpurun_out/${T}_field_${K}_ncu.csv
done
timeout 600 python tools/write_bench.py > gpurun_out/${T}_write_bench.json 2> gpurun_out/${T}_write_bench.err
cat gpurun_out/${T}_write_bench.json
timeout 600 python tools/io_bench.py > gpurun_out/${T}_io_bench.json 2> gpurun_out/${T}_io_bench.err
cat gpurun_out/${T}_io_bench.json
fi
if has m; then
for B in tf32_bench tmem_ld_bench umma_bench umma_pipe_bench; do
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -diag-suppress 128 -o tools/$B tools/$B.cu && timeout 120 ./tools/$B > gpurun_out/${T}_$B.jsonl 2>&1
tail -3 gpurun_out/${T}_$B.jsonl
done
fi
ls -la gpurun_out | tail -5
