# quick GPU check of a kernel change: smoke, the parity tests, a timing probe (C3 at SCALE), the wait profile
cd $GRAFT_REPO_ROOT
T=${1:-q}; S=${2:-0.25}; TESTS=${3:-tests/test_gpu_parity.py tests/test_gpu_configs.py}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest $TESTS -x -q 2>&1 | tail -15
timeout 300 python tools/scan_probe.py $S 0 1 2>&1 | tail -2
DMX_SCAN_GEN=1 timeout 300 python tools/scan_probe.py $S 0 2>&1 | tail -1
if [ -f tools/libdmx_prof.so ]; then DMX_LIB=$PWD/tools/libdmx_prof.so timeout 300 python tools/scan_probe.py $S 0 2>&1 | tail -5; fi
