# wait-cycle profile of the fp32 scan roles (tools/build_variant.sh prof "-DDMX_PROF=1"): usage run_prof.sh TAG SCALE DBG...
cd $GRAFT_REPO_ROOT
T=${1:-prof}; S=${2:-0.25}; shift; shift
mkdir -p gpurun_out
for L in tools/libdmx_prof*.so; do
echo "== $L" >> gpurun_out/${T}_prof.txt
DMX_LIB=$PWD/$L python tools/scan_probe.py $S "$@" >> gpurun_out/${T}_prof.txt 2>&1
done
tail -120 gpurun_out/${T}_prof.txt
