# timing probe of every experiment build tools/libdmx_*.so (and the product library)
cd $GRAFT_REPO_ROOT
T=${1:-var}
shift
mkdir -p gpurun_out
: > gpurun_out/${T}_variants.txt
for L in demuxlet_b200/libdmx.so tools/libdmx_*.so; do
  echo "== $L" | tee -a gpurun_out/${T}_variants.txt
  DMX_LIB=$PWD/$L python tools/scan_probe.py 1.0 "$@" 2>&1 | tail -4 | tee -a gpurun_out/${T}_variants.txt
done
