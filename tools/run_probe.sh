cd $GRAFT_REPO_ROOT
T=${1:-probe}
shift
mkdir -p gpurun_out
python tools/scan_probe.py 1.0 "$@" 2>&1 | tail -12 | tee gpurun_out/${T}_probe.txt
