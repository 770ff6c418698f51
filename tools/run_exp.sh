cd $GRAFT_REPO_ROOT
python tools/exp.py 0.3 0,1 2>&1 | grep -v "^$" | tail -2
DMX_LIB=$GRAFT_REPO_ROOT/tools/libdmx_tw.so python tools/exp.py 0.3 0,1 2>&1 | grep -v "^$" | tail -2
