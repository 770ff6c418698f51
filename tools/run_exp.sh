cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python tools/exp.py 0.3 0,1 2>&1 | grep -v "^$" | tail -2
