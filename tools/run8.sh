cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/exp.py 0.3 0,9,8
python tools/exp.py 0.3 0 f64
