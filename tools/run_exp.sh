cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_vcf.py tests/test_field.py -m gpu -x -q 2>&1 | tail -8
