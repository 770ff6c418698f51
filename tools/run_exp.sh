cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_bam.py tests/test_vcf.py -m gpu -x -q 2>&1 | tail -12
