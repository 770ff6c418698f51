"""Writes tests/golden/cigar_lookup.npz: what the REFERENCE's own bam_get_base_and_qual_and_read_and_qual
(/root/reference/hts_utils.cpp:279-359, run through oracle/_ref/ref_cigar) returns for the reads and query positions of
tests/bamio.py:cigar_cases(seed 31).  Run in the build container: python tests/golden/make_cigar_golden.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import bamio  # noqa: E402

reads, queries = bamio.cigar_cases(np.random.default_rng(31))
with tempfile.TemporaryDirectory() as td:
    res = bamio.run_ref_cigar(reads, queries, td)
flat = np.array([(ord(b), q, rp) for row in res for (b, q, rp) in row], np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cigar_lookup.npz"), results=flat)
print(len(reads), "reads,", len(flat), "queries")
