"""writes tests/golden/pack1.rds (a read list) and pack1.csr: the pileup the REFERENCE's own add_read
(sc_drop_seq.cpp:34-77, through oracle/_ref/ref_pack) builds from it.  Run in the build container:
    python tests/golden/make_pack_golden.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import packio  # noqa: E402

cell, snp, umis, allele, bq = packio.make_reads(6000, 40, 300, seed=20261017)
rds = os.path.join(HERE, "pack1.rds")
packio.write_rds(rds, cell, snp, umis, allele, bq, 40, 300)
print(packio.run_ref_pack(rds, os.path.join(HERE, "pack1.csr")))
