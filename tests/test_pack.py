"""CPU tests of dmx_pack_reads: the sort-based packer against (1) a golden pileup the REFERENCE's own
add_read built (tests/golden/pack1.*, make_pack_golden.py), (2) the live reference container where oracle/_ref/ref_pack
is built, (3) the python restatement used by the other tests; bit-exact in every array."""
import os

import numpy as np
import pytest

from demuxlet_b200 import api
from demuxlet_b200.plp import build_from_reads
from tests import packio

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("cell_ptr", "ent_snp", "ent_ptr", "reads", "npass", "uniq")


def same(a, b):
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f


def test_against_reference_golden_pileup():
    cell, snp, umis, allele, bq, ncells, nsnps = packio.read_rds(os.path.join(GOLD, "pack1.rds"))
    want = packio.read_csr(os.path.join(GOLD, "pack1.csr"))
    for threads in (1, 4):
        same(api.pack_reads(cell, snp, umis, allele, bq, ncells, threads=threads), want)
    assert want["npass"].sum() == len(cell) and want["uniq"].sum() == len(want["reads"])
    assert (want["npass"][-2:] == 0).all()          # empty cells stay empty rows


@pytest.mark.skipif(not os.path.exists(packio.REF_PACK), reason="oracle/_ref/ref_pack not built")
@pytest.mark.parametrize("n,ncells,nsnps,seed", [(200000, 500, 20000, 1), (50000, 3, 40, 2), (1000, 1, 1, 3)])
def test_against_live_reference_container(n, ncells, nsnps, seed, tmp_path):
    cell, snp, umis, allele, bq = packio.make_reads(n, ncells, nsnps, seed)
    rds, csr = str(tmp_path / "a.rds"), str(tmp_path / "a.csr")
    packio.write_rds(rds, cell, snp, umis, allele, bq, ncells, nsnps)
    packio.run_ref_pack(rds, csr)
    same(api.pack_reads(cell, snp, umis, allele, bq, ncells), packio.read_csr(csr))


def test_against_python_restatement_and_edge_cases():
    cell, snp, umis, allele, bq = packio.make_reads(3000, 25, 60, seed=5)
    got = api.pack_reads(cell, snp, umis, allele, bq, 25)
    gp = np.full((60, 2, 3), 1 / 3, dtype=np.float32)
    ref = build_from_reads(gp, cell, snp, umis, allele, bq, ["C%d" % i for i in range(25)], ["S0", "S1"])
    for f in ("cell_ptr", "ent_snp", "ent_ptr", "reads"):
        assert np.array_equal(got[f], getattr(ref, f)), f
    assert np.array_equal(got["npass"], ref.npass) and np.array_equal(got["uniq"], ref.uniq)
    # no reads at all
    e = api.pack_reads(np.zeros(0, np.int32), np.zeros(0, np.int32), [], np.zeros(0, np.uint8), np.zeros(0, np.uint8), 4)
    assert e["cell_ptr"].tolist() == [0, 0, 0, 0, 0] and e["ent_ptr"].tolist() == [0] and e["reads"].size == 0
    # a cell id outside [0, ncells) is refused
    with pytest.raises(api.DmxError) as ex:
        api.pack_reads(np.array([0, 7], np.int32), np.array([1, 1], np.int32), ["A", "C"], np.array([0, 1], np.uint8),
                       np.array([30, 30], np.uint8), 4)
    assert ex.value.status == -1
