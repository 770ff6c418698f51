"""CPU tests of the oracle (oracle/): pinned byte for byte on the golden outputs that the REFERENCE's own code
(lines 386-881 of cmd_cram_demuxlet.cpp, compiled into oracle/_ref/ref_harness) wrote for tests/golden/*.dmxp,
on the live reference harness where it exists, and on hand-computable cases."""
import filecmp
import json
import math
import os

import numpy as np
import pytest

from demuxlet_b200.plp import Pileup, build_from_reads
from demuxlet_b200.synth import make_synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


def oracle_for(name, po):
    m = MANIFEST[name]
    p = Pileup.load(os.path.join(GOLD, name + ".dmxp"))
    return p, po.Oracle(p, m["alphas"], m.get("doublet_prior", 0.5), m.get("min_total", 0), m.get("min_uniq", 0),
                        m.get("min_snp", 0))


@pytest.mark.parametrize("name", sorted(MANIFEST))
def test_oracle_reproduces_reference_golden_files(name, oracle_lib, tmp_path):
    p, orc = oracle_for(name, oracle_lib)
    orc.write_all(str(tmp_path / "o"), write_pair=True)
    for ext in ("single", "sing2", "best", "pair"):
        assert filecmp.cmp(tmp_path / ("o." + ext), os.path.join(GOLD, name + "." + ext), shallow=False), ext


@pytest.mark.parametrize("kw,alphas", [
    (dict(ncells=25, nv=3, nsnps=200, field="GP", seed=7, cover_frac=0.1), (0.0, 0.5)),
    (dict(ncells=20, nv=7, nsnps=150, field="GT", seed=8, cover_frac=0.1), (0.0, 0.3, 0.5)),
    (dict(ncells=12, nv=16, nsnps=500, field="PL", seed=9, cover_frac=0.05), (0.05, 0.5)),
])
def test_oracle_matches_live_reference(kw, alphas, oracle_lib, tmp_path):
    if not oracle_lib.ref_available():
        pytest.skip("oracle/_ref/ref_harness not built (no /root/reference on this machine)")
    p = make_synth(**kw)
    oracle_lib.Oracle(p, alphas).write_all(str(tmp_path / "o"), write_pair=True)
    oracle_lib.run_ref(p, str(tmp_path / "r"), alphas=alphas, write_pair=True)
    for ext in ("single", "sing2", "best", "pair"):
        assert filecmp.cmp(tmp_path / ("o." + ext), tmp_path / ("r." + ext), shallow=False), ext


def tiny(reads_per_entry, gp, alleles_bqs):
    """one cell, len(reads_per_entry) SNPs"""
    nsnps = len(reads_per_entry)
    cell, snp, umi, al, bq = [], [], [], [], []
    it = iter(alleles_bqs)
    for s, n in enumerate(reads_per_entry):
        for r in range(n):
            a, q = next(it)
            cell.append(0); snp.append(s); umi.append("U%03d" % r); al.append(a); bq.append(q)
    gp = np.asarray(gp, dtype=np.float32).reshape(nsnps, -1, 3)
    return build_from_reads(gp, np.array(cell), np.array(snp), umi, np.array(al), np.array(bq), ["AAAA-1"],
                            ["S%d" % i for i in range(gp.shape[1])])


def test_kat_one_read(oracle_lib):
    """1 cell, 1 SNP, 1 ref read at Q30: closed form of :597-684"""
    gp = [[[0.9, 0.05, 0.05], [0.1, 0.8, 0.1]]]
    p = tiny([1], gp, [(0, 30)])
    orc = oracle_lib.Oracle(p, (0.0, 0.5))
    AB, A0, L00 = orc.pass2_cell(0)
    err = 0.1 ** 3
    pR, pA = 1 - err, err / 3
    g = np.asarray(gp[0], dtype=np.float32).astype(np.float64)
    for n, alpha in enumerate((0.0, 0.5)):
        Q = np.zeros((3, 3))
        for l in range(3):
            for m in range(3):
                pp = 0.5 * l + (m - l) * 0.5 * alpha
                Q[l, m] = pR * (1 - pp) + pA * pp
        # joint normalisation over BOTH alphas: the maximum is pR at p = 0
        Q = (Q / pR + 1e-6) / (1 + 1e-6)
        for j in range(2):
            for k in range(2):
                want = math.log(sum(g[j, l] * g[k, m] * Q[l, m] for l in range(3) for m in range(3)))
                assert abs(AB[j, k, n] - want) < 1e-12


def test_quirks(oracle_lib, tmp_path):
    p = make_synth(ncells=15, nv=4, nsnps=120, field="GP", seed=3, cover_frac=0.2)
    orc = oracle_lib.Oracle(p, (0.0, 0.5))
    res = orc.run()
    b = res["best"]
    gp0 = orc.gp0()
    for c in range(p.ncells):
        AB, A0, L00 = orc.pass2_cell(c, gp0)
        # Q3: the singlet likelihood is literally pair (j, sample 0, alpha index 0)
        assert np.array_equal(res["sing"][c], AB[:, 0, 0])
        # Q4: LLK10 / LLK20 use sample 0 as the partner at the best alpha
        assert b["pair_llk10"][c] == AB[b["j_best"][c], 0, b["a_best"][c]]
        assert b["pair_llk20"][c] == AB[b["k_best"][c], 0, b["a_best"][c]]
        # separable plane: alpha = 0 makes Q'[l][m] independent of m, so AB[j][k][0] - AB[j][0][0] does not
        # depend on j (up to rounding)
        d = AB[:, :, 0] - AB[:, :1, 0]
        assert np.abs(d - d[0]).max() < 1e-9
    # Q7 / R17: .pair header has 5 names, rows have 6 columns
    orc.write_all(str(tmp_path / "q"), write_pair=True)
    lines = open(tmp_path / "q.pair").read().splitlines()
    assert len(lines[0].split("\t")) == 5 and len(lines[1].split("\t")) == 6
    # R17: at alpha = 0.5 only j < k pairs are printed
    nv = p.nv
    assert len(lines) - 1 == p.ncells * (nv + nv * (nv - 1) // 2)


def test_symmetric_pairs_differ_only_by_rounding(oracle_lib):
    """at alpha = 0.5 AB[j][k] and AB[k][j] are the same terms in a different order"""
    p = make_synth(ncells=10, nv=6, nsnps=400, field="GP", seed=5, cover_frac=0.3)
    orc = oracle_lib.Oracle(p, (0.0, 0.5))
    asym = 0
    for c in range(p.ncells):
        AB, _, _ = orc.pass2_cell(c)
        assert np.abs(AB[:, :, 1] - AB[:, :, 1].T).max() < 1e-9
        asym += orc.count_asym(AB, 1)
    assert asym > 0   # ... and the rounding noise is real: this is why the tie finaliser exists


def test_read_order_and_allele2(oracle_lib):
    # allele-2 reads carry no likelihood (:435, :604) but the SNP still counts in N.SNP (:835)
    gp = np.tile(np.array([[0.7, 0.2, 0.1], [0.2, 0.5, 0.3]], dtype=np.float32), (2, 1, 1))
    p1 = tiny([1, 1], gp, [(2, 30), (1, 25)])
    p2 = tiny([0, 1], gp, [(1, 25)])
    o1, o2 = oracle_lib.Oracle(p1).run(), oracle_lib.Oracle(p2).run()
    # an entry with only allele-2 reads has Q' == 1 everywhere: it multiplies every hypothesis by
    # sum_l sum_m gp_j[l] gp_k[m] ~ 1, so only rounding-size differences
    assert o1["best"]["n_snp"][0] == 2 and o2["best"]["n_snp"][0] == 1
    assert abs(o1["best"]["pair_llk12"][0] - o2["best"]["pair_llk12"][0]) < 1e-6
