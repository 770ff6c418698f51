"""GPU parity: libdmx (through the C ABI) against the CPU oracle on the same seeded pileups."""
import os

import numpy as np
import pytest

from tests.util import compare, diff_tsv, run_both

pytestmark = pytest.mark.gpu

BEST_EXACT = {"BARCODE", "RD.TOTL", "RD.PASS", "RD.UNIQ", "N.SNP", "BEST", "SNG.1ST", "SNG.2ND", "DBL.1ST",
              "DBL.2ND", "ALPHA"}
ROW_EXACT = {"BARCODE", "SM_ID", "SM1.ID", "SM2.ID", "RD.TOTL", "RD.PASS", "RD.UNIQ", "N.SNP"}


def synth(*a, **k):
    from demuxlet_b200.synth import make_synth
    return make_synth(*a, **k)


@pytest.mark.parametrize("ncells,nv,nsnps,field,alphas", [
    (300, 8, 2000, "GP", (0.0, 0.5)),            # C2-like
    (200, 64, 5000, "GP", (0.0, 0.5)),           # C3-like (headline shape, fewer cells)
    (500, 2, 54000, "GT", (0.0, 0.5)),           # C1' (tutorial shape: 2 samples, GT field)
    (150, 6, 3000, "GP", (0.0, 0.25, 0.5)),      # a non-symmetric plane
    (150, 5, 3000, "GP", (0.1, 0.5)),            # alpha[0] != 0: "singlet" is pair (j, 0, alpha_0)
    (60, 70, 4000, "GP", (0.0, 0.5)),            # nv > 64, not a multiple of 4: multi-block tasks
    (40, 12, 3000, "PL", tuple(round(0.05 * i, 2) for i in range(11))),   # C5-like grid
    (200, 16, 300, "GT", (0.0, 0.5)),            # GT field, low coverage: massive exact ties
    (24, 256, 6000, "GP", (0.0, 0.3, 0.5)),      # C5-like width: 4 x 4 sample blocks, symmetric and full planes
    (64, 3, 40, "GP", (0.0, 0.5)),               # tiny: cells with one or two entries, nv not a multiple of 4
])
@pytest.mark.parametrize("fp32", [False, True], ids=["f64pairs", "f32pairs"])
def test_records_match_oracle(ncells, nv, nsnps, field, alphas, fp32):
    """identity columns bit-exact in both arithmetic modes of the pair scan; LLK columns within 1e-4
    (north_star tolerance); fp32 mode: posteriors within 1e-4 relative"""
    p = synth(ncells, nv, nsnps, field, seed=11 + nv)
    orc, ref, d = run_both(p, alphas, fp32_pairs=fp32)
    worst = compare(ref, d, prb_rtol=1e-4 if fp32 else 1e-6)
    st = d.stats()
    assert st["gpu_launches"] >= 4
    print("max LLK err %.3g rescored %d fullcell %d" % (worst, st["n_rescored"], st["n_fullcell"]))


def test_filters_and_empty_cells():
    p = synth(120, 8, 1500, "GP", seed=5)
    # make a few cells empty
    keep = np.ones(p.ncells, dtype=bool)
    keep[[3, 50, 119]] = False
    from demuxlet_b200.plp import Pileup
    n_e = np.diff(p.cell_ptr)
    n_e2 = np.where(keep, n_e, 0)
    sub = p.subset_cells(np.nonzero(keep)[0])
    cell_ptr = np.concatenate([[0], np.cumsum(n_e2)])
    q = Pileup(p.gp, cell_ptr, sub.ent_snp, sub.ent_ptr, sub.reads, p.totl, p.npass, p.uniq, p.barcodes, p.samples)
    orc, ref, d = run_both(q, min_total=int(np.median(q.totl)), min_uniq=5, min_snp=3)
    compare(ref, d)
    st = d.best()["status"]
    assert (st == 1).any() and (st == 0).any()


def test_output_files(tmp_path):
    p = synth(80, 6, 2500, "GP", seed=3)
    orc, ref, d = run_both(p)
    orc.write_all(str(tmp_path / "orc"), write_pair=True)
    d.write_outputs(str(tmp_path / "gpu"), write_pair=True)
    diff_tsv(tmp_path / "orc.best", tmp_path / "gpu.best", BEST_EXACT)
    for ext in ("single", "sing2", "pair"):
        diff_tsv(tmp_path / ("orc." + ext), tmp_path / ("gpu." + ext), ROW_EXACT)


@pytest.mark.parametrize("fp32", [False, True], ids=["f64pairs", "f32pairs"])
def test_dense_pair_block_matches_oracle(fp32):
    """dmx_get_pair: dense llksAB blocks come from the fp64 scan in both modes (the fp32 scan lets hopeless hypotheses
    underflow), so they are exact to 1e-8 either way"""
    p = synth(30, 10, 2000, "GP", seed=9)
    alphas = (0.0, 0.3, 0.5)
    orc, ref, d = run_both(p, alphas, fp32_pairs=fp32)
    AB = d.pair(0, p.ncells)
    gp0 = orc.gp0()
    for c in (0, 7, 29):
        oAB, _, _ = orc.pass2_cell(c, gp0)
        assert np.abs(AB[c] - oAB).max() < 1e-8


def test_rescore_is_bit_exact():
    p = synth(20, 6, 1500, "GP", seed=4)
    orc, ref, d = run_both(p)
    for (c, j, k, n) in [(0, 1, 2, 1), (5, 2, 1, 1), (19, 0, 0, 0), (7, 3, 0, 0)]:
        assert d.rescore(c, j, k, n) == orc.pass2_one(c, j, k, n)


def test_device_rescorer_matches_host_libm():
    """the device transcription of glibc log() must be verified bit-identical to the host libm on this box,
    and the device re-scorer must then return the very bits the host scorer (and the oracle) computes"""
    from demuxlet_b200.api import Demuxlet
    p = synth(300, 16, 3000, "GP", seed=21)
    orc, ref, d = run_both(p)
    st = d.stats()
    assert st["gpu_exact"] == 1, "device log() differs from the host libm: device re-scorer disabled"
    assert st["n_rescored_gpu"] == st["n_rescored"] and st["n_rescored"] >= 2 * 300
    h = Demuxlet(p.nv, (0.0, 0.5), host_finalise=True)
    h.set_pileup(p)
    h.run()
    assert h.stats()["n_rescored_gpu"] == 0
    a, b = d.best(), h.best()
    for f in ("sng1", "sng2", "dbl1", "dbl2", "alpha_idx", "call"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["llk12"], b["llk12"])          # bit-identical doubles
    assert np.array_equal(a["llk12"], ref["best"]["pair_llk12"])   # ... and equal to the oracle's bits


def test_host_scorer_fallback_when_libm_table_is_missing(monkeypatch):
    """a host whose libm does not carry glibc's log table (DMX_NO_LIBM_TABLE simulates it) disables the device
    re-scorer at dmx_create; the host scorer then decides the same records, bit for bit in every exact field"""
    from demuxlet_b200.api import Demuxlet
    p = synth(120, 16, 2500, "GP", seed=77)
    d = Demuxlet(p.nv, (0.0, 0.5))
    d.set_pileup(p)
    d.run()
    assert d.stats()["gpu_exact"] == 1
    monkeypatch.setenv("DMX_NO_LIBM_TABLE", "1")
    h = Demuxlet(p.nv, (0.0, 0.5))
    monkeypatch.delenv("DMX_NO_LIBM_TABLE")
    h.set_pileup(p)
    h.run()
    st = h.stats()
    assert st["gpu_exact"] == 0 and st["n_rescored_gpu"] == 0 and st["n_rescored"] >= 2 * 120
    a, b = d.best(), h.best()
    for f in ("sng1", "sng2", "dbl1", "dbl2", "alpha_idx", "call", "n_snp"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["llk12"], b["llk12"])          # the same exact bits from either scorer


def _golden_cases():
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    return sorted(json.load(open(os.path.join(here, "golden", "manifest.json"))).items())


@pytest.mark.parametrize("name,m", _golden_cases())
def test_against_reference_golden_files(name, m, tmp_path):
    """libdmx vs the files the REFERENCE's own code wrote (tests/golden, made by make_golden.py): identity
    columns byte-identical, every numeric column within 1e-4"""
    from demuxlet_b200.api import Demuxlet
    from demuxlet_b200.plp import Pileup
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)
    p = Pileup.load(gold + ".dmxp")
    d = Demuxlet(p.nv, m["alphas"], m.get("doublet_prior", 0.5), m.get("min_total", 0), m.get("min_uniq", 0),
                 m.get("min_snp", 0))
    d.set_pileup(p)
    d.run()
    d.write_outputs(str(tmp_path / "g"), write_pair=True)
    diff_tsv(gold + ".best", tmp_path / "g.best", BEST_EXACT)
    for ext in ("single", "sing2", "pair"):
        diff_tsv(gold + "." + ext, tmp_path / ("g." + ext), ROW_EXACT)


def test_full_size_properties():
    """headline shape (64 samples, 100k SNPs; 1500 barcodes): size-independent properties instead of the oracle:
    relabelling invariance (reversing the barcode order permutes the records), symmetric-plane consistency
    LLK12 >= LLK1/LLK2 structure is NOT assumed; exact-tie resolution agrees between device and host scorers"""
    from demuxlet_b200.api import Demuxlet
    from demuxlet_b200.synth import make_config
    p, alphas = make_config("C3", scale=0.15)
    d = Demuxlet(p.nv, alphas)
    d.set_pileup(p)
    d.run()
    a = d.best()
    perm = np.arange(p.ncells)[::-1].copy()
    q = p.subset_cells(perm)
    e = Demuxlet(p.nv, alphas, host_finalise=True)
    e.set_pileup(q)
    e.run()
    b = e.best()[np.argsort(perm)]
    for f in ("sng1", "sng2", "dbl1", "dbl2", "alpha_idx", "call", "n_snp"):
        assert np.array_equal(a[f], b[f]), f
    assert np.array_equal(a["llk12"], b["llk12"])                       # exact values are order independent
    for f in ("sng_llk1", "sng_llk2", "llk10", "llk20", "llk00", "max_llk"):
        assert np.abs(a[f] - b[f]).max() < 1e-8
    # spot-check 24 barcodes against the oracle at full per-barcode size
    from oracle import pyoracle as po
    cells = np.linspace(0, p.ncells - 1, 24).astype(int)
    sub = p.subset_cells(cells)
    ref = po.Oracle(sub, alphas).run(nthreads=os.cpu_count() or 1, want_pass1=False)["best"]
    for of, gf in (("i_sing1", "sng1"), ("j_best", "dbl1"), ("k_best", "dbl2"), ("call", "call")):
        assert np.array_equal(ref[of], a[gf][cells]), gf
    assert np.array_equal(ref["pair_llk12"], a["llk12"][cells])
    assert np.abs(ref["pair_llk10"] - a["llk10"][cells]).max() < 1e-4


@pytest.mark.parametrize("name", ["gp6", "gp6filt", "gt2"])
def test_cli_reproduces_reference_files(name, tmp_path):
    """dmx-demuxlet (C++ driver over the C ABI) with demuxlet's options on a golden pileup: the four files against
    the ones the reference's code wrote; then --group-list keeps exactly the listed barcodes"""
    import json
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(os.path.dirname(here), "demuxlet_b200", "dmx-demuxlet")
    m = json.load(open(os.path.join(here, "golden", "manifest.json")))[name]
    gold = os.path.join(here, "golden", name)
    cmd = [exe, "--plp", gold + ".dmxp", "--out", str(tmp_path / "c"), "--write-pair", "--fp64-pairs"]
    for a in m["alphas"]:
        cmd += ["--alpha", repr(float(a))]
    for opt, key in (("--doublet-prior", "doublet_prior"), ("--min-total", "min_total"), ("--min-uniq", "min_uniq"),
                     ("--min-snp", "min_snp")):
        if key in m:
            cmd += [opt, str(m[key])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    diff_tsv(gold + ".best", tmp_path / "c.best", BEST_EXACT)
    for ext in ("single", "sing2", "pair"):
        diff_tsv(gold + "." + ext, tmp_path / ("c." + ext), ROW_EXACT)
    # --group-list: every second barcode of the .best file
    rows = open(gold + ".best").read().splitlines()
    bcs = [ln.split("\t")[0] for ln in rows[1:]][::2]
    (tmp_path / "groups.txt").write_text("\n".join(bcs) + "\n")
    r = subprocess.run(cmd[:4] + [str(tmp_path / "g")] + cmd[5:] + ["--group-list", str(tmp_path / "groups.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = open(tmp_path / "g.best").read().splitlines()
    want = [rows[0]] + [ln for ln in rows[1:] if ln.split("\t")[0] in set(bcs)]
    assert [ln.split("\t")[0] for ln in got] == [ln.split("\t")[0] for ln in want]
    for a, b in zip(got[1:], want[1:]):
        fa, fb = a.split("\t"), b.split("\t")
        for i in (0, 1, 2, 3, 4, 5, 6, 8, 11, 12, 13):   # identity columns incl. BEST / SNG.1ST / DBL.1ST / DBL.2ND / ALPHA
            assert fa[i] == fb[i], (i, a, b)


def test_error_paths_and_context_reuse():
    """call-order and input errors come back as status codes (never an abort), and one context can be fed several
    pileups through both entry points (dmx_run: pipelined copies; dmx_upload + dmx_compute: resident inputs)"""
    from demuxlet_b200.api import Demuxlet, DmxError
    p = synth(90, 8, 1200, "GP", seed=31)
    d = Demuxlet(p.nv, (0.0, 0.5), fp32_pairs=True)
    with pytest.raises(DmxError) as e:
        d.run()                                   # nothing set yet
    assert e.value.status == -4
    d.set_gp(p.gp)
    bad = p.cell_ptr.copy()
    bad[5], bad[6] = bad[6], bad[5] + 0           # not monotone
    if bad[6] < bad[5]:
        with pytest.raises(DmxError) as e:
            d.set_pileup_arrays(bad, p.ent_snp, p.ent_ptr, p.reads, p.totl, p.npass, p.uniq)
        assert e.value.status == -1
    snp_bad = p.ent_snp.copy()
    snp_bad[3] = p.nsnps + 7                      # SNP id outside the GP matrix
    d.set_pileup_arrays(p.cell_ptr, snp_bad, p.ent_ptr, p.reads, p.totl, p.npass, p.uniq)
    with pytest.raises(DmxError) as e:
        d.run()
    assert e.value.status == -1
    # now the good inputs, three ways; results must be identical
    d.set_pileup(p)
    d.run()
    a = d.best().copy()
    d.upload()
    d.compute()
    b = d.best().copy()
    q = synth(40, 8, 1200, "GP", seed=32)          # another pileup through the same context
    d.set_pileup(q)
    d.run()
    assert d.best().shape[0] == 40
    d.set_pileup(p)
    d.run()
    c2 = d.best().copy()
    d.set_pileup(p)                                # the same droplets again: the schedule is kept, the data re-uploaded
    d.run()
    assert d.best().tobytes() == c2.tobytes()
    r2 = p.reads.copy()
    r2[::7] = (r2[::7] & 0xff00) | 20              # same droplets, other qualities: not a cached result
    d.set_pileup_arrays(p.cell_ptr, p.ent_snp, p.ent_ptr, r2, p.totl, p.npass, p.uniq)
    d.run()
    assert d.best().tobytes() != c2.tobytes()
    d.set_pileup(p)
    d.run()
    assert d.best().tobytes() == c2.tobytes()
    for f in ("sng1", "sng2", "dbl1", "dbl2", "alpha_idx", "call", "n_snp"):
        assert np.array_equal(a[f], b[f]) and np.array_equal(a[f], c2[f]), f
    # runs are reproducible bit for bit: the per-sample products are accumulated in stage order by one warp, whichever
    # transform warp produced a stage (the chunked schedule of dmx_run visits the same cells with the same arithmetic)
    for f in ("llk12", "sng_llk1", "sng_llk2", "llk1", "llk2", "llk10", "llk20", "llk00", "prb_dbl", "prb_sng1"):
        assert np.array_equal(a[f], b[f]) and np.array_equal(a[f], c2[f]), f
    orc, ref, e2 = run_both(p, (0.0, 0.5), fp32_pairs=True)
    compare(ref, d, prb_rtol=1e-4)
