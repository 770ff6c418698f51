"""GPU parity at the shapes BASELINE.json names (all identity columns + every LLK column on C2 at
full size and on a C3 subset at full per-barcode size), plus C5's per-barcode shape (256 samples, the 11-point alpha
grid of cmd_cram_demuxlet.cpp:79-89, PL field, >= 5k covered SNPs per barcode) and C4's --write-pair rows.
Every case runs both arithmetic modes of the pair scan against the same oracle results and prints the largest
LLK error it saw."""
import os

import numpy as np
import pytest

from tests.util import compare, diff_tsv

pytestmark = pytest.mark.gpu

_ORACLE = {}


def oracle_of(key, plp, alphas):
    """oracle results are shared between the two scan modes of a case"""
    if key not in _ORACLE:
        from oracle import pyoracle as po
        orc = po.Oracle(plp, alphas)
        _ORACLE[key] = (orc, orc.run(nthreads=os.cpu_count() or 1))
    return _ORACLE[key]


def gpu_run(plp, alphas, fp32, **kw):
    from demuxlet_b200.api import Demuxlet
    d = Demuxlet(plp.nv, alphas, fp32_pairs=fp32, **kw)
    d.set_pileup(plp)
    d.run()
    return d


def config(name, scale):
    from demuxlet_b200.synth import make_config
    key = (name, scale)
    if key not in _ORACLE:
        _ORACLE[key] = make_config(name, scale=scale)
    return _ORACLE[key]


@pytest.mark.parametrize("fp32", [False, True], ids=["f64pairs", "f32pairs"])
def test_c2_full_size(fp32):
    """configs[1]: 10k barcodes x 8 samples x 20k SNPs, GP field, alpha {0, 0.5}: every barcode against the oracle"""
    p, alphas = config("C2", 1.0)
    assert p.ncells == 10000 and p.nv == 8 and p.nsnps == 20000
    orc, ref = oracle_of("C2", p, alphas)
    d = gpu_run(p, alphas, fp32)
    worst = compare(ref, d, prb_rtol=1e-4 if fp32 else 1e-6)
    st = d.stats()
    print("C2 full (%s): max LLK err %.3g, %d hypotheses re-scored, %d dense re-scans" %
          ("fp32" if fp32 else "fp64", worst, st["n_rescored"], st["n_fullcell"]))


@pytest.mark.parametrize("fp32", [False, True], ids=["f64pairs", "f32pairs"])
def test_c3_subset_full_compare(fp32, tmp_path):
    """configs[2] (headline): 250 barcodes of the 10k x 64 samples x 100k SNPs pileup at full per-barcode size
    (~1.1k covered SNPs each): all record fields, .sing2 / .single values, and the four files"""
    p, alphas = config("C3", 0.025)
    assert p.ncells == 250 and p.nv == 64 and p.nsnps == 100000
    assert np.diff(p.cell_ptr).mean() > 900
    orc, ref = oracle_of("C3", p, alphas)
    d = gpu_run(p, alphas, fp32)
    worst = compare(ref, d, prb_rtol=1e-4 if fp32 else 1e-6)
    print("C3 subset (%s): max LLK err %.3g over %d barcodes, mean %.0f entries" %
          ("fp32" if fp32 else "fp64", worst, p.ncells, np.diff(p.cell_ptr).mean()))
    orc.write_all(str(tmp_path / "orc"))
    d.write_outputs(str(tmp_path / "gpu"))
    from tests.test_gpu_parity import BEST_EXACT, ROW_EXACT
    diff_tsv(tmp_path / "orc.best", tmp_path / "gpu.best", BEST_EXACT)
    for ext in ("single", "sing2"):
        diff_tsv(tmp_path / ("orc." + ext), tmp_path / ("gpu." + ext), ROW_EXACT)


def c5_cells():
    """C5's per-barcode shape: 256 samples, PL field, >= 5k covered SNPs per barcode (the SNP axis is cut to 40k:
    only the size of the GP matrix depends on it, not the per-barcode work)"""
    if "c5p" not in _ORACLE:
        from demuxlet_b200.synth import make_synth, CONFIGS
        p = make_synth(3, 256, 40000, "PL", seed=20261022, cover_frac=0.16, sigma=0.05, cover_min=5600)
        _ORACLE["c5p"] = (p, CONFIGS["C5"]["alphas"])
    return _ORACLE["c5p"]


@pytest.mark.parametrize("fp32", [False, True], ids=["f64pairs", "f32pairs"])
def test_c5_barcode_shape(fp32):
    """configs[4]: nv = 256 (4 x 4 sample blocks), the 11-point alpha grid (one separable, one symmetric and nine full
    planes), PL-field genotype probabilities, > 5k entries per barcode -- where the fp32 scan's rounding error is
    largest.  The error actually seen is printed; the bound is the north-star 1e-4."""
    p, alphas = c5_cells()
    assert len(alphas) == 11 and p.nv == 256
    assert np.diff(p.cell_ptr).min() >= 5000
    orc, ref = oracle_of("C5", p, alphas)
    d = gpu_run(p, alphas, fp32)
    worst = compare(ref, d, prb_rtol=1e-4 if fp32 else 1e-6)
    print("C5 barcode shape (%s): max LLK err %.3g, entries per barcode %s" %
          ("fp32" if fp32 else "fp64", worst, np.diff(p.cell_ptr).tolist()))


def test_c4_write_pair_rows(tmp_path):
    """configs[3] prints --write-pair: the .pair rows of 120 barcodes at C4's per-barcode shape (64 samples, 2 080 rows
    each) against the oracle's, through dmx_write_outputs"""
    from demuxlet_b200.synth import make_config
    p, alphas = make_config("C4", scale=120 / 50000)
    from oracle import pyoracle as po
    orc = po.Oracle(p, alphas)
    orc.run(nthreads=os.cpu_count() or 1)
    orc.write_all(str(tmp_path / "orc"), write_pair=True)
    d = gpu_run(p, alphas, True)
    d.write_outputs(str(tmp_path / "gpu"), write_pair=True)
    from tests.test_gpu_parity import BEST_EXACT, ROW_EXACT
    diff_tsv(tmp_path / "orc.best", tmp_path / "gpu.best", BEST_EXACT)
    diff_tsv(tmp_path / "orc.pair", tmp_path / "gpu.pair", ROW_EXACT)
    rows = sum(1 for _ in open(tmp_path / "gpu.pair")) - 1
    assert rows == p.ncells * (64 + 64 * 63 // 2)


# ---- sharded execution of ONE job (cmd_cram_demuxlet.cpp:68, :92-99) ----------------------------------
def _device_list(n):
    """n shards over the GPUs of this box; with fewer GPUs than shards the devices repeat (two shards on one GPU run
    the same code path: separate contexts, streams and host threads)"""
    from demuxlet_b200 import api
    have = max(1, api.device_count())
    return [i % have for i in range(n)]


@pytest.mark.parametrize("nshards", [2, 3])
@pytest.mark.parametrize("fp32", [False, True], ids=["f64pairs", "f32pairs"])
def test_sharded_equals_single_device(nshards, fp32, tmp_path):
    """dmx_create_sharded: barcode ranges balanced by entries, one context per shard, records gathered into one set of
    arrays: the four files must equal the single-device run's byte for byte"""
    from demuxlet_b200.api import Demuxlet
    from demuxlet_b200.synth import make_synth
    p = make_synth(400, 12, 5000, "GP", seed=77)
    alphas = (0.0, 0.25, 0.5)
    one = gpu_run(p, alphas, fp32)
    many = Demuxlet(p.nv, alphas, fp32_pairs=fp32, devices=_device_list(nshards))
    many.set_pileup(p)
    many.run()
    sh = many.shards()
    assert len(sh) == nshards and sh[0][0] == 0 and sh[-1][1] == p.ncells
    ent = [int(p.cell_ptr[hi] - p.cell_ptr[lo]) for lo, hi, _ in sh]
    assert max(ent) - min(ent) <= 2 * int(np.diff(p.cell_ptr).max())       # balanced by entries, not by count
    a, b = one.best(), many.best()
    assert a.tobytes() == b.tobytes()
    for x, y in zip(one.single() + one.sing2(), many.single() + many.sing2()):
        assert np.array_equal(x, y)
    assert np.array_equal(one.pair(3, 390), many.pair(3, 390))
    assert many.rescore(333, 2, 5, 1) == one.rescore(333, 2, 5, 1)
    one.write_outputs(str(tmp_path / "one"), write_pair=True)
    many.write_outputs(str(tmp_path / "many"), write_pair=True)
    for ext in ("best", "single", "sing2", "pair"):
        assert open(tmp_path / ("one." + ext), "rb").read() == open(tmp_path / ("many." + ext), "rb").read(), ext
    # resident inputs: upload once, compute twice
    many.upload()
    many.compute()
    assert many.best().tobytes() == a.tobytes()
    st = many.stats()
    assert st["n_entries"] == p.n_entries and st["gpu_launches"] >= 4 * nshards
    orc, ref = oracle_of("shard", p, alphas)
    compare(ref, many, prb_rtol=1e-4 if fp32 else 1e-6)


def test_sharded_cli_and_field(tmp_path):
    """dmx-demuxlet --gpus / --gpu-list writes the files of the single-device run; dmx_set_field on a sharded context"""
    import subprocess
    from demuxlet_b200.api import Demuxlet
    from demuxlet_b200.synth import make_synth
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(os.path.dirname(here), "demuxlet_b200", "dmx-demuxlet")
    p = make_synth(300, 8, 4000, "GP", seed=78)
    p.save(str(tmp_path / "in.dmxp"))
    devs = _device_list(3)
    for tag, extra in (("one", []), ("many", ["--gpu-list", ",".join(map(str, devs))])):
        r = subprocess.run([exe, "--plp", str(tmp_path / "in.dmxp"), "--out", str(tmp_path / tag), "--verbose"] + extra,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    assert "shard 2" in r.stderr
    for ext in ("best", "single", "sing2"):
        assert open(tmp_path / ("one." + ext), "rb").read() == open(tmp_path / ("many." + ext), "rb").read(), ext
    # the matrix from a VCF field, built on every shard's device
    from tests.test_field import make_field
    g = make_field("GP", p.nsnps, p.nv, 9)
    one = Demuxlet(p.nv)
    one.set_field("GP", g)
    one.set_pileup_arrays(p.cell_ptr, p.ent_snp, p.ent_ptr, p.reads, p.totl, p.npass, p.uniq)
    one.run()
    many = Demuxlet(p.nv, devices=devs)
    many.set_field("GP", g)
    many.set_pileup_arrays(p.cell_ptr, p.ent_snp, p.ent_ptr, p.reads, p.totl, p.npass, p.uniq)
    many.run()
    assert one.best().tobytes() == many.best().tobytes()


# ---- the documented call site, compiled against the reference's real container --------
def _golden():
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    return sorted(json.load(open(os.path.join(here, "golden", "manifest.json"))).items())


@pytest.mark.parametrize("name,m", _golden())
def test_reference_side_binding(name, m, tmp_path):
    """oracle/_ref/ref_bind = the reference's sc_dropseq_lib_t (its own add_snp / add_cell / add_read) + the flattening
    block of INTEGRATION.md + libdmx: its four files against the ones the reference's lines 386-881 wrote for the same
    pileup (tests/golden)"""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = os.path.join(os.path.dirname(here), "oracle", "_ref", "ref_bind")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_bind is built only where /root/reference exists (it travels with the snapshot)")
    gold = os.path.join(here, "golden", name)
    cmd = [exe, gold + ".dmxp", str(tmp_path / "b"), "--write-pair"]
    for a in m["alphas"]:
        cmd += ["--alpha", repr(float(a))]
    for opt, key in (("--doublet-prior", "doublet_prior"), ("--min-total", "min_total"), ("--min-uniq", "min_uniq"),
                     ("--min-snp", "min_snp")):
        if key in m:
            cmd += [opt, str(m[key])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from tests.test_gpu_parity import BEST_EXACT, ROW_EXACT
    diff_tsv(gold + ".best", tmp_path / "b.best", BEST_EXACT)
    for ext in ("single", "sing2", "pair"):
        diff_tsv(gold + "." + ext, tmp_path / ("b." + ext), ROW_EXACT)


def test_packed_offsets_upload(monkeypatch):
    """dmx_run sends one-byte read counts instead of ent_ptr and rebuilds the offsets on the device (pack_chunk /
    entptr_kernel): same records as the plain upload, as dmx_upload + dmx_compute (never packed), and a pileup with an
    entry of more than 255 reads (which cannot be packed) still runs"""
    from demuxlet_b200.api import Demuxlet
    from demuxlet_b200.synth import make_synth
    p = make_synth(700, 8, 20000, "GP", seed=91)
    assert p.n_entries > (1 << 16)                     # large enough for the packed path ...
    monkeypatch.setenv("DMX_PTR_PACK_MIN", "1")        # ... once the "large pileups only" rule is out of the way
    d = Demuxlet(p.nv, (0.0, 0.5))
    d.set_pileup(p)
    d.run()
    a = d.best().copy()
    assert d.stats()["h2d_bytes"] < p.gp.nbytes + p.n_entries * 8     # the 8-byte offsets did not travel
    d.upload()
    d.compute()
    assert d.best().tobytes() == a.tobytes()
    monkeypatch.setenv("DMX_NO_PTR_PACK", "1")
    e = Demuxlet(p.nv, (0.0, 0.5))
    e.set_pileup(p)
    e.run()
    monkeypatch.delenv("DMX_NO_PTR_PACK")
    assert e.best().tobytes() == a.tobytes()
    assert e.stats()["h2d_bytes"] > p.gp.nbytes + p.n_entries * 8
    # one entry with 300 reads: its count does not fit a byte, the chunk goes unpacked
    k = int(p.ent_ptr[5])
    reads = np.concatenate([p.reads[:k], np.full(300, (1 << 8) | 30, np.uint16), p.reads[k:]])
    ent_ptr = p.ent_ptr.copy()
    ent_ptr[6:] += 300
    f = Demuxlet(p.nv, (0.0, 0.5))
    f.set_gp(p.gp)
    f.set_pileup_arrays(p.cell_ptr, p.ent_snp, ent_ptr, reads, p.totl, p.npass, p.uniq)
    f.run()
    g = Demuxlet(p.nv, (0.0, 0.5))
    g.set_gp(p.gp)
    g.set_pileup_arrays(p.cell_ptr, p.ent_snp, ent_ptr, reads, p.totl, p.npass, p.uniq)
    g.upload()
    g.compute()
    assert f.best().tobytes() == g.best().tobytes()
    assert not np.array_equal(f.best()["llk12"][:1], a["llk12"][:1])     # the extra reads changed barcode 0
