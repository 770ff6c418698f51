"""VCF field -> genotype-probability matrix (dmx_build_gp / dmx_set_field).

CPU: the oracle (oracle/dmx_field_oracle.c) against golden matrices the REFERENCE's own parse_posteriors /
parse_genotypes / parse_likelihoods wrote (tests/golden/field_*.npz, make_field_golden.py), against the live reference
functions where oracle/_ref/ref_field is built, and against the numpy restatement the generator uses; argument checks
of the C ABI.  GPU (-m gpu): the CUDA kernels against the oracle and the golden files, bit for bit (uint32 views)."""
import os

import numpy as np
import pytest

from demuxlet_b200 import api, synth
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("GT", "GP", "PL")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def make_field(field, nsnps, nv, seed):
    rng = np.random.default_rng(seed)
    if field == "GT":
        a = rng.choice(np.array([-1, 0, 1], np.int8), size=(nsnps, nv, 2), p=[0.05, 0.6, 0.35])
        if nsnps > 2:
            a[0] = -1
            a[1, :, 1] = -1
        return a
    if field == "GP":
        g = rng.integers(0, 3, size=(nsnps, nv))
        return (np.eye(3, dtype=np.float32)[g] + rng.random((nsnps, nv, 3), dtype=np.float32) * np.float32(0.05))
    pl = rng.integers(0, 100, size=(nsnps, nv, 3)).astype(np.int32)
    pl[np.arange(nsnps)[:, None], np.arange(nv)[None, :], rng.integers(0, 3, size=(nsnps, nv))] = 0
    if nsnps > 2:
        pl[0, 0] = np.iinfo(np.int32).min
        pl[1, :, 2] = 999
    return pl


def load_gold(name):
    z = np.load(os.path.join(GOLD, "field_%s.npz" % name))
    return str(z["field"]), float(z["gt_error"]), z["input"], z["gp"], z["af"]


# ---- CPU: the oracle is pinned on the reference ------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gt", "gp", "pl"])
def test_oracle_against_reference_golden(name):
    field, ge, arr, gp, af = load_gold(name)
    got, gaf = po.build_gp(field, arr, ge, want_af=True)
    assert np.array_equal(bits(got), bits(gp))
    if field != "GP":
        assert np.array_equal(gaf, af)
    assert np.array_equal(bits(po.build_gp(field, arr, ge, nthreads=4)), bits(gp))


@pytest.mark.skipif(not po.ref_field_available(), reason="oracle/_ref/ref_field not built")
@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("nsnps,nv,seed", [(400, 8, 1), (50, 64, 2), (7, 1, 3), (3, 131, 4)])
def test_oracle_against_live_reference(field, nsnps, nv, seed):
    arr = make_field(field, nsnps, nv, seed)
    for ge in (0.01, 1e-4):
        want, waf = po.run_ref_field(field, arr, ge, want_af=True)
        got, gaf = po.build_gp(field, arr, ge, want_af=True)
        assert np.array_equal(bits(got), bits(want))
        if field != "GP":
            assert np.array_equal(gaf, waf)


def test_oracle_against_generator_restatement():
    # demuxlet_b200/synth.py holds the numpy form the synthetic generator uses; same numbers up to float rounding order
    rng = np.random.default_rng(5)
    g = rng.integers(0, 3, size=(200, 12))
    raw = (np.eye(3, dtype=np.float32)[g] + rng.random((200, 12, 3), dtype=np.float32) * np.float32(0.05))
    assert np.abs(po.build_gp("GP", raw, 0.01) - synth.gp_from_gp(raw, 0.01)).max() < 1e-6
    gt = np.stack([(g >= 1), (g >= 2)], axis=-1).astype(np.int8)
    assert np.abs(po.build_gp("GT", gt, 0.01) - synth.gp_from_gt(g, 0.01)).max() < 1e-7
    pl = make_field("PL", 100, 12, 6)
    assert np.abs(po.build_gp("PL", pl, 0.01) - synth.gp_from_pl(np.clip(pl, 0, 255))).max() < 1e-6


def test_site_filter_oracle_against_reference():
    z = np.load(os.path.join(GOLD, "field_gt.npz"))
    assert np.array_equal(po.site_filter(z["input"], 1, 0.5), z["pass_1_50"])
    assert np.array_equal(po.site_filter(z["input"], 2, 0.8), z["pass_2_80"])
    assert 0 < z["pass_1_50"].sum() < len(z["pass_1_50"])
    if po.ref_field_available():
        gt = make_field("GT", 500, 9, 13)
        gt[10:60, :6] = -1
        for mm, cr in ((1, 0.5), (0, 0.0), (3, 0.9)):
            assert np.array_equal(po.site_filter(gt, mm, cr), po.run_ref_site_filter(gt, mm, cr))


def test_rows_are_distributions():
    for field in FIELDS:
        gp = po.build_gp(field, make_field(field, 300, 9, 7), 0.01)
        assert np.all(gp >= 0) and np.abs(gp.sum(axis=2) - 1).max() < 1e-5


def test_abi_argument_checks():
    lib = api.load_library()
    buf = np.zeros((2, 3, 3), np.float32)
    out = np.zeros((2, 3, 3), np.float32)
    assert lib.dmx_build_gp(0, 7, 3, 2, buf.ctypes.data, 0.01, out.ctypes.data, None, None) == -1     # unknown field
    assert lib.dmx_build_gp(0, 1, 0, 2, buf.ctypes.data, 0.01, out.ctypes.data, None, None) == -1     # nv < 1
    assert lib.dmx_build_gp(0, 1, 3, 2, None, 0.01, out.ctypes.data, None, None) == -1                # null input
    assert lib.dmx_build_gp_device(0, 1, 3, 2, buf.ctypes.data, 0.01, out.ctypes.data, 8, None, None) == -1   # stride < nv * 3
    assert lib.dmx_set_field(None, 1, buf.ctypes.data, 2, 0.01) == -1
    assert lib.dmx_site_filter(0, 0, 2, buf.ctypes.data, 1, 0.5, out.ctypes.data) == -1
    if api.device_count() == 0:   # no CPU fallback
        assert lib.dmx_build_gp(0, 1, 3, 2, buf.ctypes.data, 0.01, out.ctypes.data, None, None) == -3
        with pytest.raises(api.DmxError):
            api.build_gp("GP", buf)


# ---- GPU: kernels against oracle and golden, bit for bit ---------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gt", "gp", "pl"])
def test_gpu_against_reference_golden(name):
    field, ge, arr, gp, af = load_gold(name)
    got, gaf = api.build_gp(field, arr, ge, want_af=True)
    assert np.array_equal(bits(got), bits(gp))
    if field != "GP":
        assert np.array_equal(gaf, af)


@pytest.mark.gpu
@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("nsnps,nv,seed", [(3000, 8, 1), (1000, 64, 2), (33, 1, 3), (257, 2, 4), (200, 131, 5),
                                           (40, 700, 6), (0, 4, 7), (1, 5, 8)])
def test_gpu_against_oracle(field, nsnps, nv, seed):
    arr = make_field(field, nsnps, nv, seed)
    for ge in (0.01, 1e-4):
        want, waf = po.build_gp(field, arr, ge, nthreads=8, want_af=True)
        (got, gaf), st = api.build_gp(field, arr, ge, want_af=True, want_stats=True)
        assert np.array_equal(bits(got), bits(want))
        if field != "GP":
            assert np.array_equal(gaf, waf)
        assert st["gpu_launches"] == (1 if nsnps else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("field", FIELDS)
def test_gpu_full_size_chunked(field):
    # the headline shape (100 k SNPs x 64 samples): five SNP chunks on two streams, pinned and pageable buffers
    nsnps, nv = 100000, 64
    arr = make_field(field, nsnps, nv, 11)
    want = po.build_gp(field, arr, 0.01, nthreads=os.cpu_count() or 1)
    got, st = api.build_gp(field, arr, 0.01, want_stats=True)
    assert st["gpu_launches"] > 1
    assert np.array_equal(bits(got), bits(want))
    pin_in = api.pinned_empty(arr.shape, arr.dtype)
    pin_in[...] = arr
    pin_out = api.pinned_empty((nsnps, nv, 3), np.float32)
    api.build_gp(field, pin_in, 0.01, out=pin_out)
    assert np.array_equal(bits(pin_out), bits(want))
    assert np.abs(got.sum(axis=2) - 1).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("nsnps,nv", [(48, 5), (5000, 64), (300, 131), (1, 2)])
def test_gpu_site_filter(nsnps, nv):
    if (nsnps, nv) == (48, 5):
        z = np.load(os.path.join(GOLD, "field_gt.npz"))
        assert np.array_equal(api.site_filter(z["input"], 1, 0.5), z["pass_1_50"])
        assert np.array_equal(api.site_filter(z["input"], 2, 0.8), z["pass_2_80"])
    gt = make_field("GT", nsnps, nv, 17)
    gt[: nsnps // 3, : (2 * nv) // 3] = -1
    for mm, cr in ((1, 0.5), (0, 0.0), (3, 0.9), (1, 1.0)):
        assert np.array_equal(api.site_filter(gt, mm, cr), po.site_filter(gt, mm, cr))


@pytest.mark.gpu
def test_gpu_rejects_non_biallelic_gt():
    gt = np.zeros((10, 4, 2), np.int8)
    gt[3, 1, 0] = 2
    with pytest.raises(api.DmxError):
        api.build_gp("GT", gt)


@pytest.mark.gpu
@pytest.mark.parametrize("field,nv", [("GT", 6), ("PL", 64), ("GP", 70)])
def test_set_field_matches_set_gp(field, nv):
    # dmx_set_field (matrix built into the context's device copy) gives the run dmx_set_gp gives on the oracle's matrix
    p = synth.make_synth(40, nv, 3000, "GP", seed=9)
    arr = make_field(field, 3000, nv, 12)
    gp = po.build_gp(field, arr, 0.01)
    res = []
    for use_field in (False, True):
        d = api.Demuxlet(nv, (0.0, 0.5))
        if use_field:
            d.set_field(field, arr, 0.01)
        else:
            d.set_gp(gp)
        d.set_pileup_arrays(p.cell_ptr, p.ent_snp, p.ent_ptr, p.reads, p.totl, p.npass, p.uniq)
        d.run()
        res.append((d.best().copy(), d.stats()["h2d_bytes"]))
        d.close()
    a, b = res[0][0], res[1][0]
    for f in a.dtype.names:
        if f not in ("n_rescored", "exact_mask"):
            assert np.array_equal(a[f], b[f]), f
    assert res[1][1] == res[0][1] - gp.nbytes      # the matrix was not uploaded a second time


@pytest.mark.gpu
@pytest.mark.parametrize("field", FIELDS)
def test_cli_field_matrix(field, tmp_path):
    """dmx-demuxlet --field F --field-matrix FILE (matrix built on the GPU from the FORMAT values, like demuxlet's
    --field / --geno-error per VCF record) writes the files --plp alone writes with the oracle's matrix inside"""
    import dataclasses
    import filecmp
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "demuxlet_b200", "dmx-demuxlet")
    nv, nsnps = 6, 2500
    p = synth.make_synth(30, nv, nsnps, "GP", seed=4)
    arr = make_field(field, nsnps, nv, 15)
    with_gp = dataclasses.replace(p, gp=po.build_gp(field, arr, 0.02))
    without = dataclasses.replace(p, gp=np.zeros_like(p.gp))
    with_gp.save(str(tmp_path / "a.dmxp"))
    without.save(str(tmp_path / "b.dmxp"))
    po.write_field_file(str(tmp_path / "b.fld"), field, arr, 0.01)
    ra = subprocess.run([exe, "--plp", str(tmp_path / "a.dmxp"), "--out", str(tmp_path / "a")], capture_output=True, text=True)
    rb = subprocess.run([exe, "--plp", str(tmp_path / "b.dmxp"), "--out", str(tmp_path / "b"), "--field", field,
                         "--field-matrix", str(tmp_path / "b.fld"), "--geno-error", "0.02"], capture_output=True, text=True)
    assert ra.returncode == 0 and rb.returncode == 0, ra.stderr + rb.stderr
    for ext in ("best", "single", "sing2"):
        assert filecmp.cmp(str(tmp_path / ("a." + ext)), str(tmp_path / ("b." + ext)), shallow=False), ext
    bad = subprocess.run([exe, "--plp", str(tmp_path / "b.dmxp"), "--out", str(tmp_path / "c"), "--field",
                          "GP" if field != "GP" else "PL", "--field-matrix", str(tmp_path / "b.fld")], capture_output=True, text=True)
    assert bad.returncode == 1 and "does not match" in bad.stderr
