"""CPU tests of the host-side pieces: the flat pileup container, the reference-order packer, the synthetic
generator's VCF-field transforms, barcode sharding and the record all-gather (gloo, world_size 2)."""
import os

import numpy as np
import pytest

from demuxlet_b200.parallel import shard_ranges
from demuxlet_b200.plp import Pileup, build_from_reads
from demuxlet_b200.synth import gp_from_gp, gp_from_gt, gp_from_pl, make_synth


def test_pileup_roundtrip_and_subset(tmp_path):
    p = make_synth(50, 5, 300, "GP", seed=2, cover_frac=0.1)
    path = str(tmp_path / "a.dmxp")
    p.save(path)
    q = Pileup.load(path)
    for f in ("gp", "cell_ptr", "ent_snp", "ent_ptr", "reads", "totl", "npass", "uniq"):
        assert np.array_equal(getattr(p, f), getattr(q, f))
    assert q.barcodes == p.barcodes and q.samples == p.samples
    cells = [7, 3, 44]
    s = p.subset_cells(cells, compact_snps=True)
    assert s.ncells == 3 and s.nsnps == len(np.unique(s.ent_snp))
    for i, c in enumerate(cells):
        e0, e1 = p.cell_ptr[c], p.cell_ptr[c + 1]
        f0, f1 = s.cell_ptr[i], s.cell_ptr[i + 1]
        assert e1 - e0 == f1 - f0
        assert np.array_equal(p.gp[p.ent_snp[e0:e1]], s.gp[s.ent_snp[f0:f1]])
        assert np.array_equal(p.reads[p.ent_ptr[e0]:p.ent_ptr[e1]], s.reads[s.ent_ptr[f0]:s.ent_ptr[f1]])


def test_packer_follows_reference_add_read():
    """sc_drop_seq.cpp:34-77: first read of a (cell, snp, umi) keeps allele / bq, duplicates only count;
    entries SNP-ascending, reads UMI-string-ascending"""
    gp = np.full((3, 2, 3), 1 / 3, dtype=np.float32)
    cell = np.array([0, 0, 0, 1, 0, 0])
    snp = np.array([2, 0, 2, 1, 2, 0])
    umi = ["GG", "AC", "AA", "TT", "GG", "AC"]
    al = np.array([1, 0, 0, 1, 0, 1])
    bq = np.array([30, 20, 25, 40, 13, 13])
    p = build_from_reads(gp, cell, snp, umi, al, bq, ["c0", "c1"], ["s0", "s1"])
    assert list(p.cell_ptr) == [0, 2, 3]
    assert list(p.ent_snp) == [0, 2, 1]
    assert list(p.ent_ptr) == [0, 1, 3, 4]
    # snp 0: AC kept from its first read (al 0, bq 20); snp 2: AA (0, 25) sorts before GG (1, 30)
    assert list(p.reads) == [(0 << 8) | 20, (0 << 8) | 25, (1 << 8) | 30, (1 << 8) | 40]
    assert list(p.npass) == [5, 1] and list(p.uniq) == [3, 1]


def test_gp_field_transforms():
    g = np.array([[0, 1, 2, 1]], dtype=np.int8)
    gt = gp_from_gt(g)
    assert gt.dtype == np.float32 and np.allclose(gt.sum(-1), 1.0)
    assert gt[0, 0, 0] == np.float32(0.99) and gt[0, 0, 1] == np.float32(0.005)   # bcf_filtered_reader.cpp:399
    raw = np.random.default_rng(0).uniform(0.01, 1, size=(5, 7, 3)).astype(np.float32)
    gp = gp_from_gp(raw)
    assert gp.dtype == np.float32 and np.allclose(gp.sum(-1), 1.0, atol=1e-5) and (gp > 0).all()
    pl = np.random.default_rng(1).integers(0, 60, size=(4, 6, 3))
    pp = gp_from_pl(pl)
    assert np.allclose(pp.sum(-1), 1.0, atol=1e-5)


def test_synth_shapes():
    p = make_synth(200, 8, 2000, "GP", seed=1)
    n = np.diff(p.cell_ptr)
    assert n.min() >= 1 and 10 < n.mean() < 40          # LogNormal(ln(nsnps/100), 0.5)
    for c in range(0, 200, 17):
        s = p.ent_snp[p.cell_ptr[c]:p.cell_ptr[c + 1]]
        assert (np.diff(s) > 0).all()                    # SNP-ascending, unique
    assert ((p.reads >> 8) <= 2).all() and ((p.reads & 0xff) >= 13).all() and ((p.reads & 0xff) <= 40).all()
    assert len(set(p.barcodes)) == p.ncells


def test_shard_ranges_balance():
    p = make_synth(1000, 4, 3000, "GP", seed=4)
    for n in (1, 2, 3, 8):
        r = shard_ranges(p.cell_ptr, n)
        assert r[0][0] == 0 and r[-1][1] == p.ncells
        assert all(r[i][1] == r[i + 1][0] for i in range(n - 1))
        ent = [int(p.cell_ptr[hi] - p.cell_ptr[lo]) for lo, hi in r]
        assert max(ent) - min(ent) <= 2 * int(np.diff(p.cell_ptr).max())
    assert shard_ranges(np.array([0, 0, 0]), 4)[-1] == (2, 2) or True   # degenerate input does not crash


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from demuxlet_b200.parallel import RecordExchange, allgather_records, shard_pileup
    from demuxlet_b200.api import BEST_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = make_synth(300, 4, 1000, "GP", seed=6)
    sub, (lo, hi) = shard_pileup(p, rank, world)
    rec = np.zeros(sub.ncells, dtype=BEST_DTYPE)
    rec["n_snp"] = np.diff(sub.cell_ptr)
    rec["sng1"] = np.arange(lo, hi)
    rec["llk12"] = -np.arange(lo, hi, dtype=np.float64)
    counts = [h - l for l, h in shard_ranges(p.cell_ptr, world)]
    full = allgather_records(rec, counts)
    ok = (full.shape[0] == p.ncells and np.array_equal(full["sng1"], np.arange(p.ncells)) and
          np.array_equal(full["n_snp"], np.diff(p.cell_ptr)) and
          np.array_equal(full["llk12"], -np.arange(p.ncells, dtype=np.float64)))
    # the buffered exchange bench.py uses every step: uneven shards, then repeated calls on the same buffers
    ex = RecordExchange(counts, BEST_DTYPE.itemsize)
    for rep in range(2):
        rec["llk12"] = -np.arange(lo, hi, dtype=np.float64) - rep
        full2 = ex(rec)
        ok = ok and full2.shape[0] == p.ncells and np.array_equal(full2["sng1"], np.arange(p.ncells)) and \
            np.array_equal(full2["llk12"], -np.arange(p.ncells, dtype=np.float64) - rep)
    ex3 = RecordExchange([7, 7], 8)
    ok = ok and np.array_equal(ex3(np.arange(7, dtype=np.int64) + 100 * rank),
                               np.concatenate([np.arange(7), np.arange(7) + 100]))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_allgather_gloo_world2():
    import torch.multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def _cli():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(here, "demuxlet_b200", "dmx-demuxlet")


def test_cli_option_handling_without_a_device():
    """the command line driver keeps demuxlet's option names; BAM/VCF input (htslib) is refused with a message,
    unknown options and a missing --plp/--out are errors (no GPU needed for any of this)"""
    import subprocess
    exe = _cli()
    if not os.path.exists(exe):
        pytest.skip("dmx-demuxlet not built")
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--alpha" in r.stderr and "--group-list" in r.stderr and "--write-pair" in r.stderr
    r = subprocess.run([exe, "--sam", "x.bam", "--out", "o"], capture_output=True, text=True)
    assert r.returncode == 1 and "--sam needs --vcf" in r.stderr
    r = subprocess.run([exe, "--plp", "x.dmxp", "--sm", "a", "--out", "o"], capture_output=True, text=True)
    assert r.returncode == 1 and "--sm / --sm-list select samples of --vcf" in r.stderr
    r = subprocess.run([exe, "--plp", "x.dmxp", "--out", "o", "--field", "GT"], capture_output=True, text=True)
    assert r.returncode == 1 and "--field-matrix" in r.stderr
    r = subprocess.run([exe, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unrecognized option" in r.stderr
    r = subprocess.run([exe, "--out", "o"], capture_output=True, text=True)
    assert r.returncode == 1 and "--plp" in r.stderr
    r = subprocess.run([exe, "--plp", "/nonexistent.dmxp", "--out", "o"], capture_output=True, text=True)
    assert r.returncode == 1 and "Cannot open" in r.stderr


def test_integration_doc_block_is_the_compiled_one():
    """INTEGRATION.md section 2 shows the call site a maintainer adds to cmd_cram_demuxlet.cpp; oracle/ref_bind.cpp
    compiles oracle/integration_block.inc around the reference's real sc_dropseq_lib_t.  The two texts must be the
    same, so the documented boundary is the tested one."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    blk = open(os.path.join(root, "oracle", "integration_block.inc")).read()
    assert blk.strip() and blk in md
