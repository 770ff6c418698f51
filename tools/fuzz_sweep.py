"""dmx_sweep_bam against the literal python restatement of the reference's loop (tests/bamio.py) on random SNP sets, reads,\noptions, thread counts and chunk sizes.  usage: python tools/fuzz_sweep.py SEED0 SEED1   (no GPU needed)"""
import os, sys, numpy as np, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demuxlet_b200 import api
from tests import bamio
from tests.test_bam import REFS, same
td = tempfile.mkdtemp()
nbad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(5, 400))
    chroms = ["chrM", "chr1", "chr2", "chrUn"][: int(rng.integers(2, 5))] if rng.random() < 0.5 else ["chr1", "chr2"]
    nch = len(chroms)
    rid = np.sort(rng.integers(0, nch, n)).astype(np.int32)
    if rng.random() < 0.3:
        rid[rid == 0] = min(1, nch - 1)          # a declared contig without SNPs
        rid = np.sort(rid)
    pos = np.zeros(n, np.int32)
    span = int(rng.choice([300, 2000, 45000]))
    for r in range(nch):
        m = rid == r
        if m.sum():
            pos[m] = np.sort(rng.choice(np.arange(0, span), m.sum(), replace=m.sum() > span))
    ref = rng.choice(np.array([b"A", b"C", b"G", b"T"]), n).astype("S1")
    alt = rng.choice(np.array([b"A", b"C", b"G", b"T"]), n).astype("S1")
    name2tid = {nm: i for i, (nm, _) in enumerate(REFS)}
    by_tid = {name2tid[c]: pos[rid == i] for i, c in enumerate(chroms)}
    reads, barcodes = bamio.random_reads(rng, REFS, by_tid, int(rng.integers(1, 1500)), ncells=int(rng.integers(1, 30)))
    for r in reads:
        if span < 45000:
            r["pos"] = int(r["pos"] % (span + 100))
        if rng.random() < 0.05 and r["cigar"] == [("M", 90)]:
            r["cigar"] = [("M", 30), ("X", 1), ("=", 19), ("M", 40)]
    reads.sort(key=lambda r: (r["tid"], r["pos"]))
    path = os.path.join(td, "f.bam")
    bamio.write_bam(path, REFS, reads, member=int(rng.integers(200, 5000)), bgzf=bool(rng.integers(0, 2)))
    opts = {}
    if rng.random() < 0.3: opts["min_td"] = int(rng.integers(0, 10))
    if rng.random() < 0.2: opts["tag_umi"] = ""
    if rng.random() < 0.2: opts["tag_group"] = ""
    if rng.random() < 0.3: opts.update(exclude_flag=int(rng.choice([0, 0x100, 0x0f00])), min_mq=0)   # unmapped reads go through
    groups = barcodes[::2] if rng.random() < 0.3 else None
    os.environ["DMX_BAM_THREADS"] = str(int(rng.choice([1, 2, 3, 8])))
    os.environ["DMX_BAM_MIN_CHUNK"] = str(int(rng.choice([1, 5, 50, 2048])))
    want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt, groups=groups, **opts)
    try:
        got = api.sweep_bam(path, chroms, rid, pos, ref, alt, groups=groups, **opts)
        same(got, want)
    except Exception as e:
        nbad += 1
        print("seed", seed, "FAIL", type(e).__name__, str(e)[:200])
print("done", nbad, "failures")
