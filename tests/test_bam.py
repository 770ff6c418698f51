"""dmx_sweep_bam (htslib-free BAM reader + the reference's BAM sweep, SURVEY 8f N4 / BAM half; host only) against the
plain-python restatement in tests/bamio.py on generated BAM files; then the whole command line,
dmx-demuxlet --sam --vcf --field, against the oracle chain on the GPU."""
import os

import numpy as np
import pytest

from demuxlet_b200 import api
from tests import bamio, vcfio

REFS = [("chrM", 3000), ("chr1", 60000), ("chr2", 50000), ("chrUn", 5000)]


def make_snps(rng, n=400):
    chroms = ["chr1", "chr2"]                 # the VCF has no chrM / chrUn: reads there are skipped (:197-200)
    rid = np.sort(rng.integers(0, 2, n)).astype(np.int32)
    pos = np.zeros(n, np.int32)
    for r in (0, 1):
        m = rid == r
        pos[m] = np.sort(rng.choice(np.arange(100, 45000), m.sum(), replace=False))
    ref = rng.choice(np.array([b"A", b"C", b"G", b"T"]), n).astype("S1")
    alt = np.array([rng.choice([b for b in (b"A", b"C", b"G", b"T") if b != x]) for x in ref]).astype("S1")
    return chroms, rid, pos, ref, alt


def same(a, b):
    assert a["barcodes"] == b["barcodes"] and a["counts"] == b["counts"]
    for k in ("totl", "cell", "snp", "allele", "bq"):
        assert np.array_equal(a[k], b[k]), k
    assert a["umis"] == b["umis"]


@pytest.mark.parametrize("seed,opts", [(1, {}), (2, {"min_bq": 0, "cap_bq": 30, "min_td": 5}),
                                       (3, {"min_mq": 0, "exclude_flag": 0x4}), (4, {"tag_group": "", "tag_umi": "RG"}),
                                       (5, {"tag_umi": ""}), (6, {"exclude_flag": 0, "min_mq": 0})])
def test_sweep_matches_restatement(seed, opts, tmp_path):
    rng = np.random.default_rng(seed)
    chroms, rid, pos, ref, alt = make_snps(rng)
    by_tid = {1: pos[rid == 0], 2: pos[rid == 1]}
    reads, barcodes = bamio.random_reads(rng, REFS, by_tid, 3000)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, REFS, reads)
    got = api.sweep_bam(path, chroms, rid, pos, ref, alt, **opts)
    want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt, **opts)
    same(got, want)
    assert len(got["cell"]) > 300 and got["counts"]["off_vcf"] > 0
    assert (got["counts"]["filtered"] > 0) == (opts.get("exclude_flag") != 0)    # seed 6: unmapped reads with CIGARs go through
    if not opts:
        assert "." in got["barcodes"] and "." in got["umis"] and set(got["allele"]) == {0, 1, 2}
        keep = barcodes[:5]
        same(api.sweep_bam(path, chroms, rid, pos, ref, alt, groups=keep),
             bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt, groups=keep))


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_bgzf_blocks_inflate_in_parallel(threads, tmp_path, monkeypatch):
    """a real BGZF file (blocks with the BC size field) is inflated block-parallel, in batches, records walked in file
    order: same tuples as the gzread path on the same records, for any thread count; a corrupt block is an error"""
    monkeypatch.setenv("DMX_BAM_THREADS", threads)
    rng = np.random.default_rng(8)
    chroms, rid, pos, ref, alt = make_snps(rng)
    by_tid = {1: pos[rid == 0], 2: pos[rid == 1]}
    reads, _ = bamio.random_reads(rng, REFS, by_tid, 9000)
    a, b = str(tmp_path / "plain.bam"), str(tmp_path / "bgzf.bam")
    bamio.write_bam(a, REFS, reads)
    bamio.write_bam(b, REFS, reads, member=700, bgzf=True)       # ~1500 blocks: more than one batch of 1024
    want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt)
    same(api.sweep_bam(a, chroms, rid, pos, ref, alt), want)
    same(api.sweep_bam(b, chroms, rid, pos, ref, alt), want)
    raw = bytearray(open(b, "rb").read())
    raw[len(raw) // 2] ^= 0x55
    (tmp_path / "bad.bam").write_bytes(bytes(raw))
    with pytest.raises(api.DmxError):
        api.sweep_bam(str(tmp_path / "bad.bam"), chroms, rid, pos, ref, alt)


def _read(tid, pos, cigar, cb="c1-1", umi="u", base="C"):
    lseq = sum(ln for op, ln in cigar if op in "MIS=X")
    return {"tid": tid, "pos": pos, "mapq": 60, "flag": 0, "cigar": cigar, "seq": base * lseq, "qual": [30] * lseq,
            "tags": [("CB", "Z", cb), ("UB", "Z", umi)]}


def test_sweep_buffer_tests_by_position_alone(tmp_path):
    """the reference tests a read against every SNP in its buffer by position only (cmd_cram_demuxlet.cpp:306-308,
    hts_utils.cpp:279-359 has no chromosome): the look-ahead SNP of the next chromosome and, right after a chromosome
    change, the not yet dropped tail of the previous one give add_read calls for reads at the same coordinate of another
    chromosome; reads with = / X operations find a base for any buffered SNP outside their M / D / N segments"""
    chroms = ["chrM", "chr1", "chr2"]
    #           chrM: none   chr1: 1000, 5000, 5050          chr2: 1020, 5040
    rid = np.array([1, 1, 1, 2, 2], np.int32)
    pos = np.array([1000, 5000, 5050, 1020, 5040], np.int32)
    ref, alt = np.full(5, b"A", "S1"), np.full(5, b"C", "S1")
    reads = [
        _read(0, 990, [("M", 50)], umi="m"),       # chrM (declared, no SNPs): the buffer holds SNP 0 = chr1:1000 -> hit
        _read(1, 980, [("M", 50)], umi="a"),       # chr1:1000 itself
        _read(1, 5030, [("M", 50)], umi="b"),      # chr1:5050, and the look-ahead chr2:1020 is not under it
        _read(1, 5060, [("M", 10)], umi="x"),      # behind the last chr1 SNP: look-ahead = chr2:1020, no overlap
        _read(2, 1010, [("M", 50)], umi="c"),      # chr2:1020
        _read(2, 5035, [("=", 10), ("M", 20)], umi="d"),   # = ignored: chr2:5040 is outside the M walk -> base at offset 20
    ]
    path = str(tmp_path / "q.bam")
    bamio.write_bam(path, [("chrM", 16000), ("chr1", 60000), ("chr2", 50000)], reads)
    got = api.sweep_bam(path, chroms, rid, pos, ref, alt)
    want = bamio.sweep([("chrM", 16000), ("chr1", 60000), ("chr2", 50000)], reads, chroms, rid, pos, ref, alt)
    same(got, want)
    assert list(zip(got["umis"], got["snp"].tolist())) == [("m", 0), ("a", 0), ("b", 2), ("c", 3), ("d", 4)]
    # the tail of the previous chromosome: after a chr1 read far in front of chr1's last SNPs, the first chr2 read drops
    # what was buffered (chr1:1000, chr1:5000), reads ahead through chr1:5050 and is tested against it
    reads2 = [_read(1, 980, [("M", 50)], umi="a"), _read(2, 4990, [("M", 80)], umi="t")]
    bamio.write_bam(path, [("chrM", 16000), ("chr1", 60000), ("chr2", 50000)], reads2)
    got = api.sweep_bam(path, chroms, rid, pos, ref, alt)
    same(got, bamio.sweep([("chrM", 16000), ("chr1", 60000), ("chr2", 50000)], reads2, chroms, rid, pos, ref, alt))
    assert list(zip(got["umis"], got["snp"].tolist())) == [("a", 0), ("t", 2), ("t", 4)]


@pytest.mark.parametrize("threads,chunk", [("1", "1"), ("2", "40"), ("5", "7"), ("16", "300")])
def test_sweep_is_independent_of_the_thread_count(threads, chunk, tmp_path, monkeypatch):
    """records are walked in chunks on all host threads and merged in file order: same tuples, barcode ids and counts
    for any number of threads and any chunk size, on chromosomes whose coordinates overlap (so that chunks hit the
    exact second walk) and with = / X reads in between"""
    monkeypatch.setenv("DMX_BAM_THREADS", threads)
    monkeypatch.setenv("DMX_BAM_MIN_CHUNK", chunk)
    rng = np.random.default_rng(21)
    chroms, rid, pos, ref, alt = make_snps(rng, 1500)
    reads, barcodes = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 7000, ncells=40)
    for r in reads[::97]:
        if r["cigar"] == [("M", 90)]:
            r["cigar"] = [("M", 30), ("X", 1), ("=", 19), ("M", 40)]
    path = str(tmp_path / "t.bam")
    bamio.write_bam(path, REFS, reads, member=900, bgzf=True)
    want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt)
    same(api.sweep_bam(path, chroms, rid, pos, ref, alt), want)
    keep = barcodes[::3]
    same(api.sweep_bam(path, chroms, rid, pos, ref, alt, groups=keep),
         bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt, groups=keep))


@pytest.mark.parametrize("gz,threads", [(False, "1"), (True, "4")])
def test_sam_text_gives_the_tuples_of_its_bam(gz, threads, tmp_path, monkeypatch):
    """--sam takes SAM text as well in the reference (htslib opens either): every line is turned into the BAM record
    sam_parse1 would build and goes through the same walk -- same tuples, barcodes and counts as the BAM of the reads
    (qualities above 93 cannot be written as text and are capped in both files)"""
    monkeypatch.setenv("DMX_BAM_THREADS", threads)
    monkeypatch.setenv("DMX_BAM_MIN_CHUNK", "100")
    rng = np.random.default_rng(33)
    chroms, rid, pos, ref, alt = make_snps(rng)
    reads, barcodes = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 3000)
    for r in reads:
        if r["qual"] is not None:
            r["qual"] = [min(q, 93) for q in r["qual"]]
    pb, ps = str(tmp_path / "a.bam"), str(tmp_path / ("a.sam.gz" if gz else "a.sam"))
    bamio.write_bam(pb, REFS, reads, bgzf=True)
    bamio.write_sam(ps, REFS, reads, gz=gz, final_newline=not gz)
    want = api.sweep_bam(pb, chroms, rid, pos, ref, alt)
    same(api.sweep_bam(ps, chroms, rid, pos, ref, alt), want)
    same(want, bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt))
    keep = barcodes[:4]
    same(api.sweep_bam(ps, chroms, rid, pos, ref, alt, groups=keep, tag_umi=""),
         bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt, groups=keep, tag_umi=""))
    (tmp_path / "bad.sam").write_text("@SQ\tSN:chr1\tLN:100\nr0\t0\tchr1\t5\t60\t3Q\t*\t0\t0\tACG\tIII\n")
    with pytest.raises(api.DmxError, match="malformed SAM"):
        api.sweep_bam(str(tmp_path / "bad.sam"), chroms, rid, pos, ref, alt)


@pytest.mark.parametrize("blocks,chunk", [("1", "7"), ("2", "5000")])
def test_sweep_streams_window_by_window(blocks, chunk, tmp_path, monkeypatch):
    """windows of a few bytes (every header field and record straddles several): same tuples from the gzip, BGZF and
    SAM-text form of the reads"""
    rng = np.random.default_rng(43)
    chroms, rid, pos, ref, alt = make_snps(rng)
    reads, _ = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 400)
    for r in reads:
        if r["qual"] is not None:
            r["qual"] = [min(q, 93) for q in r["qual"]]
    pa, pb, ps = str(tmp_path / "a.bam"), str(tmp_path / "b.bam"), str(tmp_path / "c.sam")
    bamio.write_bam(pa, REFS, reads)
    bamio.write_bam(pb, REFS, reads, member=300, bgzf=True)
    bamio.write_sam(ps, REFS, reads)
    want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt)
    monkeypatch.setenv("DMX_IO_BATCH_BLOCKS", blocks)
    monkeypatch.setenv("DMX_IO_CHUNK_BYTES", chunk)
    for p in (pa, pb, ps):
        same(api.sweep_bam(p, chroms, rid, pos, ref, alt), want)


def test_sam_specification_example(tmp_path):
    """the example alignment of the SAM specification (section 1.1; qualities added, its QUAL column is '*'): a SNP at
    every reference position must find the base the specification's picture shows there --

        Coor     12345678901234  5678901234567890123456789012345
        ref      AGCATGTTAGATAA**GATAGCTGTGCTAGTAGGCAGTCAGCGCCAT
        +r001/1        TTAGATAAAGGATA*CTG
        +r002         aaaAGATAA*GGATA
        +r003       gcctaAGCTAA
        +r004                     ATAGCT..............TCAGC
        -r003                            ttagctTAGGC          (supplementary: dropped by --excl-flag 0x0f04)
        -r001/2                                        CAGCGGCAT

    -- through M / I / D / N / S / H / P operations, from the SAM text and from the BAM of the same reads"""
    ref = "AGCATGTTAGATAAGATAGCTGTGCTAGTAGGCAGTCAGCGCCAT"
    sam = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:ref\tLN:45\n"
           "r001\t99\tref\t7\t30\t8M2I4M1D3M\t=\t37\t39\tTTAGATAAAGGATACTG\t" + "I" * 17 + "\n"
           "r002\t0\tref\t9\t30\t3S6M1P1I4M\t*\t0\t0\tAAAAGATAAGGATA\t" + "I" * 14 + "\n"
           "r003\t0\tref\t9\t30\t5S6M\t*\t0\t0\tGCCTAAGCTAA\t" + "I" * 11 + "\tSA:Z:ref,29,-,6H5M,17,0;\n"
           "r004\t0\tref\t16\t30\t6M14N5M\t*\t0\t0\tATAGCTTCAGC\t" + "I" * 11 + "\n"
           "r003\t2064\tref\t29\t17\t6H5M\t*\t0\t0\tTAGGC\t" + "I" * 5 + "\tSA:Z:ref,9,+,5S6M,30,1;\n"
           "r001\t147\tref\t37\t30\t9M\t=\t7\t-39\tCAGCGGCAT\t" + "I" * 9 + "\tNM:i:1\n")
    ps = tmp_path / "spec.sam"
    ps.write_text(sam)
    n = len(ref)
    rid, pos = np.zeros(n, np.int32), np.arange(n, dtype=np.int32)
    snp_ref = np.array([c.encode() for c in ref], "S1")
    snp_alt = np.array([b"C" if c != "C" else b"G" for c in ref], "S1")      # ref 11 A>C (r003), ref 42 C>G (r001/2)
    # what the picture shows: (read, 1-based reference positions of its aligned bases, those bases)
    shown = [("r001/1", list(range(7, 15)) + list(range(15, 19)) + list(range(20, 23)), "TTAGATAA" "GATA" "CTG"),
             ("r002", list(range(9, 15)) + list(range(15, 19)), "AGATAA" "GATA"),
             ("r003", list(range(9, 15)), "AGCTAA"),
             ("r004", list(range(16, 22)) + list(range(36, 41)), "ATAGCT" "TCAGC"),
             ("r001/2", list(range(37, 46)), "CAGCGGCAT")]
    want = []
    for _, where, bases in shown:
        for p1, b in zip(where, bases):
            want.append((p1 - 1, 0 if b == ref[p1 - 1] else (1 if b.encode() == snp_alt[p1 - 1] else 2), 40))
    assert sum(a for _, a, _ in want) == 2                 # the two mismatches of the example
    got = api.sweep_bam(str(ps), ["ref"], rid, pos, snp_ref, snp_alt)
    assert list(zip(got["snp"].tolist(), got["allele"].tolist(), got["bq"].tolist())) == want
    assert got["barcodes"] == ["."] and got["totl"].tolist() == [5] and got["counts"] == {
        "records": 6, "filtered": 1, "off_vcf": 0, "skip_bcd": 0} and set(got["umis"]) == {"."}
    # the same reads as BAM
    ops = {"r001/1": [("M", 8), ("I", 2), ("M", 4), ("D", 1), ("M", 3)], "r002": [("S", 3), ("M", 6), ("P", 1), ("I", 1), ("M", 4)],
           "r003": [("S", 5), ("M", 6)], "r004": [("M", 6), ("N", 14), ("M", 5)], "r003s": [("H", 6), ("M", 5)], "r001/2": [("M", 9)]}
    rows = [("r001/1", 99, 6, 30, "TTAGATAAAGGATACTG"), ("r002", 0, 8, 30, "AAAAGATAAGGATA"), ("r003", 0, 8, 30, "GCCTAAGCTAA"),
            ("r004", 0, 15, 30, "ATAGCTTCAGC"), ("r003s", 2064, 28, 17, "TAGGC"), ("r001/2", 147, 36, 30, "CAGCGGCAT")]
    reads = [{"tid": 0, "pos": p0, "mapq": mq, "flag": fl, "cigar": ops[nm], "seq": sq, "qual": [40] * len(sq), "tags": []}
             for nm, fl, p0, mq, sq in rows]
    pb = str(tmp_path / "spec.bam")
    bamio.write_bam(pb, [("ref", 45)], reads, bgzf=True)
    same(api.sweep_bam(pb, ["ref"], rid, pos, snp_ref, snp_alt), got)


def test_bgzf_framing_matches_the_specification_constant():
    """the empty BGZF block the test writer emits is, byte for byte, the end-of-file marker printed in the SAM
    specification (section 4.1.2) -- the one fixed BGZF byte string there is: header fields, BC subfield, BSIZE, CRC, ISIZE"""
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    assert bamio.bgzf_block(b"") == eof and len(eof) == 28


def test_sweep_reads_standard_input(tmp_path):
    """--sam - (htslib's name for standard input): a BAM piped in gives the tuples of the file"""
    import subprocess
    import sys
    rng = np.random.default_rng(47)
    chroms, rid, pos, ref, alt = make_snps(rng, 100)
    reads, _ = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 500)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, REFS, reads, bgzf=True)
    np.savez(str(tmp_path / "snps.npz"), rid=rid, pos=pos, ref=ref, alt=alt)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from demuxlet_b200 import api; z = np.load(%r); "
            "s = api.sweep_bam('-', %r, z['rid'], z['pos'], z['ref'], z['alt']); "
            "print(len(s['cell']), int(s['snp'].sum()), int(s['totl'].sum()), len(s['barcodes']))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "snps.npz"), chroms))
    with open(path, "rb") as f:
        r = subprocess.run([sys.executable, "-c", code], stdin=f, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    w = api.sweep_bam(path, chroms, rid, pos, ref, alt)
    assert r.stdout.split() == [str(len(w["cell"])), str(int(w["snp"].sum())), str(int(w["totl"].sum())), str(len(w["barcodes"]))]


@pytest.mark.parametrize("level,member", [(0, 5000), (1, 60000), (9, 60000), (6, 40)])
def test_block_decoder_on_every_kind_of_deflate_block(level, member, tmp_path, monkeypatch):
    """BGZF blocks go through the block decoder of csrc/dmx_inflate.h first and through zlib when it cannot vouch for
    its output (length, end of stream, CRC-32): stored blocks (level 0), fixed-Huffman blocks (tiny payloads), dynamic
    ones at the fastest and the best setting -- the same tuples with the decoder and with zlib alone (DMX_IO_ZLIB)"""
    rng = np.random.default_rng(51)
    chroms, rid, pos, ref, alt = make_snps(rng)
    reads, _ = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 1500)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, REFS, reads, member=member, bgzf=True, level=level)
    want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt)
    same(api.sweep_bam(path, chroms, rid, pos, ref, alt), want)
    monkeypatch.setenv("DMX_IO_ZLIB", "1")
    same(api.sweep_bam(path, chroms, rid, pos, ref, alt), want)


def test_sweep_feeds_the_packer(tmp_path):
    """sweep tuples -> dmx_pack_reads == the python container restatement on the same tuples"""
    from demuxlet_b200.plp import build_from_reads
    rng = np.random.default_rng(5)
    chroms, rid, pos, ref, alt = make_snps(rng)
    reads, _ = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 4000)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, REFS, reads)
    s = api.sweep_bam(path, chroms, rid, pos, ref, alt)
    packed = api.pack_reads(s["cell"], s["snp"], s["umis"], s["allele"], s["bq"], len(s["barcodes"]))
    gp = np.full((len(pos), 2, 3), 1 / 3, np.float32)
    want = build_from_reads(gp, s["cell"], s["snp"], s["umis"], s["allele"], s["bq"], s["barcodes"], ["a", "b"], s["totl"])
    for k in ("cell_ptr", "ent_snp", "ent_ptr", "reads", "npass", "uniq"):
        assert np.array_equal(packed[k], getattr(want, k)), k
    assert packed["npass"].sum() == len(s["cell"]) and (s["totl"] >= packed["npass"]).all()


def test_sweep_errors(tmp_path):
    rng = np.random.default_rng(6)
    chroms, rid, pos, ref, alt = make_snps(rng, 20)
    with pytest.raises(api.DmxError):
        api.sweep_bam(str(tmp_path / "none.bam"), chroms, rid, pos, ref, alt)
    p = tmp_path / "bad.bam"
    p.write_bytes(b"not a bam")
    with pytest.raises(api.DmxError, match="not a BAM"):
        api.sweep_bam(str(p), chroms, rid, pos, ref, alt)
    path = str(tmp_path / "a.bam")
    bamio.write_bam(path, REFS, [])
    with pytest.raises(api.DmxError, match="sorted"):
        api.sweep_bam(path, chroms, rid[::-1].copy(), pos, ref, alt)
    with pytest.raises(api.DmxError, match="orderings"):
        api.sweep_bam(path, chroms[::-1], rid, pos, ref, alt)
    assert api.sweep_bam(path, chroms, rid, pos, ref, alt)["barcodes"] == []


def make_case(tmp_path, seed, field, nreads=6000, nv=5):
    """a consistent VCF + BAM pair: records on chr1 / chr2, reads carrying the alleles of one donor per barcode"""
    rng = np.random.default_rng(seed)
    samples = ["donor%d" % i for i in range(nv)]
    recs = vcfio.random_records(rng, 500, nv, field)
    recs = [(("chr1" if i < 250 else "chr2"), 200 + 37 * (i % 250), r[2], r[3], r[4], r[5]) for i, r in enumerate(recs)]
    vcf = str(tmp_path / "a.vcf.gz")
    # chrM is declared in the VCF header but carries no record: its reads still count for RD.TOTL (the reference maps
    # read chromosomes through the header dictionary, cmd_cram_demuxlet.cpp:197-199, :295); chrUn is not declared at all
    vcfio.write_vcf(vcf, samples, recs, bgzip_like=True, contigs=["chrM", "chr1", "chr2"])
    v = vcfio.parse_vcf(vcf, field)
    assert v["chroms"] == ["chrM", "chr1", "chr2"]
    by_tid = {1: v["pos"][v["rid"] == 1], 2: v["pos"][v["rid"] == 2]}
    reads, barcodes = bamio.random_reads(rng, REFS, by_tid, nreads, ncells=20)
    bam = str(tmp_path / "a.bam")
    bamio.write_bam(bam, REFS, reads)
    return vcf, bam, v, reads, samples, barcodes


@pytest.mark.gpu
@pytest.mark.parametrize("field,seed", [("GT", 11), ("PL", 12)])
def test_cli_sam_vcf_end_to_end(field, seed, tmp_path):
    """dmx-demuxlet --sam BAM --vcf VCF --field F --out P, demuxlet's own command line, against the oracle chain
    (python VCF / BAM restatements -> oracle filter and matrix -> python add_read container -> oracle likelihoods and
    writers): identity columns byte-identical, LLK columns within 1e-4"""
    import subprocess
    from demuxlet_b200.plp import build_from_reads
    from oracle import pyoracle as po
    from tests.util import diff_tsv
    from tests.test_gpu_parity import BEST_EXACT, ROW_EXACT
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "demuxlet_b200", "dmx-demuxlet")
    vcf, bam, v, reads, samples, barcodes = make_case(tmp_path, seed, field)
    keep = po.site_filter(v["gt"], 1, 0.5)
    vals = np.ascontiguousarray(v["values"][keep])
    gp = po.build_gp(field, vals, 0.01)
    groups = barcodes[:15] + ["."]
    s = bamio.sweep(REFS, reads, v["chroms"], v["rid"][keep], v["pos"][keep], v["ref"][keep], v["alt"][keep], groups=groups)
    plp = build_from_reads(gp, s["cell"], s["snp"], s["umis"], s["allele"], s["bq"], s["barcodes"], samples, s["totl"])
    po.Oracle(plp, (0.0, 0.5)).write_all(str(tmp_path / "orc"))
    (tmp_path / "groups.txt").write_text("\n".join(groups) + "\n")
    r = subprocess.run([exe, "--sam", bam, "--vcf", vcf, "--field", field, "--out", str(tmp_path / "cli"), "--fp64-pairs",
                        "--group-list", str(tmp_path / "groups.txt"), "--verbose"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "observations" in r.stderr
    diff_tsv(tmp_path / "orc.best", tmp_path / "cli.best", BEST_EXACT)
    for ext in ("single", "sing2"):
        diff_tsv(tmp_path / ("orc." + ext), tmp_path / ("cli." + ext), ROW_EXACT)
    assert len(open(tmp_path / "cli.best").read().splitlines()) > 10
    # reads on the declared, record-free contig are in RD.TOTL; without the ##contig dictionary they would be dropped
    no_m = bamio.sweep(REFS, reads, ["chr1", "chr2"], v["rid"][keep] - 1, v["pos"][keep], v["ref"][keep], v["alt"][keep],
                       groups=groups)
    assert s["counts"]["off_vcf"] < no_m["counts"]["off_vcf"] and int(np.sum(s["totl"])) > int(np.sum(no_m["totl"]))


def _reference_lookups(tmp_path):
    """(reads, queries, results) of the CIGAR look-up cases: results from the committed golden file written by the
    reference's function, cross-checked against the live function where oracle/_ref/ref_cigar is built"""
    reads, queries = bamio.cigar_cases(np.random.default_rng(31))
    flat = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cigar_lookup.npz"))["results"]
    res, k = [], 0
    for qs in queries:
        res.append([(chr(int(b)), int(q), int(rp)) for b, q, rp in flat[k:k + len(qs)]])
        k += len(qs)
    assert k == len(flat)
    if os.path.exists(bamio.REF_CIGAR):
        assert bamio.run_ref_cigar(reads, queries, str(tmp_path)) == res
    return reads, queries, res


def test_cigar_lookup_restatement_against_reference_function(tmp_path):
    reads, queries, res = _reference_lookups(tmp_path)
    n_valid = 0
    for r, qs, row in zip(reads, queries, res):
        for q, want in zip(qs, row):
            got = bamio.lookup(r, q)
            assert got[2] == want[2], (r["cigar"], r["pos"], q, got, want)
            if want[2] >= 0:                      # base / qual are only read when the position is found (:304-306)
                assert got == want, (r["cigar"], r["pos"], q, got, want)
                n_valid += 1
    assert n_valid > 5000


def test_sweep_lookup_against_reference_function(tmp_path):
    """dmx_sweep_bam on the same reads with a SNP at every position: the emitted tuples are exactly the reference
    function's results put through the thresholds of cmd_cram_demuxlet.cpp:304-325, for every SNP the reference's buffer
    holds when the read is tested (:202-237: from one base before the read start to the furthest read end so far) --
    which, for reads with = / X operations (ignored by the reference's walk), includes bases "found" outside the read"""
    reads, queries, res = _reference_lookups(tmp_path)
    L = 700
    rid, pos = np.zeros(L, np.int32), np.arange(L, dtype=np.int32)
    ref, alt = np.full(L, b"A", "S1"), np.full(L, b"C", "S1")
    path = str(tmp_path / "c.bam")
    bamio.write_bam(path, [("chr1", 5000)], reads)
    got = api.sweep_bam(path, ["chr1"], rid, pos, ref, alt)
    want = []
    cells = {}
    s, hi, outside = 0, 0, 0
    for r, qs, row in zip(reads, queries, res):
        cb = dict((t, v) for t, _, v in r["tags"])["CB"]
        ic = cells.setdefault(cb, len(cells))
        span = sum(ln for op, ln in r["cigar"] if op in "MDN=X")
        s = max(s, min(max(r["pos"] - 1, 0), hi + 1))            # SNP index == position here
        hi = max(hi, min(r["pos"] + max(span, 1), L - 1))
        by_q = dict(zip(qs, row))
        for q in range(s, hi + 1):
            # beyond the positions the reference function was asked about, its restatement (checked against it above)
            base, qual, rpos = by_q[q] if q in by_q else bamio.lookup(r, q)
            if rpos < 0 or base == "N" or qual - 33 < 13:
                continue
            if not (r["pos"] <= q < r["pos"] + max(span, 1)):
                assert any(op in "=X" for op, _ in r["cigar"])
                outside += 1
            want.append((ic, q, 0 if base == "A" else (1 if base == "C" else 2), min(qual - 33, 40)))
    have = list(zip(got["cell"].tolist(), got["snp"].tolist(), got["allele"].tolist(), got["bq"].tolist()))
    assert have == want and len(want) > 3000 and outside > 100
    assert same(got, bamio.sweep([("chr1", 5000)], reads, ["chr1"], rid, pos, ref, alt)) is None


def test_random_umis_are_glibc_rand():
    """--tag-UMI "": the reference appends hex(rand()) to "." per read and never seeds rand() (cmd_cram_demuxlet.cpp:276):
    the python restatement of glibc's generator gives the numbers the C library gives after srand(1)"""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)
    want = [libc.rand() for _ in range(2000)]
    rnd = bamio.GlibcRand()
    assert [rnd() for _ in range(2000)] == want and want[0] == 1804289383
