"""dmx_vcf_read (htslib-free VCF text reader, SURVEY 8f N4 / VCF half; host only) against the plain-python restatement
in tests/vcfio.py on generated files: plain, gzip and multi-member (bgzip-like) gzip, the three fields, missing and
short values, multi-allelic and ALT-less records; then the chain VCF -> dmx_site_filter -> dmx_build_gp on the GPU
against the oracle chain."""
import numpy as np
import pytest

from demuxlet_b200 import api
from oracle import pyoracle as po
from tests import vcfio


def same(a, b, field):
    assert a["samples"] == b["samples"] and a["chroms"] == b["chroms"] and a["n_skipped"] == b["n_skipped"]
    for k in ("rid", "pos", "ref", "alt", "gt"):
        assert np.array_equal(a[k], b[k]), k
    if field == "GP":      # missing values are NaNs on both sides; the rest are the same floats
        assert np.array_equal(np.isnan(a["values"]), np.isnan(b["values"]))
        assert np.array_equal(np.nan_to_num(a["values"], nan=-1.0), np.nan_to_num(b["values"], nan=-1.0))
    else:
        assert np.array_equal(a["values"], b["values"])


@pytest.mark.parametrize("field", ["GT", "GP", "PL"])
@pytest.mark.parametrize("ext,bg", [(".vcf", False), (".vcf.gz", False), (".vcf.gz", True)])
def test_reader_matches_restatement(field, ext, bg, tmp_path):
    rng = np.random.default_rng(7)
    samples = ["S%02d" % i for i in range(9)]
    recs = vcfio.random_records(rng, 400, len(samples), field)
    path = str(tmp_path / ("a" + ext))
    vcfio.write_vcf(path, samples, recs, bgzip_like=bg)
    got, want = api.read_vcf(path, field), vcfio.parse_vcf(path, field)
    same(got, want, field)
    assert got["n_skipped"] > 0 and len(got["pos"]) > 300 and len(got["chroms"]) == 3


@pytest.mark.parametrize("threads", ["1", "3", "16"])
def test_bgzf_vcf_any_thread_count(threads, tmp_path, monkeypatch):
    """a bgzipped VCF (BGZF blocks with the BC size field) is inflated block-parallel and its lines are parsed in
    contiguous ranges on the host threads, straight into the arrays of the whole file: same records for any thread
    count, skipped records closed up, chromosome ids by first appearance; python's gzip module reads the same text"""
    monkeypatch.setenv("DMX_VCF_THREADS", threads)
    rng = np.random.default_rng(17)
    samples = ["S%02d" % i for i in range(5)]
    recs = vcfio.random_records(rng, 1500, len(samples), "GP")
    path = str(tmp_path / "a.vcf.gz")
    vcfio.write_vcf(path, samples, recs, bgzf=True, contigs=("chrM", "chr2"))
    got, want = api.read_vcf(path, "GP"), vcfio.parse_vcf(path, "GP")
    same(got, want, "GP")
    assert got["chroms"] == ["chrM", "chr2", "chr1", "chr3"] and got["n_skipped"] > 50
    raw = bytearray(open(path, "rb").read())
    raw[len(raw) // 2] ^= 0x55
    (tmp_path / "bad.vcf.gz").write_bytes(bytes(raw))
    with pytest.raises(api.DmxError):
        api.read_vcf(str(tmp_path / "bad.vcf.gz"), "GP")


def test_any_other_field_name_is_read_like_gp(tmp_path):
    """the reference reads any --field other than GT and PL as posteriors like GP (parse_posteriors' last branch,
    bcf_filtered_reader.cpp:408): dmx_vcf_read_key with the key named"""
    rng = np.random.default_rng(23)
    samples = ["a", "b", "c"]
    recs = vcfio.random_records(rng, 200, len(samples), "GP")
    renamed = [(c, p, r, a, [("POST" if k == "GP" else k) for k in keys], [{("POST" if k == "GP" else k): x for k, x in col.items()} for col in cols])
               for c, p, r, a, keys, cols in recs]
    pa, pb = str(tmp_path / "a.vcf"), str(tmp_path / "b.vcf")
    vcfio.write_vcf(pa, samples, recs)
    vcfio.write_vcf(pb, samples, renamed)
    a, b = api.read_vcf(pa, "GP"), api.read_vcf(pb, "GP", key="POST")
    assert np.array_equal(a["values"].view(np.uint32), b["values"].view(np.uint32)) and not b["no_field"].any()
    assert api.read_vcf(pb, "GP")["no_field"].all()


@pytest.mark.parametrize("field", ["GT", "GP", "PL"])
@pytest.mark.parametrize("int_type,idx", [(0, False), (3, True)])
def test_bcf_gives_the_arrays_of_its_vcf_text(field, int_type, idx, tmp_path):
    """--vcf takes VCF or BCF in the reference (htslib opens either): the BCF2 reader delivers, record by record, the
    arrays the text reader delivers for the same records -- GT with missing / haploid / phased calls, GP / PL with
    missing values and short vectors, int8 / int16 / int32 vectors, multi-allelic and ALT-less records skipped,
    dictionary numbering by appearance or by IDX=.  (The writer is tests/vcfio.py:write_bcf, from the specification:
    unpinned against htslib, like the text reader.)"""
    rng = np.random.default_rng(29)
    samples = ["S%02d" % i for i in range(7)]
    recs = vcfio.random_records(rng, 700, len(samples), field)
    contigs = ["chrM", "chr1", "chr2", "chr3"]
    pt, pb = str(tmp_path / "a.vcf"), str(tmp_path / "a.bcf")
    vcfio.write_vcf(pt, samples, recs, contigs=contigs)
    vcfio.write_bcf(pb, samples, recs, contigs, int_type=int_type, idx=idx)
    a, b = api.read_vcf(pt, field), api.read_vcf(pb, field)
    same(b, a, field)
    assert a["chroms"] == b["chroms"] == contigs and b["n_skipped"] == a["n_skipped"] > 0 and len(b["pos"]) > 500
    assert not b["no_field"].any()
    if field != "GT":      # the other field is not in these records: flagged, values missing
        other = "PL" if field == "GP" else "GP"
        c = api.read_vcf(pb, other)
        assert c["no_field"].all() and np.array_equal(c["gt"], a["gt"])
    raw = open(pb, "rb").read()
    (tmp_path / "cut.bcf").write_bytes(raw[:len(raw) // 2])
    with pytest.raises(api.DmxError):
        api.read_vcf(str(tmp_path / "cut.bcf"), field)


@pytest.mark.parametrize("blocks,chunk", [("1", "7"), ("3", "1000")])
def test_reader_streams_window_by_window(blocks, chunk, tmp_path, monkeypatch):
    """the file is read window by window (a batch of BGZF blocks / a chunk of any other stream behind the unconsumed tail
    of the last one) and parsed as it comes: windows of a few bytes -- every line, BCF header and BCF record straddles
    several -- give the records of the one-window read, in text, gzip, BGZF and BCF form"""
    rng = np.random.default_rng(41)
    samples = ["S%d" % i for i in range(4)]
    recs = vcfio.random_records(rng, 150, len(samples), "PL")
    contigs = ["chr1", "chr2", "chr3"]
    paths = [str(tmp_path / n) for n in ("a.vcf", "b.vcf.gz", "c.vcf.gz", "d.bcf")]
    vcfio.write_vcf(paths[0], samples, recs, contigs=contigs)
    vcfio.write_vcf(paths[1], samples, recs, contigs=contigs)
    vcfio.write_vcf(paths[2], samples, recs, contigs=contigs, bgzf=True)
    vcfio.write_bcf(paths[3], samples, recs, contigs)
    want = [api.read_vcf(p, "PL") for p in paths]
    monkeypatch.setenv("DMX_IO_BATCH_BLOCKS", blocks)
    monkeypatch.setenv("DMX_IO_CHUNK_BYTES", chunk)
    for p, w in zip(paths, want):
        same(api.read_vcf(p, "PL"), w, "PL")
        same(w, want[0], "PL")


def test_vcf_specification_example(tmp_path):
    """the example file of the VCF specification (section 1.1): two biallelic records are kept, the two multi-allelic
    ones and the monomorphic one are skipped; genotypes, phased or not, positions and alleles as printed there; a
    sample column that drops its trailing subfields (NA00003 of the second record) reads as missing"""
    text = """##fileformat=VCFv4.3
##fileDate=20090805
##source=myImputationProgramV3.1
##reference=file:///seq/references/1000GenomesPilot-NCBI36.fasta
##contig=<ID=20,length=62435964,assembly=B36,md5=f126cdf8a6e0c7f379d618ff66beb2da,species="Homo sapiens",taxonomy=x>
##phasing=partial
##INFO=<ID=NS,Number=1,Type=Integer,Description="Number of Samples With Data">
##INFO=<ID=DP,Number=1,Type=Integer,Description="Total Depth">
##INFO=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">
##INFO=<ID=AA,Number=1,Type=String,Description="Ancestral Allele">
##INFO=<ID=DB,Number=0,Type=Flag,Description="dbSNP membership, build 129">
##INFO=<ID=H2,Number=0,Type=Flag,Description="HapMap2 membership">
##FILTER=<ID=q10,Description="Quality below 10">
##FILTER=<ID=s50,Description="Less than 50% of samples have data">
##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">
##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read Depth">
##FORMAT=<ID=HQ,Number=2,Type=Integer,Description="Haplotype Quality">
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tNA00001\tNA00002\tNA00003
20\t14370\trs6054257\tG\tA\t29\tPASS\tNS=3;DP=14;AF=0.5;DB;H2\tGT:GQ:DP:HQ\t0|0:48:1:51,51\t1|0:48:8:51,51\t1/1:43:5:.,.
20\t17330\t.\tT\tA\t3\tq10\tNS=3;DP=11;AF=0.017\tGT:GQ:DP:HQ\t0|0:49:3:58,50\t0|1:3:5:65,3\t0/0:41:3
20\t1110696\trs6040355\tA\tG,T\t67\tPASS\tNS=2;DP=10;AF=0.333,0.667;AA=T;DB\tGT:GQ:DP:HQ\t1|2:21:6:23,27\t2|1:2:0:18,2\t2/2:35:4
20\t1230237\t.\tT\t.\t47\tPASS\tNS=3;DP=13;AA=T\tGT:GQ:DP:HQ\t0|0:54:7:56,60\t0|0:48:4:51,51\t0/0:61:2
20\t1234567\tmicrosat1\tGTC\tG,GTCT\t50\tPASS\tNS=3;DP=9;AA=G\tGT:GQ:DP\t0/1:35:4\t0/2:17:2\t1/1:40:3
"""
    p = tmp_path / "spec.vcf"
    p.write_text(text)
    v = api.read_vcf(str(p), "GT")
    assert v["samples"] == ["NA00001", "NA00002", "NA00003"] and v["chroms"] == ["20"] and v["n_skipped"] == 3
    assert v["pos"].tolist() == [14369, 17329] and v["ref"].tolist() == [b"G", b"T"] and v["alt"].tolist() == [b"A", b"A"]
    assert v["gt"].tolist() == [[[0, 0], [1, 0], [1, 1]], [[0, 0], [0, 1], [0, 0]]]
    # HQ read the way the reference reads any other --field: float posteriors (two values and a missing third)
    h = api.read_vcf(str(p), "GP", key="HQ")
    nan = float("nan")
    want = np.array([[[51, 51, nan], [51, 51, nan], [nan, nan, nan]], [[58, 50, nan], [65, 3, nan], [nan, nan, nan]]], np.float32)
    assert not h["no_field"].any() and np.array_equal(np.nan_to_num(h["values"], nan=-1), np.nan_to_num(want, nan=-1))
    # the reference's filter on these genotypes (an = 6; record 1: ac = 3, record 2: ac = 1): --min-mac 1 keeps both, 2 only the first
    assert po.site_filter(v["gt"], 1, 0.5).tolist() == [True, True] and po.site_filter(v["gt"], 2, 0.5).tolist() == [True, False]


def test_format_floats_are_strtod_narrowed(tmp_path):
    """GP text -> float as htslib does it, (float)strtod(text): the reader's one-operation conversion of plain decimals
    (mantissa < 2^53, |decimal exponent| <= 22) and its strtod fall-back give that bit pattern for every spelling"""
    rng = np.random.default_rng(12)
    texts = ["0", "-0", "1", "0.1", "0.3", ".5", "5.", "1e-3", "1E5", "2.5e+3", "1e22", "1e23", "1e-22", "1e-23", "1e-45", "3e-46",
             "1e39", "-1e39", "inf", "-inf", "nan", "0.30000000000000004", "0.1234567890123456789", "123456789012345678",
             "999999999999999", "9999999999999999", "16777217", "16777216.5", "0.000001", "1e", "1.5e+", "+4.25",
             "3.4028235e38", "3.4028236e38", "1.17549435e-38", "7.006492321624085e-46", "8.5e-46", "1.00000006", "1.00000005960464"]
    for _ in range(3000):
        kind = rng.integers(0, 4)
        if kind == 0:
            texts.append("%.*f" % (int(rng.integers(0, 12)), rng.random() * 10 ** int(rng.integers(-3, 6))))
        elif kind == 1:
            texts.append("%.*g" % (int(rng.integers(1, 20)), rng.random() * 10.0 ** int(rng.integers(-30, 30))))
        elif kind == 2:
            texts.append("%.*e" % (int(rng.integers(0, 18)), rng.standard_normal() * 10.0 ** int(rng.integers(-45, 39))))
        else:       # halfway cases of the float grid written with many digits
            f = np.float32(rng.random() * 10.0 ** int(rng.integers(-10, 10)))
            mid = (float(f) + float(np.nextafter(f, np.float32(np.inf)))) / 2
            texts.append(repr(mid))
    while len(texts) % 3:
        texts.append("0.25")
    recs = [("1", 10 + i, "A", "C", ["GT", "GP"], [{"GT": "0/1", "GP": ",".join(texts[3 * i:3 * i + 3])}])
            for i in range(len(texts) // 3)]
    path = str(tmp_path / "f.vcf")
    vcfio.write_vcf(path, ["s"], recs)
    got = api.read_vcf(path, "GP")["values"].reshape(-1)
    with np.errstate(over="ignore"):
        want = np.array([np.float32(float(t)) if t not in ("1e", "1.5e+") else np.float32(float(t.split("e")[0])) for t in texts], np.float32)
    bad = [(t, g, w) for t, g, w in zip(texts, got, want) if g.view(np.uint32) != w.view(np.uint32) and not (np.isnan(g) and np.isnan(w))]
    assert not bad, bad[:5]


def test_sample_subset(tmp_path):
    """--sm / --sm-list: columns of the named samples in the byte order of their names (the reference's std::set)"""
    rng = np.random.default_rng(10)
    samples = ["zeta", "Alpha", "mid", "beta", "S1", "s1"]
    path = str(tmp_path / "a.vcf")
    vcfio.write_vcf(path, samples, vcfio.random_records(rng, 120, len(samples), "PL"))
    full = api.read_vcf(path, "PL")
    pick = ["mid", "zeta", "Alpha", "mid", "S1"]
    got = api.read_vcf(path, "PL", samples=pick)
    order = sorted(set(pick))                       # byte order: "Alpha" < "S1" < "mid" < "zeta"
    assert got["samples"] == order == ["Alpha", "S1", "mid", "zeta"]
    cols = [samples.index(x) for x in order]
    assert np.array_equal(got["gt"], full["gt"][:, cols]) and np.array_equal(got["values"], full["values"][:, cols])
    with pytest.raises(api.DmxError, match="Cannot find sample ID nobody"):
        api.read_vcf(path, "PL", samples=["mid", "nobody"])


def test_reader_errors(tmp_path):
    with pytest.raises(api.DmxError):
        api.read_vcf(str(tmp_path / "missing.vcf"))
    p = tmp_path / "nohdr.vcf"
    p.write_text("chr1\t10\t.\tA\tT\t.\t.\t.\tGT\t0/1\n")
    with pytest.raises(api.DmxError, match="#CHROM"):
        api.read_vcf(str(p))
    p = tmp_path / "nogp.vcf"
    vcfio.write_vcf(str(p), ["a", "b"], [("1", 5, "A", "C", ["GT"], [{"GT": "0/1"}, {"GT": "1/1"}])])
    assert len(api.read_vcf(str(p), "GT")["pos"]) == 1
    # a record without the requested key is flagged, not refused: the reference fails on it only if it passes the
    # genotype filter (parse_posteriors comes after passed_vfilter, cmd_cram_demuxlet.cpp:153,:211)
    vcfio.write_vcf(str(p), ["a", "b"], [("1", 5, "A", "C", ["GT"], [{"GT": "0/0"}, {"GT": "0/0"}]),
                                         ("1", 9, "A", "C", ["GT", "GP"], [{"GT": "0/1", "GP": "0,1,0"}, {"GT": "1/1", "GP": "0,0,1"}]),
                                         ("1", 12, "A", "C", ["GT"], [{"GT": "0/1"}, {"GT": "1/1"}])])
    v = api.read_vcf(str(p), "GP")
    assert v["no_field"].tolist() == [1, 0, 1] and np.isnan(v["values"][[0, 2]]).all() and not np.isnan(v["values"][1]).any()
    assert api.read_vcf(str(p), "GT")["no_field"].tolist() == [0, 0, 0]
    p = tmp_path / "cols.vcf"
    p.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\n1\t5\t.\tA\tC\t.\t.\t.\tGT\t0/1\n")
    with pytest.raises(api.DmxError, match="columns"):
        api.read_vcf(str(p))


@pytest.mark.gpu
@pytest.mark.parametrize("field", ["GT", "GP", "PL"])
def test_vcf_to_matrix_chain(field, tmp_path):
    """VCF text -> record filter -> genotype-probability matrix, device chain against the oracle chain"""
    rng = np.random.default_rng(8)
    samples = ["S%02d" % i for i in range(12)]
    path = str(tmp_path / "a.vcf.gz")
    vcfio.write_vcf(path, samples, vcfio.random_records(rng, 600, len(samples), field), bgzip_like=True)
    v = api.read_vcf(path, field)
    assert np.array_equal(api.site_filter(v["gt"], 1, 0.5), po.site_filter(v["gt"], 1, 0.5))
    keep = api.site_filter(v["gt"], 4, 0.9)
    assert np.array_equal(keep, po.site_filter(v["gt"], 4, 0.9)) and 0 < keep.sum() < len(keep)
    vals = np.ascontiguousarray(v["values"][keep])
    if field == "GP":       # a record with a missing GP row yields NaNs: compare the finite records
        ok = ~np.isnan(vals).any(axis=(1, 2))
        vals = np.ascontiguousarray(vals[ok])
    got = api.build_gp(field, vals, 0.01)
    want = po.build_gp(field, vals, 0.01)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("field", ["GT", "PL"])
def test_cli_vcf(field, tmp_path):
    """dmx-demuxlet --plp P --vcf V --field F: VCF text -> record filter -> matrix on the GPU -> the files --plp alone
    writes when the pileup file carries the oracle's matrix of the same records; sample names come from the VCF"""
    import dataclasses
    import filecmp
    import os
    import subprocess
    from demuxlet_b200 import synth
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "demuxlet_b200", "dmx-demuxlet")
    rng = np.random.default_rng(9)
    samples = ["donor%d" % i for i in range(6)]
    path = str(tmp_path / "a.vcf.gz")
    vcfio.write_vcf(path, samples, vcfio.random_records(rng, 900, len(samples), field), bgzip_like=True)
    v = vcfio.parse_vcf(path, field)
    keep = po.site_filter(v["gt"], 2, 0.6)
    gp = po.build_gp(field, np.ascontiguousarray(v["values"][keep]), 0.03)
    p = synth.make_synth(25, len(samples), int(keep.sum()), "GP", seed=5)
    dataclasses.replace(p, gp=gp, samples=samples).save(str(tmp_path / "a.dmxp"))
    dataclasses.replace(p, gp=np.zeros_like(gp)).save(str(tmp_path / "b.dmxp"))
    ra = subprocess.run([exe, "--plp", str(tmp_path / "a.dmxp"), "--out", str(tmp_path / "a")], capture_output=True, text=True)
    rb = subprocess.run([exe, "--plp", str(tmp_path / "b.dmxp"), "--out", str(tmp_path / "b"), "--vcf", path, "--field", field,
                         "--geno-error", "0.03", "--min-mac", "2", "--min-callrate", "0.6", "--verbose"],
                        capture_output=True, text=True)
    assert ra.returncode == 0 and rb.returncode == 0, ra.stderr + rb.stderr
    assert "pass --min-mac 2" in rb.stderr
    for ext in ("best", "single", "sing2"):
        assert filecmp.cmp(str(tmp_path / ("a." + ext)), str(tmp_path / ("b." + ext)), shallow=False), ext
    assert "donor3" in open(tmp_path / "b.single").read()
    bad = subprocess.run([exe, "--plp", str(tmp_path / "b.dmxp"), "--out", str(tmp_path / "c"), "--vcf", path, "--field", field],
                         capture_output=True, text=True)          # default filter keeps another number of records
    assert bad.returncode == 1 and "not the SNPs of the pileup" in bad.stderr


@pytest.mark.gpu
def test_cli_vcf_missing_field_only_matters_after_the_filter(tmp_path):
    """a record whose FORMAT lacks --field stops the reference only when it passes the genotype filter
    (parse_posteriors follows passed_vfilter, cmd_cram_demuxlet.cpp:153,:211-212): a monomorphic record without PL is
    dropped silently, a polymorphic one is the reference's error"""
    import os
    import subprocess
    from demuxlet_b200 import synth
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "demuxlet_b200", "dmx-demuxlet")
    rng = np.random.default_rng(10)
    samples = ["d%d" % i for i in range(4)]
    recs = vcfio.random_records(rng, 300, len(samples), "PL")
    mono = ("1", 1, "A", "C", ["GT"], [{"GT": "0/0"}] * 4)
    poly = ("1", 2, "A", "C", ["GT"], [{"GT": "0/1"}, {"GT": "1/1"}, {"GT": "0/0"}, {"GT": "0/1"}])
    ok_path, bad_path = str(tmp_path / "ok.vcf"), str(tmp_path / "bad.vcf")
    vcfio.write_vcf(ok_path, samples, [mono] + recs)
    vcfio.write_vcf(bad_path, samples, [mono, poly] + recs)
    v = vcfio.parse_vcf(ok_path, "GT")
    keep = int(po.site_filter(v["gt"], 1, 0.5).sum())
    synth.make_synth(10, len(samples), keep, "GP", seed=6).save(str(tmp_path / "a.dmxp"))
    r = subprocess.run([exe, "--plp", str(tmp_path / "a.dmxp"), "--out", str(tmp_path / "a"), "--vcf", ok_path, "--field", "PL"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "--plp", str(tmp_path / "a.dmxp"), "--out", str(tmp_path / "b"), "--vcf", bad_path, "--field", "PL"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "Cannot parse posterior probability at 1:2" in r.stderr


@pytest.mark.gpu
def test_cli_vcf_sample_subset(tmp_path):
    """--sm / --sm-list on the command line: the kept samples (name order) are the samples of the whole run"""
    import dataclasses
    import filecmp
    import os
    import subprocess
    from demuxlet_b200 import synth
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "demuxlet_b200", "dmx-demuxlet")
    rng = np.random.default_rng(19)
    samples = ["d5", "d1", "d4", "d2", "d3", "d0"]
    path = str(tmp_path / "a.vcf")
    vcfio.write_vcf(path, samples, vcfio.random_records(rng, 700, len(samples), "PL"))
    v = vcfio.parse_vcf(path, "PL")
    pick = ["d4", "d0", "d2", "d5"]
    order = sorted(pick)
    cols = [samples.index(x) for x in order]
    gt, vals = np.ascontiguousarray(v["gt"][:, cols]), np.ascontiguousarray(v["values"][:, cols])
    keep = po.site_filter(gt, 1, 0.5)
    gp = po.build_gp("PL", np.ascontiguousarray(vals[keep]), 0.01)
    p = synth.make_synth(20, len(order), int(keep.sum()), "GP", seed=6)
    dataclasses.replace(p, gp=gp, samples=order).save(str(tmp_path / "a.dmxp"))
    dataclasses.replace(p, gp=np.zeros_like(gp)).save(str(tmp_path / "b.dmxp"))
    (tmp_path / "sm.txt").write_text("d4\textra\nd0\n")
    ra = subprocess.run([exe, "--plp", str(tmp_path / "a.dmxp"), "--out", str(tmp_path / "a")], capture_output=True, text=True)
    rb = subprocess.run([exe, "--plp", str(tmp_path / "b.dmxp"), "--out", str(tmp_path / "b"), "--vcf", path, "--field", "PL", "--sm", "d2",
                         "--sm", "d5", "--sm-list", str(tmp_path / "sm.txt")], capture_output=True, text=True)
    assert ra.returncode == 0 and rb.returncode == 0, ra.stderr + rb.stderr
    for ext in ("best", "single", "sing2"):
        assert filecmp.cmp(str(tmp_path / ("a." + ext)), str(tmp_path / ("b." + ext)), shallow=False), ext


def test_contig_dictionary_defines_the_chromosome_ids(tmp_path):
    """##contig lines come first (header order), then chromosomes in order of first appearance in ANY record, also a
    skipped multi-allelic one (htslib's dictionary; the BAM sweep maps read chromosomes through it,
    cmd_cram_demuxlet.cpp:197-199)"""
    rng = np.random.default_rng(3)
    samples = ["a", "b", "c"]
    recs = vcfio.random_records(rng, 60, 3, "GT")                      # chr1..chr3
    recs.insert(0, ("chrZ", 5, "A", "T,G", ["GT"], [{"GT": "0/1"}] * 3))   # skipped record on an undeclared chromosome
    path = str(tmp_path / "c.vcf")
    vcfio.write_vcf(path, samples, recs, contigs=["chrM", "chr2", "chr1", "chrEmpty"])
    got, want = api.read_vcf(path, "GT"), vcfio.parse_vcf(path, "GT")
    same(got, want, "GT")
    assert got["chroms"] == ["chrM", "chr2", "chr1", "chrEmpty", "chrZ", "chr3"]
    assert got["n_skipped"] >= 1 and set(np.unique(got["rid"])) == {1, 2, 5}


def test_tutorial_vcf_pin():
    """the one VCF the reference holds (tutorial/jurkat_293T_exons_only.vcf.gz: 54 424 records x 2 samples, plain gzip,
    chromosome order 1, 10, 11, ...): dmx_vcf_read against an independent python parse, live where /root/reference
    exists, and against the summary committed by tests/golden/make_tutorial_golden.py"""
    import hashlib
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "tutorial_vcf.json")))
    assert gold["n_records"] == 54424 and gold["n_skipped"] == 308 and gold["n_biallelic"] == 54116
    assert gold["samples"] == ["jurkat", "293T_RTG"] and gold["record_chroms"][:3] == ["1", "10", "11"]
    src = gold["source"]
    if not os.path.exists(src):
        pytest.skip("the tutorial VCF lives in /root/reference (build container only)")
    got = api.read_vcf(src, "GT")

    def digest(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    assert got["samples"] == gold["samples"] and got["chroms"] == gold["chroms"]
    assert got["n_skipped"] == gold["n_skipped"] and len(got["pos"]) == gold["n_biallelic"]
    assert digest(got["rid"]) == gold["sha_rid"] and digest(got["pos"]) == gold["sha_pos"]
    assert digest(got["gt"]) == gold["sha_gt"]
    assert hashlib.sha256(got["ref"].tobytes()).hexdigest()[:16] == gold["sha_ref"]
    assert hashlib.sha256(got["alt"].tobytes()).hexdigest()[:16] == gold["sha_alt"]
    # and live against the independent parse itself
    from tests.golden.make_tutorial_golden import independent_parse
    ind = independent_parse(src)
    assert np.array_equal(ind["gt"], got["gt"]) and np.array_equal(ind["pos"], got["pos"]) and np.array_equal(ind["rid"], got["rid"])
    npz = np.load(os.path.join(here, "golden", "tutorial_gt.npz"))
    assert np.array_equal(npz["gt"], got["gt"])


@pytest.mark.gpu
def test_tutorial_gt_chain_on_gpu():
    """the tutorial VCF's genotype calls (committed tests/golden/tutorial_gt.npz) through dmx_site_filter and the
    --field GT matrix on the device, against the oracle (pinned on the reference's own passed_vfilter / parse_posteriors)"""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gt = np.load(os.path.join(here, "golden", "tutorial_gt.npz"))["gt"]
    assert gt.shape == (54116, 2, 2)
    # demuxlet's defaults (--min-mac 1 --min-callrate 0.5) and a setting that drops records of this two-sample file
    for mac, cr in ((1, 0.5), (2, 1.0)):
        assert np.array_equal(api.site_filter(gt, mac, cr), po.site_filter(gt, mac, cr))
    strict = po.site_filter(gt, 2, 1.0)
    assert 0 < strict.sum() < len(strict)
    keep = api.site_filter(gt)
    g = api.build_gp("GT", gt[keep], 0.01)
    assert np.array_equal(g.view(np.uint32), po.build_gp("GT", gt[keep], 0.01).view(np.uint32))
