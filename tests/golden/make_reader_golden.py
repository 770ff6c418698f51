"""Fixed input files of the htslib-free readers and what they must deliver (tests/test_reader_golden.py): the same 400
reads as BGZF BAM and SAM text over 300 SNPs on three contigs, the same 120 records x 5 samples as bgzipped VCF and BCF2.
The files come from the spec-following writers in tests/bamio.py / tests/vcfio.py (not from htslib: the image has none);
the expected arrays from the python restatements (tests/bamio.py:sweep -- the reference's loop transcribed --,
tests/vcfio.py:parse_vcf), NOT from the library under test.  Committed so that a later change of writer, restatement
or reader that shifts any byte or value shows up.
usage: python tests/golden/make_reader_golden.py   (writes tests/golden/readers/)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import bamio, vcfio                      # noqa: E402

OUT = os.path.join(HERE, "readers")
os.makedirs(OUT, exist_ok=True)
REFS = [("chrM", 3000), ("chr1", 60000), ("chr2", 50000), ("chrUn", 5000)]
rng = np.random.default_rng(20261017)
chroms = ["chrM", "chr1", "chr2"]
rid = np.sort(rng.integers(1, 3, 300)).astype(np.int32)
pos = np.zeros(300, np.int32)
for r in (1, 2):
    m = rid == r
    pos[m] = np.sort(rng.choice(np.arange(100, 6000), m.sum(), replace=False))
ref = rng.choice(np.array([b"A", b"C", b"G", b"T"]), 300).astype("S1")
alt = np.array([rng.choice([b for b in (b"A", b"C", b"G", b"T") if b != x]) for x in ref]).astype("S1")
reads, barcodes = bamio.random_reads(rng, REFS, {1: pos[rid == 1], 2: pos[rid == 2]}, 400, ncells=6)
for r in reads:
    r["pos"] = r["pos"] % 6100
    if r["qual"] is not None:
        r["qual"] = [min(q, 93) for q in r["qual"]]        # SAM text cannot carry more
for r in reads[::29]:
    if r["cigar"] == [("M", 90)]:
        r["cigar"] = [("M", 30), ("X", 1), ("=", 19), ("M", 40)]
reads.sort(key=lambda r: (r["tid"], r["pos"]))
bamio.write_bam(os.path.join(OUT, "reads.bam"), REFS, reads, member=1500, bgzf=True)
bamio.write_sam(os.path.join(OUT, "reads.sam"), REFS, reads)
want = bamio.sweep(REFS, reads, chroms, rid, pos, ref, alt)
np.savez_compressed(os.path.join(OUT, "sweep.npz"), rid=rid, pos=pos, ref=ref, alt=alt, chroms=np.array(chroms),
                    barcodes=np.array(want["barcodes"]), totl=want["totl"], cell=want["cell"], snp=want["snp"],
                    allele=want["allele"], bq=want["bq"], umis=np.array(want["umis"]),
                    counts=np.array([want["counts"][k] for k in ("records", "filtered", "off_vcf", "skip_bcd")]))
samples = ["S%d" % i for i in range(5)]
contigs = ["chrM", "chr1", "chr2", "chr3"]
for field in ("GP", "PL"):
    recs = vcfio.random_records(rng, 120, len(samples), field)
    vcfio.write_vcf(os.path.join(OUT, field + ".vcf.gz"), samples, recs, contigs=contigs, bgzf=True)
    vcfio.write_bcf(os.path.join(OUT, field + ".bcf"), samples, recs, contigs, idx=(field == "PL"))
    v = vcfio.parse_vcf(os.path.join(OUT, field + ".vcf.gz"), field)
    np.savez_compressed(os.path.join(OUT, field + ".npz"), samples=np.array(v["samples"]), chroms=np.array(v["chroms"]),
                        rid=v["rid"], pos=v["pos"], ref=v["ref"], alt=v["alt"], gt=v["gt"], values=v["values"],
                        n_skipped=v["n_skipped"])
print(sorted(os.listdir(OUT)))
