"""inputs of tools/san_readers.sh: BAM (gzip / BGZF / corrupt / truncated / garbage), SAM, VCF text / BGZF / BCF (whole, truncated, garbage)"""
import sys, os, struct, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = sys.argv[1]
from tests import bamio, vcfio
from tests.test_bam import REFS, make_snps
rng = np.random.default_rng(77)
chroms, rid, pos, ref, alt = make_snps(rng, 600)
reads, _ = bamio.random_reads(rng, REFS, {1: pos[rid == 0], 2: pos[rid == 1]}, 6000)
for r in reads[::53]:
    if r["cigar"] == [("M", 90)]:
        r["cigar"] = [("M", 30), ("X", 1), ("=", 19), ("M", 40)]
bamio.write_bam(os.path.join(OUT, "a.bam"), REFS, reads)
bamio.write_bam(os.path.join(OUT, "b.bam"), REFS, reads, member=500, bgzf=True)
for r in reads:
    if r["qual"] is not None: r["qual"] = [min(q, 93) for q in r["qual"]]
bamio.write_sam(os.path.join(OUT, "c.sam"), REFS, reads)
raw = bytearray(open(os.path.join(OUT, "b.bam"), "rb").read()); raw[len(raw)//2] ^= 0x55; open(os.path.join(OUT, "bad.bam"), "wb").write(bytes(raw))
open(os.path.join(OUT, "cut.bam"), "wb").write(open(os.path.join(OUT, "b.bam"), "rb").read()[:70000])
with open(os.path.join(OUT, "snps.bin"), "wb") as f:
    f.write(struct.pack("<qi", len(pos), len(chroms)) + b"".join(c.encode() + b"\0" for c in chroms))
    f.write(rid.tobytes() + pos.tobytes() + ref.tobytes() + alt.tobytes())
samples = ["S%d" % i for i in range(6)]
for field in ("GP", "PL"):
    recs = vcfio.random_records(rng, 2000, len(samples), field)
    contigs = ["chrM", "chr1", "chr2", "chr3"]
    vcfio.write_vcf(os.path.join(OUT, "%s.vcf" % field), samples, recs, contigs=contigs)
    vcfio.write_vcf(os.path.join(OUT, "%s.vcf.gz" % field), samples, recs, contigs=contigs, bgzf=True)
    vcfio.write_bcf(os.path.join(OUT, "%s.bcf" % field), samples, recs, contigs)
    raw = open(os.path.join(OUT, "%s.bcf" % field), "rb").read(); open(os.path.join(OUT, "%s_cut.bcf" % field), "wb").write(raw[:len(raw)//2])
open(os.path.join(OUT, "garbage.vcf"), "wb").write(os.urandom(5000))
open(os.path.join(OUT, "garbage.bam"), "wb").write(b"BAM\1" + os.urandom(5000))
