"""dmx_vcf_read on VCF text (plain / gzip / multi-member / BGZF) and BCF2 of the same random records against the python\nrestatement (tests/vcfio.py), any thread count.  usage: python tools/fuzz_vcf.py SEED0 SEED1   (no GPU needed)"""
import os, sys, numpy as np, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demuxlet_b200 import api
from tests import vcfio
from tests.test_vcf import same
td = tempfile.mkdtemp()
nbad = 0
for seed in range(int(sys.argv[1]), int(sys.argv[2])):
    rng = np.random.default_rng(5000 + seed)
    field = ["GT", "GP", "PL"][seed % 3]
    nv = int(rng.integers(1, 12)); nrec = int(rng.integers(1, 900))
    samples = ["s%d" % i for i in range(nv)]
    recs = vcfio.random_records(rng, nrec, nv, field)
    contigs = ["chrM", "chr1", "chr2", "chr3"]
    os.environ["DMX_VCF_THREADS"] = str(int(rng.choice([1, 2, 5, 16])))
    kind = int(rng.integers(0, 4))
    pt = os.path.join(td, "a.vcf" + (".gz" if kind else ""))
    vcfio.write_vcf(pt, samples, recs, contigs=contigs, bgzip_like=(kind == 2), bgzf=(kind == 3))
    pb = os.path.join(td, "a.bcf")
    vcfio.write_bcf(pb, samples, recs, contigs, int_type=int(rng.choice([0, 2, 3])), idx=bool(rng.integers(0, 2)))
    try:
        want = vcfio.parse_vcf(pt, field)
        same(api.read_vcf(pt, field), want, field)
        same(api.read_vcf(pb, field), want, field)
    except Exception as e:
        nbad += 1
        print("seed", seed, field, "FAIL", type(e).__name__, str(e)[:300])
print("done", nbad, "failures")
