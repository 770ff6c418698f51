"""The committed reader fixtures (tests/golden/readers/, written by tests/golden/make_reader_golden.py): fixed BAM / SAM /
VCF / BCF bytes and the arrays the python restatements derived from them -- the library must deliver exactly those, at
any thread count and window size.  No GPU needed."""
import os

import numpy as np
import pytest

from demuxlet_b200 import api

D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "readers")


@pytest.mark.parametrize("name", ["reads.bam", "reads.sam"])
@pytest.mark.parametrize("threads,blocks", [("1", None), ("4", "2")])
def test_sweep_golden(name, threads, blocks, monkeypatch):
    monkeypatch.setenv("DMX_BAM_THREADS", threads)
    monkeypatch.setenv("DMX_BAM_MIN_CHUNK", "16")
    if blocks:
        monkeypatch.setenv("DMX_IO_BATCH_BLOCKS", blocks)
        monkeypatch.setenv("DMX_IO_CHUNK_BYTES", "777")
    z = np.load(os.path.join(D, "sweep.npz"))
    got = api.sweep_bam(os.path.join(D, name), [str(c) for c in z["chroms"]], z["rid"], z["pos"], z["ref"], z["alt"])
    assert got["barcodes"] == [str(b) for b in z["barcodes"]] and got["umis"] == [str(u) for u in z["umis"]]
    for k in ("totl", "cell", "snp", "allele", "bq"):
        assert np.array_equal(got[k], z[k]), k
    assert [got["counts"][k] for k in ("records", "filtered", "off_vcf", "skip_bcd")] == z["counts"].tolist()
    assert len(z["cell"]) > 100 and z["counts"][1] > 0 and z["counts"][2] > 0


@pytest.mark.parametrize("field", ["GP", "PL"])
@pytest.mark.parametrize("ext", [".vcf.gz", ".bcf"])
@pytest.mark.parametrize("threads", ["1", "5"])
def test_vcf_golden(field, ext, threads, monkeypatch):
    monkeypatch.setenv("DMX_VCF_THREADS", threads)
    z = np.load(os.path.join(D, field + ".npz"))
    got = api.read_vcf(os.path.join(D, field + ext), field)
    assert got["samples"] == [str(s) for s in z["samples"]] and got["chroms"] == [str(c) for c in z["chroms"]]
    assert got["n_skipped"] == int(z["n_skipped"]) > 0 and len(got["pos"]) > 90
    for k in ("rid", "pos", "ref", "alt", "gt"):
        assert np.array_equal(got[k], z[k]), k
    if field == "GP":      # missing values are NaNs on both sides, the rest the same float bits
        nan = np.isnan(z["values"])
        assert np.array_equal(np.isnan(got["values"]), nan)
        assert np.array_equal(got["values"][~nan].view(np.uint32), z["values"][~nan].view(np.uint32))
    else:
        assert np.array_equal(got["values"], z["values"])
