"""Writes tests/golden/field_{gt,gp,pl}.npz: small VCF-field arrays and the genotype-probability matrix the
REFERENCE's own parse_posteriors / parse_genotypes / parse_likelihoods produce for them
(/root/reference/bcf_filtered_reader.cpp:186-320, :360-454, run through oracle/_ref/ref_field).

Run in the build container (needs /root/reference): python tests/golden/make_field_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs():
    rng = np.random.default_rng(20261017)
    nv, ns = 5, 48
    gt = rng.choice(np.array([-1, 0, 1], np.int8), size=(ns, nv, 2), p=[0.1, 0.6, 0.3])
    gt[0] = -1          # every call missing: acs = an = 0
    gt[1] = 0
    gt[2] = 1
    gt[3, :, 0] = -1    # half-missing calls: the other allele still counts (:234)
    raw = rng.random((ns, nv, 3)).astype(np.float32) ** 4
    raw[0] = np.float32(1.0)
    raw[1, :, :] = [0, 0, 1]
    raw[2, 0] = [1e-30, 1e-30, 1e-30]
    pl = rng.integers(0, 120, size=(ns, nv, 3)).astype(np.int32)
    pl[0] = 0
    pl[1, 2] = np.iinfo(np.int32).min      # missing: reads as phred 255 (PhredHelper.h:40)
    pl[2] = [0, 30, 300]
    pl[3, :, :] = 255
    return {"gt": ("GT", gt), "gp": ("GP", raw), "pl": ("PL", pl)}


def main():
    if not po.ref_field_available():
        po.build()
    for name, (field, arr) in inputs().items():
        gp, af = po.run_ref_field(field, arr, 0.01, want_af=True)
        extra = {}
        if field == "GT":   # the reference's genotype-count site filter (:540-573) at demuxlet's defaults and a stricter pair
            extra = {"pass_1_50": po.run_ref_site_filter(arr, 1, 0.5), "pass_2_80": po.run_ref_site_filter(arr, 2, 0.8)}
        np.savez_compressed(os.path.join(HERE, "field_%s.npz" % name), field=field, gt_error=0.01, input=arr, gp=gp, af=af,
                            **extra)
        print(name, arr.shape, gp.shape)


if __name__ == "__main__":
    main()
