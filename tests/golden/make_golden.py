#!/usr/bin/env python
"""Regenerates tests/golden/*: small pileups (DMXPLP01 containers) and the four TSVs that the REFERENCE's own
lines 386-881 (oracle/_ref/ref_harness, built from /root/reference by oracle/Makefile) write for them.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The outputs are committed; tests/test_oracle.py checks the CPU oracle against them byte for byte, and the GPU
tests check libdmx against them, so parity stays pinned on machines where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from demuxlet_b200.plp import Pileup, build_from_reads  # noqa: E402
from demuxlet_b200.synth import make_synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

CASES = {
    # name: (make_synth kwargs, alphas, extra ref args)
    "gp6": (dict(ncells=40, nv=6, nsnps=400, field="GP", seed=101, cover_frac=0.08), (0.0, 0.5), {}),
    "gt2": (dict(ncells=60, nv=2, nsnps=600, field="GT", seed=102, cover_frac=0.03), (0.0, 0.5), {}),
    "gt8ties": (dict(ncells=50, nv=8, nsnps=120, field="GT", seed=103, cover_frac=0.05), (0.0, 0.5), {}),
    "pl5grid": (dict(ncells=30, nv=5, nsnps=300, field="PL", seed=104, cover_frac=0.1),
                (0.0, 0.1, 0.25, 0.4, 0.5), {}),
    "gp4alpha0": (dict(ncells=30, nv=4, nsnps=300, field="GP", seed=105, cover_frac=0.1), (0.2, 0.5), {}),
    "gp6filt": (dict(ncells=40, nv=6, nsnps=400, field="GP", seed=106, cover_frac=0.05), (0.0, 0.5),
                dict(min_total=120, min_uniq=12, min_snp=10, doublet_prior=0.3)),
}


def main():
    po.build()
    assert po.ref_available(), "oracle/_ref/ref_harness missing: /root/reference not present?"
    manifest = {}
    for name, (kw, alphas, extra) in CASES.items():
        p = make_synth(**kw)
        if name == "gp6filt":   # one barcode without any SNP (cmd_cram_demuxlet.cpp:592)
            n_e = np.diff(p.cell_ptr)
            keep = np.ones(p.ncells, dtype=bool)
            keep[7] = False
            sub = p.subset_cells(np.nonzero(keep)[0])
            cell_ptr = np.concatenate([[0], np.cumsum(np.where(keep, n_e, 0))])
            p = Pileup(p.gp, cell_ptr, sub.ent_snp, sub.ent_ptr, sub.reads, p.totl, p.npass, p.uniq, p.barcodes,
                       p.samples)
        path = os.path.join(HERE, name + ".dmxp")
        p.save(path)
        po.run_ref(p, os.path.join(HERE, name), alphas=alphas, write_pair=True, plp_path=path,
                   doublet_prior=extra.get("doublet_prior"), min_total=extra.get("min_total", 0),
                   min_uniq=extra.get("min_uniq", 0), min_snp=extra.get("min_snp", 0))
        manifest[name] = dict(alphas=list(alphas), **extra)
        print(name, p.ncells, p.nv, p.nsnps, p.n_entries, os.path.getsize(path))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
