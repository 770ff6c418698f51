"""Pins of the htslib-free VCF reader on the one fixture the reference holds:
/root/reference/tutorial/jurkat_293T_exons_only.vcf.gz (2 samples, GT only, plain gzip, chromosome order 1,10,11,...).
An independent parse (python gzip + str.split, nothing from demuxlet_b200) writes
  tutorial_vcf.json  record counts, sample ids, contig / chromosome order, per-array checksums
  tutorial_gt.npz    the allele indices of the biallelic records (int8 [nrec][2][2]) + rid / pos: the input of the
                     -m gpu chain test (dmx_site_filter, dmx_build_gp --field GT), which cannot read /root/reference
Run here (the build container): python tests/golden/make_tutorial_golden.py"""
import gzip
import hashlib
import json
import os
import re

import numpy as np

SRC = "/root/reference/tutorial/jurkat_293T_exons_only.vcf.gz"
HERE = os.path.dirname(os.path.abspath(__file__))


def independent_parse(path=SRC):
    samples, contigs, chroms = [], [], []
    rid, pos, ref, alt, gt = [], [], [], [], []
    n_total = n_skip = 0
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith("##contig=<"):
                m = re.search(r"[<,]ID=([^,>]+)", line)
                if m:
                    contigs.append(m.group(1))
                continue
            if line.startswith("#CHROM"):
                samples = line.rstrip("\n").split("\t")[9:]
                continue
            if line.startswith("#") or not line.strip():
                continue
            c = line.rstrip("\n").split("\t")
            n_total += 1
            if c[0] not in chroms:
                chroms.append(c[0])
            if c[4] == "." or "," in c[4]:
                n_skip += 1
                continue
            keys = c[8].split(":")
            g = []
            for col in c[9:]:
                t = col.split(":")[keys.index("GT")]
                a = [(-1 if x in (".", "") else int(x)) for x in re.split(r"[/|]", t)][:2]
                g.append((a + [-1, -1])[:2])
            order = contigs + [x for x in chroms if x not in contigs]
            rid.append(order.index(c[0]) if c[0] in contigs else None)
            pos.append(int(c[1]) - 1)
            ref.append(c[3][0])
            alt.append(c[4][0])
            gt.append(g)
    order = contigs + [x for x in chroms if x not in contigs]
    # rid against the final dictionary (contigs first, then first appearance)
    rid = []
    with gzip.open(path, "rt") as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            c = line.split("\t", 5)
            if c[4] == "." or "," in c[4]:
                continue
            rid.append(order.index(c[0]))
    return {"samples": samples, "contigs": contigs, "record_chroms": chroms, "chroms": order, "n_records": n_total,
            "n_skipped": n_skip, "rid": np.array(rid, np.int32), "pos": np.array(pos, np.int32),
            "ref": "".join(ref), "alt": "".join(alt), "gt": np.array(gt, np.int8)}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


if __name__ == "__main__":
    p = independent_parse()
    summary = {"source": SRC, "samples": p["samples"], "n_contig_lines": len(p["contigs"]),
               "record_chroms": p["record_chroms"], "chroms": p["chroms"], "n_records": p["n_records"],
               "n_skipped": p["n_skipped"], "n_biallelic": int(len(p["pos"])),
               "sha_rid": digest(p["rid"]), "sha_pos": digest(p["pos"]), "sha_gt": digest(p["gt"]),
               "sha_ref": hashlib.sha256(p["ref"].encode()).hexdigest()[:16],
               "sha_alt": hashlib.sha256(p["alt"].encode()).hexdigest()[:16]}
    json.dump(summary, open(os.path.join(HERE, "tutorial_vcf.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "tutorial_gt.npz"), gt=p["gt"], rid=p["rid"], pos=p["pos"])
    print(json.dumps(summary, indent=1))
