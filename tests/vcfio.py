"""VCF text fixtures and a plain-python restatement of demuxlet_b200/csrc/dmx_vcf.cpp (what htslib's FORMAT accessors
hand to the reference per record, /root/reference/bcf_filtered_reader.cpp:194, :250, :408): the checker of the C++
reader.  htslib is not in this image, so parity with htslib's own parse is unpinned; the conventions restated here
(missing GT -> negative allele, missing / short GP -> NaN, missing / short PL -> INT32_MIN, n_allele != 2 skipped)
are htslib's published ones."""
import gzip

import numpy as np

INT_MISSING = np.iinfo(np.int32).min


def write_vcf(path, samples, records, bgzip_like=False, contigs=(), bgzf=False):
    """records: list of (chrom, pos1, ref, alt, fmt_keys, [per-sample dict key -> text]); contigs: ##contig header lines"""
    lines = ["##fileformat=VCFv4.2", "##source=dmx-b200 tests"]
    for i, c in enumerate(contigs):
        lines.append(("##contig=<ID=%s,length=%d>" % (c, 1000000 + i)) if i % 2 == 0 else
                     ('##contig=<length=%d,assembly="a,b",ID=%s>' % (1000000 + i, c)))
    lines.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(samples)))
    for chrom, pos1, ref, alt, keys, cols in records:
        lines.append("\t".join([chrom, str(pos1), ".", ref, alt, ".", "PASS", ".", ":".join(keys)] +
                               [":".join(c[k] for k in keys if k in c) for c in cols]))
    text = ("\n".join(lines) + "\n").encode()
    if path.endswith(".gz"):
        if bgzf:            # real BGZF: blocks with the BC size field and the empty end-of-file block (SAM spec 4.1)
            from tests.bamio import bgzf_block
            with open(path, "wb") as f:
                for i in range(0, len(text), 900):
                    f.write(bgzf_block(text[i:i + 900]))
                f.write(bgzf_block(b""))
        elif bgzip_like:      # several gzip members back to back, as bgzip writes them
            with open(path, "wb") as f:
                for i in range(0, len(text), 1500):
                    f.write(gzip.compress(text[i:i + 1500]))
        else:
            with gzip.open(path, "wb") as f:
                f.write(text)
    else:
        with open(path, "wb") as f:
            f.write(text)


def parse_vcf(path, field="GT"):
    op = gzip.open if path.endswith(".gz") else open
    samples, chroms = [], []
    rid, pos, ref, alt, gt, vals = [], [], [], [], [], []
    nskip = 0
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if not line:
                continue
            if line.startswith("#"):
                if line.startswith("#CHROM"):
                    samples = line.split("\t")[9:]
                elif line.startswith("##contig=<"):      # the header's contig dictionary comes first, in header order
                    import re
                    m = re.search(r"[<,]ID=([^,>]+)", line)
                    if m and m.group(1) not in chroms:
                        chroms.append(m.group(1))
                continue
            c = line.split("\t")
            if c[0] not in chroms:                       # every record's chromosome, also of the skipped ones
                chroms.append(c[0])
            if c[4] == "." or len(c[4].split(",")) != 1:
                nskip += 1
                continue
            keys = c[8].split(":")
            rid.append(chroms.index(c[0]))
            pos.append(int(c[1]) - 1)
            ref.append(c[3][0].encode())
            alt.append(c[4][0].encode())
            g, v = [], []
            for col in c[9:]:
                sub = col.split(":")
                t = sub[keys.index("GT")] if keys.index("GT") < len(sub) else ""
                a = [x for x in t.replace("|", "/").split("/")][:2] if t else []
                a = [(-1 if (x == "." or x == "") else int(x)) for x in a] + [-1, -1]
                g.append(a[:2])
                if field != "GT":
                    i = keys.index(field)
                    t = sub[i] if i < len(sub) else ""
                    parts = t.split(",")[:3] if t else []
                    if field == "GP":
                        x = [np.float32(float(p)) if p not in (".", "") else np.float32("nan") for p in parts]   # (float)strtod, as htslib
                        x += [np.float32("nan")] * (3 - len(x))
                    else:
                        x = [int(p) if p not in (".", "") else INT_MISSING for p in parts]
                        x += [INT_MISSING] * (3 - len(x))
                    v.append(x)
            gt.append(g)
            vals.append(v)
    nv = len(samples)
    gt = np.array(gt, np.int8).reshape(len(pos), nv, 2)
    if field == "GT":
        values = gt
    else:
        values = np.array(vals, np.float32 if field == "GP" else np.int32).reshape(len(pos), nv, 3)
    return {"samples": samples, "chroms": chroms, "rid": np.array(rid, np.int32), "pos": np.array(pos, np.int32),
            "ref": np.array(ref, "S1"), "alt": np.array(alt, "S1"), "gt": gt, "values": values, "n_skipped": nskip}


def random_records(rng, nrec, nv, field):
    recs = []
    for r in range(nrec):
        chrom = "chr%d" % (1 + r * 3 // max(nrec, 1))
        kind = rng.random()
        alt = "T" if kind > 0.1 else ("T,G" if kind > 0.05 else ".")
        keys = ["GT", "DP"] + ([field] if field != "GT" else [])
        if rng.random() < 0.3 and field != "GT":
            keys = [field, "GT", "DP"]
        cols = []
        for s in range(nv):
            u = rng.random()
            if u < 0.05:
                g = "./."
            elif u < 0.08:
                g = "."
            elif u < 0.1:
                g = "1"
            elif u < 0.12:
                g = "./1"
            else:
                g = "%d%s%d" % (rng.integers(0, 2), "/|"[rng.integers(0, 2)], rng.integers(0, 2))
            c = {"GT": g, "DP": str(rng.integers(0, 60))}
            if field == "GP":
                w = rng.random()
                if w < 0.04:
                    c["GP"] = "."
                elif w < 0.06:
                    c["GP"] = "0.5,0.5"
                else:
                    x = rng.dirichlet([0.3, 0.3, 0.3])
                    c["GP"] = ",".join(("%.4g" % t) for t in x)
            if field == "PL":
                w = rng.random()
                if w < 0.04:
                    c["PL"] = "."
                elif w < 0.06:
                    c["PL"] = "0,.,40"
                else:
                    c["PL"] = ",".join(str(t) for t in rng.integers(0, 400, size=3))
            cols.append(c)
        recs.append((chrom, 1000 + 17 * r, "ACGT"[r % 4] + ("A" if r % 50 == 7 else ""), alt, keys, cols))
    return recs
