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


# ---- BCF2 writer (VCF specification 6.3), for the reader's tests ----------------------------------------------------
import struct as _struct

_INT_MISSING = {1: -128, 2: -32768, 3: -2147483648}
_FLOAT_MISSING, _FLOAT_END = 0x7F800001, 0x7F800002


def _typed_header(t, n):
    if n < 15:
        return bytes([(n << 4) | t])
    return bytes([0xF0 | t]) + _typed_int(n)


def _int_type(vals):
    lo, hi = (min(vals), max(vals)) if vals else (0, 0)
    if lo >= -120 and hi <= 127:
        return 1
    if lo >= -32760 and hi <= 32767:
        return 2
    return 3


def _typed_int(x):
    t = _int_type([x])
    return bytes([0x10 | t]) + _struct.pack("<" + "bhi"[t - 1], x)


def _typed_str(txt):
    b = txt.encode()
    return _typed_header(7, len(b)) + b


def write_bcf(path, samples, records, contigs, int_type=0, idx=False):
    """the records of write_vcf as a BCF2 file (BGZF blocks).  contigs: every chromosome, in dictionary order (BCF needs
    them declared).  int_type 1 / 2 / 3 forces the width of the PL / DP vectors (0: smallest that fits, like htslib);
    idx=True writes IDX= into the header lines, in an order different from their appearance."""
    from tests.bamio import bgzf_block
    fmt_ids = ["GT", "DP", "GP", "PL"]
    ids = ["PASS"] + fmt_ids
    order = list(range(len(fmt_ids)))
    if idx:
        order = order[::-1]                      # header lines in reverse, IDX keeps the numbering
    lines = ["##fileformat=VCFv4.2", '##FILTER=<ID=PASS,Description="All filters passed"%s>' % (",IDX=0" if idx else "")]
    for i, c in enumerate(contigs):
        lines.append("##contig=<ID=%s,length=%d%s>" % (c, 1000000 + i, (",IDX=%d" % i) if idx else ""))
    types = {"GT": ("1", "String"), "DP": ("1", "Integer"), "GP": ("G", "Float"), "PL": ("G", "Integer")}
    for k in order:
        f = fmt_ids[k]
        lines.append('##FORMAT=<ID=%s,Number=%s,Type=%s,Description="x, y"%s>' % (f, types[f][0], types[f][1],
                                                                                    (",IDX=%d" % (k + 1)) if idx else ""))
    lines.append("\t".join(["#CHROM", "POS", "ID", "REF", "ALT", "QUAL", "FILTER", "INFO", "FORMAT"] + list(samples)))
    text = ("\n".join(lines) + "\n").encode() + b"\0"
    out = bytearray(b"BCF\2\2" + _struct.pack("<I", len(text)) + text)
    for chrom, pos1, ref, alt, keys, cols in records:
        alleles = [ref] + ([] if alt == "." else alt.split(","))
        shared = bytearray(_struct.pack("<iiifII", contigs.index(chrom), pos1 - 1, len(ref), float("nan"),
                                        (len(alleles) << 16) | 0, (len(keys) << 24) | len(samples)))
        _struct.pack_into("<I", shared, 12, 0x7F800001)          # QUAL missing
        shared += _typed_str(".")                                 # ID
        for a in alleles:
            shared += _typed_str(a)
        shared += _typed_header(1, 1) + _struct.pack("<b", 0)     # FILTER = PASS
        indiv = bytearray()
        for k in keys:
            indiv += _typed_int(ids.index(k))
            if k == "GT":
                rows = []
                for c in cols:
                    t = c["GT"]
                    vals, phased = [], 0
                    for ch in t.replace("|", "/|").split("/"):      # keeps the separator in front of a phased allele
                        ph = ch.startswith("|")
                        a = ch.lstrip("|")
                        vals.append((0 if a in (".", "") else (int(a) + 1) << 1) | (1 if ph else 0))
                    rows.append(vals[:2])
                width = max(len(r) for r in rows)
                indiv += _typed_header(1, width)
                for r in rows:
                    indiv += _struct.pack("<%db" % width, *(r + [-127] * (width - len(r))))
            elif k == "GP":
                rows = []
                for c in cols:
                    parts = [] if c["GP"] == "." else c["GP"].split(",")
                    vals = [_FLOAT_MISSING if x == "." else _struct.unpack("<I", _struct.pack("<f", float(x)))[0] for x in parts]
                    rows.append(vals or [_FLOAT_MISSING])
                width = max(len(r) for r in rows)
                indiv += _typed_header(5, width)
                for r in rows:
                    indiv += _struct.pack("<%dI" % width, *(r + [_FLOAT_END] * (width - len(r))))
            else:                                                  # DP, PL
                rows = []
                for c in cols:
                    parts = [] if c[k] == "." else c[k].split(",")
                    rows.append([None if x == "." else int(x) for x in parts] or [None])
                width = max(len(r) for r in rows)
                t = int_type or _int_type([x for r in rows for x in r if x is not None])
                indiv += _typed_header(t, width)
                for r in rows:
                    vals = [(_INT_MISSING[t] if x is None else x) for x in r] + [_INT_MISSING[t] + 1] * (width - len(r))
                    indiv += _struct.pack("<%d%s" % (width, "bhi"[t - 1]), *vals)
        out += _struct.pack("<II", len(shared), len(indiv)) + shared + indiv
    with open(path, "wb") as f:
        for i in range(0, len(out), 60000):
            f.write(bgzf_block(bytes(out[i:i + 60000])))
        f.write(bgzf_block(b""))
