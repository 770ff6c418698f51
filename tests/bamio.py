"""BAM fixtures and a plain-python restatement of demuxlet_b200/csrc/dmx_bam.cpp (the reference's BAM sweep,
/root/reference/cmd_cram_demuxlet.cpp:186-338 with hts_utils.cpp:279-359 and sam_filtered_reader.cpp:284-296): the
checker of the C++ sweep.  The BAM writer follows the SAM specification (section 4.2; BGZF = concatenated gzip members).
htslib is not in this image, so parity with htslib's own decoding is unpinned."""
import gzip
import struct

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def bgzf_block(data: bytes, level: int = 6) -> bytes:
    """one BGZF block (SAM spec 4.1): gzip member with the BC extra subfield = total block size - 1"""
    import zlib
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    c = co.compress(data) + co.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(c) + 25) + c +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam(path, refs, reads, member=4000, bgzf=False, level=6):
    """refs: [(name, length)]; reads: dicts tid, pos, mapq, flag, cigar [(op, len)], seq, qual (list of ints or None),
    tags [(tag, type, value)] with type in Z, i, A, B (B: list of uint8)"""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs)))
    for name, ln in refs:
        out += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    for i, r in enumerate(reads):
        name = ("r%d" % i).encode() + b"\0"
        seq = r["seq"]
        cig = b"".join(struct.pack("<I", (ln << 4) | CIGAR_OPS.index(op)) for op, ln in r["cigar"])
        sq = bytearray((len(seq) + 1) // 2)
        for k, c in enumerate(seq):
            sq[k >> 1] |= SEQ_CODE[c] << (4 if k % 2 == 0 else 0)
        ql = bytes(r["qual"]) if r["qual"] is not None else b"\xff" * len(seq)
        aux = b""
        for tag, typ, val in r["tags"]:
            if typ == "Z":
                aux += tag.encode() + b"Z" + val.encode() + b"\0"
            elif typ == "i":
                aux += tag.encode() + b"i" + struct.pack("<i", val)
            elif typ == "A":
                aux += tag.encode() + b"A" + val.encode()
            elif typ == "B":
                aux += tag.encode() + b"BC" + struct.pack("<I", len(val)) + bytes(val)
        body = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(name), r["mapq"], 0, len(r["cigar"]), r["flag"],
                           len(seq), -1, -1, 0) + name + cig + bytes(sq) + ql + aux
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        if bgzf:                        # real BGZF blocks: the reader inflates them block-parallel
            for i in range(0, len(out), member):
                f.write(bgzf_block(bytes(out[i:i + member]), level))
            f.write(bgzf_block(b""))
            return
        for i in range(0, len(out), member):
            f.write(gzip.compress(bytes(out[i:i + member])))
        f.write(gzip.compress(b""))     # BGZF end-of-file marker: an empty member


def write_sam(path, refs, reads, gz=False, final_newline=True):
    """the same reads as SAM text (SAM spec 1.4) with an @HD / @SQ header"""
    lines = ["@HD\tVN:1.6\tSO:coordinate"] + ["@SQ\tSN:%s\tLN:%d" % (n, ln) for n, ln in refs] + ["@PG\tID:tests\tPN:bamio"]
    for i, r in enumerate(reads):
        tags = []
        for t, ty, v in r["tags"]:
            if ty == "B":
                tags.append("%s:B:i,%s" % (t, ",".join(str(x) for x in v)))
            else:
                tags.append("%s:%s:%s" % (t, ty, v))
        qual = "*" if r["qual"] is None else "".join(chr(min(q, 93) + 33) for q in r["qual"])
        lines.append("\t".join(["r%d" % i, str(r["flag"]), refs[r["tid"]][0] if r["tid"] >= 0 else "*", str(r["pos"] + 1),
                                 str(r["mapq"]), "".join("%d%s" % (ln, op) for op, ln in r["cigar"]) or "*", "*", "0", "0",
                                 r["seq"] or "*", qual] + tags))
    text = ("\n".join(lines) + ("\n" if final_newline else "")).encode()
    if gz:
        import gzip
        with gzip.open(path, "wb") as f:
            f.write(text)
    else:
        with open(path, "wb") as f:
            f.write(text)


class GlibcRand:
    """glibc's rand() / random() of a process that never calls srand(): the TYPE_3 additive feedback generator
    (r[i] = r[i-3] + r[i-31], 310 values discarded) seeded with 1 -- what the reference's `rand()` at
    cmd_cram_demuxlet.cpp:276 returns"""

    def __init__(self, seed=1):
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            r[i] = (16807 * r[i - 1]) % 2147483647
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = r
        for _ in range(310):
            self._step()

    def _step(self):
        v = (self.r[-31] + self.r[-3]) & 0xffffffff
        self.r.append(v)
        self.r.pop(0)
        return v

    def __call__(self):
        return self._step() >> 1


def sweep(refs, reads, chroms, snp_rid, snp_pos, snp_ref, snp_alt, groups=None, tag_group="CB", tag_umi="UB", min_mq=20,
          exclude_flag=0x0f04, min_bq=13, cap_bq=40, min_td=0):
    """cmd_cram_demuxlet.cpp:186-338 read by read, with the reference's SNP buffer transcribed literally: `ibeg` / `nbuf`
    are its variables (buffer = SNPs [ibeg, ibeg + nbuf), the newest one is the VCF cursor), clear_buffer_before is
    bcf_filtered_reader.cpp:649-669, and a read is tested against every buffered SNP by position alone (:306-308)."""
    rnd = GlibcRand()
    name2rid = {c: i for i, c in enumerate(chroms)}
    tid2rid = [name2rid.get(n, -1) for n, _ in refs]
    bc, totl = {}, []
    out = {"cell": [], "snp": [], "allele": [], "bq": [], "umis": []}
    cnt = {"records": 0, "filtered": 0, "off_vcf": 0, "skip_bcd": 0}
    groups = set(groups or [])
    nsnps = len(snp_pos)
    ibeg, nbuf, eof = 0, (1 if nsnps else 0), nsnps == 0     # :150 reads the first SNP before the loop
    for r in reads:
        cnt["records"] += 1
        if r["mapq"] < min_mq or (r["flag"] & exclude_flag):
            cnt["filtered"] += 1
            continue
        rid = tid2rid[r["tid"]] if 0 <= r["tid"] < len(refs) else -1
        if rid < 0:
            cnt["off_vcf"] += 1
            continue
        span = sum(ln for op, ln in r["cigar"] if op in "MDN=X")
        if r["flag"] & 4:                                  # bam_endpos: an unmapped read ends one base behind its start
            span = 0
        endpos = r["pos"] + (span if span > 0 else 1)
        n_rm = 0                                           # clear_buffer_before(chrom, read start), rlen = 1
        for i in range(ibeg, ibeg + nbuf):
            if snp_rid[i] < rid or (snp_rid[i] == rid and snp_pos[i] + 1 < r["pos"]):
                n_rm += 1
            else:
                break
        nbuf -= n_rm
        ibeg += n_rm                                       # :206
        while not eof:                                     # :209: the cursor is the newest SNP
            cur = ibeg + nbuf - 1 if nbuf else ibeg - 1
            if not (snp_rid[cur] < rid or (snp_rid[cur] == rid and snp_pos[cur] < endpos)):
                break
            if cur + 1 < nsnps:
                nbuf += 1
            else:
                eof = True
        tags = {t: v for t, ty, v in r["tags"] if ty == "Z"}
        sb = tags.get(tag_group, ".") if tag_group else "."
        if tag_group and groups and sb not in groups:
            cnt["skip_bcd"] += 1
            continue
        if sb not in bc:
            bc[sb] = len(bc)
            totl.append(0)
        ib = bc[sb]
        umi = (".%x" % rnd()) if tag_umi == "" else tags.get(tag_umi, ".")     # :274-276
        totl[ib] += 1
        for i in range(ibeg, ibeg + nbuf):                 # :306
            if not r["cigar"]:
                continue
            cpos, rpos, sp = r["pos"], 0, int(snp_pos[i])
            for op, ln in r["cigar"]:                      # hts_utils.cpp:296-330
                if op == "M":
                    if cpos <= sp <= cpos + ln - 1:
                        rpos += sp - cpos
                        break
                    cpos += ln
                    rpos += ln
                elif op in "DN":
                    if cpos <= sp <= cpos + ln - 1:
                        rpos = -1
                        break
                    cpos += ln
                elif op in "SI":
                    rpos += ln
            if rpos < 0 or rpos >= len(r["seq"]):
                continue
            base = r["seq"][rpos]
            if base == "N":
                continue
            raw = 0xff if r["qual"] is None else r["qual"][rpos]
            q = (raw + 33) & 0xff                          # the reference's signed char qual + 33
            q = (q - 256 if q >= 128 else q) - 33
            if q < min_bq or rpos < min_td - 1 or rpos + min_td > len(r["seq"]):
                continue
            out["cell"].append(ib)
            out["snp"].append(i)
            out["allele"].append(0 if base == snp_ref[i].decode() else (1 if base == snp_alt[i].decode() else 2))
            out["bq"].append(min(q, cap_bq))
            out["umis"].append(umi)
    return {"barcodes": list(bc), "totl": np.array(totl, np.int32), "cell": np.array(out["cell"], np.int32),
            "snp": np.array(out["snp"], np.int32), "allele": np.array(out["allele"], np.uint8),
            "bq": np.array(out["bq"], np.uint8), "umis": out["umis"], "counts": cnt}


def random_reads(rng, refs, snp_by_tid, n, ncells=12):
    """position-sorted reads with a mix of CIGARs, flags, qualities and tags; every third read covers a SNP"""
    barcodes = ["".join(rng.choice(list("ACGT"), 10)) + "-1" for _ in range(ncells)]
    reads = []
    for _ in range(n):
        tid = int(rng.integers(0, len(refs)))
        snps = snp_by_tid.get(tid, [])
        if len(snps) and rng.random() < 0.7:
            pos = max(0, int(snps[rng.integers(0, len(snps))]) - int(rng.integers(0, 60)))
        else:
            pos = int(rng.integers(0, refs[tid][1] - 200))
        kind = rng.random()
        if kind < 0.4:
            cigar = [("M", 90)]
        elif kind < 0.55:
            cigar = [("S", 7), ("M", 40), ("I", 3), ("M", 40)]
        elif kind < 0.7:
            cigar = [("M", 30), ("D", 4), ("M", 60)]
        elif kind < 0.85:
            cigar = [("M", 25), ("N", 150), ("M", 65)]
        elif kind < 0.95:
            cigar = [("H", 5), ("S", 3), ("M", 50), ("S", 37)]
        else:
            cigar = []
        lseq = sum(ln for op, ln in cigar if op in "MIS=X") or 50
        seq = "".join(rng.choice(list("ACGTN"), lseq, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        if rng.random() < 0.03:
            qual = None
        else:
            qual = [int(x) for x in rng.choice([2, 8, 12, 13, 14, 25, 37, 40, 41, 60, 93, 100], lseq)]
        flag = int(rng.choice([0, 16, 0x4, 0x100, 0x400, 0x800, 1 | 2 | 64], p=[0.45, 0.3, 0.03, 0.04, 0.04, 0.04, 0.1]))
        tags = [("NH", "i", 1), ("xf", "B", [1, 2, 3])]
        u = rng.random()
        if u < 0.9:
            tags.append(("CB", "Z", barcodes[rng.integers(0, ncells)]))
        elif u < 0.95:
            tags.append(("CB", "A", "x"))                 # wrong type: treated as missing
        if rng.random() < 0.92:
            tags.append(("UB", "Z", "".join(rng.choice(list("ACGT"), 3))))
        tags.append(("RG", "Z", "grp"))
        reads.append({"tid": tid, "pos": pos, "mapq": int(rng.choice([0, 3, 19, 20, 60, 255])), "flag": flag, "cigar": cigar,
                      "seq": seq, "qual": qual, "tags": tags})
    reads.sort(key=lambda r: (r["tid"], r["pos"]))
    return reads, barcodes


# ---- the reference's own CIGAR look-up (oracle/_ref/ref_cigar) ---------------------------------------------------
import os as _os
import subprocess as _subprocess

REF_CIGAR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "oracle", "_ref", "ref_cigar")


def lookup(read, spos):
    """python restatement of bam_get_base_and_qual_and_read_and_qual (hts_utils.cpp:279-359): (base, qual char as a
    signed value, rpos); rpos -1 = BAM_READ_INDEX_NA"""
    rlen = len(read["seq"])
    base, qual, rpos = "N", 0, 0
    if read["cigar"]:
        cpos = read["pos"]
        for op, ln in read["cigar"]:
            if op == "M":
                if cpos <= spos <= cpos + ln - 1:
                    rpos += spos - cpos
                    break
                cpos += ln
                rpos += ln
            elif op in "DN":
                if cpos <= spos <= cpos + ln - 1:
                    rpos = -1
                    break
                cpos += ln
            elif op in "SI":
                rpos += ln
        if 0 <= rpos <= rlen:
            base = read["seq"][rpos] if rpos < rlen else "\0"
            raw = 0 if rpos >= rlen else (0xff if read["qual"] is None else read["qual"][rpos]) + 33
            raw &= 0xff
            qual = raw - 256 if raw >= 128 else raw
        else:
            rpos = -1
    if rpos >= rlen:
        rpos, base = -1, "."
    return base, qual, rpos


def run_ref_cigar(reads, queries, tmpdir):
    """[(base, qual, rpos)] per read and query position from the reference's function"""
    fin, fout = _os.path.join(tmpdir, "cases.bin"), _os.path.join(tmpdir, "cases.out")
    with open(fin, "wb") as f:
        f.write(struct.pack("<i", len(reads)))
        for r, qs in zip(reads, queries):
            f.write(struct.pack("<ii", r["pos"], len(r["cigar"])))
            for op, ln in r["cigar"]:
                f.write(struct.pack("<I", (ln << 4) | CIGAR_OPS.index(op)))
            seq = r["seq"]
            sq = bytearray((len(seq) + 1) // 2)
            for k, c in enumerate(seq):
                sq[k >> 1] |= SEQ_CODE[c] << (4 if k % 2 == 0 else 0)
            f.write(struct.pack("<i", len(seq)) + bytes(sq))
            f.write(bytes(r["qual"]) if r["qual"] is not None else b"\xff" * len(seq))
            f.write(struct.pack("<i", len(qs)) + b"".join(struct.pack("<I", q) for q in qs))
    _subprocess.run([REF_CIGAR, fin, fout], check=True)
    raw = open(fout, "rb").read()
    out, k = [], 0
    for qs in queries:
        row = []
        for _ in qs:
            b, q, rp = struct.unpack_from("<bbi", raw, k)
            k += 6
            row.append((chr(b & 0xff), q, rp))
        out.append(row)
    return out


def cigar_cases(rng, n=300):
    """reads with every CIGAR operation, also ones the reference's walk ignores (=, X, H, P), and their query positions"""
    shapes = [[("M", 40)], [("S", 5), ("M", 20), ("I", 2), ("M", 13)], [("M", 10), ("D", 3), ("M", 30)],
              [("M", 12), ("N", 40), ("M", 28)], [("H", 4), ("S", 2), ("M", 30), ("S", 8)], [],
              [("=", 10), ("X", 1), ("M", 29)], [("M", 10), ("P", 2), ("I", 1), ("M", 29)], [("S", 40)], [("M", 1)],
              [("I", 3), ("M", 37)], [("M", 20), ("D", 1), ("I", 1), ("M", 19)]]
    reads, queries = [], []
    for i in range(n):
        cigar = shapes[i % len(shapes)]
        lseq = sum(ln for op, ln in cigar if op in "MIS=X") or 20
        pos = int(rng.integers(0, 500))
        seq = "".join(rng.choice(list("ACGTN"), lseq, p=[0.24, 0.24, 0.24, 0.24, 0.04]))
        qual = None if i % 17 == 3 else [int(x) for x in rng.choice([0, 12, 13, 30, 40, 41, 94, 95, 120, 222, 223, 254], lseq)]
        span = sum(ln for op, ln in cigar if op in "MDN=X")
        reads.append({"tid": 0, "pos": pos, "mapq": 60, "flag": 0, "cigar": cigar, "seq": seq, "qual": qual,
                      "tags": [("CB", "Z", "cell%d" % (i % 7)), ("UB", "Z", "u%d" % i)]})
        queries.append(list(range(max(0, pos - 3), pos + max(span, 1) + 4)))
    order = sorted(range(n), key=lambda i: reads[i]["pos"])
    return [reads[i] for i in order], [queries[i] for i in order]
