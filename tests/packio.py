"""read-list / CSR files exchanged with oracle/_ref/ref_pack (the reference's own add_read) and synthetic read lists"""
import json
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PACK = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ref_pack")


def make_reads(n, ncells, nsnps, seed, dup_frac=0.3, umi_lens=(6, 12)):
    """reads in arrival order with repeated (cell, snp, umi) keys whose allele / quality differ between arrivals,
    UMIs of several lengths (std::string order: a prefix sorts first) and a few empty cells"""
    rng = np.random.default_rng(seed)
    cell = rng.integers(0, max(1, ncells - 2), size=n).astype(np.int32)     # the last two cells stay empty
    snp = rng.integers(0, nsnps, size=n).astype(np.int32)
    alphabet = np.frombuffer(b"ACGT", dtype=np.uint8)
    umis = []
    for i in range(n):
        ln = int(rng.integers(umi_lens[0], umi_lens[1] + 1))
        umis.append(bytes(alphabet[rng.integers(0, 4, size=ln)]).decode())
    ndup = int(n * dup_frac)
    src = rng.integers(0, n, size=ndup)
    dst = rng.integers(0, n, size=ndup)
    for s, d in zip(src, dst):            # same key, (usually) different allele / quality, later arrival loses
        cell[d], snp[d], umis[d] = cell[s], snp[s], umis[s]
    for d in dst[: ndup // 4]:            # and some keys that differ only by a UMI suffix
        umis[d] = umis[d][: umi_lens[0]]
    allele = rng.integers(0, 3, size=n).astype(np.uint8)
    bq = rng.integers(13, 41, size=n).astype(np.uint8)
    return cell, snp, umis, allele, bq


def write_rds(path, cell, snp, umis, allele, bq, ncells, nsnps):
    enc = [u.encode() for u in umis]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in enc])
    with open(path, "wb") as f:
        f.write(b"DMXRDS01")
        f.write(struct.pack("<3q", len(enc), ncells, nsnps))
        for a in (np.asarray(cell, np.int32), np.asarray(snp, np.int32), np.asarray(allele, np.uint8),
                  np.asarray(bq, np.uint8), off):
            f.write(a.tobytes())
        f.write(b"".join(enc))


def read_rds(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"DMXRDS01"
        n, ncells, nsnps = struct.unpack("<3q", f.read(24))
        cell = np.frombuffer(f.read(4 * n), np.int32)
        snp = np.frombuffer(f.read(4 * n), np.int32)
        allele = np.frombuffer(f.read(n), np.uint8)
        bq = np.frombuffer(f.read(n), np.uint8)
        off = np.frombuffer(f.read(8 * (n + 1)), np.int64)
        buf = f.read(int(off[-1]))
    umis = [buf[off[i]:off[i + 1]].decode() for i in range(n)]
    return cell, snp, umis, allele, bq, ncells, nsnps


def read_csr(path):
    with open(path, "rb") as f:
        assert f.read(8) == b"DMXCSR01"
        ncells, E, R = struct.unpack("<3q", f.read(24))
        out = {"cell_ptr": np.frombuffer(f.read(8 * (ncells + 1)), np.int64),
               "ent_snp": np.frombuffer(f.read(4 * E), np.int32),
               "ent_ptr": np.frombuffer(f.read(8 * (E + 1)), np.int64),
               "reads": np.frombuffer(f.read(2 * R), np.uint16),
               "npass": np.frombuffer(f.read(4 * ncells), np.int32),
               "uniq": np.frombuffer(f.read(4 * ncells), np.int32)}
    return out


def run_ref_pack(rds, csr):
    r = subprocess.run([REF_PACK, rds, csr], check=True, stdout=subprocess.PIPE)
    return json.loads(r.stdout.decode().strip().splitlines()[-1])
