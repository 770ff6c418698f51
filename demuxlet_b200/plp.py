"""Flat pileup container (the data that crosses the C-ABI in include/dmx.h).

It is the CSR flattening of the reference's ``sc_dropseq_lib_t``
(/root/reference/sc_drop_seq.h:34-58): ``cell_umis[cell][snp] -> {umi -> allele<<24|bq<<16|n}``
becomes

    cell_ptr[ncells+1] -> entries;  ent_snp[E] (ascending within a cell, like std::map<int32_t,..>)
    ent_ptr[E+1]       -> reads;    reads[R] = allele<<8 | bq  (UMI-string order inside an entry,
                                    like std::map<std::string,uint32_t>, sc_drop_seq.h:21)

plus the per-SNP x per-sample genotype probability matrix ``gp[nsnps][nv][3]`` (float32: the
reference stores doubles converted from htslib floats, cmd_cram_demuxlet.cpp:227-231) and the three
per-cell read counters (sc_drop_seq.h:47-49).
"""
from __future__ import annotations

import dataclasses
import struct
from typing import List, Sequence

import numpy as np

MAGIC = b"DMXPLP01"


@dataclasses.dataclass
class Pileup:
    gp: np.ndarray          # float32 [nsnps, nv, 3]
    cell_ptr: np.ndarray    # int64 [ncells+1]
    ent_snp: np.ndarray     # int32 [E]
    ent_ptr: np.ndarray     # int64 [E+1]
    reads: np.ndarray       # uint16 [R]   allele<<8 | bq
    totl: np.ndarray        # int32 [ncells]
    npass: np.ndarray       # int32 [ncells]
    uniq: np.ndarray        # int32 [ncells]
    barcodes: List[str]
    samples: List[str]

    def __post_init__(self):
        self.gp = np.ascontiguousarray(self.gp, dtype=np.float32)
        self.cell_ptr = np.ascontiguousarray(self.cell_ptr, dtype=np.int64)
        self.ent_snp = np.ascontiguousarray(self.ent_snp, dtype=np.int32)
        self.ent_ptr = np.ascontiguousarray(self.ent_ptr, dtype=np.int64)
        self.reads = np.ascontiguousarray(self.reads, dtype=np.uint16)
        self.totl = np.ascontiguousarray(self.totl, dtype=np.int32)
        self.npass = np.ascontiguousarray(self.npass, dtype=np.int32)
        self.uniq = np.ascontiguousarray(self.uniq, dtype=np.int32)
        self.validate()

    # ---- shapes -------------------------------------------------------------------------------
    @property
    def nsnps(self) -> int:
        return int(self.gp.shape[0])

    @property
    def nv(self) -> int:
        return int(self.gp.shape[1])

    @property
    def ncells(self) -> int:
        return int(self.cell_ptr.shape[0] - 1)

    @property
    def n_entries(self) -> int:
        return int(self.ent_snp.shape[0])

    @property
    def n_reads(self) -> int:
        return int(self.reads.shape[0])

    def validate(self) -> None:
        assert self.gp.ndim == 3 and self.gp.shape[2] == 3, "gp must be [nsnps][nv][3]"
        assert self.cell_ptr[0] == 0 and self.cell_ptr[-1] == self.n_entries
        assert self.ent_ptr.shape[0] == self.n_entries + 1
        assert self.ent_ptr[0] == 0 and self.ent_ptr[-1] == self.n_reads
        assert len(self.barcodes) == self.ncells and len(self.samples) == self.nv
        assert self.totl.shape[0] == self.ncells == self.npass.shape[0] == self.uniq.shape[0]
        if self.n_entries:
            assert int(self.ent_snp.min()) >= 0 and int(self.ent_snp.max()) < self.nsnps

    # ---- on-disk form (shared with oracle/ref_harness.cpp and the dmx_demuxlet CLI) -----------
    def save(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write(struct.pack("<5q", self.nv, self.nsnps, self.ncells, self.n_entries, self.n_reads))
            for a in (self.gp, self.cell_ptr, self.ent_snp, self.ent_ptr, self.reads,
                      self.totl, self.npass, self.uniq):
                f.write(a.tobytes())
            for s in list(self.barcodes) + list(self.samples):
                f.write(s.encode() + b"\0")

    @staticmethod
    def load(path: str) -> "Pileup":
        with open(path, "rb") as f:
            if f.read(8) != MAGIC:
                raise ValueError("not a DMXPLP01 file: %s" % path)
            nv, nsnps, ncells, E, R = struct.unpack("<5q", f.read(40))

            def arr(dt, n):
                return np.frombuffer(f.read(np.dtype(dt).itemsize * n), dtype=dt).copy()

            gp = arr(np.float32, nsnps * nv * 3).reshape(nsnps, nv, 3)
            cell_ptr = arr(np.int64, ncells + 1)
            ent_snp = arr(np.int32, E)
            ent_ptr = arr(np.int64, E + 1)
            reads = arr(np.uint16, R)
            totl = arr(np.int32, ncells)
            npass = arr(np.int32, ncells)
            uniq = arr(np.int32, ncells)
            names = f.read().split(b"\0")
            names = [n.decode() for n in names[: ncells + nv]]
        return Pileup(gp, cell_ptr, ent_snp, ent_ptr, reads, totl, npass, uniq,
                      names[:ncells], names[ncells:])

    # ---- views --------------------------------------------------------------------------------
    def subset_cells(self, cells: Sequence[int], compact_snps: bool = False) -> "Pileup":
        """Pileup restricted to ``cells`` (in the given order).  ``compact_snps`` renumbers the
        SNPs that remain (ascending, so per-cell SNP order is preserved): this is how a bounded
        sample of a big configuration is fed to the unmodified reference code, whose int32 table
        index overflows at nsnps*nv*nv*9 >= 2^31."""
        cells = np.asarray(cells, dtype=np.int64)
        e_lo, e_hi = self.cell_ptr[cells], self.cell_ptr[cells + 1]
        n_e = e_hi - e_lo
        cell_ptr = np.concatenate([[0], np.cumsum(n_e)]).astype(np.int64)
        eidx = _ranges(e_lo, n_e)
        ent_snp = self.ent_snp[eidx]
        r_lo, r_hi = self.ent_ptr[eidx], self.ent_ptr[eidx + 1]
        n_r = r_hi - r_lo
        ent_ptr = np.concatenate([[0], np.cumsum(n_r)]).astype(np.int64)
        reads = self.reads[_ranges(r_lo, n_r)]
        gp = self.gp
        if compact_snps:
            used = np.unique(ent_snp)
            remap = np.full(self.nsnps, -1, dtype=np.int32)
            remap[used] = np.arange(used.shape[0], dtype=np.int32)
            ent_snp = remap[ent_snp]
            gp = self.gp[used]
        return Pileup(gp, cell_ptr, ent_snp, ent_ptr, reads, self.totl[cells], self.npass[cells],
                      self.uniq[cells], [self.barcodes[int(c)] for c in cells], list(self.samples))


def _ranges(starts: np.ndarray, lengths: np.ndarray) -> np.ndarray:
    """concatenate [s, s+len) ranges, vectorised"""
    lengths = np.asarray(lengths, dtype=np.int64)
    total = int(lengths.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    offs = np.repeat(np.asarray(starts, dtype=np.int64) - np.concatenate([[0], np.cumsum(lengths)[:-1]]),
                     lengths)
    return offs + np.arange(total, dtype=np.int64)


def build_from_reads(nsnps_gp: np.ndarray, cell: np.ndarray, snp: np.ndarray, umi: Sequence[str],
                     allele: np.ndarray, bq: np.ndarray, barcodes: List[str], samples: List[str],
                     totl: np.ndarray | None = None) -> Pileup:
    """Reference-order pileup builder (pure Python, small inputs; the CSR packer that mirrors
    sc_dropseq_lib_t::add_read, /root/reference/sc_drop_seq.cpp:34-77): reads arrive in BAM order;
    the first read of a (cell,snp,umi) keeps its allele/bq, later ones only bump the count (:51-58);
    entries are SNP-ascending per cell and UMI-string-ascending per entry."""
    ncells = len(barcodes)
    table: dict = {}
    npass = np.zeros(ncells, dtype=np.int32)
    uniq = np.zeros(ncells, dtype=np.int32)
    for c, s, u, a, q in zip(cell, snp, umi, allele, bq):
        npass[c] += 1
        d = table.setdefault((int(c), int(s)), {})
        if u not in d:
            d[u] = (int(a) << 8) | int(q)
            uniq[c] += 1
    cell_ptr = [0]
    ent_snp, ent_ptr, reads = [], [0], []
    keys = sorted(table.keys())
    ki = 0
    for c in range(ncells):
        while ki < len(keys) and keys[ki][0] == c:
            d = table[keys[ki]]
            ent_snp.append(keys[ki][1])
            for u in sorted(d.keys(), key=lambda x: x.encode()):
                reads.append(d[u])
            ent_ptr.append(len(reads))
            ki += 1
        cell_ptr.append(len(ent_snp))
    if totl is None:
        totl = npass.copy()
    return Pileup(nsnps_gp, np.array(cell_ptr), np.array(ent_snp, dtype=np.int32), np.array(ent_ptr),
                  np.array(reads, dtype=np.uint16), totl, npass, uniq, barcodes, samples)
