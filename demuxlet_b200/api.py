"""ctypes binding of libdmx.so (include/dmx.h) and the host-side driver object.

`Demuxlet` mirrors what the reference's main() does after the pileup is built
(/root/reference/cmd_cram_demuxlet.cpp:386-875): it takes the flat pileup (demuxlet_b200.plp.Pileup),
the alpha grid / doublet prior / min-filters of the reference CLI (:54-56, :69-71), runs the likelihood
scan on the GPU and writes <out>.single/.sing2/.best[/.pair].  There is no CPU fallback: a missing
library or a machine without an sm_100 GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from .plp import Pileup

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMX_LIB") or os.path.join(HERE, "libdmx.so")   # DMX_LIB: experiment builds (tools/)

DMX_EXACT_TIES = 0x1
DMX_BORROW_INPUTS = 0x2
DMX_FP32_PAIRS = 0x4
DMX_SKIP_SINGLE = 0x8
DMX_TRUST_INPUTS = 0x10
DMX_HOST_FINALISE = 0x20

# every symbol include/dmx.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "dmx_config_default", "dmx_device_count", "dmx_create", "dmx_destroy", "dmx_set_gp", "dmx_set_pileup",
    "dmx_create_sharded", "dmx_shard_count", "dmx_shard_range", "dmx_shard_stats",
    "dmx_upload", "dmx_compute", "dmx_run", "dmx_get_best", "dmx_get_single", "dmx_get_sing2", "dmx_get_pair",
    "dmx_rescore", "dmx_exact_llk", "dmx_write_outputs", "dmx_get_stats", "dmx_alloc_host", "dmx_free_host", "dmx_strerror",
    "dmx_last_error", "dmx_abi_version", "dmx_pack_reads", "dmx_pileup_sizes", "dmx_pileup_copy", "dmx_pileup_free",
    "dmx_build_gp", "dmx_build_gp_device", "dmx_set_field", "dmx_site_filter",
    "dmx_sweep_opts_default", "dmx_sweep_bam", "dmx_sweep_sizes", "dmx_sweep_counts", "dmx_sweep_copy", "dmx_sweep_barcode",
    "dmx_sweep_free", "dmx_sweep_last_error", "dmx_vcf_read", "dmx_vcf_read_key", "dmx_vcf_select", "dmx_vcf_sizes", "dmx_vcf_sample", "dmx_vcf_chrom", "dmx_vcf_copy", "dmx_vcf_no_field", "dmx_vcf_free", "dmx_vcf_last_error",
]

DMX_FIELD = {"GT": 0, "GP": 1, "PL": 2}            # --field (cmd_cram_demuxlet.cpp:28, :45)
_FIELD_DTYPE = {0: np.int8, 1: np.float32, 2: np.int32}


class DmxFieldStats(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("gpu_launches", C.c_int64),
                ("algo_bytes", C.c_double)]


class DmxConfig(C.Structure):
    _fields_ = [
        ("nv", C.c_int32), ("n_alpha", C.c_int32), ("alpha", C.POINTER(C.c_double)),
        ("doublet_prior", C.c_double),
        ("min_total", C.c_int32), ("min_uniq", C.c_int32), ("min_snp", C.c_int32),
        ("device", C.c_int32), ("flags", C.c_uint32), ("n_host_threads", C.c_int32),
    ]


BEST_DTYPE = np.dtype([
    ("status", "<i4"), ("n_snp", "<i4"), ("sng1", "<i4"), ("sng2", "<i4"),
    ("dbl1", "<i4"), ("dbl2", "<i4"), ("alpha_idx", "<i4"), ("call", "<i4"),
    ("sng_llk1", "<f8"), ("sng_llk2", "<f8"), ("sng_llk0", "<f8"),
    ("llk12", "<f8"), ("llk1", "<f8"), ("llk2", "<f8"), ("llk10", "<f8"), ("llk20", "<f8"), ("llk00", "<f8"),
    ("prb_dbl", "<f8"), ("prb_sng1", "<f8"),
    ("max_llk", "<f8"), ("sum_single", "<f8"), ("sum_double", "<f8"),
    ("n_rescored", "<i4"), ("exact_mask", "<u4"),
], align=True)


class DmxStats(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_double), ("prep_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
        ("finalise_ms", C.c_double), ("total_ms", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("n_entries", C.c_int64), ("n_reads", C.c_int64), ("n_tasks", C.c_int64),
        ("gpu_launches", C.c_int64), ("n_rescored", C.c_int64), ("n_fullcell", C.c_int64),
        ("n_rescored_gpu", C.c_int64), ("gpu_exact", C.c_int32), ("pad_", C.c_int32),
        ("algo_bytes", C.c_double), ("pair_kernel_ms", C.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class DmxError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str = ""):
        self.status = status
        super().__init__("%s failed: %s (%d)%s" % (where, _strerror(status), status, (" -- " + detail) if detail else ""))


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libdmx.so; raises OSError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    lib = C.CDLL(path or LIB_PATH)
    lib.dmx_strerror.restype = C.c_char_p
    lib.dmx_last_error.restype = C.c_char_p
    lib.dmx_alloc_host.restype = C.c_void_p
    lib.dmx_alloc_host.argtypes = [C.c_int64]
    lib.dmx_free_host.argtypes = [C.c_void_p]
    lib.dmx_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(DmxConfig)]
    lib.dmx_create_sharded.argtypes = [C.POINTER(C.c_void_p), C.POINTER(DmxConfig), C.c_int32, C.POINTER(C.c_int32)]
    lib.dmx_shard_count.argtypes = [C.c_void_p]
    lib.dmx_shard_range.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    lib.dmx_shard_stats.argtypes = [C.c_void_p, C.c_int32, C.POINTER(DmxStats)]
    lib.dmx_destroy.argtypes = [C.c_void_p]
    lib.dmx_destroy.restype = None
    lib.dmx_set_gp.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.dmx_set_pileup.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7
    for f in ("dmx_upload", "dmx_compute", "dmx_run"):
        getattr(lib, f).argtypes = [C.c_void_p]
    lib.dmx_get_best.argtypes = [C.c_void_p, C.c_void_p]
    lib.dmx_get_single.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dmx_get_sing2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dmx_get_pair.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.dmx_rescore.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    lib.dmx_write_outputs.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_char_p, C.c_int]
    lib.dmx_get_stats.argtypes = [C.c_void_p, C.POINTER(DmxStats)]
    lib.dmx_config_default.argtypes = [C.POINTER(DmxConfig)]
    lib.dmx_config_default.restype = None
    lib.dmx_build_gp.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_double, C.c_void_p,
                                 C.c_void_p, C.POINTER(DmxFieldStats)]
    lib.dmx_build_gp_device.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_double, C.c_void_p,
                                        C.c_int64, C.c_void_p, C.c_void_p]
    lib.dmx_set_field.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_double]
    lib.dmx_site_filter.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_double, C.c_void_p]
    if path is None:
        _lib = lib
    return lib


def _strerror(status: int) -> str:
    try:
        return load_library().dmx_strerror(status).decode()
    except OSError:
        return "status %d" % status


def device_count() -> int:
    return int(load_library().dmx_device_count())


def exact_llk(plp: Pileup, alphas: Sequence[float], cell: int, j: int, k: int, n: int) -> float:
    """llksAB[j][k][n] of one cell in the reference's exact arithmetic (host only; needs no GPU)"""
    lib = load_library()
    al = np.ascontiguousarray(alphas, dtype=np.float64)
    out = C.c_double()
    lib.dmx_exact_llk.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]
    rc = lib.dmx_exact_llk(plp.nv, len(al), al.ctypes.data, plp.gp.ctypes.data, plp.cell_ptr.ctypes.data,
                           plp.ent_snp.ctypes.data, plp.ent_ptr.ctypes.data, plp.reads.ctypes.data, cell, j, k, n,
                           C.byref(out))
    if rc != 0:
        raise DmxError(rc, "dmx_exact_llk")
    return out.value


def _field_array(field, arr):
    code = DMX_FIELD[field] if isinstance(field, str) else int(field)
    a = np.ascontiguousarray(arr, dtype=_FIELD_DTYPE[code])
    if a.ndim != 3 or a.shape[2] != (2 if code == 0 else 3):
        raise ValueError("field array must be [nsnps][nv][%d]" % (2 if code == 0 else 3))
    return code, a


def build_gp(field, arr, gt_error: float = 0.01, device: int = 0, want_af: bool = False, out=None,
             want_stats: bool = False):
    """VCF FORMAT field -> float32[nsnps][nv][3] on the GPU, bit-identical to the reference's
    BCFFilteredReader::parse_posteriors (/root/reference/bcf_filtered_reader.cpp:360-454).
    field "GT": int8[nsnps][nv][2] allele indices (< 0 missing); "GP": float32[nsnps][nv][3]; "PL": int32[nsnps][nv][3].
    No CPU fallback: raises DmxError without an sm_100 device."""
    lib = load_library()
    code, a = _field_array(field, arr)
    nsnps, nv = a.shape[0], a.shape[1]
    if out is None:
        out = np.empty((nsnps, nv, 3), np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (nsnps, nv, 3)
    af = np.zeros(nsnps, np.float64) if want_af else None
    st = DmxFieldStats()
    rc = lib.dmx_build_gp(device, code, nv, nsnps, a.ctypes.data, gt_error, out.ctypes.data,
                          af.ctypes.data if want_af else None, C.byref(st))
    if rc != 0:
        raise DmxError(rc, "dmx_build_gp")
    res = (out, af) if want_af else out
    if want_stats:
        return res, {k: getattr(st, k) for k, _ in DmxFieldStats._fields_}
    return res


def site_filter(gt, min_mac: int = 1, min_callrate: float = 0.5, device: int = 0) -> np.ndarray:
    """the reference's per-record variant filter (passed_vfilter, /root/reference/bcf_filtered_reader.cpp:540-573)
    with demuxlet's defaults (cmd_cram_demuxlet.cpp:24-25): bool[nsnps] from int8[nsnps][nv][2] allele indices"""
    lib = load_library()
    _, a = _field_array("GT", gt)
    out = np.zeros(a.shape[0], np.uint8)
    rc = lib.dmx_site_filter(device, a.shape[1], a.shape[0], a.ctypes.data, min_mac, min_callrate, out.ctypes.data)
    if rc != 0:
        raise DmxError(rc, "dmx_site_filter")
    return out.astype(bool)


def read_vcf(path: str, field: str = "GT", samples: Optional[Sequence[str]] = None, key: Optional[str] = None) -> dict:
    """dmx_vcf_read: the biallelic records of a VCF (plain / gzip / bgzip) as the flat arrays site_filter / build_gp
    take.  Returns samples, chroms, rid, pos (0-based), ref, alt, gt int8[nrec][nv][2], values (GT: gt; GP float32;
    PL int32 [nrec][nv][3]), n_skipped, no_field (uint8[nrec]: the record's FORMAT lacks `field`; an error for the
    reference only if the record passes site_filter, /root/reference/cmd_cram_demuxlet.cpp:153,:211).
    key: FORMAT key to read as `field` (the reference treats any --field but GT / PL as posteriors like GP)."""
    lib = load_library()
    lib.dmx_vcf_read_key.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.dmx_vcf_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int32)]
    lib.dmx_vcf_sample.argtypes = [C.c_void_p, C.c_int32]
    lib.dmx_vcf_sample.restype = C.c_char_p
    lib.dmx_vcf_chrom.argtypes = [C.c_void_p, C.c_int32]
    lib.dmx_vcf_chrom.restype = C.c_char_p
    lib.dmx_vcf_copy.argtypes = [C.c_void_p] * 7
    lib.dmx_vcf_no_field.argtypes = [C.c_void_p] * 2
    lib.dmx_vcf_free.argtypes = [C.c_void_p]
    lib.dmx_vcf_free.restype = None
    lib.dmx_vcf_last_error.restype = C.c_char_p
    code = DMX_FIELD[field]
    h = C.c_void_p()
    rc = lib.dmx_vcf_read_key(os.fsencode(path), code, key.encode() if key else None, C.byref(h))
    if rc != 0:
        raise DmxError(rc, "dmx_vcf_read", lib.dmx_vcf_last_error().decode())
    try:
        if samples:      # --sm / --sm-list
            lib.dmx_vcf_select.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p)]
            arr = (C.c_char_p * len(samples))(*[x.encode() for x in samples])
            rc = lib.dmx_vcf_select(h, len(samples), arr)
            if rc != 0:
                raise DmxError(rc, "dmx_vcf_select", lib.dmx_vcf_last_error().decode())
        nv, nrec, nskip, nchrom = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32()
        lib.dmx_vcf_sizes(h, C.byref(nv), C.byref(nrec), C.byref(nskip), C.byref(nchrom))
        nv, nrec = nv.value, nrec.value
        gt = np.zeros((nrec, nv, 2), np.int8)
        vals = gt if code == 0 else np.zeros((nrec, nv, 3), _FIELD_DTYPE[code])
        rid, pos = np.zeros(nrec, np.int32), np.zeros(nrec, np.int32)
        ref, alt = np.zeros(nrec, "S1"), np.zeros(nrec, "S1")
        lib.dmx_vcf_copy(h, gt.ctypes.data, None if code == 0 else vals.ctypes.data, rid.ctypes.data, pos.ctypes.data,
                         ref.ctypes.data, alt.ctypes.data)
        no_field = np.zeros(nrec, np.uint8)
        lib.dmx_vcf_no_field(h, no_field.ctypes.data)
        return {"samples": [lib.dmx_vcf_sample(h, i).decode() for i in range(nv)], "no_field": no_field,
                "chroms": [lib.dmx_vcf_chrom(h, i).decode() for i in range(nchrom.value)],
                "rid": rid, "pos": pos, "ref": ref, "alt": alt, "gt": gt, "values": vals, "n_skipped": nskip.value}
    finally:
        lib.dmx_vcf_free(h)


class DmxSweepOpts(C.Structure):
    _fields_ = [("tag_group", C.c_char_p), ("tag_umi", C.c_char_p), ("min_mq", C.c_int32), ("exclude_flag", C.c_uint32),
                ("min_bq", C.c_int32), ("cap_bq", C.c_int32), ("min_td", C.c_int32), ("n_groups", C.c_int64),
                ("groups", C.POINTER(C.c_char_p))]


def sweep_bam(path: str, chroms: Sequence[str], snp_rid, snp_pos, snp_ref, snp_alt, groups: Optional[Sequence[str]] = None,
              raw_umis: bool = False, **opts) -> dict:
    """dmx_sweep_bam: the reference's BAM sweep (/root/reference/cmd_cram_demuxlet.cpp:186-338) as flat arrays: one
    (cell, snp, umi, allele, bq) per add_read call in call order, barcodes in add_cell order, RD.TOTL.
    opts: tag_group, tag_umi, min_mq, exclude_flag, min_bq, cap_bq, min_td.  raw_umis: "umis" as the (uint8 buffer,
    int64 offsets) pair pack_reads takes instead of a list of strings (no per-read python object)."""
    lib = load_library()
    lib.dmx_sweep_opts_default.argtypes = [C.POINTER(DmxSweepOpts)]
    lib.dmx_sweep_opts_default.restype = None
    lib.dmx_sweep_bam.argtypes = [C.c_char_p, C.POINTER(DmxSweepOpts), C.c_int32, C.POINTER(C.c_char_p), C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.dmx_sweep_sizes.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 3
    lib.dmx_sweep_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int64)] * 4
    lib.dmx_sweep_copy.argtypes = [C.c_void_p] * 8
    lib.dmx_sweep_barcode.argtypes = [C.c_void_p, C.c_int64]
    lib.dmx_sweep_barcode.restype = C.c_char_p
    lib.dmx_sweep_free.argtypes = [C.c_void_p]
    lib.dmx_sweep_free.restype = None
    lib.dmx_sweep_last_error.restype = C.c_char_p
    o = DmxSweepOpts()
    lib.dmx_sweep_opts_default(C.byref(o))
    for k, v in opts.items():
        setattr(o, k, v.encode() if isinstance(v, str) else v)
    keep = [g.encode() for g in (groups or [])]
    garr = (C.c_char_p * max(len(keep), 1))(*keep)
    o.n_groups, o.groups = len(keep), garr
    carr = (C.c_char_p * max(len(chroms), 1))(*[c.encode() for c in chroms])
    rid = np.ascontiguousarray(snp_rid, np.int32)
    pos = np.ascontiguousarray(snp_pos, np.int32)
    ref = np.ascontiguousarray(snp_ref, "S1")
    alt = np.ascontiguousarray(snp_alt, "S1")
    h = C.c_void_p()
    rc = lib.dmx_sweep_bam(os.fsencode(path), C.byref(o), len(chroms), carr, len(pos), rid.ctypes.data, pos.ctypes.data,
                           ref.ctypes.data, alt.ctypes.data, C.byref(h))
    if rc != 0:
        raise DmxError(rc, "dmx_sweep_bam", lib.dmx_sweep_last_error().decode())
    try:
        nc, nr, ub = C.c_int64(), C.c_int64(), C.c_int64()
        lib.dmx_sweep_sizes(h, C.byref(nc), C.byref(nr), C.byref(ub))
        cnt = [C.c_int64() for _ in range(4)]
        lib.dmx_sweep_counts(h, *[C.byref(c) for c in cnt])
        nc, nr, ub = nc.value, nr.value, ub.value
        cell, snp = np.zeros(nr, np.int32), np.zeros(nr, np.int32)
        allele, bq = np.zeros(nr, np.uint8), np.zeros(nr, np.uint8)
        umi_buf, umi_off, totl = np.zeros(max(ub, 1), np.uint8), np.zeros(nr + 1, np.int64), np.zeros(nc, np.int32)
        lib.dmx_sweep_copy(h, cell.ctypes.data, snp.ctypes.data, allele.ctypes.data, bq.ctypes.data, umi_buf.ctypes.data,
                           umi_off.ctypes.data, totl.ctypes.data)
        raw = umi_buf.tobytes()
        return {"barcodes": [lib.dmx_sweep_barcode(h, i).decode() for i in range(nc)], "totl": totl, "cell": cell,
                "snp": snp, "allele": allele, "bq": bq,
                "umis": (umi_buf, umi_off) if raw_umis else [raw[umi_off[i]:umi_off[i + 1]].decode() for i in range(nr)],
                "counts": dict(zip(("records", "filtered", "off_vcf", "skip_bcd"), [c.value for c in cnt]))}
    finally:
        lib.dmx_sweep_free(h)


def encode_umis(umis: Sequence[str]):
    """list of UMI strings -> (byte buffer, int64 offsets) as dmx_pack_reads takes them"""
    enc = [u.encode() for u in umis]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        off[1:] = np.cumsum([len(b) for b in enc])
    buf = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
    return buf, off


def pack_reads(cell, snp, umis, allele, bq, ncells: int, threads: int = 0) -> dict:
    """Sort-based pileup packer (host only, needs no GPU): reads in arrival order -> the CSR arrays of
    dmx_set_pileup plus cell_pass_reads / cell_uniq_reads, with the semantics of the reference's add_read
    (sc_drop_seq.cpp:34-77).  `umis`: list of strings, or an (uint8 buffer, int64 offsets) pair."""
    lib = load_library()
    lib.dmx_pack_reads.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
    lib.dmx_pileup_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.dmx_pileup_copy.argtypes = [C.c_void_p] * 7
    lib.dmx_pileup_free.argtypes = [C.c_void_p]
    lib.dmx_pileup_free.restype = None
    cell = np.ascontiguousarray(cell, dtype=np.int32)
    snp = np.ascontiguousarray(snp, dtype=np.int32)
    allele = np.ascontiguousarray(allele, dtype=np.uint8)
    bq = np.ascontiguousarray(bq, dtype=np.uint8)
    buf, off = umis if isinstance(umis, tuple) else encode_umis(umis)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.int64)
    n = int(cell.shape[0])
    assert snp.shape[0] == n and allele.shape[0] == n and bq.shape[0] == n and off.shape[0] == n + 1
    h = C.c_void_p()
    rc = lib.dmx_pack_reads(n, cell.ctypes.data, snp.ctypes.data, allele.ctypes.data, bq.ctypes.data, buf.ctypes.data,
                            off.ctypes.data, int(ncells), int(threads), C.byref(h))
    if rc != 0:
        raise DmxError(rc, "dmx_pack_reads")
    try:
        nc, E, R = C.c_int64(), C.c_int64(), C.c_int64()
        lib.dmx_pileup_sizes(h, C.byref(nc), C.byref(E), C.byref(R))
        out = {"cell_ptr": np.zeros(nc.value + 1, np.int64), "ent_snp": np.zeros(E.value, np.int32),
               "ent_ptr": np.zeros(E.value + 1, np.int64), "reads": np.zeros(R.value, np.uint16),
               "npass": np.zeros(nc.value, np.int32), "uniq": np.zeros(nc.value, np.int32)}
        rc = lib.dmx_pileup_copy(h, out["cell_ptr"].ctypes.data, out["ent_snp"].ctypes.data, out["ent_ptr"].ctypes.data,
                                 out["reads"].ctypes.data, out["npass"].ctypes.data, out["uniq"].ctypes.data)
        if rc != 0:
            raise DmxError(rc, "dmx_pileup_copy")
    finally:
        lib.dmx_pileup_free(h)
    return out


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over cudaMallocHost memory (for DMX_BORROW_INPUTS); freed with the array."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.dmx_alloc_host(max(n, 1))
    if not ptr:
        raise MemoryError("dmx_alloc_host(%d)" % n)
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[ptr] = buf
    import weakref
    weakref.finalize(buf, lib.dmx_free_host, ptr)
    return arr


_PINNED: dict = {}


class Demuxlet:
    """The likelihood scan of one pileup on one GPU, or -- with `devices=[...]` -- on several GPUs of this host over
    contiguous barcode ranges balanced by pileup entries (dmx_create_sharded; same results, same files).

    alphas        --alpha (multi-valued; default {0, 0.5}, cmd_cram_demuxlet.cpp:78-90)
    doublet_prior --doublet-prior (:56)
    min_total/min_uniq/min_snp  --min-total/--min-uniq/--min-snp (:69-71)
    """

    def __init__(self, nv: int, alphas: Sequence[float] = (0.0, 0.5), doublet_prior: float = 0.5,
                 min_total: int = 0, min_uniq: int = 0, min_snp: int = 0, device: int = 0,
                 exact_ties: bool = True, borrow: bool = False, fp32_pairs: bool = False,
                 skip_single: bool = False, trust_inputs: bool = False, host_threads: int = 0,
                 host_finalise: bool = False, devices: Optional[Sequence[int]] = None):
        self.lib = load_library()
        self.nv = int(nv)
        self.alpha = np.ascontiguousarray(alphas, dtype=np.float64)
        self.nA = int(self.alpha.shape[0])
        flags = ((DMX_EXACT_TIES if exact_ties else 0) | (DMX_BORROW_INPUTS if borrow else 0) |
                 (DMX_FP32_PAIRS if fp32_pairs else 0) | (DMX_SKIP_SINGLE if skip_single else 0) |
                 (DMX_TRUST_INPUTS if trust_inputs else 0) | (DMX_HOST_FINALISE if host_finalise else 0))
        self.cfg = DmxConfig(self.nv, self.nA, self.alpha.ctypes.data_as(C.POINTER(C.c_double)), doublet_prior,
                             min_total, min_uniq, min_snp, device, flags, host_threads)
        self._h = C.c_void_p()
        if devices is None:
            self._chk(self.lib.dmx_create(C.byref(self._h), C.byref(self.cfg)), "dmx_create")
        else:
            dv = (C.c_int32 * len(devices))(*[int(x) for x in devices])
            self._chk(self.lib.dmx_create_sharded(C.byref(self._h), C.byref(self.cfg), len(devices), dv),
                      "dmx_create_sharded")
        self._keep = []
        self.ncells = 0
        self.plp: Optional[Pileup] = None

    def _chk(self, rc: int, where: str):
        if rc != 0:
            raise DmxError(rc, where, self.lib.dmx_last_error().decode() if rc in (-2, -3) else "")

    def shards(self):
        """[(cell_lo, cell_hi, device)] of the barcode ranges (one entry for a single-device context)"""
        out = []
        for i in range(int(self.lib.dmx_shard_count(self._h))):
            lo, hi, dev = C.c_int64(), C.c_int64(), C.c_int32()
            self._chk(self.lib.dmx_shard_range(self._h, i, C.byref(lo), C.byref(hi), C.byref(dev)), "dmx_shard_range")
            out.append((lo.value, hi.value, dev.value))
        return out

    def shard_stats(self, i: int) -> dict:
        st = DmxStats()
        self._chk(self.lib.dmx_shard_stats(self._h, i, C.byref(st)), "dmx_shard_stats")
        return st.asdict()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.dmx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs -----------------------------------------------------------------------------------
    def set_gp(self, gp: np.ndarray):
        assert gp.dtype == np.float32 and gp.flags.c_contiguous and gp.shape[1:] == (self.nv, 3)
        self._gp = gp
        self._chk(self.lib.dmx_set_gp(self._h, gp.ctypes.data, gp.shape[0]), "dmx_set_gp")

    def set_field(self, field, arr, gt_error: float = 0.01):
        """set_gp from a VCF field (GT / GP / PL arrays as build_gp takes them): the matrix is built on the device"""
        code, a = _field_array(field, arr)
        assert a.shape[1] == self.nv
        self._chk(self.lib.dmx_set_field(self._h, code, a.ctypes.data, a.shape[0], gt_error), "dmx_set_field")

    def set_pileup_arrays(self, cell_ptr, ent_snp, ent_ptr, reads, totl, npass, uniq):
        self._keep = [cell_ptr, ent_snp, ent_ptr, reads, totl, npass, uniq]
        self.ncells = int(cell_ptr.shape[0] - 1)
        self._chk(self.lib.dmx_set_pileup(self._h, self.ncells, cell_ptr.ctypes.data, ent_snp.ctypes.data,
                                          ent_ptr.ctypes.data, reads.ctypes.data, totl.ctypes.data,
                                          npass.ctypes.data, uniq.ctypes.data), "dmx_set_pileup")

    def set_pileup(self, plp: Pileup):
        assert plp.nv == self.nv
        self.plp = plp
        self.set_gp(plp.gp)
        self.set_pileup_arrays(plp.cell_ptr, plp.ent_snp, plp.ent_ptr, plp.reads, plp.totl, plp.npass, plp.uniq)

    # ---- run --------------------------------------------------------------------------------------
    def upload(self):
        self._chk(self.lib.dmx_upload(self._h), "dmx_upload")

    def compute(self):
        self._chk(self.lib.dmx_compute(self._h), "dmx_compute")

    def run(self):
        self._chk(self.lib.dmx_run(self._h), "dmx_run")

    # ---- results ----------------------------------------------------------------------------------
    def best(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """the per-barcode records; `out`: a caller-owned BEST_DTYPE array of ncells records to fill (e.g. a pinned
        staging buffer of the record exchange) instead of a new one"""
        if out is None:
            out = np.zeros(self.ncells, dtype=BEST_DTYPE)
        assert out.dtype == BEST_DTYPE and out.shape[0] == self.ncells and out.flags["C_CONTIGUOUS"]
        self._chk(self.lib.dmx_get_best(self._h, out.ctypes.data), "dmx_get_best")
        return out

    def single(self):
        llk = np.zeros((self.ncells, self.nv))
        llk0 = np.zeros(self.ncells)
        self._chk(self.lib.dmx_get_single(self._h, llk.ctypes.data, llk0.ctypes.data), "dmx_get_single")
        return llk, llk0

    def sing2(self):
        a = np.zeros((self.ncells, self.nv))
        b = np.zeros((self.ncells, self.nA))
        self._chk(self.lib.dmx_get_sing2(self._h, a.ctypes.data, b.ctypes.data), "dmx_get_sing2")
        return a, b

    def pair(self, lo: int, hi: int) -> np.ndarray:
        out = np.zeros((hi - lo, self.nv, self.nv, self.nA))
        self._chk(self.lib.dmx_get_pair(self._h, lo, hi, out.ctypes.data), "dmx_get_pair")
        return out

    def rescore(self, cell: int, j: int, k: int, n: int) -> float:
        v = C.c_double()
        self._chk(self.lib.dmx_rescore(self._h, cell, j, k, n, C.byref(v)), "dmx_rescore")
        return v.value

    def stats(self) -> dict:
        st = DmxStats()
        self._chk(self.lib.dmx_get_stats(self._h, C.byref(st)), "dmx_get_stats")
        return st.asdict()

    def write_outputs(self, prefix: str, barcodes: Optional[Sequence[str]] = None,
                      samples: Optional[Sequence[str]] = None, write_pair: bool = False):
        barcodes = barcodes if barcodes is not None else self.plp.barcodes
        samples = samples if samples is not None else self.plp.samples
        assert len(barcodes) == self.ncells and len(samples) == self.nv
        bcs = (C.c_char_p * self.ncells)(*[b.encode() for b in barcodes])
        sms = (C.c_char_p * self.nv)(*[s.encode() for s in samples])
        self._chk(self.lib.dmx_write_outputs(self._h, bcs, sms, prefix.encode(), 1 if write_pair else 0),
                  "dmx_write_outputs")
