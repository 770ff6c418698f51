"""ctypes binding of oracle/liboracle.so and a runner for oracle/_ref/ref_harness.

TEST INFRASTRUCTURE ONLY (see oracle/dmx_oracle.h): imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never by demuxlet_b200/.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ref_harness")


class OrcInput(C.Structure):
    _fields_ = [
        ("nv", C.c_int32), ("n_alpha", C.c_int32), ("alpha", C.POINTER(C.c_double)),
        ("doublet_prior", C.c_double),
        ("min_total", C.c_int32), ("min_uniq", C.c_int32), ("min_snp", C.c_int32),
        ("nsnps", C.c_int64), ("gp", C.POINTER(C.c_float)),
        ("ncells", C.c_int64), ("cell_ptr", C.POINTER(C.c_int64)), ("ent_snp", C.POINTER(C.c_int32)),
        ("ent_ptr", C.POINTER(C.c_int64)), ("reads", C.POINTER(C.c_uint16)),
        ("totl", C.POINTER(C.c_int32)), ("npass", C.POINTER(C.c_int32)), ("uniq", C.POINTER(C.c_int32)),
    ]


BEST_DTYPE = np.dtype([
    ("skipped", "<i4"), ("n_snp", "<i4"), ("i_sing1", "<i4"), ("i_sing2", "<i4"),
    ("j_best", "<i4"), ("k_best", "<i4"), ("a_best", "<i4"), ("call", "<i4"),
    ("max_llk", "<f8"), ("sum_single", "<f8"), ("sum_double", "<f8"),
    ("sing_llk1", "<f8"), ("sing_llk2", "<f8"), ("sing_llk0", "<f8"),
    ("pair_llk12", "<f8"), ("pair_llk1", "<f8"), ("pair_llk2", "<f8"),
    ("pair_llk10", "<f8"), ("pair_llk20", "<f8"), ("pair_llk00", "<f8"),
    ("post_doublet", "<f8"), ("post_singlet", "<f8"),
], align=True)

_lib = None


def build() -> None:
    subprocess.run(["make", "-C", HERE, "all"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_time_cells.restype = C.c_double
        _lib.orc_pass2_one.restype = C.c_double
        _lib.orc_count_asym.restype = C.c_int64
        _lib.orc_write_all.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Holds an orc_input over a Pileup (keeps the numpy arrays alive)."""

    def __init__(self, plp, alphas=(0.0, 0.5), doublet_prior=0.5, min_total=0, min_uniq=0, min_snp=0):
        self.plp = plp
        self.alpha = np.ascontiguousarray(alphas, dtype=np.float64)
        self.inp = OrcInput(
            plp.nv, len(self.alpha), _p(self.alpha, C.c_double), doublet_prior,
            min_total, min_uniq, min_snp, plp.nsnps, _p(plp.gp, C.c_float), plp.ncells,
            _p(plp.cell_ptr, C.c_int64), _p(plp.ent_snp, C.c_int32), _p(plp.ent_ptr, C.c_int64),
            _p(plp.reads, C.c_uint16), _p(plp.totl, C.c_int32), _p(plp.npass, C.c_int32),
            _p(plp.uniq, C.c_int32))
        assert BEST_DTYPE.itemsize == 8 * 4 + 14 * 8
        self.nv, self.nA = plp.nv, len(self.alpha)

    def gp0(self):
        out = np.zeros((self.plp.nsnps, 3))
        lib().orc_gp0(C.byref(self.inp), _p(out, C.c_double))
        return out

    def run(self, nthreads=1, want_pass1=True):
        n, nv, nA = self.plp.ncells, self.nv, self.nA
        best = np.zeros(n, dtype=BEST_DTYPE)
        sing = np.zeros((n, nv)); l00 = np.zeros((n, nA))
        p1 = np.zeros((n, nv)); p10 = np.zeros(n)
        lib().orc_run(C.byref(self.inp), C.c_int32(nthreads), best.ctypes.data_as(C.c_void_p),
                      _p(sing, C.c_double), _p(l00, C.c_double),
                      _p(p1, C.c_double) if want_pass1 else None,
                      _p(p10, C.c_double) if want_pass1 else None)
        return dict(best=best, sing=sing, l00=l00, p1=p1, p1_0=p10)

    def pass2_cell(self, cell, gp0=None):
        nv, nA = self.nv, self.nA
        if gp0 is None:
            gp0 = self.gp0()
        AB = np.zeros((nv, nv, nA)); A0 = np.zeros((nv, nA)); L00 = np.zeros(nA)
        lib().orc_pass2_cell(C.byref(self.inp), _p(gp0, C.c_double), C.c_int64(cell),
                             _p(AB, C.c_double), _p(A0, C.c_double), _p(L00, C.c_double))
        return AB, A0, L00

    def pass2_one(self, cell, j, k, n):
        return lib().orc_pass2_one(C.byref(self.inp), C.c_int64(cell), C.c_int32(j), C.c_int32(k),
                                   C.c_int32(n))

    def count_asym(self, AB, n):
        AB = np.ascontiguousarray(AB)
        return int(lib().orc_count_asym(C.byref(self.inp), _p(AB, C.c_double), C.c_int32(n)))

    def time_cells(self, cells, nthreads=1, with_pass1=True):
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        return float(lib().orc_time_cells(C.byref(self.inp), _p(cells, C.c_int64),
                                          C.c_int64(len(cells)), C.c_int32(nthreads),
                                          C.c_int32(1 if with_pass1 else 0)))

    def write_all(self, prefix, write_pair=False):
        bcs = (C.c_char_p * self.plp.ncells)(*[b.encode() for b in self.plp.barcodes])
        sms = (C.c_char_p * self.nv)(*[s.encode() for s in self.plp.samples])
        rc = lib().orc_write_all(C.byref(self.inp), bcs, sms, prefix.encode(), C.c_int32(1 if write_pair else 0))
        if rc != 0:
            raise OSError("orc_write_all failed for %s" % prefix)


def ref_available() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(plp, prefix, alphas=None, write_pair=False, doublet_prior=None, min_total=0, min_uniq=0,
            min_snp=0, plp_path=None):
    """Run the reference's own lines 386-881 (oracle/_ref/ref_harness) on a Pileup; writes
    <prefix>.single/.sing2/.best[/.pair]; returns the harness' timing dict."""
    if not ref_available():
        raise FileNotFoundError(REF_BIN)
    tmp = None
    if plp_path is None:
        tmp = tempfile.NamedTemporaryFile(suffix=".dmxp", delete=False)
        tmp.close()
        plp_path = tmp.name
        plp.save(plp_path)
    cmd = [REF_BIN, plp_path, prefix]
    for a in (alphas or []):
        cmd += ["--alpha", repr(float(a))]
    if write_pair:
        cmd.append("--write-pair")
    if doublet_prior is not None:
        cmd += ["--doublet-prior", repr(float(doublet_prior))]
    cmd += ["--min-total", str(min_total), "--min-uniq", str(min_uniq), "--min-snp", str(min_snp)]
    try:
        out = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    finally:
        if tmp is not None:
            os.unlink(plp_path)
    return json.loads(out.stdout.decode().strip().splitlines()[-1])


# ---- VCF field -> genotype probabilities ----------------------------------------------------------
REF_FIELD_BIN = os.path.join(HERE, "_ref", "ref_field")
FIELD_CODE = {"GT": 0, "GP": 1, "PL": 2}
FIELD_DTYPE = {0: np.int8, 1: np.float32, 2: np.int32}


def _field_array(field, arr):
    code = FIELD_CODE[field] if isinstance(field, str) else int(field)
    a = np.ascontiguousarray(arr, dtype=FIELD_DTYPE[code])
    if a.ndim != 3 or a.shape[2] != (2 if code == 0 else 3):
        raise ValueError("field array must be [nsnps][nv][%d]" % (2 if code == 0 else 3))
    return code, a


def build_gp(field, arr, gt_error: float = 0.01, nthreads: int = 1, want_af: bool = False):
    """oracle/dmx_field_oracle.c: float32[nsnps][nv][3] (and get_af(1) per SNP) from a GT / GP / PL array."""
    code, a = _field_array(field, arr)
    nsnps, nv = a.shape[0], a.shape[1]
    out = np.empty((nsnps, nv, 3), np.float32)
    af = np.zeros(nsnps, np.float64)
    rc = lib().orc_build_gp(C.c_int32(code), C.c_int32(nv), C.c_int64(nsnps), a.ctypes.data_as(C.c_void_p),
                            C.c_double(gt_error), out.ctypes.data_as(C.c_void_p), af.ctypes.data_as(C.c_void_p),
                            C.c_int32(nthreads))
    if rc != 0:
        raise ValueError("orc_build_gp: %d" % rc)
    return (out, af) if want_af else out


def ref_field_available() -> bool:
    return os.path.exists(REF_FIELD_BIN) and os.access(REF_FIELD_BIN, os.X_OK)


def write_field_file(path, field, arr, gt_error: float):
    code, a = _field_array(field, arr)
    with open(path, "wb") as f:
        f.write(np.array([code, a.shape[1]], np.int32).tobytes())
        f.write(np.array([a.shape[0]], np.int64).tobytes())
        f.write(np.array([gt_error], np.float64).tobytes())
        f.write(a.tobytes())
    return code, a


def run_ref_field(field, arr, gt_error: float = 0.01, want_af: bool = False):
    """The reference's own parse_posteriors / parse_genotypes / parse_likelihoods (oracle/_ref/ref_field)."""
    if not ref_field_available():
        raise FileNotFoundError(REF_FIELD_BIN)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.fld"), os.path.join(td, "out.gp")
        code, a = write_field_file(fin, field, arr, gt_error)
        subprocess.run([REF_FIELD_BIN, fin, fout], check=True)
        nsnps, nv = a.shape[0], a.shape[1]
        raw = np.fromfile(fout, np.uint8)
    gp = raw[: nsnps * nv * 12].view(np.float32).reshape(nsnps, nv, 3).copy()
    af = raw[nsnps * nv * 12:].view(np.float64).copy()
    return (gp, af) if want_af else gp


def site_filter(gt, min_mac: int = 1, min_callrate: float = 0.5) -> np.ndarray:
    """oracle restatement of passed_vfilter's genotype-count tests (bcf_filtered_reader.cpp:540-573)"""
    _, a = _field_array("GT", gt)
    out = np.zeros(a.shape[0], np.uint8)
    lib().orc_site_filter(C.c_int32(a.shape[1]), C.c_int64(a.shape[0]), a.ctypes.data_as(C.c_void_p), C.c_int32(min_mac),
                          C.c_double(min_callrate), out.ctypes.data_as(C.c_void_p))
    return out.astype(bool)


def run_ref_site_filter(gt, min_mac: int = 1, min_callrate: float = 0.5) -> np.ndarray:
    """the reference's own lines 540-573 (oracle/_ref/ref_field, mode 3)"""
    if not ref_field_available():
        raise FileNotFoundError(REF_FIELD_BIN)
    _, a = _field_array("GT", gt)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.fld"), os.path.join(td, "out.pass")
        with open(fin, "wb") as f:
            f.write(np.array([3, a.shape[1]], np.int32).tobytes())
            f.write(np.array([a.shape[0]], np.int64).tobytes())
            f.write(np.array([0.0], np.float64).tobytes())
            f.write(np.array([min_mac], np.int32).tobytes())
            f.write(np.array([min_callrate], np.float64).tobytes())
            f.write(a.tobytes())
        subprocess.run([REF_FIELD_BIN, fin, fout], check=True)
        return np.fromfile(fout, np.uint8).astype(bool)
