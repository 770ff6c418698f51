"""CPU tests of the C-ABI library: it loads, exports every symbol include/dmx.h declares, refuses to run
without a GPU (no CPU fallback), and its host-side exact scorer is bit-identical to the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from demuxlet_b200 import api
from demuxlet_b200.plp import Pileup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "dmx.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(api.LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "demuxlet_b200", "csrc")], check=True)
    return api.load_library()


def declared_symbols():
    txt = open(HDR).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dmx_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(api.ABI_SYMBOLS) == syms


def test_abi_version_and_strerror(lib):
    assert lib.dmx_abi_version() == 2
    assert lib.dmx_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.dmx_strerror(-3)


def test_struct_sizes_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dmx.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(dmx_best_rec), '
                   'sizeof(dmx_stats), sizeof(dmx_config), sizeof(dmx_field_stats), sizeof(dmx_sweep_opts));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    a, b, c, d, e = map(int, subprocess.run([str(exe)], check=True, stdout=subprocess.PIPE).stdout.split())
    assert a == api.BEST_DTYPE.itemsize
    assert b == C.sizeof(api.DmxStats)
    assert c == C.sizeof(api.DmxConfig)
    assert d == C.sizeof(api.DmxFieldStats)
    assert e == C.sizeof(api.DmxSweepOpts)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert api.device_count() == 0
    with pytest.raises(api.DmxError) as e:
        api.Demuxlet(4)
    assert e.value.status == -3


def test_invalid_configs_rejected(lib):
    # nv < 2 or fewer than 2 alphas index out of bounds in the reference; the shim refuses them
    for nv, alphas in ((1, (0.0, 0.5)), (4, (0.5,)), (4, tuple([0.1] * 17))):
        al = np.ascontiguousarray(alphas, dtype=np.float64)
        cfg = api.DmxConfig(nv, len(al), al.ctypes.data_as(C.POINTER(C.c_double)), 0.5, 0, 0, 0, 0, 1, 0)
        h = C.c_void_p()
        assert lib.dmx_create(C.byref(h), C.byref(cfg)) == -1


def test_host_exact_scorer_is_bit_identical_to_oracle(lib, oracle_lib):
    for name, alphas in (("gp6", (0.0, 0.5)), ("pl5grid", (0.0, 0.1, 0.25, 0.4, 0.5)), ("gt8ties", (0.0, 0.5))):
        p = Pileup.load(os.path.join(ROOT, "tests", "golden", name + ".dmxp"))
        orc = oracle_lib.Oracle(p, alphas)
        rng = np.random.default_rng(1)
        for _ in range(40):
            c, j, k, n = (int(rng.integers(p.ncells)), int(rng.integers(p.nv)), int(rng.integers(p.nv)),
                          int(rng.integers(len(alphas))))
            assert api.exact_llk(p, alphas, c, j, k, n) == orc.pass2_one(c, j, k, n)
