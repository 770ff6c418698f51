"""dmx_build_gp (VCF field -> genotype-probability matrix) on the headline matrix shape
(100 k SNPs x 64 samples) for --field GT / GP / PL:
  kernel   dmx_build_gp_device on inputs resident in HBM, CUDA events on the launching stream, input + output
           (> 126 MB of L2 per launch only for nsnps >= 100 k: the default runs 400 k SNPs so every launch streams
           > L2), achieved GB/s against MEASURED_PEAKS.json
  e2e      dmx_build_gp from pinned host buffers (copies inside the timed region)
  cpu      the oracle restatement on one thread (how the reference runs it: once per record, inline in the BAM sweep)
           and on all cores, and the reference's own functions (oracle/_ref/ref_field) where built
usage: python tools/field_bench.py [NSNPS=400000] [NV=64]      prints one JSON line per field"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                        # noqa: E402  (device buffers and events only)
from demuxlet_b200 import api                       # noqa: E402
from oracle import pyoracle as po                   # noqa: E402
from tests.test_field import make_field, bits       # noqa: E402

nsnps = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 64
peak = 6550.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass
lib = api.load_library()
dev = torch.device("cuda:0")
ncores = os.cpu_count() or 1
for field in ("GT", "GP", "PL"):
    code = api.DMX_FIELD[field]
    arr = make_field(field, nsnps, nv, 21)
    d_in = torch.from_numpy(arr).to(dev)
    d_out = torch.empty((nsnps, nv, 3), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def launch():
        rc = lib.dmx_build_gp_device(0, code, nv, nsnps, d_in.data_ptr(), 0.01, d_out.data_ptr(), nv * 3, None, st)
        assert rc == 0, rc

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    reps = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        launch()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    kernel_ms = float(np.mean(ms))
    algo = nsnps * nv * ((2 if field == "GT" else 12) + 12)
    # end to end from pinned host memory
    pin_in = api.pinned_empty(arr.shape, arr.dtype)
    pin_in[...] = arr
    pin_out = api.pinned_empty((nsnps, nv, 3), np.float32)
    for _ in range(2):
        api.build_gp(field, pin_in, 0.01, out=pin_out)
    t0 = time.perf_counter()
    for _ in range(3):
        _, fst = api.build_gp(field, pin_in, 0.01, out=pin_out, want_stats=True)
    e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    # CPU: bounded sample of the same records
    ns_cpu = min(nsnps, 50000)
    t0 = time.perf_counter()
    want = po.build_gp(field, arr[:ns_cpu], 0.01, nthreads=1)
    cpu1_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    po.build_gp(field, arr[:ns_cpu], 0.01, nthreads=ncores)
    cpun_s = time.perf_counter() - t0
    assert np.array_equal(bits(pin_out[:ns_cpu]), bits(want))
    assert np.array_equal(bits(d_out[:ns_cpu].cpu().numpy()), bits(want))
    ref_s = None
    if po.ref_field_available():
        t0 = time.perf_counter()
        po.run_ref_field(field, arr[:ns_cpu], 0.01)
        ref_s = time.perf_counter() - t0       # includes the harness' file I/O
    print(json.dumps({
        "field": field, "nsnps": nsnps, "nv": nv, "kernel_ms": round(kernel_ms, 4), "kernel_ms_min": round(ms[0], 4),
        "algo_bytes": algo, "achieved_gbs": round(algo / kernel_ms / 1e6, 1), "peak_gbs": peak,
        "frac": round(algo / kernel_ms / 1e6 / peak, 3),
        "snps_per_s_kernel": round(nsnps / kernel_ms * 1e3), "e2e_ms": round(e2e_ms, 3),
        "e2e_kernel_ms": round(fst["kernel_ms"], 3), "snps_per_s_e2e": round(nsnps / e2e_ms * 1e3),
        "cpu": {"sample_snps": ns_cpu, "snps_per_s_1thread": round(ns_cpu / cpu1_s),
                "snps_per_s_allcores": round(ns_cpu / cpun_s), "cores": ncores,
                "snps_per_s_reference_harness": round(ns_cpu / ref_s) if ref_s else None},
        "bit_identical": True}))
