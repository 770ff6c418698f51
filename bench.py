#!/usr/bin/env python
"""bench.py -- cells/sec of the demuxlet likelihood hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

A "step" is one pass of the hot path (cmd_cram_demuxlet.cpp:386-875 equivalent: entry tables, pair scan,
reductions, read-back, exact tie finaliser) over one batch of synthetic pileups of the headline shape
(10k barcodes x 64 samples x 100k SNPs, alpha {0, 0.5}; SURVEY 8d "synth-v1").  One process per GPU;
with N > 1 every rank owns its own 10k-barcode shard (weak scaling: barcodes are independent) and the
ranks exchange the per-barcode best records with one NCCL all-gather inside the timed region.

  value  cells/s with the inputs resident in HBM (dmx_compute), whole job, max over ranks
  e2e    cells/s through the C ABI with pinned HOST buffers (dmx_run: H2D + kernels + D2H + finaliser)
  roofline   pair-scan kernel: SURVEY 8(d) gather-model bytes / CUDA-event time vs measured HBM peak
  cpu_baseline   the CPU oracle (oracle/, OpenMP over cells) on a bounded sample, rank 0, N = 1 only

--impl reference times the reference's own lines 386-881 (oracle/_ref/ref_harness, built from
/root/reference by oracle/Makefile) as nproc parallel processes over disjoint barcode groups (its
--group-list sharding model), each step a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec (10k barcodes, 64 samples, 100k SNPs) at 1/2/4/8 B200 vs CPU ref"
UNIT = "cells/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(name: str, rank: int, scale: float):
    """synth-v1 workload of this rank; generation takes about a minute at the headline shape, so the flat
    pileup is cached under $TMPDIR for repeated runs on the same box"""
    from demuxlet_b200.synth import make_config, CONFIGS
    from demuxlet_b200.plp import Pileup
    path = os.path.join(tempfile.gettempdir(), "dmx_synth_%s_%g_%d.dmxp" % (name, scale, rank))
    if os.path.exists(path):
        try:
            return Pileup.load(path), CONFIGS[name]["alphas"]
        except Exception:
            pass
    plp, alphas = make_config(name, scale=scale, seed_offset=1000 * rank)
    try:
        plp.save(path + ".tmp%d" % os.getpid())
        os.replace(path + ".tmp%d" % os.getpid(), path)
    except Exception:
        pass
    return plp, alphas


# -------------------------------------------------------------------------------------------------
def fp_roofline(plp, alphas, pair_s, fp32, clocks):
    """the pair scan against the arithmetic roofline that actually binds it: useful multiply-adds per second vs
    the SM's issue peak (148 SMs x 128 FP32 or 64 FP64 lanes x 2 flop x SM clock)"""
    nv = plp.nv
    hyp = 0
    for n, a in enumerate(alphas):
        if n == 0 and a == 0.0:
            continue                      # separable plane: no pair work
        hyp += nv * (nv + 1) // 2 if a == 0.5 else nv * nv
    flop = float(plp.n_entries) * hyp * 7.0          # 3 mul + 2 add for the dot, 1 mul into the product (+1 FMA slot)
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * (128 if fp32 else 64) * 2 * mhz * 1e6 / 1e12
    ach = flop / pair_s / 1e12
    return {"bound": "fp32 pipe" if fp32 else "fp64 pipe", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak, "flop_per_launch": flop,
            "note": "7 flop per hypothesis and pileup entry (4 issue slots); peak = 148 SMs x lanes x 2 x SM clock"}


def cpu_baseline(plp, alphas, budget_s: float = 15.0):
    """the CPU oracle (restatement of the reference loop) on a bounded sample of the same cells"""
    from oracle import pyoracle as po
    ncores = os.cpu_count() or 1
    orc = po.Oracle(plp, alphas)
    rng = np.random.default_rng(7)
    probe = rng.choice(plp.ncells, size=min(plp.ncells, ncores), replace=False)
    t_probe = orc.time_cells(probe, nthreads=ncores)
    per_cell = t_probe / len(probe)
    n = int(max(len(probe), min(plp.ncells, budget_s / max(per_cell, 1e-9))))
    cells = rng.choice(plp.ncells, size=n, replace=False)
    t = orc.time_cells(cells, nthreads=ncores)
    return {"value": n / t, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": "%d random cells of the workload (%d entries), OpenMP over cells, %.1f s" %
                      (n, int(np.diff(plp.cell_ptr)[cells].sum()), t)}


def run_reference(args):
    """--impl reference: the reference's own code (oracle/_ref/ref_harness) on host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    from demuxlet_b200.synth import CONFIGS
    plp, alphas = make_workload(args.workload, 0, args.scale)
    cfg = CONFIGS[args.workload]
    ncores = os.cpu_count() or 1
    use_ref = po.ref_available()
    cells_per_proc = 4 if plp.nv >= 32 else 64
    # memory guard: the reference materialises nsnps*nv^2*9 doubles per process (:542)
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 8 << 30
    est_per_proc = cells_per_proc * float(np.diff(plp.cell_ptr).mean()) * plp.nv * plp.nv * 9 * 8 * 1.3
    nproc = int(max(1, min(ncores, avail * 0.6 // max(est_per_proc, 1))))
    rng = np.random.default_rng(11)
    vals, walls = [], []
    sample_desc = ""
    tmpd = tempfile.mkdtemp(prefix="dmxref")
    for step in range(args.warmup + args.steps):
        t_wall0 = time.perf_counter()
        if use_ref:
            groups = [rng.choice(plp.ncells, size=cells_per_proc, replace=False) for _ in range(nproc)]
            subs = [plp.subset_cells(np.sort(g), compact_snps=True) for g in groups]
            paths = []
            for i, s in enumerate(subs):
                pth = os.path.join(tmpd, "g%d.dmxp" % i)
                s.save(pth)
                paths.append(pth)
            procs = []
            for i, pth in enumerate(paths):
                cmd = [po.REF_BIN, pth, os.path.join(tmpd, "out%d" % i)]
                for a in alphas:
                    cmd += ["--alpha", repr(float(a))]
                procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL))
            t_worst = 0.0
            for s, pr in zip(subs, procs):
                out = pr.communicate()[0].decode().strip().splitlines()[-1]
                r = json.loads(out)
                # lines 386-881: pass 1 + pair loop, with the gpAB table (:542-560) charged at the
                # full-size job's SNPs-per-cell rather than the compacted sample's
                t_tab = r["t_table_s"] / max(s.nsnps, 1) * (cfg["nsnps"] / cfg["ncells"]) * s.ncells
                t_worst = max(t_worst, r["t_pass1_s"] + r["t_loop_s"] + t_tab)
            v = nproc * cells_per_proc / t_worst
            sample_desc = ("%d processes x %d random cells per step (--group-list sharding), lines 386-881 only; "
                           "gpAB table cost amortised to full-job SNPs/cell" % (nproc, cells_per_proc))
        else:
            orc = po.Oracle(plp, alphas)
            cells = rng.choice(plp.ncells, size=min(plp.ncells, ncores * cells_per_proc), replace=False)
            t = orc.time_cells(cells, nthreads=ncores)
            v = len(cells) / t
            sample_desc = "%d random cells per step, oracle port with OpenMP over cells" % len(cells)
        if step >= args.warmup:
            vals.append(v)
            walls.append(time.perf_counter() - t_wall0)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (synth-v1)",
        "config": {"workload": workload_name(args), "alpha": list(map(float, alphas))},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc if use_ref else ncores,
                         "kind": "reference" if use_ref else "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_name(args):
    from demuxlet_b200.synth import CONFIGS
    c = CONFIGS[args.workload]
    n = max(2, int(round(c["ncells"] * args.scale)))
    return "%s: synthetic %d barcodes x %d samples x %d SNPs, %s field, alpha %s" % (
        args.workload, n, c["nv"], c["nsnps"], c["field"], "/".join("%g" % a for a in c["alphas"]))


# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from demuxlet_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    plp, alphas = make_workload(args.workload, rank, args.scale)
    ncells = plp.ncells

    # pinned host copies of the inputs: what a caller that wants full PCIe speed hands to the C ABI
    def pin(a):
        out = api.pinned_empty(a.shape, a.dtype)
        out[...] = a
        return out
    gp = pin(plp.gp)
    arrays = [pin(x) for x in (plp.cell_ptr, plp.ent_snp, plp.ent_ptr, plp.reads, plp.totl, plp.npass, plp.uniq)]

    fp32 = not args.fp64_pairs
    # the ranks of one box share its host cores: each finaliser pool gets its share
    host_threads = max(1, (os.cpu_count() or 1) // max(1, world))
    d = api.Demuxlet(plp.nv, alphas, device=local_rank, borrow=True, trust_inputs=True, fp32_pairs=fp32,
                     host_threads=host_threads)
    d.set_gp(gp)
    d.set_pileup_arrays(*arrays)

    rec_bytes = api.BEST_DTYPE.itemsize
    gather_in = torch.empty(ncells * rec_bytes, dtype=torch.uint8, device="cuda")
    gather_out = torch.empty(world * ncells * rec_bytes, dtype=torch.uint8, device="cuda") if world > 1 else None
    host_rec = torch.empty(ncells * rec_bytes, dtype=torch.uint8).pin_memory()

    def exchange():
        """the one collective of the path: all-gather of the per-barcode best records"""
        if world == 1:
            return
        host_rec.numpy()[:] = d.best().view(np.uint8)
        gather_in.copy_(host_rec, non_blocking=True)
        dist.all_gather_into_tensor(gather_out, gather_in)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        sync()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    def step_resident():
        d.compute()
        exchange()

    def step_e2e():
        d._chk(d.lib.dmx_set_gp(d._h, gp.ctypes.data, gp.shape[0]), "dmx_set_gp")   # marks inputs dirty:
        d.set_pileup_arrays(*arrays)                                                 # every step re-uploads
        d.run()
        exchange()

    # ---- value: inputs resident in HBM ---------------------------------------------------------------
    d.upload()
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kern_ms, pair_ms, fin_ms, prep_ms = [], [], [], []

    def step_resident_rec():
        step_resident()
        st = d.stats()
        kern_ms.append(st["kernel_ms"]); pair_ms.append(st["pair_kernel_ms"]); fin_ms.append(st["finalise_ms"])
        prep_ms.append(st["prep_ms"])
    dt = timed(step_resident_rec, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    st = d.stats()
    value = world * ncells * args.steps / dt
    ms_per_step = dt / args.steps * 1e3

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    dt2 = timed(step_e2e, args.steps)
    st2 = d.stats()
    e2e_value = world * ncells * args.steps / dt2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    pair_s = float(np.mean(pair_ms)) * 1e-3
    achieved = st["algo_bytes"] / pair_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "pair_scan_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 pair products + f64 logs / exact re-score" if fp32 else "f64", "data": "synthetic (synth-v1, SURVEY 8d)",
        "timing": "host wall clock between device synchronisations (a step ends with the host-side tie finaliser), max "
                  "over ranks; kernel times below from CUDA events on the library's own stream",
        "config": {"workload": workload_name(args), "alpha": list(map(float, alphas)),
                   "cells_per_gpu": ncells, "entries_per_gpu": int(plp.n_entries), "parallelism": "barcode shards x%d" % world,
                   "l2": "no flush: inputs larger than L2 (%.2f GB of CSR + entry tables streamed per step)" %
                         ((plp.n_entries * (8 * 8 * 2 + 4 + 8) + plp.n_reads * 2) / 1e9),
                   "exact_ties": True},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(st2["h2d_bytes"]) * world,
                "d2h_bytes_per_step": int(st2["d2h_bytes"]) * world, "ms_per_step": dt2 / args.steps * 1e3,
                "h2d_ms": st2["h2d_ms"], "kernel_ms": st2["kernel_ms"], "finalise_ms": st2["finalise_ms"],
                "run_ms": st2["total_ms"]},
        "gpu_launches": int(st["gpu_launches"]) * args.steps,
        "roofline": {"bound": "hbm", "kernel": "pair_scan_f32_kernel" if fp32 else "pair_scan_kernel",
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algo_bytes_per_launch": st["algo_bytes"], "kernel_ms": pair_s * 1e3,
                     "note": "gather-model bytes (SURVEY 8d) / CUDA-event time; at nv=64 the kernel is bound by the SM's "
                             "%s issue rate, not by HBM: the 77 MB GP matrix is L2 resident, DRAM traffic << gather-model "
                             "bytes (DESIGN.md 6)" % ("FP32" if fp32 else "FP64")},
        "roofline_fp": fp_roofline(plp, alphas, pair_s, fp32, clocks),
        "breakdown_ms": {"prep": float(np.mean(prep_ms)), "pair_scan": float(np.mean(pair_ms)),
                         "all_kernels": float(np.mean(kern_ms)), "d2h": st["d2h_ms"], "finalise_host": float(np.mean(fin_ms)),
                         "rescored_hypotheses": int(st["n_rescored"]), "host_threads": host_threads},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(plp, alphas)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line goes to the real stdout; everything else a library prints there (NCCL's version banner ...)
    has been sent to stderr"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the workload's barcodes (debug)")
    ap.add_argument("--fp64-pairs", action="store_true",
                    help="pair scan on the FP64 tensor path (DMMA) instead of fp32 products")
    ap.add_argument("--fp32-pairs", action="store_true", help="(default; kept for old command lines)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
