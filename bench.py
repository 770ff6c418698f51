#!/usr/bin/env python
"""bench.py -- cells/sec of the demuxlet likelihood hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C3]

A "step" is one pass of the hot path (cmd_cram_demuxlet.cpp:386-875 equivalent: entry tables, pair scan,
reductions, exact re-scoring of near ties, decision logic, read-back) over one batch of synthetic pileups of the headline shape
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
            "note": "7 flop per hypothesis and pileup entry (the K = 3 form; entries with one informative read are "
                    "evaluated in an equivalent 2-flop form); peak = 148 SMs x lanes x 2 x SM clock"}


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
    if t_probe >= 0.5 * budget_s or n == len(probe):     # expensive barcodes (C5's shape): the probe is the sample
        n, cells, t = len(probe), probe, t_probe
    else:
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
    # the reference indexes its gpAB table with int32 products (:542, :553, :677): a sample it can run at all
    # must keep (SNPs of the sample) * nv^2 * 9 below 2^31
    mean_ent = float(np.diff(plp.cell_ptr).mean())
    while cells_per_proc > 1 and cells_per_proc * mean_ent * plp.nv * plp.nv * 9 >= 2 ** 31:
        cells_per_proc //= 2
    why_port = ""
    if use_ref and mean_ent * plp.nv * plp.nv * 9 >= 2 ** 31:
        use_ref = False
        why_port = ("; the unmodified reference cannot run one barcode of this shape (int32 table index overflow at "
                    "nsnps*nv^2*9 >= 2^31, cmd_cram_demuxlet.cpp:542)")
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
            sample_desc = "%d random cells per step, oracle port with OpenMP over cells%s" % (len(cells), why_port)
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
def hyps_per_cell(nv, alphas):
    h = 0
    for n, a in enumerate(alphas):
        if n == 0 and a == 0.0:
            continue
        h += nv * (nv + 1) // 2 if a == 0.5 else nv * nv
    return h


class Case:
    """one pileup on this rank's GPU, timed both ways (inputs resident in HBM; host buffers through dmx_run), with
    the record exchange of the N > 1 path inside both timed regions"""

    def __init__(self, args, plp, alphas, counts, world, rank, local_rank):
        import torch
        import torch.distributed as dist
        from demuxlet_b200 import api
        from demuxlet_b200.parallel import RecordExchange
        self.torch, self.dist, self.api = torch, dist, api
        self.args, self.plp, self.alphas, self.world, self.rank = args, plp, alphas, world, rank

        def pin(a):   # pinned host copies: what a caller that wants full PCIe speed hands to the C ABI
            out = api.pinned_empty(a.shape, a.dtype)
            out[...] = a
            return out
        self.gp = pin(plp.gp)
        self.arrays = [pin(x) for x in (plp.cell_ptr, plp.ent_snp, plp.ent_ptr, plp.reads, plp.totl, plp.npass, plp.uniq)]
        self.fp32 = not args.fp64_pairs
        # the ranks of one box share its host cores: each finaliser pool gets its share
        self.host_threads = max(1, (os.cpu_count() or 1) // max(1, world))
        self.d = api.Demuxlet(plp.nv, alphas, device=local_rank, borrow=True, trust_inputs=True, fp32_pairs=self.fp32,
                              host_threads=self.host_threads)
        self.d.set_gp(self.gp)
        self.d.set_pileup_arrays(*self.arrays)
        self.ex = RecordExchange(counts, api.BEST_DTYPE.itemsize) if world > 1 else None
        self.total_cells = int(sum(counts))
        self.gathered = None
        self.t_exchange, self.n_exchange = 0.0, 0
        self.pair_bytes = 0
        self.pair_buf = None
        if args.write_pair:   # --write-pair: the dense llksAB block of every barcode comes back too (:772-797)
            per = plp.nv * plp.nv * len(alphas)
            self.pair_slab = max(1, min(plp.ncells, (8 << 30) // (per * 8)))     # whole job if it fits 8 GB of host memory
            self.pair_buf = np.zeros((self.pair_slab, plp.nv, plp.nv, len(alphas)))
            self.pair_bytes = plp.ncells * per * 8

    def exchange(self):
        if self.ex is not None:
            t0 = time.perf_counter()
            mine = self.d.best(out=self.ex.send_view(self.api.BEST_DTYPE))   # records straight into the pinned send buffer
            self.gathered = self.ex(mine)
            self.t_exchange += time.perf_counter() - t0
            self.n_exchange += 1

    def pairs(self):
        if self.pair_buf is None:
            return
        n = self.plp.ncells
        for lo in range(0, n, self.pair_slab):
            hi = min(n, lo + self.pair_slab)
            self.d._chk(self.d.lib.dmx_get_pair(self.d._h, lo, hi, self.pair_buf.ctypes.data), "dmx_get_pair")

    def sync(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        self.sync()
        self.t_exchange, self.n_exchange = 0.0, 0      # of this timed region only (the warm-up holds NCCL's set-up)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.sync()
        dt = time.perf_counter() - t0
        if self.world > 1:
            t = self.torch.tensor([dt], dtype=self.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    def step_resident(self):
        self.d.compute()
        self.pairs()
        self.exchange()

    def step_e2e(self):
        d = self.d
        d._chk(d.lib.dmx_set_gp(d._h, self.gp.ctypes.data, self.gp.shape[0]), "dmx_set_gp")   # marks inputs dirty:
        d.set_pileup_arrays(*self.arrays)                                                     # every step re-uploads
        d.run()
        self.pairs()
        self.exchange()

    def run(self, sampler=None):
        args, d = self.args, self.d
        d.upload()
        for _ in range(args.warmup):
            self.step_resident()
        if sampler is not None:
            sampler.start()
        rec = {"kern": [], "pair": [], "fin": [], "prep": []}

        def step_rec():
            self.step_resident()
            st = d.stats()
            rec["kern"].append(st["kernel_ms"]); rec["pair"].append(st["pair_kernel_ms"])
            rec["fin"].append(st["finalise_ms"]); rec["prep"].append(st["prep_ms"])
        dt = self.timed(step_rec, args.steps)
        exchange_ms = self.t_exchange / max(self.n_exchange, 1) * 1e3
        clocks = sampler.stop() if sampler is not None else None
        st = d.stats()
        for _ in range(max(1, args.warmup // 2)):
            self.step_e2e()
        dt2 = self.timed(self.step_e2e, args.steps)
        st2 = d.stats()
        if self.ex is not None and self.rank == 0:
            assert self.gathered is not None and self.gathered.shape[0] == self.total_cells
        return {"dt": dt, "dt2": dt2, "st": st, "st2": st2, "rec": rec, "clocks": clocks,
                "exchange_ms": exchange_ms}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from demuxlet_b200.parallel import shard_pileup, shard_ranges

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- weak scaling (the line's value): every rank owns one whole workload of the named shape ---------------------
    plp, alphas = make_workload(args.workload, rank, args.scale)
    ncells = plp.ncells
    sampler = ClockSampler(local_rank) if rank == 0 else None
    weak = None
    if args.scaling in ("weak", "both") or world == 1:
        case = Case(args, plp, alphas, [ncells] * world, world, rank, local_rank)
        weak = case.run(sampler)
        weak["pair_bytes"], weak["host_threads"] = case.pair_bytes, case.host_threads
        case.d.close()
        del case

    # ---- strong scaling: ONE workload (rank 0's), its barcodes split over the ranks in contiguous ranges balanced by
    # pileup entries (the reference's --group-list model, cmd_cram_demuxlet.cpp:68), one all-gather of the records -----
    strong = None
    if world > 1 and args.scaling in ("strong", "both"):
        if rank == 0:
            full = plp
        if world > 1:
            dist.barrier()                      # rank 0's workload file is in the cache now
        if rank != 0:
            full, _ = make_workload(args.workload, 0, args.scale)
        ranges = shard_ranges(full.cell_ptr, world)
        counts = [hi - lo for lo, hi in ranges]
        sub, _ = shard_pileup(full, rank, world)
        case = Case(args, sub, alphas, counts, world, rank, local_rank)
        strong = case.run(sampler if weak is None else None)
        strong["counts"], strong["total"] = counts, full.ncells
        strong["entries"] = [int(full.cell_ptr[hi] - full.cell_ptr[lo]) for lo, hi in ranges]
        strong["pair_bytes"], strong["host_threads"] = case.pair_bytes * world, case.host_threads
        case.d.close()
        del case

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    fp32 = not args.fp64_pairs
    main, scaling = (weak, "weak") if weak is not None else (strong, "strong")
    total_cells = world * ncells if scaling == "weak" else strong["total"]
    st, st2 = main["st"], main["st2"]
    value = total_cells * args.steps / main["dt"]
    e2e_value = total_cells * args.steps / main["dt2"]
    peak, peak_src = load_peaks()
    pair_s = float(np.mean(main["rec"]["pair"])) * 1e-3
    achieved = st["algo_bytes"] / pair_s / 1e9
    kernel = ("pair_scan_f32_kernel" if os.environ.get("DMX_SCAN_GEN") == "1" else "pair_scan_v2_kernel") if fp32 else "pair_scan_kernel"
    traffic = None   # DRAM bytes of one launch of this kernel on this workload, from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "traffic_%s_%s.json" % (kernel, args.workload))
    if os.path.exists(tpath) and args.scale == 1.0 and scaling == "weak":
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    hyp = hyps_per_cell(plp.nv, alphas)
    line = {
        "metric": METRIC if args.workload == "C3" else "cells/sec (%s) at 1/2/4/8 B200 vs CPU ref" % args.workload,
        "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["dt"] / args.steps * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32 pair products + f64 logs / exact re-score" if fp32 else "f64", "data": "synthetic (synth-v1, SURVEY 8d)",
        "timing": "host wall clock between device synchronisations (a step ends with the records in host memory), max "
                  "over ranks; kernel times below from CUDA events on the library's own stream",
        "config": {"workload": workload_name(args), "alpha": list(map(float, alphas)),
                   "cells_per_gpu": ncells if scaling == "weak" else strong["counts"],
                   "cells_total": total_cells, "entries_per_gpu": int(plp.n_entries) if scaling == "weak" else strong["entries"],
                   "hypotheses_per_cell": hyp,
                   "parallelism": ("every rank one whole workload of this shape (x%d)" % world) if scaling == "weak" else
                                  "one workload, barcode ranges balanced by entries over %d ranks" % world,
                   "exchange": "none" if world == 1 else "one NCCL all-gather of the per-barcode records per step",
                   "l2": "no flush: inputs larger than L2 (%.2f GB of CSR + entry tables streamed per step)" %
                         ((plp.n_entries * (8 * 8 * 2 + 4 + 8) + plp.n_reads * 2) / 1e9),
                   "write_pair": bool(args.write_pair), "exact_ties": True},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(st2["h2d_bytes"]) * world,
                "d2h_bytes_per_step": (int(st2["d2h_bytes"]) + main["pair_bytes"]) * world,
                "ms_per_step": main["dt2"] / args.steps * 1e3,
                "h2d_ms": st2["h2d_ms"], "kernel_ms": st2["kernel_ms"], "finalise_ms": st2["finalise_ms"],
                "run_ms": st2["total_ms"]},
        "gpu_launches": int(st["gpu_launches"]) * args.steps,
        "roofline": {"bound": "hbm", "kernel": kernel,
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algo_bytes_per_launch": st["algo_bytes"], "kernel_ms": pair_s * 1e3,
                     "note": "gather-model bytes (SURVEY 8d) / CUDA-event time of the scan launches of one step; the GP "
                             "matrix is re-used out of L2, so DRAM traffic is below the gather model (DESIGN.md 6)"},
        "roofline_fp": fp_roofline(plp, alphas, pair_s, fp32, main["clocks"]),
        "breakdown_ms": {"prep": float(np.mean(main["rec"]["prep"])), "pair_scan": float(np.mean(main["rec"]["pair"])),
                         "all_kernels": float(np.mean(main["rec"]["kern"])), "d2h": st["d2h_ms"],
                         "finalise_host": float(np.mean(main["rec"]["fin"])),
                         "rescored_hypotheses": int(st["n_rescored"]), "host_threads": main["host_threads"],
                         "record_exchange": main["exchange_ms"]},
        "clocks": main["clocks"],
    }
    if args.write_pair:
        line["write_pair"] = {"dense_llk_bytes_per_step": main["pair_bytes"] * (world if scaling == "weak" else 1),
                              "note": "dmx_get_pair of every barcode inside both timed regions (fp64 scan with dense "
                                      "stores + read-back); the text of .pair is formatted by dmx_write_outputs, "
                                      "timed by tools/write_bench.py"}
    if strong is not None and main is not strong:
        s1, s2 = strong["st"], strong["st2"]
        line["strong"] = {
            "scaling": "strong", "value": strong["total"] * args.steps / strong["dt"], "unit": UNIT,
            "ms_per_step": strong["dt"] / args.steps * 1e3,
            "e2e": {"value": strong["total"] * args.steps / strong["dt2"], "unit": UNIT,
                    "ms_per_step": strong["dt2"] / args.steps * 1e3},
            "cells_total": strong["total"], "cells_per_gpu": strong["counts"], "entries_per_gpu": strong["entries"],
            "rank0_ms": {"pair_scan": float(np.mean(strong["rec"]["pair"])), "all_kernels": float(np.mean(strong["rec"]["kern"])),
                         "finalise_host": float(np.mean(strong["rec"]["fin"])), "e2e_h2d": s2["h2d_ms"]},
            "gpu_launches": int(s1["gpu_launches"]) * args.steps,
            "note": "the SAME %d-barcode job as N = 1, split over the ranks; value = barcodes of the job / max-over-ranks "
                    "step time, all-gather of the records inside the step" % strong["total"]}
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
    ap.add_argument("--scaling", default="both", choices=["weak", "strong", "both"],
                    help="N > 1: weak = every rank one whole workload (the line's value); strong = one workload split "
                         "over the ranks (reported under 'strong'); both by default")
    ap.add_argument("--write-pair", action="store_true",
                    help="also fetch the dense llksAB block of every barcode in every step (what --write-pair prints)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
