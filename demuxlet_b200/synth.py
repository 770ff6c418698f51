"""synth-v1: the synthetic pileup generator of SURVEY.md section 8(d).

The oracle and the GPU path read the SAME generated arrays, so the generator only has to be
deterministic for a seed, not portable across numpy versions.

Genotype-probability builders restate the reference's VCF field handling
(/root/reference/bcf_filtered_reader.cpp): GT :378-401, GP :411-450, PL (10-iteration EM) :254-312.
Their output -- float32 [nsnps][nv][3] -- is exactly what the reference then widens to double
(cmd_cram_demuxlet.cpp:227-231) and what the kernels consume.
"""
from __future__ import annotations

import numpy as np

from .plp import Pileup

# configuration shapes named in BASELINE.json: (ncells, nv, nsnps, field, alphas)
CONFIGS = {
    "C1p": dict(ncells=500, nv=2, nsnps=54000, field="GT", alphas=(0.0, 0.5)),
    "C2": dict(ncells=10000, nv=8, nsnps=20000, field="GP", alphas=(0.0, 0.5)),
    "C3": dict(ncells=10000, nv=64, nsnps=100000, field="GP", alphas=(0.0, 0.5)),
    "C4": dict(ncells=50000, nv=64, nsnps=100000, field="GP", alphas=(0.0, 0.5)),
    "C5": dict(ncells=100000, nv=256, nsnps=500000, field="PL",
               alphas=tuple(round(0.05 * i, 2) for i in range(11))),
}
BASE_SEED = 20261017


def gp_from_gt(g: np.ndarray, gt_error: float = 0.01) -> np.ndarray:
    """bcf_filtered_reader.cpp:398-400 (called genotypes only)"""
    hit = np.float32(1.0 - gt_error)
    miss = np.float32(gt_error / (3 - 1.0))
    out = np.full(g.shape + (3,), miss, dtype=np.float32)
    np.put_along_axis(out, g[..., None].astype(np.int64), hit, axis=-1)
    return out


def gp_from_gp(raw: np.ndarray, gt_error: float = 0.01) -> np.ndarray:
    """bcf_filtered_reader.cpp:411-450: float normalise, blend with the cohort mean that includes a
    (1/4, 1/2, 1/4) pseudo-sample"""
    raw = raw.astype(np.float32)
    nsnps, nv, _ = raw.shape
    s = raw[..., 0].copy()
    s += raw[..., 1]
    s += raw[..., 2]                                 # float accumulation order of :430-434
    g = raw / s[..., None]                           # :436
    init = np.array([0.25, 0.5, 0.25], dtype=np.float32)   # :421-425 for nalleles == 2
    stack = np.concatenate([np.broadcast_to(init, (nsnps, 1, 3)), g], axis=1)
    sums = np.cumsum(stack, axis=1, dtype=np.float32)[:, -1, :]   # sequential float adds, :437
    sums = sums / np.float32(nv + 1)                 # :442
    out = (1.0 - gt_error) * g.astype(np.float64) + gt_error * sums[:, None, :].astype(np.float64)
    return out.astype(np.float32)                    # :448 stores to float


def gp_from_pl(pl: np.ndarray, niter: int = 10) -> np.ndarray:
    """bcf_filtered_reader.cpp:254-312, diploid biallelic: EM on the allele frequency, posterior
    from the final iteration stored as float"""
    pl = np.minimum(pl, 255)
    prob = np.power(0.1, pl * 0.1)                   # phredConv.toProb, PhredHelper.cpp:31
    nsnps, nv, _ = pl.shape
    af = np.full((nsnps, 2), 0.5)
    gp = None
    for _ in range(niter):
        w = np.stack([af[:, 0] * af[:, 0], 2 * af[:, 1] * af[:, 0], af[:, 1] * af[:, 1]], axis=-1)
        gp = w[:, None, :] * prob
        gp = gp / gp.sum(axis=-1, keepdims=True)
        new = np.stack([2 * gp[..., 0] + gp[..., 1], gp[..., 1] + 2 * gp[..., 2]], axis=-1).sum(axis=1)
        af = new / (2.0 * nv)
    return gp.astype(np.float32)


def make_gp(rng: np.random.Generator, nsnps: int, nv: int, field: str):
    af = rng.uniform(0.05, 0.5, size=nsnps)
    g = rng.binomial(2, af[:, None], size=(nsnps, nv)).astype(np.int8)
    if field == "GT":
        gp = gp_from_gt(g)
    elif field == "GP":
        raw = rng.uniform(0.0, 0.05, size=(nsnps, nv, 3)).astype(np.float32)
        np.put_along_axis(raw, g[..., None].astype(np.int64),
                          np.take_along_axis(raw, g[..., None].astype(np.int64), -1) + np.float32(1.0), -1)
        gp = gp_from_gp(raw)
    elif field == "PL":
        # 10 simulated reads per site and sample, error 0.01 -> phred-scaled genotype likelihoods
        nalt = rng.binomial(10, np.clip(g / 2.0, 0.01, 0.99))
        nref = 10 - nalt
        e = 0.01
        ll = np.stack([nref * np.log10(1 - e) + nalt * np.log10(e),
                       10 * np.log10(0.5) * np.ones_like(nalt, dtype=np.float64),
                       nalt * np.log10(1 - e) + nref * np.log10(e)], axis=-1)
        pl = np.rint(-10.0 * (ll - ll.max(axis=-1, keepdims=True))).astype(np.int32)
        gp = gp_from_pl(pl)
    else:
        raise ValueError("field must be GT, GP or PL")
    return g, gp


def make_synth(ncells: int, nv: int, nsnps: int, field: str = "GP", seed: int = BASE_SEED,
               cover_frac: float = 0.01, sigma: float = 0.5, doublet_frac: float = 0.1,
               cover_min: int = 1) -> Pileup:
    """Pileup for `ncells` droplets over `nsnps` SNPs and `nv` genotyped samples."""
    rng = np.random.default_rng(seed)
    g, gp = make_gp(rng, nsnps, nv, field)

    # barcodes: unique 16-mers + "-1", first-seen (cell-id) order is random w.r.t. lexicographic
    codes = set()
    barcodes = []
    while len(barcodes) < ncells:
        need = ncells - len(barcodes)
        arr = rng.integers(0, 4, size=(need + 16, 16))
        for row in arr:
            s = "".join("ACGT"[i] for i in row) + "-1"
            if s not in codes:
                codes.add(s)
                barcodes.append(s)
                if len(barcodes) == ncells:
                    break
    samples = ["S%03d" % j for j in range(nv)]

    # covered SNPs per cell: round(LogNormal(ln(nsnps*cover_frac), sigma)) in [cover_min, nsnps]
    want = np.rint(rng.lognormal(np.log(max(nsnps * cover_frac, 1.0)), sigma, size=ncells))
    want = np.clip(want, cover_min, nsnps).astype(np.int64)
    cell_of = np.repeat(np.arange(ncells, dtype=np.int64), want)
    snp_of = rng.integers(0, nsnps, size=cell_of.shape[0], dtype=np.int64)
    key = cell_of * nsnps + snp_of                   # dedupe; sorted by (cell, snp)
    key.sort()
    if key.shape[0]:
        key = key[np.concatenate([[True], key[1:] != key[:-1]])]
    cell_of, snp_of = key // nsnps, (key % nsnps).astype(np.int32)
    n_ent = np.bincount(cell_of, minlength=ncells)
    cell_ptr = np.concatenate([[0], np.cumsum(n_ent)]).astype(np.int64)
    E = snp_of.shape[0]

    # donors: singlet (one sample) or doublet (two distinct samples, reads 50/50)
    is_dbl = rng.random(ncells) < doublet_frac
    s1 = rng.integers(0, nv, size=ncells)
    s2 = (s1 + 1 + rng.integers(0, max(nv - 1, 1), size=ncells)) % nv
    s2 = np.where(is_dbl, s2, s1)

    n_reads = rng.geometric(0.8, size=E).astype(np.int64)       # 1 + Geom(0.8), mean 1.25
    ent_ptr = np.concatenate([[0], np.cumsum(n_reads)]).astype(np.int64)
    R = int(ent_ptr[-1])
    ent_of = np.repeat(np.arange(E, dtype=np.int64), n_reads)
    r_cell = cell_of[ent_of]
    donor = np.where(rng.random(R) < 0.5, s1[r_cell], s2[r_cell])
    geno = g[snp_of[ent_of], donor]
    allele = (rng.random(R) < geno / 2.0).astype(np.uint16)
    bq = rng.integers(13, 41, size=R).astype(np.uint16)
    corrupt = rng.random(R) < np.power(10.0, -bq.astype(np.float64) / 10.0)
    kind = rng.random(R)
    allele = np.where(corrupt & (kind < 1 / 3), 1 - allele, allele)
    allele = np.where(corrupt & (kind >= 1 / 3), 2, allele).astype(np.uint16)
    reads = ((allele << 8) | bq).astype(np.uint16)

    uniq = np.bincount(r_cell, minlength=ncells).astype(np.int32)
    npass = (uniq + uniq // 3).astype(np.int32)
    totl = (npass * 3 + 7).astype(np.int32)
    return Pileup(gp, cell_ptr, snp_of, ent_ptr, reads, totl, npass, uniq, barcodes, samples)


def make_config(name: str, scale: float = 1.0, seed_offset: int = 0) -> tuple[Pileup, tuple]:
    """One of the BASELINE.json shapes; `scale` < 1 shrinks the number of cells only (the per-cell
    work -- nv, nsnps, coverage -- stays that of the named configuration)."""
    cfg = CONFIGS[name]
    idx = list(CONFIGS).index(name) + 1
    ncells = max(2, int(round(cfg["ncells"] * scale)))
    p = make_synth(ncells, cfg["nv"], cfg["nsnps"], cfg["field"], seed=BASE_SEED + idx + seed_offset)
    return p, cfg["alphas"]
