"""How exact would the pair scan be with 3xTF32 tensor-core tiles (DESIGN.md 6, round-2 lever)?  Host-side emulation on
cells of the C3 shape (64 samples, alpha = 0.5): per (entry, hypothesis) S = sum_l a_j[l] w_k[l] computed
  f32     as today: FMUL + 2 FFMA in fp32
  3xtf32  a = ah + al, w = wh + wl (10-bit mantissas, cvt.rna), S = ah.wh + ah.wl + al.wh, products exact, fp32
          accumulation over two k = 8 steps, rounded to nearest (rn) or truncated (rz: the pessimistic model of the
          tensor core's adder)
then multiplied into an fp32 running product; the error of log prod against the fp64 sum of logs, per hypothesis.
usage: python tools/tf32_precision.py [NCELLS=6]   (host only)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demuxlet_b200.synth import make_config     # noqa: E402

ncells = int(sys.argv[1]) if len(sys.argv) > 1 else 6
plp, _ = make_config("C3", scale=ncells / 10000.0)
err = np.array([0.75, 0.75] + [10 ** (-q / 10) for q in range(2, 256)])


def tf32(x):
    b = np.asarray(x, np.float32).view(np.uint32)
    return (((b + np.uint32(0x1000)) & np.uint32(0xffffe000))).view(np.float32)


def to_f32(x, rz):
    y = np.asarray(x, np.float64).astype(np.float32)
    if rz:      # move one ulp towards zero where rounding went away from zero
        away = np.abs(y.astype(np.float64)) > np.abs(x)
        y = np.where(away, np.nextafter(y, np.float32(0)), y)
    return y


p_lm = np.array([[0.5 * l + (m - l) * 0.25 for m in range(3)] for l in range(3)])
worst = {"f32": 0.0, "3xtf32_rn": 0.0, "3xtf32_rz": 0.0}
bias = {k: [] for k in worst}
for c in range(plp.ncells):
    e0, e1 = plp.cell_ptr[c], plp.cell_ptr[c + 1]
    exact = np.zeros((64, 64))
    acc = {k: np.ones((64, 64), np.float32) for k in worst}
    ex = {k: np.zeros((64, 64)) for k in worst}
    for e in range(e0, e1):
        g = plp.gp[plp.ent_snp[e]].astype(np.float64)                 # [64][3]
        Q = np.ones((3, 3))
        for r in plp.reads[plp.ent_ptr[e]:plp.ent_ptr[e + 1]]:
            al, bq = int(r) >> 8, int(r) & 0xff
            if al == 2:
                continue
            pR = (1 - err[bq]) if al == 0 else err[bq] / 3
            pA = (1 - err[bq]) if al == 1 else err[bq] / 3
            Q = Q * (pR * (1 - p_lm) + pA * p_lm)
            Q = Q / Q.max()
        Q = (Q + 1e-6) / (1 + 1e-6)
        w = g @ Q.T                                                    # w[k][l] = sum_m Q[l][m] g[k][m]
        exact += np.log(g @ w.T)
        a32, w32 = g.astype(np.float32), w.astype(np.float32)
        S = {}
        s = (a32[:, None, 0].astype(np.float64) * w32[None, :, 0]).astype(np.float32)
        s = (a32[:, None, 1].astype(np.float64) * w32[None, :, 1] + s).astype(np.float32)
        S["f32"] = (a32[:, None, 2].astype(np.float64) * w32[None, :, 2] + s).astype(np.float32)
        ah, wh = tf32(a32), tf32(w32)
        al_, wl = tf32(a32 - ah), tf32(w32 - wh)
        ah, wh, al_, wl = (x.astype(np.float64) for x in (ah, wh, al_, wl))
        hh = ah @ wh.T
        step1 = hh + (ah[:, None, 0] * wl[None, :, 0]) + (ah[:, None, 1] * wl[None, :, 1])      # k = 0..7 (3 + 3 + 2 terms)
        rest = ah[:, None, 2] * wl[None, :, 2] + al_ @ wh.T
        for rz in (False, True):
            S["3xtf32_rz" if rz else "3xtf32_rn"] = to_f32(to_f32(step1, rz).astype(np.float64) + rest, rz)
        for k in worst:
            acc[k] = (acc[k].astype(np.float64) * S[k]).astype(np.float32)
            m, x = np.frexp(acc[k])
            acc[k] = m.astype(np.float32)
            ex[k] += x
    iu = np.triu_indices(64, 1)
    for k in worst:
        d = (np.log(acc[k].astype(np.float64)) + ex[k] * np.log(2.0) - exact)[iu]
        worst[k] = max(worst[k], float(np.abs(d).max()))
        bias[k].append(float(d.mean()))
print(json.dumps({"cells": int(plp.ncells), "entries_per_cell": float(np.diff(plp.cell_ptr).mean()),
                  "max_abs_llk_err": worst, "mean_llk_err": {k: float(np.mean(v)) for k, v in bias.items()}}))
