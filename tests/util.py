"""helpers shared by the parity tests: run oracle and GPU path on one pileup and compare"""
import numpy as np

LLK_TOL = 1e-4          # BASELINE.json north_star: log-likelihood columns within 1e-4
IDX_FIELDS = [("i_sing1", "sng1"), ("i_sing2", "sng2"), ("j_best", "dbl1"), ("k_best", "dbl2"),
              ("a_best", "alpha_idx"), ("call", "call")]
LLK_FIELDS = [("sing_llk1", "sng_llk1"), ("sing_llk2", "sng_llk2"), ("sing_llk0", "sng_llk0"),
              ("pair_llk12", "llk12"), ("pair_llk1", "llk1"), ("pair_llk2", "llk2"), ("pair_llk10", "llk10"),
              ("pair_llk20", "llk20"), ("pair_llk00", "llk00"), ("max_llk", "max_llk")]
PRB_FIELDS = [("post_doublet", "prb_dbl"), ("post_singlet", "prb_sng1")]


def run_both(plp, alphas=(0.0, 0.5), doublet_prior=0.5, min_total=0, min_uniq=0, min_snp=0, nthreads=8, **kw):
    from demuxlet_b200.api import Demuxlet
    from oracle import pyoracle as po
    orc = po.Oracle(plp, alphas, doublet_prior, min_total, min_uniq, min_snp)
    ref = orc.run(nthreads=nthreads)
    d = Demuxlet(plp.nv, alphas, doublet_prior, min_total, min_uniq, min_snp, **kw)
    d.set_pileup(plp)
    d.run()
    return orc, ref, d


def compare(ref, d, tol=LLK_TOL, check_single=True, prb_rtol=1e-6):
    """ref: oracle.run() dict; d: Demuxlet after run().  Returns max abs LLK error."""
    ob, gb = ref["best"], d.best()
    assert np.array_equal(ob["skipped"] != 0, gb["status"] != 0)
    live = gb["status"] == 0
    assert np.array_equal(ob["n_snp"], gb["n_snp"])
    for of, gf in IDX_FIELDS:
        bad = np.nonzero(live & (ob[of] != gb[gf]))[0]
        assert bad.size == 0, "%s differs at cells %s: oracle %s gpu %s" % (gf, bad[:5], ob[of][bad[:5]], gb[gf][bad[:5]])
    worst = 0.0
    for of, gf in LLK_FIELDS:
        err = np.abs(ob[of][live] - gb[gf][live])
        if err.size:
            worst = max(worst, float(err.max()))
            assert err.max() <= tol, "%s max err %g" % (gf, err.max())
    for of, gf in PRB_FIELDS:
        assert np.allclose(ob[of][live], gb[gf][live], rtol=prb_rtol, atol=1e-12), gf
    sing, l00 = d.sing2()
    err = np.abs(sing - ref["sing"])[live]
    if err.size:
        worst = max(worst, float(err.max()))
        assert err.max() <= tol
    assert np.abs(l00 - ref["l00"])[live].max(initial=0) <= tol
    if check_single:
        p1, p10 = d.single()
        notfilt = gb["status"] != 1
        assert np.abs(p1 - ref["p1"])[notfilt].max(initial=0) <= tol
        assert np.abs(p10 - ref["p1_0"])[notfilt].max(initial=0) <= tol
    return worst


def diff_tsv(path_a, path_b, exact_cols, tol=LLK_TOL):
    """compare two demuxlet TSVs: `exact_cols` byte-identical, every other column numerically within tol
    (absolute for |x| >= 1e-3 style LLKs, relative 1e-2 for %.3lg probabilities)"""
    with open(path_a) as fa, open(path_b) as fb:
        la, lb = fa.read().splitlines(), fb.read().splitlines()
    assert len(la) == len(lb), "row count %d vs %d" % (len(la), len(lb))
    assert la[0] == lb[0]
    hdr = la[0].split("\t")
    for ra, rb in zip(la[1:], lb[1:]):
        ca, cb = ra.split("\t"), rb.split("\t")
        assert len(ca) == len(cb)
        for i, (x, y) in enumerate(zip(ca, cb)):
            name = hdr[i] if i < len(hdr) else "col%d" % i
            if name in exact_cols or x == y:
                assert x == y, "%s: %r vs %r in %r" % (name, x, y, ra)
                continue
            fx, fy = float(x), float(y)
            if name.startswith("PRB") or name.startswith("POSTPRB") or name == "col5":
                assert abs(fx - fy) <= 1e-2 * max(abs(fx), abs(fy)) + 1e-300, "%s: %r vs %r" % (name, x, y)
            else:
                assert abs(fx - fy) <= tol + 1.0001e-4, "%s: %r vs %r" % (name, x, y)   # + print rounding
