"""CPU tests: the oracle against the reference's known answers and against the reference's own
object code; host-compiled device arithmetic against the oracle; C-ABI exports."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from helpers import load_fixture, sane_starts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
def test_golden_single_subject(k):
    """sub_samples@log_likelihoods[,1] / @summed_log_prior[,1] of the reference fixtures."""
    fx = load_fixture(k)
    d = fx.odata("sub")
    pr = fx.oprior("sub_prior")
    th = fx.g["sub_theta"]
    n_clean = 0
    for c in range(th.shape[0]):
        ld = ob.trial_logdens(fx.om, d, th[c])
        v = ob.sumloglike_rinit(fx.om, d, th[c])
        gold = fx.g["sub_ll"][c]
        if np.all(ld > np.log(1e-12)):  # no (1 - cdf) ~ 0 trials: must match to rounding
            assert abs(v - gold) <= 1e-12 * abs(gold), (k, c, v, gold)
            n_clean += 1
        else:  # cancellation-dominated trials present: same to ~1 %
            assert abs(v - gold) <= 0.3 * abs(gold)
        assert abs(ob.sumlogprior(pr, th[c]) - fx.g["sub_lp"][c]) <= 1e-13 * max(1.0, abs(fx.g["sub_lp"][c]))
    assert n_clean >= 0.8 * th.shape[0]


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
def test_golden_hierarchical(k):
    """pop_samples: subject ll / lp goldens, phi hyper-likelihood and hyper-prior goldens."""
    fx = load_fixture(k)
    pp, hp = fx.oprior("p_prior"), fx.oprior("h_prior")
    n_clean = n_all = 0
    for s in range(fx.n_pop):
        d = fx.odata(f"pop{s}")
        ths = fx.g["pop_theta_all"][s]
        for c in range(ths.shape[0]):
            ld = ob.trial_logdens(fx.om, d, ths[c])
            v = ob.sumloglike_rinit(fx.om, d, ths[c])
            gold = fx.g[f"pop{s}_ll"][c]
            n_all += 1
            if np.all(ld > np.log(1e-12)):
                assert abs(v - gold) <= 1e-12 * abs(gold), (k, s, c, v, gold)
                n_clean += 1
            else:
                assert abs(v - gold) <= 0.3 * abs(gold)
            lp = ob.sumlogprior(pp, ths[c])
            assert abs(lp - fx.g[f"pop{s}_lp"][c]) <= 1e-13 * max(1.0, abs(lp))
    assert n_clean >= 0.8 * n_all
    phi = fx.g["phi_theta"]
    npar = phi.shape[1] // 2
    allth = fx.g["pop_theta_all"]
    for c in range(phi.shape[0]):
        tot = sum(ob.sumlogprior(pp, allth[s, c], phi[c, :npar], phi[c, npar:]) for s in range(allth.shape[0]))
        assert abs(tot - fx.g["phi_ll"][c]) <= 1e-13 * abs(fx.g["phi_ll"][c])
        assert abs(ob.sumlogprior(hp, phi[c]) - fx.g["phi_lp"][c]) <= 1e-13 * abs(fx.g["phi_lp"][c])


needs_ref = pytest.mark.skipif(ob.ref_lib() is None, reason="oracle/_ref (reference object code) was never built")


@needs_ref
@pytest.mark.parametrize("k", [3, 6])
def test_density_bitwise_vs_reference_object_code(k):
    """orc_lba_cell == lba_class::dlba of /root/reference/src/de.o, bit for bit (same Phi/phi shim)."""
    fx = load_fixture(k)
    R, L = ob.ref_lib(), ob.lib()
    na = fx.ct.n_acc
    rng = np.random.default_rng(k)
    n = 0
    for s in range(2):
        rt, cell = fx.g[f"pop{s}_rt"], fx.g[f"pop{s}_cell"]
        thetas = list(fx.g["pop_theta_all"][s][::6]) + list(sane_starts(fx, 6, rng))
        for th in thetas:
            th = ob.f64(th)
            for cc in np.unique(cell):
                P = np.zeros((6, na))
                L.orc_cell_params(C.byref(fx.om.c), ob.ptr(th), int(cc), ob.ptr(P))
                r = ob.f64(rt[cell == cc])
                o1, o2 = np.zeros_like(r), np.zeros_like(r)
                u = np.zeros(4 * na + 8)
                R.ref_set_uniform_stream(ob.ptr(u), len(u))
                R.ref_lba_cell(ob.ptr(P), na, ob.ptr(fx.om.posdrift, ob.c_u8p), ob.ptr(r), len(r), ob.ptr(o2))
                L.orc_lba_cell(ob.ptr(P), na, ob.ptr(fx.om.posdrift, ob.c_u8p), None, ob.ptr(r), len(r), ob.ptr(o1))
                assert np.array_equal(o1, o2)
                n += len(r)
    assert n > 10000


@needs_ref
def test_selection_bitwise_vs_reference_object_code():
    """get_chains / get_subchains of de.o vs the restatement, same injected uniforms (tie-free keys)."""
    R, L = ob.ref_lib(), ob.lib()
    rng = np.random.default_rng(11)
    for _ in range(300):
        nchain = int(rng.integers(3, 130))
        k = int(rng.integers(0, nchain))
        u = ob.f64(rng.uniform(size=2 * nchain + 5))
        keys = (u * 2147483647.0).astype(np.int64)
        if len(np.unique(keys[:nchain + 1])) != nchain + 1:
            continue
        R.ref_set_uniform_stream(ob.ptr(u), len(u))
        o = (C.c_uint * 2)()
        R.ref_get_chains(nchain, k, 2, o)
        r = ob.make_rng(stream=u)
        a, o2 = ob.Addr(), (C.c_uint * 2)()
        L.orc_get_chains(nchain, k, 2, C.byref(r), C.byref(a), o2)
        assert list(o) == list(o2) and r.pos == R.ref_uniform_stream_pos() == nchain - 1
        R.ref_set_uniform_stream(ob.ptr(u), len(u))
        s1 = (C.c_uint * nchain)()
        n1 = R.ref_get_subchains(nchain, s1)
        r = ob.make_rng(stream=u)
        s2 = (C.c_uint * nchain)()
        n2 = L.orc_get_subchains(nchain, C.byref(r), C.byref(a), s2)
        assert n1 == n2 and list(s1[:n1]) == list(s2[:n2]) and r.pos == R.ref_uniform_stream_pos() == nchain + 1


@needs_ref
def test_tnorm_bitwise_vs_reference_object_code():
    R, L = ob.ref_lib(), ob.lib()
    rng = np.random.default_rng(3)
    for _ in range(20000):
        x, mean, sd = rng.uniform(-1, 12), rng.uniform(-2, 8), rng.uniform(-0.2, 4)
        lo, up, lg = rng.choice([0.0, -np.inf, 0.5]), rng.choice([np.inf, 10.0]), int(rng.integers(0, 2))
        a, b = R.ref_tnorm_d(x, mean, sd, lo, up, lg), L.orc_tnorm_d(x, mean, sd, lo, up, lg)
        assert a == b or (np.isnan(a) and np.isnan(b))


def _subject_state(fx, od, oprior, nchain, rng, center=None, phi=None):
    th = sane_starts(fx, nchain, rng, center=center)
    D = th.shape[1]
    lp = np.array([ob.sumlogprior(oprior, th[c], None if phi is None else phi[c, :D], None if phi is None else phi[c, D:])
                   for c in range(nchain)])
    ll = np.array([ob.sumloglike(fx.om, od, th[c]) for c in range(nchain)])
    return th, lp, ll


def _hier_state(fx, S, nchain, rng):
    D = fx.ct.npar
    opp, ohp = fx.oprior("p_prior"), fx.oprior("h_prior")
    center = np.concatenate([fx.g["pop_mean"], fx.g["pop_scale"]])
    phi0 = center[None, :] * (1.0 + 0.1 * rng.standard_normal((nchain, 2 * D)))
    subj = [_subject_state(fx, fx.odata(f"pop{s}"), opp, nchain, rng, center=fx.g["ps"][s], phi=phi0) for s in range(S)]
    lp0 = np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(nchain)])
    ll0 = np.array([sum(ob.sumlogprior(opp, subj[s][0][c], phi0[c, :D], phi0[c, D:]) for s in range(S)) for c in range(nchain)])
    return (phi0, lp0, ll0), subj


@needs_ref
def test_sampler_sweeps_bitwise_vs_reference_object_code():
    """de_class::crossover / migration of src/de.o (src/de.cpp:111-199), driven on hand-built objects with an
    injected uniform stream, against the restatement fed the same stream: theta, log prior and log
    likelihood of every chain bit-identical, the same number of uniforms consumed."""
    ob.ref2_prime()
    R = ob.ref_lib()
    for k, prior_name in ((2, "sub_prior"), (6, "sub_prior"), (3, "p_prior")):
        fx = load_fixture(k)
        rng = np.random.default_rng(30 + k)
        od, op, D = fx.odata("sub"), fx.oprior(prior_name), fx.ct.npar
        nchain = 3 * D
        th, lp, ll = _subject_state(fx, od, op, nchain, rng)
        for kind, para in ((0, -1), (1, -1), (0, 1), (1, 2), (0, -1)):
            u = ob.ref2_set_stream(rng.uniform(size=400000))
            a = ob.ref2_sweep_subject(kind, para, D, fx.om, od, op, th, lp, ll)
            used = R.ref_uniform_stream_pos()
            pop = ob.OPop(th, lp, ll, 2, 1)
            r = ob.make_rng(stream=u)
            f = ob.lib().orc_crossover_subject if kind == 0 else ob.lib().orc_migration_subject
            f(C.byref(ob.make_de(D, nchain)), C.byref(pop.c), C.byref(op.c), C.byref(fx.om.c), C.byref(od.c), C.byref(r), C.c_uint(0),
              C.c_uint(1), C.c_int(para))
            assert r.pos == used and used > nchain
            assert np.array_equal(a[0], pop.theta) and np.array_equal(a[1], pop.lp) and np.array_equal(a[2], pop.ll)
            assert not np.array_equal(a[0], th)  # something was accepted
            th, lp, ll = a


@needs_ref
@pytest.mark.parametrize("pblocked", [False, True])
def test_run_chains_bitwise_vs_reference_object_code(pblocked):
    """The reference's whole 1-level driver de_class::run_chains (src/de.cpp:201-242: migration decision,
    blocked / unblocked crossover, theta_phi::store thinning) from src/de.o vs orc_run_subject."""
    ob.ref2_prime()
    fx = load_fixture(2)
    rng = np.random.default_rng(5)
    od, op, D = fx.odata("sub"), fx.oprior("sub_prior"), fx.ct.npar
    nchain, nmc, thin = 3 * D, 5, 3
    th, lp, ll = _subject_state(fx, od, op, nchain, rng)
    u = ob.ref2_set_stream(rng.uniform(size=1500000))
    ot, olp, oll = ob.ref2_run_chains(D, fx.om, od, op, th, lp, ll, nmc, thin, sub_migration_prob=0.3, is_pblocked=pblocked)
    used = ob.ref_lib().ref_uniform_stream_pos()
    pop = ob.OPop(th, lp, ll, nmc, thin)
    r = ob.make_rng(stream=u)
    ob.run_subject(ob.make_de(D, nchain, sub_migration_prob=0.3, is_pblocked=pblocked), pop, op, fx.om, od, r, 0, (nmc - 1) * thin)
    assert r.pos == used
    assert np.array_equal(ot, pop.out_theta) and np.array_equal(olp, pop.out_lp) and np.array_equal(oll, pop.out_ll)
    assert not np.array_equal(ot[0], ot[-1])


@needs_ref
@pytest.mark.parametrize("blocked", [False, True])
def test_run_hchains_bitwise_vs_reference_object_code(blocked):
    """The reference's hierarchical driver de_class::run_hchains (src/de.cpp:272-383) from src/de.o -- phi
    crossover / migration with refreshed hyper-likelihood, subject steps with phi-driven priors and the stale
    log prior, per-parameter blocking, storage -- vs orc_run_hier: every stored sample bit-identical."""
    ob.ref2_prime()
    fx = load_fixture(2)
    rng = np.random.default_rng(9)
    S, D = fx.n_pop, fx.ct.npar
    nchain = 6 * D
    nmc, thin = (3, 1) if blocked else (4, 2)
    phi_s, subj_s = _hier_state(fx, S, nchain, rng)
    opp, ohp = fx.oprior("p_prior"), fx.oprior("h_prior")
    datas = [fx.odata(f"pop{s}") for s in range(S)]
    kw = dict(pop_migration_prob=0.3, sub_migration_prob=0.3, is_hblocked=blocked, is_pblocked=blocked)
    u = ob.ref2_set_stream(rng.uniform(size=4000000))
    (pt, plp, pll), subs = ob.ref2_run_hchains(2 * D, fx.om, datas, opp, ohp, phi_s, subj_s, nmc, thin, **kw)
    used = ob.ref_lib().ref_uniform_stream_pos()
    phi = ob.OPop(*phi_s, nmc, thin)
    pops = [ob.OPop(*s, nmc, thin) for s in subj_s]
    r = ob.make_rng(stream=u)
    ob.run_hier(ob.make_de(2 * D, nchain, **kw), phi, pops, opp, ohp, fx.om, datas, r, (nmc - 1) * thin)
    assert r.pos == used
    assert np.array_equal(pt, phi.out_theta) and np.array_equal(plp, phi.out_lp) and np.array_equal(pll, phi.out_ll)
    for s in range(S):
        assert np.array_equal(subs[s][0], pops[s].out_theta) and np.array_equal(subs[s][1], pops[s].out_lp)
        assert np.array_equal(subs[s][2], pops[s].out_ll)
    assert not np.array_equal(pt[0], pt[-1])


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    L = ob.lib()
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        o = (C.c_uint * 4)()
        L.orc_philox4x32_10((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), o)
        assert list(o) == want


@pytest.fixture(scope="module")
def hostmath():
    """The engine's device math headers compiled as host C++ (arithmetic check without a GPU)."""
    out = os.path.join(ROOT, "tests", "host", "libhostmath.so")
    src = os.path.join(ROOT, "tests", "host", "host_math_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", out], check=True)
    H = C.CDLL(out)
    for f in ("hm_pnorm_std", "hm_dnorm_std", "hm_pnorm5", "hm_dnorm4", "hm_dprior1", "hm_draw_uniform"):
        getattr(H, f).restype = C.c_double
    H.hm_pnorm_std.argtypes = [C.c_double]
    H.hm_dnorm_std.argtypes = [C.c_double]
    H.hm_pnorm5.argtypes = [C.c_double] * 3 + [C.c_int]
    H.hm_dnorm4.argtypes = [C.c_double] * 3 + [C.c_int]
    H.hm_dprior1.argtypes = [C.c_int] + [C.c_double] * 5 + [C.c_int]
    H.hm_draw_uniform.argtypes = [C.c_ulonglong] + [C.c_uint] * 6
    return H


def test_device_math_on_host_matches_oracle(hostmath):
    """gg_lba.cuh / gg_math.cuh (host build) vs the oracle on fixture cells: <= 1e-10 rel on log density."""
    H, L = hostmath, ob.lib()
    for k in (3, 6):
        fx = load_fixture(k)
        na = fx.ct.n_acc
        rng = np.random.default_rng(k)
        rt, cell = fx.g["pop0_rt"], fx.g["pop0_cell"]
        thetas = list(fx.g["pop_theta_all"][0][::4]) + list(sane_starts(fx, 10, rng))
        worst = 0.0
        for th in thetas:
            th = ob.f64(th)
            for cc in np.unique(cell):
                P = np.zeros((6, na))
                L.orc_cell_params(C.byref(fx.om.c), ob.ptr(th), int(cc), ob.ptr(P))
                r = ob.f64(rt[cell == cc])
                o1, o2 = np.zeros_like(r), np.zeros_like(r)
                L.orc_lba_cell(ob.ptr(P), na, ob.ptr(fx.om.posdrift, ob.c_u8p), None, ob.ptr(r), len(r), ob.ptr(o1))
                P2 = P.copy()
                P2[1] -= P2[0]
                H.hm_lba_cell(ob.ptr(ob.f64(P2)), na, ob.ptr(fx.om.posdrift, ob.c_u8p), ob.ptr(r), len(r), ob.ptr(o2))
                ok = (o1 > 1e-9) & (o2 > 1e-9)
                if ok.any():
                    rel = np.abs(np.log(o1[ok]) - np.log(o2[ok])) / np.maximum(np.abs(np.log(o1[ok])), 1.0)
                    worst = max(worst, rel.max())
                assert np.all(np.abs(o1 - o2) <= 1e-9 * np.maximum(o1, 1e-10) + 1e-14)
        assert worst <= 1e-10


def test_fast_norm_pair_accuracy(hostmath):
    """gg_fastmath.cuh norm_pair (host build) against 50-digit mpmath: Phi and phi to a few ulp,
    relative accuracy kept in the lower tail, special values like R's pnorm / dnorm."""
    import mpmath as mp
    mp.mp.dps = 50
    H = hostmath
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(0)
    z = np.concatenate([rng.uniform(-37.5, 9, 3000), rng.uniform(-6, 6, 3000), rng.uniform(-1, 1, 1000)])
    cdf, pdf = np.zeros_like(z), np.zeros_like(z)
    H.hm_norm_pair(z.ctypes.data_as(dp), len(z), cdf.ctypes.data_as(dp), pdf.ctypes.data_as(dp))
    for zi, c, p in zip(z, cdf, pdf):
        rc, rp = mp.ncdf(mp.mpf(float(zi))), mp.npdf(mp.mpf(float(zi)))
        assert abs((mp.mpf(float(c)) - rc) / rc) < 8e-16, (zi, c)
        assert abs((mp.mpf(float(p)) - rp) / rp) < 6e-16, (zi, p)
    zs = np.array([0.0, np.inf, -np.inf, 40.0, -40.0, 1e300, np.nan])
    cdf, pdf = np.zeros_like(zs), np.zeros_like(zs)
    H.hm_norm_pair(zs.ctypes.data_as(dp), len(zs), cdf.ctypes.data_as(dp), pdf.ctypes.data_as(dp))
    assert list(cdf[:6]) == [0.5, 1.0, 0.0, 1.0, 0.0, 1.0] and np.isnan(cdf[6]) and np.isnan(pdf[6])
    assert pdf[0] == 0.3989422804014327 and list(pdf[1:6]) == [0.0] * 5
    # the batched twins the trial loop runs (table-driven exponential: 2^(j/32) table + degree-6 polynomial): same accuracy
    # class over the whole finite range, four-wide and two-wide versions bit-identical to each other
    z = np.concatenate([rng.uniform(-37.5, 9, 4000), rng.uniform(-6, 6, 3000), rng.uniform(-1, 1, 1000), np.array([0.0, -37.5, 37.5, -45.0, 8.3, 1e-300, -1e-300, 5e-324])])
    z = z[: len(z) // 4 * 4]
    c4, p4, c2, p2 = (np.zeros_like(z) for _ in range(4))
    H.hm_norm_pairs_hot(z.ctypes.data_as(dp), len(z), c4.ctypes.data_as(dp), p4.ctypes.data_as(dp), c2.ctypes.data_as(dp), p2.ctypes.data_as(dp))
    assert np.array_equal(c4, c2) and np.array_equal(p4, p2)
    worst_c = worst_p = 0.0
    for zi, c, p_ in zip(z, c4, p4):
        zc = max(min(float(zi), 37.5), -37.5)  # the loop clamps |z| at 37.5 (both values are below 1e-305 there)
        rc, rp = mp.ncdf(mp.mpf(zc)), mp.npdf(mp.mpf(zc))
        worst_c, worst_p = max(worst_c, float(abs((mp.mpf(float(c)) - rc) / rc))), max(worst_p, float(abs((mp.mpf(float(p_)) - rp) / rp)))
    assert worst_c < 8e-16 and worst_p < 6e-16, (worst_c, worst_p)
    # the Estrin-scheme twin used by the cell-table build: same accuracy class, same special values
    z = np.concatenate([rng.uniform(-37.5, 9, 2000), rng.uniform(-6, 6, 2000)])
    low = np.zeros_like(z)
    H.hm_norm_cdf_lowlatency(z.ctypes.data_as(dp), len(z), low.ctypes.data_as(dp))
    for zi, c in zip(z, low):
        rc = mp.ncdf(mp.mpf(float(zi)))
        assert abs((mp.mpf(float(c)) - rc) / rc) < 1e-15, (zi, c)
    low = np.zeros_like(zs)
    H.hm_norm_cdf_lowlatency(zs.ctypes.data_as(dp), len(zs), low.ctypes.data_as(dp))
    assert list(low[:6]) == [0.5, 1.0, 0.0, 1.0, 0.0, 1.0] and np.isnan(low[6])


def test_device_prior_math_on_host_matches_oracle(hostmath):
    H, L = hostmath, ob.lib()
    rng = np.random.default_rng(0)
    for dist in (1, 2, 3, 4, 5, 6, 7):
        for _ in range(2000):
            x, p0, p1 = rng.uniform(-1, 6), rng.uniform(0.1, 4), rng.uniform(0.1, 4)
            lo, up, lg = float(rng.choice([0.0, -np.inf, 0.5])), float(rng.choice([np.inf, 10.0])), int(rng.integers(0, 2))
            if dist == 2:
                lo, up = 0.0, 10.0
            pr = ob.OPrior([p0], [p1], [lo], [up], [dist], [lg])
            xx, out = ob.f64([x]), np.zeros(1)
            L.orc_dprior(C.byref(pr.c), ob.ptr(pr.p0), ob.ptr(pr.p1), ob.ptr(xx), ob.ptr(out))
            got = H.hm_dprior1(dist, x, p0, p1, lo, up, lg)
            a = out[0]
            assert (np.isnan(a) and np.isnan(got)) or a == got or abs(a - got) <= 1e-12 * max(1.0, abs(a)), (dist, x, p0, p1, lo, up, lg, a, got)


def test_uniform_addressing_matches_oracle(hostmath):
    """Same (seed, pop, iter, sweep, chain, purpose, slot) -> same uniform in engine code and oracle."""
    H, L = hostmath, ob.lib()
    rng = np.random.default_rng(4)
    for _ in range(500):
        seed = int(rng.integers(0, 2**63))
        pop, it, sw, ch, pu, sl = (int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)), int(rng.integers(0, 4096)),
                                   int(rng.integers(0, 65536)), int(rng.integers(0, 7)), int(rng.integers(0, 1000)))
        r = ob.make_rng(seed=seed)
        a = ob.Addr(pop, it, sw, ch, pu, sl)
        assert L.orc_uniform(C.byref(r), C.byref(a)) == H.hm_draw_uniform(seed, pop, it, sw, ch, pu, sl)


def test_cabi_library_exports_every_declared_symbol():
    """libggdmc_b200.so loads (no GPU needed) and exports everything include/ggdmc_b200.h declares."""
    from ggdmc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ggdmc_b200.h")).read()
    declared = set(re.findall(r"\b(ggdmc_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.ggdmc_b200_abi_version() == 2


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the compute entry points fail loudly (GGDMC_ERR_CUDA)."""
    from ggdmc_b200 import _lib, engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    fx = load_fixture(2)
    with pytest.raises(_lib.GgdmcError) as ei:
        engine.trial_logdens(fx.ct, fx.trials("sub"), fx.g["sub_theta"][:2])
    assert ei.value.code == _lib.ERR_CUDA


def test_rprior_draws_follow_every_prior_family():
    """ggdmc_b200.init.rprior (host side of initialise_theta / initialise_phi): support and moments of all seven
    families, checked against the oracle's density through a histogram."""
    from ggdmc_b200.init import rprior, _first_valid
    from ggdmc_b200.model import PriorTable
    #        tnorm        beta_lu     gamma_l     lnorm_l     cauchy      unif        norm
    p0 = np.array([1.0, 2.0, 3.0, 0.2, 0.5, -1.0, 2.0])
    p1 = np.array([0.8, 3.0, 0.5, 0.4, 1.5, 4.0, 0.7])
    lo = np.array([0.2, -1.0, 0.5, 1.0, -3.0, -np.inf, -np.inf])
    up = np.array([2.5, 4.0, np.inf, np.inf, 6.0, np.inf, np.inf])
    dist = np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.int32)
    tab = PriorTable(7, p0, p1, lo, up, dist, np.zeros(7, dtype=np.uint8), [f"p{i}" for i in range(7)])
    x = rprior(tab, 200_000, np.random.default_rng(5))
    assert x.shape == (200_000, 7) and np.all(np.isfinite(x))
    for i in range(7):
        op = ob.OPrior(p0[i:i + 1], p1[i:i + 1], lo[i:i + 1], up[i:i + 1], dist[i:i + 1], np.zeros(1, dtype=np.uint8))  # log_p = 0: the density
        xi = x[:, i]
        a, b = (p0[i], p1[i]) if dist[i] == 6 else (lo[i], up[i])
        assert np.all(xi >= a) and np.all(xi <= b), i
        # histogram on the central 90 % against the oracle's density (dprior of one parameter at a time)
        q = np.quantile(xi, [0.05, 0.95])
        edges = np.linspace(q[0], q[1], 21)
        h, _ = np.histogram(xi, bins=edges)
        mid = 0.5 * (edges[1:] + edges[:-1])
        dens = np.array([ob.sumlogprior(op, [m]) for m in mid])
        expect = dens * (edges[1] - edges[0]) * len(xi)
        assert np.all(np.abs(h - expect) <= 6 * np.sqrt(expect) + 0.01 * expect), (i, h, expect)
    v = np.array([[False, True, True], [False, False, False], [True, False, True]])
    assert list(_first_valid(v)) == [1, -1, 0]


def test_ctypes_mirrors_match_the_c_header_layout(tmp_path):
    """Every struct of include/ggdmc_b200.h has the size and field offsets its ctypes mirror in ggdmc_b200/_lib.py
    assumes (a C probe compiled against the header prints offsetof / sizeof)."""
    from ggdmc_b200 import _lib as B
    pairs = {"ggdmc_model_t": B.ModelT, "ggdmc_trials_t": B.TrialsT, "ggdmc_prior_t": B.PriorT, "ggdmc_config_t": B.ConfigT,
             "ggdmc_samples_t": B.SamplesT, "ggdmc_start_t": B.StartT}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "ggdmc_b200.h"', 'int main(void) {']
    for cname, T in pairs.items():
        lines.append(f'  printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for f, _ in T._fields_:
            lines.append(f'  printf("{cname} {f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    got = {tuple(l.split()[:2]): int(l.split()[2]) for l in out if l.strip()}
    for cname, T in pairs.items():
        assert got[(cname, "sizeof")] == C.sizeof(T), cname
        for f, _ in T._fields_:
            assert got[(cname, f)] == getattr(T, f).offset, (cname, f)
