"""CPU tests of the DDM ("fastdm") path: the oracle's restatement of ddm_class against the reference's own object
code (likelihood_class::ddm_likelihood and de_class::run_chains of src/de.o), against the published series it
implements, and the engine's device header compiled for the host against the oracle.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from helpers import ddm_model, ddm_prior, ddm_simulate, ddm_theta
from test_oracle_cpu import hostmath, needs_ref  # noqa: F401  (fixture + marker)

DBL_MIN = 2.2250738585072014e-308


def _grid_data(rng, n=240):
    cell = np.sort(rng.integers(0, 4, n)).astype(np.uint16)
    rt = np.concatenate([rng.uniform(0.15, 0.5, n // 3), rng.uniform(0.3, 1.5, n - 2 * (n // 3)), rng.uniform(1.0, 6.0, n // 3)])
    return ob.OData(rng.permutation(rt), cell)


def _same(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


def _oracle_density(om, d, theta):
    out = np.zeros(len(d.rt))
    for c in range(om.n_cell):
        idx = np.where(d.cell == c)[0]
        P = np.zeros(10 * om.n_acc)
        ob.lib().orc_cell_params(C.byref(om.c), ob.ptr(ob.f64(theta)), c, ob.ptr(P))
        ok, dens = ob.ddm_cell(P.reshape(10, om.n_acc)[:, 0], bool(om.posdrift[c]), d.rt[idx])
        out[idx] = dens
    return out


def _edge_thetas(rng):
    """Parameter vectors that hit validate_parameters' eight rules, rt <= t0, and extreme but valid corners."""
    base = ddm_theta(rng, 3)
    names = ["a", "st0", "sv", "sz", "t0", "v.s1", "v.s2", "z"]
    out = []
    for nm, val in (("a", -1.0), ("a", 0.0), ("st0", -0.1), ("sv", -0.5), ("sz", -0.1), ("sz", 10.0), ("t0", -1.0), ("z", 0.0),
                    ("z", 50.0), ("z", float("nan")), ("a", float("nan")), ("t0", 5.9), ("sv", 40.0), ("v.s1", 25.0), ("v.s2", -25.0),
                    ("a", 0.05), ("a", 30.0), ("st0", 1e-9), ("sz", 1e-9), ("t0", float("inf"))):
        th = base.copy()
        th[names.index(nm)] = val
        if nm == "a" and np.isfinite(val) and val > 0:
            th[names.index("z")] = 0.5 * val
            th[names.index("sz")] = 0.1 * val
        out.append(th)
    return out


@needs_ref
@pytest.mark.parametrize("precision,s", [(3.0, 1.0), (2.5, 1.0), (3.0, 0.1), (2.0, 2.0)])
def test_ddm_density_bitwise_vs_reference_object_code(precision, s):
    """likelihood_class::ddm_likelihood of src/de.o (design_class::set_parameter_values, ddm_class::set_parameters,
    validate_parameters, dddm -- @hdr/likelihood.h:129-161, @hdr/ddm.h) on a hand-built "fastdm" object vs orc_ddm_cell:
    every trial density bit-identical, for all four variability combinations and both boundaries."""
    ct, om = ddm_model(precision, s)
    rng = np.random.default_rng(int(precision * 10 + s * 100))
    d = _grid_data(rng)
    n_pos = 0
    for it in range(80):
        th = ddm_theta(rng, it % 4)
        if s != 1.0:  # a, v, sv are divided by s (@hdr/ddm.h:203-206): keep the scaled model in the same regime
            for i in (0, 2, 5, 6):
                th[i] *= s
            th[[3, 7]] *= 1.0  # sz, z are NOT rescaled by the reference (:216-218 divide by the scaled a)
            th[7] = th[0] / s * 0.5
            th[3] = min(th[3], 0.2 * th[0] / s)
        ref = ob.ref2_ddm_density(om, d, th)
        mine = _oracle_density(om, d, th)
        assert np.all(_same(ref, mine)), (it, th, np.argwhere(~_same(ref, mine))[:3])
        n_pos += int(np.sum(ref > 1e-6))
    assert n_pos > 20 * len(d.rt)  # the comparison is not about floors


@needs_ref
def test_ddm_edge_cases_bitwise_vs_reference_object_code():
    """Invalid cells (every rule of validate_parameters, NaN passing the comparisons), rt below the non-decision time,
    extreme drifts / boundaries: the restatement takes the reference's branch every time."""
    ct, om = ddm_model()
    rng = np.random.default_rng(3)
    d = _grid_data(rng)
    n_floor = 0
    for th in _edge_thetas(rng):
        ref = ob.ref2_ddm_density(om, d, th)
        mine = _oracle_density(om, d, th)
        assert np.all(_same(ref, mine)), (th, np.argwhere(~_same(ref, mine))[:3])
        n_floor += int(np.all(ref == 1e-10))
    assert n_floor >= 8  # invalid cells: every trial 1e-10 (@hdr/likelihood.h:158)


def _navarro_fuss_lower(t, v, a, w, terms=400):
    """Lower-boundary first-passage density, large-time series (Navarro & Fuss 2009, eq. 5/6), no variabilities."""
    k = np.arange(1, terms + 1)[:, None]
    return (np.pi / a**2) * np.exp(-v * a * w - v * v * t / 2) * np.sum(k * np.exp(-(k * np.pi) ** 2 * t / (2 * a * a)) * np.sin(k * np.pi * w), axis=0)


def test_ddm_density_matches_published_series():
    """Without variabilities the density is the Navarro-Fuss series; the reference truncates it at an absolute error of
    1e-6 (ddm::EPSILON).  Upper boundary = lower boundary with (v, w) -> (-v, 1 - w)."""
    rng = np.random.default_rng(11)
    for _ in range(40):
        a, v, t0 = rng.uniform(0.6, 2.5), rng.normal(0, 2), rng.uniform(0.1, 0.3)
        z = a * rng.uniform(0.3, 0.7)
        rt = t0 + rng.uniform(0.05, 3.0, 50)
        P = [a, 0.0, 3.0, 1.0, 0.0, 0.0, 0.0, t0, v, z]
        ok, lo = ob.ddm_cell(P, False, rt)
        ok2, up = ob.ddm_cell(P, True, rt)
        assert ok and ok2
        assert np.all(np.abs(lo - _navarro_fuss_lower(rt - t0, v, a, z / a)) <= 2e-6)
        assert np.all(np.abs(up - _navarro_fuss_lower(rt - t0, -v, a, 1 - z / a)) <= 2e-6)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_ddm_defective_densities_integrate_to_one(kind):
    """Size-independent property: P(lower) + P(upper) = 1, with every variability switched on in turn."""
    rng = np.random.default_rng(20 + kind)
    for _ in range(3):
        th = ddm_theta(rng, kind)
        a, st0, sv, sz, t0, v1, v2, z = th
        t = np.linspace(t0 + 1e-4, t0 + st0 + 14.0, 14001 if kind < 3 else 4001)
        P = [a, 0.0, 3.0, 1.0, st0, sv, sz, t0, v1 * 0.4, z]
        lo, up = ob.ddm_cell(P, False, t)[1], ob.ddm_cell(P, True, t)[1]
        total = np.trapezoid(lo + up, t)
        assert abs(total - 1.0) < 5e-3, (th, total)


def test_ddm_device_math_on_host_matches_oracle(hostmath):
    """gg_ddm.cuh compiled for the host vs the oracle: valid flags and NaNs equal, edge cases included.  The header
    keeps the reference's term counts and integration grids but produces the terms with reciprocals and recurrences, so
    densities agree to a few ulp, not bit for bit: <= 1e-12 relative (observed 1.4e-14 over 1.9e6 evaluations)."""
    H = hostmath
    rng = np.random.default_rng(5)
    d = _grid_data(rng, 120)
    rt = ob.f64(d.rt)
    thetas = [ddm_theta(rng, k % 4) for k in range(60)] + _edge_thetas(rng)
    n_exact = n_tot = 0
    for th in thetas:
        a, st0, sv, sz, t0, v1, v2, z = th
        for upper in (False, True):
            P = ob.f64([a, 0.0, 3.0, 1.0, st0, sv, sz, t0, v1, z])
            ok, want = ob.ddm_cell(P, upper, rt)
            got = np.zeros(len(rt))
            ok_h = H.hm_ddm_cell(ob.ptr(P), int(upper), ob.ptr(rt), len(rt), ob.ptr(got))
            assert bool(ok_h) == ok
            assert np.array_equal(np.isnan(got), np.isnan(want))
            fin = ~np.isnan(want)
            assert np.all(np.abs(got[fin] - want[fin]) <= 1e-12 * np.abs(want[fin]) + 1e-15), th
            n_exact += int(np.sum(np.abs(got[fin] - want[fin]) <= 1e-13 * np.abs(want[fin]) + 1e-18))
            n_tot += len(rt)
    assert n_exact >= 0.95 * n_tot  # edge-case vectors included (overflowing drifts, huge boundaries)


def _ddm_subject_state(om, d, oprior, nchain, rng, kind=1):
    th = np.stack([ddm_theta(rng, kind) for _ in range(nchain)])
    center = th[0]
    th = center[None, :] * (1.0 + 0.03 * rng.standard_normal(th.shape))
    lp = np.array([ob.sumlogprior(oprior, t) for t in th])
    ll = np.array([ob.lib().orc_sumloglike(C.byref(om.c), C.byref(d.c), ob.ptr(ob.f64(t)), None, None) for t in th])
    return th, lp, ll


@needs_ref
@pytest.mark.parametrize("pblocked", [False, True])
def test_ddm_run_chains_bitwise_vs_reference_object_code(pblocked):
    """de_class::run_chains of src/de.o on a "fastdm" likelihood object -- sumloglike's DDM branch
    (@hdr/likelihood.h:295-305: log(max(density, DBL_MIN))) included -- vs orc_run_subject: stored samples bit-identical
    and the same number of uniforms consumed (the DDM density draws none)."""
    ob.ref2_prime()
    ct, om = ddm_model()
    rng = np.random.default_rng(41)
    truth = ddm_theta(rng, 1)
    rt, cell = ddm_simulate(truth, 60, rng)
    d = ob.OData(rt, cell)
    pt, op = ddm_prior()
    D = ct.npar
    nchain, nmc, thin = 3 * D, 4, 2
    ob.lib().orc_sumloglike.restype = C.c_double
    th, lp, ll = _ddm_subject_state(om, d, op, nchain, rng)
    u = ob.ref2_set_stream(rng.uniform(size=400000))
    ot, olp, oll = ob.ref2_run_chains(D, om, d, op, th, lp, ll, nmc, thin, sub_migration_prob=0.3, is_pblocked=pblocked)
    used = ob.ref_lib().ref_uniform_stream_pos()
    ob.ref_lib().ref2_set_model_type(0)
    pop = ob.OPop(th, lp, ll, nmc, thin)
    r = ob.make_rng(stream=u)
    ob.run_subject(ob.make_de(D, nchain, sub_migration_prob=0.3, is_pblocked=pblocked), pop, op, om, d, r, 0, (nmc - 1) * thin)
    assert r.pos == used
    assert np.array_equal(ot, pop.out_theta) and np.array_equal(olp, pop.out_lp) and np.array_equal(oll, pop.out_ll)
    assert not np.array_equal(ot[0], ot[-1])
    assert np.all(np.isfinite(oll[-1]))


def test_ddm_model_objects_flatten_to_the_cell_table():
    """model@type "fastdm": build_cell_table (Python mirror) and the Rcpp glue's flatten_model (compiled against the Rcpp
    stand-in) turn model_boolean + constants + pnames into the 10-row table the engine takes, is_positive_drift per cell."""
    import glue_mock as G
    from ggdmc_b200.model import build_cell_table
    from helpers import ddm_objects
    ct, om = ddm_model()
    rng = np.random.default_rng(2)
    d = _grid_data(rng, 60)
    model, dmi = ddm_objects(d.rt, d.cell)
    got = build_cell_table(model, dmi.node_1_index, dmi.is_positive_drift)
    assert got.type == "fastdm" and got.param_src.shape == (4, 10, 2)
    assert np.array_equal(got.param_src, ct.param_src) and np.array_equal(got.posdrift, ct.posdrift)
    assert np.array_equal(got.const_val, ct.const_val) and got.pnames == ct.pnames
    n = ct.param_src.size
    buf, dims = (C.c_int * n)(), (C.c_int * 4)()
    assert G.lib().gh_flatten_model(G.r_dmi(dmi), buf, n, dims) == 0, G.lib().gh_last_error()
    assert list(dims) == [ct.n_acc, ct.n_cell, ct.npar, 3 + 1000 * 1]  # 3 constants, type GGDMC_MODEL_DDM
    assert np.array_equal(np.array(buf[:]).reshape(ct.param_src.shape), ct.param_src)
    with pytest.raises(ValueError, match="is_positive_drift"):  # per-accumulator flags are the LBA's convention
        build_cell_table(model, dmi.node_1_index, np.array([True, True]))


@needs_ref
def test_ddm_run_hchains_bitwise_vs_reference_object_code():
    """de_class::run_hchains of src/de.o (src/de.cpp:272-383) with "fastdm" likelihood objects for every subject --
    phi step, subject steps under phi-driven truncated-normal priors, migration at both levels -- vs orc_run_hier:
    every stored sample of phi and of every subject bit-identical, same number of uniforms consumed."""
    ob.ref2_prime()
    ct, om = ddm_model()
    rng = np.random.default_rng(8)
    S, D = 3, ct.npar
    center = ddm_theta(rng, 1)
    truths = [center * (1.0 + 0.05 * rng.standard_normal(D)) for _ in range(S)]
    datas = [ob.OData(*ddm_simulate(t, 40, rng)) for t in truths]
    lower = np.array([0.0, 0.0, 0.0, 0.0, 0.0, -20.0, -20.0, 0.0])
    p0, p1 = center.copy(), 0.3 * np.abs(center) + 0.1
    opp = ob.OPrior(p0, p1, lower, np.full(D, np.inf), np.full(D, 1, np.int32), np.ones(D, np.uint8))
    ohp = ob.OPrior(np.concatenate([center - 3.0, np.full(D, 0.01)]), np.concatenate([center + 3.0, np.full(D, 3.0)]), np.zeros(2 * D),
                    np.zeros(2 * D), np.full(2 * D, 6, np.int32), np.ones(2 * D, np.uint8))
    nchain, nmc, thin = 2 * 2 * D, 3, 2
    phi0 = np.concatenate([p0, p1])[None, :] * (1.0 + 0.05 * rng.standard_normal((nchain, 2 * D)))
    subj = []
    for s in range(S):
        th = truths[s][None, :] * (1.0 + 0.03 * rng.standard_normal((nchain, D)))
        lp = np.array([ob.sumlogprior(opp, th[c], phi0[c, :D], phi0[c, D:]) for c in range(nchain)])
        ll = np.array([ob.sumloglike(om, datas[s], th[c]) for c in range(nchain)])
        subj.append((th, lp, ll))
    lp0 = np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(nchain)])
    ll0 = np.array([sum(ob.sumlogprior(opp, subj[s][0][c], phi0[c, :D], phi0[c, D:]) for s in range(S)) for c in range(nchain)])
    kw = dict(pop_migration_prob=0.3, sub_migration_prob=0.3)
    u = ob.ref2_set_stream(rng.uniform(size=1500000))
    (pt, plp, pll), subs = ob.ref2_run_hchains(2 * D, om, datas, opp, ohp, (phi0, lp0, ll0), subj, nmc, thin, **kw)
    used = ob.ref_lib().ref_uniform_stream_pos()
    ob.ref_lib().ref2_set_model_type(0)
    phi = ob.OPop(phi0, lp0, ll0, nmc, thin)
    pops = [ob.OPop(*s, nmc, thin) for s in subj]
    r = ob.make_rng(stream=u)
    ob.run_hier(ob.make_de(2 * D, nchain, **kw), phi, pops, opp, ohp, om, datas, r, (nmc - 1) * thin)
    assert r.pos == used
    assert np.array_equal(pt, phi.out_theta) and np.array_equal(plp, phi.out_lp) and np.array_equal(pll, phi.out_ll)
    for s in range(S):
        assert np.array_equal(subs[s][0], pops[s].out_theta) and np.array_equal(subs[s][1], pops[s].out_lp)
        assert np.array_equal(subs[s][2], pops[s].out_ll)
    assert not np.array_equal(pt[0], pt[-1])


def test_ddm_integer_ceil_sqrt_equals_the_reference_expression(hostmath):
    """gg::ddm_ceil_sqrt (float root + two FP64 comparisons) against ceil(sqrt(x)) as get_N writes it (@hdr/ddm.h:413,
    421-423): equal everywhere except within one rounding of a perfect square from above -- where sqrt() rounds down onto
    the integer and the reference's ceil stays one short of the exact answer -- and with the reference's INT_MIN for NaN,
    negative and huge arguments."""
    H = hostmath
    H.hm_ddm_ceil_sqrt.argtypes = [C.c_double]
    H.hm_ddm_ceil_sqrt.restype = C.c_int
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.uniform(0, 50, 20000), 10.0 ** rng.uniform(-300, 11.9, 20000), rng.integers(0, 2000, 4000).astype(float) ** 2,
                         [0.0, 1e-320, 1.0, 4.0, 1e12, 1e13, 1e300, np.inf, -1.0, -1e-300, np.nan]])
    squares = rng.integers(1, 100000, 4000).astype(float) ** 2
    xs = np.concatenate([xs, np.nextafter(squares, 0), np.nextafter(squares, np.inf)])
    n_diff = 0
    for x in xs:
        got = H.hm_ddm_ceil_sqrt(float(x))
        r = np.ceil(np.sqrt(x)) if x >= 0 else np.nan
        want = int(r) if np.isfinite(r) and abs(r) < 2 ** 31 else -2 ** 31
        if got != want:  # only just above a perfect square k^2, where the exact answer is k + 1 and sqrt() rounds to k
            k = round(float(np.sqrt(x)))
            assert got == want + 1 == k + 1 and x > k * k and np.sqrt(x) == k, (x, got, want)
            n_diff += 1
    assert n_diff <= 4000  # the nextafter(k^2, inf) probes


@needs_ref
def test_readme_ddm_example_model_bitwise_vs_reference_object_code():
    """The DDM of the reference's second README example (README.md:247-300: free a, sz, t0, v, z; st0 = sv = 0,
    precision = 3; the matching response is the upper boundary) with subjects drawn around its population mean: every
    trial density of the restatement equals likelihood_class::ddm_likelihood of src/de.o bit for bit (start-point
    variability on: the 4..10-abscissa midpoint rule is on the path)."""
    from ggdmc_b200 import workloads as W
    ct, p_vector, pop_mean, pop_scale = W.ddm_readme_model()
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar, type=ob.MODEL_DDM)
    rng = np.random.default_rng(9032)
    n = 256
    cell = np.sort(rng.integers(0, 4, n)).astype(np.uint16)
    d = ob.OData(0.15 + rng.gamma(2.0, 0.12, n), cell)
    n_pos = 0
    for _ in range(40):
        th = pop_mean + pop_scale * rng.standard_normal(5)
        ref = ob.ref2_ddm_density(om, d, th)
        mine = _oracle_density(om, d, th)
        assert np.all(_same(ref, mine)), th
        n_pos += int(np.sum(ref > 1e-3))
    assert n_pos > 20 * n


def test_ddm_golden_vectors():
    """tests/golden/ddm_ref.npz -- trial densities written by the reference's own object code
    (tests/golden/make_ddm_golden.py; three (precision, s) settings, all variability combinations, the edge cases) --
    reproduced by the oracle bit for bit.  Needs neither /root/reference nor oracle/_ref."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddm_ref.npz"))
    n = 0
    for k, (precision, s) in enumerate(g["settings"]):
        ct, om = ddm_model(float(precision), float(s))
        d = ob.OData(g[f"rt{k}"], g[f"cell{k}"])
        assert np.array_equal(d.rt, g[f"rt{k}"])  # stored grouped by cell already
        for th, want in zip(g[f"theta{k}"], g[f"dens{k}"]):
            assert np.all(_same(_oracle_density(om, d, th), want)), (k, th)
            n += int(np.sum(want > 1e-6))
    assert n > 20000
