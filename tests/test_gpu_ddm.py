"""GPU tests of the DDM ("fastdm") path (run with -m gpu on a B200): k_trial_logdens_ddm / k_like_ddm through the C ABI
against the CPU oracle -- which is bit-identical to the reference's own object code (tests/test_ddm_cpu.py) -- and the
sampler on a DDM likelihood draw for draw against the oracle."""
import numpy as np
import pytest

from ggdmc_b200 import _lib as B
from ggdmc_b200 import engine as E
from ggdmc_b200.model import PriorTable, Trials
from oracle import binding as ob
from helpers import ddm_model, ddm_prior, ddm_simulate, ddm_theta
from test_ddm_cpu import _edge_thetas, _grid_data
from test_gpu_sampler import SCHEDULES, compare

pytestmark = pytest.mark.gpu

LOG_DBL_MIN = np.log(2.2250738585072014e-308)


def _check_logdens(got, ref, what):
    """1e-10 relative on the log density (the north-star bound for the LBA, kept for the DDM) wherever the density is
    well conditioned; the series are truncated at an absolute error of 1e-6 and sum terms of both signs, so a density
    of size d carries an absolute rounding error of ~1e-15 -> 1e-14 / d on its log.  Below 1e-12 only smallness is
    checked (there the value is truncation noise in the reference too, possibly negative -> log(DBL_MIN))."""
    assert np.array_equal(np.isnan(got), np.isnan(ref)), what
    ok = ~np.isnan(ref)
    got, ref = got[ok], ref[ok]
    dens = np.exp(ref)
    big = dens >= 1e-12
    tol = 1e-10 * np.maximum(np.abs(ref[big]), 1.0) + 1e-14 / dens[big]
    err = np.abs(got[big] - ref[big])
    assert np.all(err <= tol), (what, np.where(err > tol)[0][:5], err[err > tol][:5], ref[big][err > tol][:5])
    assert np.all(got[~big] <= np.log(1e-11)), what
    strict = dens >= 1e-4
    assert np.all(np.abs(got[strict] - ref[strict]) <= 1e-10 * np.maximum(np.abs(ref[strict]), 1.0)), what
    return int(strict.sum())


@pytest.mark.parametrize("precision,s", [(3.0, 1.0), (2.5, 1.0), (3.0, 0.1)])
def test_ddm_trial_logdens_vs_oracle(precision, s):
    """Check 1 for the DDM: per-trial log densities on fixed theta arrays, all four variability combinations, both
    boundaries, short / typical / very long response times."""
    ct, om = ddm_model(precision, s)
    rng = np.random.default_rng(int(10 * precision + 100 * s))
    od = _grid_data(rng, 360)
    tr = Trials(od.rt.copy(), od.cell.copy())
    thetas = []
    for it in range(48):
        th = ddm_theta(rng, it % 4)
        if s != 1.0:
            for i in (0, 2, 5, 6):
                th[i] *= s
            th[7] = th[0] / s * 0.5
            th[3] = min(th[3], 0.2 * th[0] / s)
        thetas.append(th)
    thetas = np.stack(thetas)
    got = E.trial_logdens(ct, tr, thetas)
    n_strict = 0
    for i, th in enumerate(thetas):
        n_strict += _check_logdens(got[i], ob.trial_logdens(om, od, th), (precision, s, i))
    assert n_strict > 3000


def test_ddm_trial_logdens_edge_cases():
    """validate_parameters' rules (invalid cell -> log 1e-10 for every trial), NaN parameters, rt below t0
    (density 0 -> log DBL_MIN), extreme drifts and boundaries."""
    ct, om = ddm_model()
    rng = np.random.default_rng(3)
    od = _grid_data(rng)
    tr = Trials(od.rt.copy(), od.cell.copy())
    thetas = np.stack(_edge_thetas(rng))
    got = E.trial_logdens(ct, tr, thetas)
    n_floor = n_min = 0
    for i, th in enumerate(thetas):
        ref = ob.trial_logdens(om, od, th)
        _check_logdens(got[i], ref, ("edge", i))
        n_floor += int(np.all(ref == np.log(1e-10)) and np.all(got[i] == ref))
        n_min += int(np.sum((ref == LOG_DBL_MIN) & (got[i] == LOG_DBL_MIN)))
    assert n_floor >= 8 and n_min > 50


def _subjects(rng, S, n_per_stim, kind):
    ct, om = ddm_model()
    truths = [ddm_theta(rng, kind) for _ in range(S)]
    data = [ddm_simulate(t, n_per_stim, rng) for t in truths]
    return ct, om, truths, [Trials(rt, cell) for rt, cell in data], [ob.OData(rt, cell) for rt, cell in data]


@pytest.mark.parametrize("kind", [0, 1, 3])
def test_ddm_sumloglike_vs_oracle(kind):
    """likelihood_class::sumloglike's DDM branch (@hdr/likelihood.h:295-305) for several subjects x parameter vectors
    near the generating values: 1e-10 relative; ragged trial counts (subjects keep only finished simulations)."""
    rng = np.random.default_rng(50 + kind)
    S, K = 3, 20
    ct, om, truths, trials, odata = _subjects(rng, S, 90 if kind < 3 else 40, kind)
    theta = np.stack([t[None, :] * (1.0 + 0.04 * rng.standard_normal((K, len(t)))) for t in truths])
    got = E.sumloglike(ct, trials, theta)
    for s in range(S):
        for c in range(K):
            ref = ob.sumloglike(om, odata[s], theta[s, c])
            assert np.isfinite(ref)
            assert abs(got[s, c] - ref) <= 1e-10 * abs(ref), (s, c, got[s, c], ref)
    # the R-side initialisation rule (R/phi.R:3-13): densities <= 0 (e.g. rt below a jittered t0) count as
    # .Machine$double.eps instead of DBL_MIN; where there are none the two sums are the same number
    got_init = E.sumloglike(ct, trials, theta, init_rule=True)
    n_same = 0
    for s in range(S):
        for c in range(K):
            ref = ob.sumloglike_rinit(om, odata[s], theta[s, c])
            assert abs(got_init[s, c] - ref) <= 1e-10 * abs(ref), (s, c, got_init[s, c], ref)
            n_same += int(got_init[s, c] == got[s, c])
    assert n_same >= S * K // 2


def _start(om, od, oprior, truth, nchain, rng, p0=None, p1=None):
    th = truth[None, :] * (1.0 + 0.03 * rng.standard_normal((nchain, len(truth))))
    if p0 is None:
        lp = np.array([ob.sumlogprior(oprior, t) for t in th])
    else:
        lp = np.array([ob.sumlogprior(oprior, th[c], p0[c], p1[c]) for c in range(nchain)])
    ll = np.array([ob.sumloglike(om, od, t) for t in th])
    return th, lp, ll


@pytest.mark.parametrize("schedule,jacobi", SCHEDULES)
@pytest.mark.parametrize("pblocked", [False, True])
def test_ddm_run_subject_trajectory(schedule, jacobi, pblocked):
    """run_subject on a DDM likelihood: every stored theta identical to the oracle's replay of the same addressed draws
    (crossover + migration sweeps, blocked and unblocked), log prior / log likelihood to 1e-9."""
    rng = np.random.default_rng(9)
    ct, om, truths, trials, odata = _subjects(rng, 1, 80, 1)
    prior, oprior = ddm_prior()
    D = ct.npar
    nchain, nmc, thin = 3 * D, 4, 2
    seeds = [9032, 5]
    starts = [_start(om, odata[0], oprior, truths[0], nchain, rng) for _ in seeds]
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, sub_migration_prob=0.35, is_pblocked=pblocked, schedule=schedule,
                   seeds=seeds)
    st = E.PopState(np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts]), np.stack([s[2] for s in starts]))
    out = E.run_subject(ct, trials[0], prior, tun, st)
    for r, seed in enumerate(seeds):
        pop = ob.OPop(*starts[r], nmc, thin)
        de = ob.make_de(D, nchain, sub_migration_prob=0.35, is_pblocked=pblocked, jacobi=jacobi)
        ob.run_subject(de, pop, oprior, om, odata[0], ob.make_rng(seed=seed), 0, (nmc - 1) * thin)
        compare(out, r, pop, f"ddm run_subject seed {seed}")
    assert not np.array_equal(out.theta[0, 0], out.theta[0, -1])


def _hier_priors(center):
    """p_prior: truncated normal per parameter (location / scale come from phi); h_prior: uniform over both."""
    D = len(center)
    lower = np.array([0.0, 0.0, 0.0, 0.0, 0.0, -20.0, -20.0, 0.0])
    upper = np.full(D, np.inf)
    dist, logp = np.full(D, 1, np.int32), np.ones(D, np.uint8)
    p0, p1 = center.copy(), 0.3 * np.abs(center) + 0.1
    names = [f"p{i}" for i in range(D)]
    pp = PriorTable(D, p0, p1, lower, upper, dist, logp, names)
    opp = ob.OPrior(p0, p1, lower, upper, dist, logp)
    hlo = np.concatenate([center - 3.0, np.full(D, 0.01)])
    hhi = np.concatenate([center + 3.0, np.full(D, 3.0)])
    hd, hl = np.full(2 * D, 6, np.int32), np.ones(2 * D, np.uint8)
    hp = PriorTable(2 * D, hlo, hhi, np.zeros(2 * D), np.zeros(2 * D), hd, hl, names + names)
    ohp = ob.OPrior(hlo, hhi, np.zeros(2 * D), np.zeros(2 * D), hd, hl)
    return pp, opp, hp, ohp


@pytest.mark.parametrize("schedule,jacobi", SCHEDULES[:2])
def test_ddm_hierarchical_trajectory(schedule, jacobi):
    """run (run_hchains) on DDM subjects: phi step, subject steps with phi-driven truncated-normal priors, migration at
    both levels -- draw for draw against the oracle."""
    rng = np.random.default_rng(17)
    S = 4
    ct, om = ddm_model()
    center = ddm_theta(rng, 1)
    truths = [center * (1.0 + 0.05 * rng.standard_normal(len(center))) for _ in range(S)]
    data = [ddm_simulate(t, 50, rng) for t in truths]
    trials, odata = [Trials(rt, cell) for rt, cell in data], [ob.OData(rt, cell) for rt, cell in data]
    pp, opp, hp, ohp = _hier_priors(center)
    D = ct.npar
    nchain, nmc, thin = 2 * 2 * D, 3, 2
    phi0 = np.concatenate([center, 0.3 * np.abs(center) + 0.1])[None, :] * (1.0 + 0.05 * rng.standard_normal((nchain, 2 * D)))
    subj = [_start(om, odata[s], opp, truths[s], nchain, rng, phi0[:, :D], phi0[:, D:]) for s in range(S)]
    lp0 = np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(nchain)])
    ll0 = np.array([sum(ob.sumlogprior(opp, subj[s][0][c], phi0[c, :D], phi0[c, D:]) for s in range(S)) for c in range(nchain)])
    seed = 2718
    kw = dict(pop_migration_prob=0.3, sub_migration_prob=0.3)
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=2 * D, schedule=schedule, seeds=[seed], **kw)
    phi_out, subj_out = E.run_hier(ct, trials, pp, hp, tun, E.PopState(phi0, lp0, ll0), [E.PopState(*s) for s in subj])
    phi = ob.OPop(phi0, lp0, ll0, nmc, thin)
    pops = [ob.OPop(*s, nmc, thin) for s in subj]
    ob.run_hier(ob.make_de(2 * D, nchain, jacobi=jacobi, **kw), phi, pops, opp, ohp, om, odata, ob.make_rng(seed=seed), (nmc - 1) * thin)
    compare(phi_out, 0, phi, "ddm phi")
    for s in range(S):
        compare(subj_out[s], 0, pops[s], f"ddm subject {s}")
    assert not np.array_equal(phi_out.theta[0, 0], phi_out.theta[0, -1])


def test_undefined_model_type_is_an_error():
    """Anything but "lba" / "fastdm" is the reference's "Undefined model type" (@hdr/likelihood.h:312)."""
    ct, om = ddm_model()
    rng = np.random.default_rng(1)
    od = _grid_data(rng, 24)
    m = E._model(ct)
    m.c.type = 7
    import ctypes as C
    t = E._trials([Trials(od.rt.copy(), od.cell.copy())])
    th = B.f64(ddm_theta(rng, 0)[None, :])
    out = np.zeros((1, len(od.rt)))
    err = C.create_string_buffer(256)
    rc = B.lib().ggdmc_b200_trial_logdens(C.byref(m.c), C.byref(t.c), B.ptr(th), 1, B.ptr(out), err)
    assert rc == B.ERR_ARG and b"Undefined model type" in err.value


def test_ddm_full_size_properties():
    """Size-independent properties at a size the oracle does not finish in seconds (64 subjects x 2048 trials x 24
    parameter vectors, drift variability on): the sum over trials is additive over a split of a subject's trials and
    invariant under a permutation of them (the engine regroups by cell and response time on upload), a subject's sums do
    not depend on its neighbours, and three spot checks against the oracle."""
    from ggdmc_b200 import workloads as W
    rng = np.random.default_rng(77)
    ct, truth, prior = W.ddm_model(fixed=("st0", "sz"))
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar, type=ob.MODEL_DDM)
    pool = W.ddm_simulate(truth, 8000, rng, pnames=ct.pnames)
    S, K, n = 64, 24, 2048
    subj = []
    for _ in range(S):
        idx = rng.choice(len(pool.rt), n, replace=False)
        subj.append(Trials(pool.rt[idx].copy(), pool.cell[idx].copy()))  # deliberately NOT grouped by cell
    theta = truth[None, None, :] * (1.0 + 0.03 * rng.uniform(-1, 1, size=(S, K, ct.npar)))
    full = E.sumloglike(ct, subj, theta)
    assert np.all(np.isfinite(full))
    halves_a = [Trials(t.rt[: n // 3].copy(), t.cell[: n // 3].copy()) for t in subj]   # ragged split
    halves_b = [Trials(t.rt[n // 3:].copy(), t.cell[n // 3:].copy()) for t in subj]
    split = E.sumloglike(ct, halves_a, theta) + E.sumloglike(ct, halves_b, theta)
    assert np.all(np.abs(split - full) <= 1e-11 * np.abs(full))
    perm = [rng.permutation(n) for _ in range(S)]
    shuffled = E.sumloglike(ct, [Trials(t.rt[p].copy(), t.cell[p].copy()) for t, p in zip(subj, perm)], theta)
    assert np.array_equal(shuffled, full)  # same device order after the upload's (cell, rt) ordering -> same bits
    alone = E.sumloglike(ct, subj[5:6], theta[5:6])
    assert np.all(np.abs(alone[0] - full[5]) <= 1e-13 * np.abs(full[5]))  # a lone subject is split into more trial chunks
    for s, c in ((0, 0), (17, 9), (63, 23)):
        ref = ob.sumloglike(om, ob.OData(subj[s].rt, subj[s].cell), theta[s, c])
        assert abs(full[s, c] - ref) <= 1e-10 * abs(ref)


def test_ddm_through_reference_interface():
    """model@type "fastdm" through the Python mirror of the reference interface: initialise_theta scores its candidates
    with the DDM likelihood under the R-side rule, run_subject returns a posterior with the reference's slots, feeding it
    back continues from its last slice, and the same call through the engine layer gives the same bits."""
    from ggdmc_b200 import api, init
    from helpers import DDM_PNAMES, ddm_objects
    rng = np.random.default_rng(23)
    ct, om = ddm_model()
    truth = ddm_theta(rng, 1)
    rt, cell = ddm_simulate(truth, 80, rng)
    model, dmi = ddm_objects(rt, cell)
    pt, op = ddm_prior()
    D, nchain, nmc, thin = ct.npar, 3 * ct.npar, 5, 2
    prior = api.Prior(nparameter=D, pnames=list(DDM_PNAMES), p_prior=api.prior_list(pt))
    ti = api.ThetaInput(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, pnames=list(DDM_PNAMES))
    de = api.DEInput(sub_migration_prob=0.06, nparameter=D, nchain=nchain)
    cfg = api.Config(prior=prior, theta_input=ti, de_input=de, seed=11)
    st = init.initialise_theta(ti, prior, dmi, seed=3)
    od = ob.OData(rt, cell)
    assert np.all(np.isfinite(st.theta[:, :, 0])) and np.all(np.isfinite(st.log_likelihoods[:, 0]))
    for c in range(nchain):  # trial-by-trial form of the R rule: a density <= 0 counts as .Machine$double.eps
        ld = ob.trial_logdens(om, od, st.theta[:, c, 0])
        ref = np.where(ld <= LOG_DBL_MIN, np.log(np.finfo(float).eps), ld).sum()
        assert abs(st.log_likelihoods[c, 0] - ref) <= 1e-9 * abs(ref)
    fit = api.run_subject(cfg, dmi, st)
    assert fit.theta.shape == (D, nchain, nmc) and fit.pnames == list(DDM_PNAMES) and fit.nmc == nmc and fit.thin == thin
    assert np.array_equal(fit.theta[:, :, 0], st.theta[:, :, 0]) and not np.array_equal(fit.theta[:, :, 0], fit.theta[:, :, -1])
    again = api.run_subject(cfg, dmi, fit)
    assert np.array_equal(again.theta[:, :, 0], fit.theta[:, :, -1])
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, sub_migration_prob=0.06, seeds=[11])
    start = E.PopState(st.theta[:, :, 0].T[None].copy(), st.summed_log_prior[:, 0][None].copy(), st.log_likelihoods[:, 0][None].copy())
    out = E.run_subject(ct, Trials(od.rt.copy(), od.cell.copy()), pt, tun, start)
    assert np.array_equal(np.transpose(out.theta[0], (2, 1, 0)), fit.theta)


def test_ddm_single_subject_posterior_recovers_generating_values():
    """Check 3 for the DDM, in the form that does not repeat what the LBA tests establish (schedule equivalence is a
    property of the sampler, tests/test_gpu_posterior.py; chain-for-chain identity with the oracle is the trajectory test
    above): a 6-parameter DDM fit (600 simulated trials, 18 chains, 4 replicates, default schedule) converges
    (R-hat < 1.05), the replicates agree with each other, and the generating values lie inside every marginal."""
    from ggdmc_b200 import workloads as W
    from test_gpu_posterior import rhat, summaries
    ct, truth, prior = W.ddm_model(fixed=("st0", "sz"))
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar, type=ob.MODEL_DDM)
    oprior = ob.OPrior(prior.p0, prior.p1, prior.lower, prior.upper, prior.dist, prior.log_p)
    rng = np.random.default_rng(404)
    tr = W.ddm_simulate(truth, 300, rng, pnames=ct.pnames)
    od = ob.OData(tr.rt, tr.cell)
    D, C, R, thin = ct.npar, 3 * ct.npar, 4, 4
    th = truth[None, None, :] * (1.0 + 0.1 * rng.standard_normal((R, C, D)))
    lp = np.array([[ob.sumlogprior(oprior, t) for t in th[r]] for r in range(R)])
    ll = E.sumloglike(ct, [tr] * R, th)
    st = E.PopState(th, lp, ll)
    burn = E.run_subject(ct, tr, prior, E.Tuning(nmc=251, nchain=C, thin=thin, nparameter=D, sub_migration_prob=0.06,
                                                  seeds=[200 + r for r in range(R)]), st)
    fit = E.run_subject(ct, tr, prior, E.Tuning(nmc=601, nchain=C, thin=thin, nparameter=D, seeds=[1200 + r for r in range(R)]),
                        E.PopState(burn.theta[:, -1], burn.lp[:, -1], burn.ll[:, -1]))
    x = fit.theta[:, 1:]  # [R, n, C, D]
    assert np.all(np.isfinite(x)) and np.all(np.isfinite(fit.ll))
    assert max(rhat(x[r]).max() for r in range(R)) < 1.05
    s = np.stack([summaries(x[r]) for r in range(R)])  # [R, 4, D]
    sd = x.reshape(-1, D).std(0)
    assert np.all(np.abs(s - s.mean(0)) <= 0.35 * sd), (s, sd)  # replicate-to-replicate spread of mean / quantiles
    flat = x.reshape(-1, D)
    lo, hi = np.quantile(flat, 0.0005, axis=0), np.quantile(flat, 0.9995, axis=0)
    assert np.all((truth > lo) & (truth < hi)), (truth, lo, hi)


def test_ddm_trial_logdens_vs_reference_golden_vectors():
    """The kernel against known answers written by the reference's own object code (tests/golden/ddm_ref.npz, made by
    tests/golden/make_ddm_golden.py from likelihood_class::ddm_likelihood of src/de.o): no oracle in between."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ddm_ref.npz"))
    n_strict = 0
    for k, (precision, s) in enumerate(g["settings"]):
        ct, om = ddm_model(float(precision), float(s))
        thetas, dens = g[f"theta{k}"][:48], g[f"dens{k}"][:48]
        got = E.trial_logdens(ct, Trials(g[f"rt{k}"].copy(), g[f"cell{k}"].copy()), thetas)
        for i in range(len(thetas)):
            ref = np.log(np.where(dens[i] < 2.2250738585072014e-308, 2.2250738585072014e-308, dens[i]))  # @hdr/likelihood.h:303
            n_strict += _check_logdens(got[i], ref, ("golden", k, i))
    assert n_strict > 9000
