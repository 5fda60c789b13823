"""Check 3 (-m gpu): posteriors from fixed-seed runs agree within Monte-Carlo error.

The REFERENCE schedule of the engine reproduces the oracle's (= the reference's) chain order draw for
draw (tests/test_gpu_sampler.py), so long REFERENCE-schedule runs stand in for the reference sampler;
the default PARALLEL schedule must give the same posterior.  Eight independent replicates per arm give
a replicate-level Monte-Carlo standard error for every summary; an independent CPU oracle run is the
third arm.  Criteria: |difference| <= 2 MCSE for the bulk of the summaries (with 60+ summaries a few
2-sigma excursions are expected by chance; none may exceed 4), R-hat < 1.05, truth recovered."""
import numpy as np
import pytest

from ggdmc_b200 import _lib as B
from ggdmc_b200 import engine as E
from oracle import binding as ob
from helpers import load_fixture, sane_starts

pytestmark = pytest.mark.gpu


def summaries(x):
    """x [n, C, D] -> dict of per-parameter summaries pooled over samples and chains."""
    flat = x.reshape(-1, x.shape[-1])
    return np.stack([flat.mean(0), np.quantile(flat, 0.05, axis=0), np.quantile(flat, 0.5, axis=0), np.quantile(flat, 0.975, axis=0)])


def rhat(x):
    """Gelman-Rubin PSRF per parameter over chains, each chain split in two halves."""
    n = x.shape[0] // 2
    y = np.concatenate([x[:n], x[n:2 * n]], axis=1)  # [n, 2C, D]
    cm = y.mean(0)
    W = y.var(0, ddof=1).mean(0)
    Bn = cm.var(0, ddof=1)
    return np.sqrt((n - 1) / n + Bn / W)


def fit_subject(fx, tr, prior, starts, schedule, seeds, burn_nmc, nmc, thin):
    D, C = fx.ct.npar, starts[0][0].shape[0]
    st = E.PopState(np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts]), np.stack([s[2] for s in starts]))
    burn = E.run_subject(fx.ct, tr, prior, E.Tuning(nmc=burn_nmc, nchain=C, thin=thin, nparameter=D, sub_migration_prob=0.06,
                                                     schedule=schedule, seeds=seeds), st)
    st2 = E.PopState(burn.theta[:, -1], burn.lp[:, -1], burn.ll[:, -1])
    return E.run_subject(fx.ct, tr, prior, E.Tuning(nmc=nmc, nchain=C, thin=thin, nparameter=D, sub_migration_prob=0.0,
                                                    schedule=schedule, seeds=[s + 1000 for s in seeds]), st2)


def test_single_subject_posterior_agrees_across_schedules_and_with_oracle():
    """BASELINE config 1: README B x v model, 13 parameters, 768 trials, 39 chains, thin 8."""
    fx = load_fixture(6)
    tr, od = fx.trials("sub"), fx.odata("sub")
    prior, oprior = fx.prior("sub_prior"), fx.oprior("sub_prior")
    D, C, R, thin, nmc = fx.ct.npar, 3 * fx.ct.npar, 8, 8, 401
    rng = np.random.default_rng(2026)
    starts = []
    for _ in range(R):
        th = sane_starts(fx, C, rng, jitter=0.1)
        starts.append((th, np.array([ob.sumlogprior(oprior, t) for t in th]), np.array([ob.sumloglike(fx.om, od, t) for t in th])))
    arms = {}
    for name, sched, seed0 in (("reference", B.SCHEDULE_REFERENCE, 100), ("parallel", B.SCHEDULE_PARALLEL, 200)):
        out = fit_subject(fx, tr, prior, starts, sched, [seed0 + r for r in range(R)], 301, nmc, thin)
        arms[name] = out.theta[:, 1:]  # [R, n, C, D]
        for r in range(R):
            assert rhat(arms[name][r]).max() < 1.05, (name, r, rhat(arms[name][r]))
    stat = {k: np.stack([summaries(v[r]) for r in range(R)]) for k, v in arms.items()}  # [R, 4, D]
    mean = {k: v.mean(0) for k, v in stat.items()}
    mcse = {k: v.std(0, ddof=1) / np.sqrt(R) for k, v in stat.items()}
    z = np.abs(mean["reference"] - mean["parallel"]) / np.sqrt(mcse["reference"] ** 2 + mcse["parallel"] ** 2)
    assert np.mean(z <= 2.0) >= 0.85 and z.max() < 4.0, z
    # posterior spread identical too (ratio of pooled sds)
    sd = {k: v.reshape(-1, D).std(0) for k, v in arms.items()}
    assert np.all(np.abs(sd["reference"] / sd["parallel"] - 1.0) < 0.08)
    # truth recovered (the generating values lie inside the central 99.9 % region of each marginal)
    truth = fx.g["p_vector"]
    flat = arms["parallel"].reshape(-1, D)
    lo, hi = np.quantile(flat, 0.0005, axis=0), np.quantile(flat, 0.9995, axis=0)
    assert np.all((truth > lo) & (truth < hi)), (truth, lo, hi)
    # third arm: the CPU oracle (reference chain order), one replicate, shorter
    th0, lp0, ll0 = starts[0]
    pop = ob.OPop(th0, lp0, ll0, 201, thin)
    ob.run_subject(ob.make_de(D, C, sub_migration_prob=0.06), pop, oprior, fx.om, od, ob.make_rng(seed=5), 0, 200 * thin)
    pop2 = ob.OPop(pop.theta, pop.lp, pop.ll, 201, thin)
    ob.run_subject(ob.make_de(D, C, sub_migration_prob=0.0), pop2, oprior, fx.om, od, ob.make_rng(seed=6), 0, 200 * thin)
    so = summaries(pop2.out_theta[1:])
    # a single replicate of half the length: its standard error is ~ sqrt(R * 2) x the arm's MCSE
    zo = np.abs(so - mean["parallel"]) / (mcse["parallel"] * np.sqrt(2.0 * R) * np.sqrt(1 + 1 / (2.0 * R)))
    assert np.mean(zo <= 2.0) >= 0.85 and zo.max() < 4.5, zo


def test_hierarchical_posterior_agrees_across_schedules():
    """Hierarchical fit (8-parameter model, 4 subjects x 256 trials, 48 chains): phi and subject-level
    posteriors of the two schedules agree within Monte-Carlo error."""
    from test_gpu_sampler import hier_setup
    fx = load_fixture(2)
    S, D, R, thin, nmc = fx.n_pop, fx.ct.npar, 6, 4, 301
    C = 6 * D
    trials = [fx.trials(f"pop{s}") for s in range(S)]
    pp, hp = fx.prior("p_prior"), fx.prior("h_prior")
    rng = np.random.default_rng(7)
    setups = [hier_setup(fx, S, C, rng) for _ in range(R)]
    phi_st = E.PopState(np.stack([s[0][0] for s in setups]), np.stack([s[0][1] for s in setups]), np.stack([s[0][2] for s in setups]))
    sub_st = [E.PopState(np.stack([s[1][i][0] for s in setups]), np.stack([s[1][i][1] for s in setups]),
                         np.stack([s[1][i][2] for s in setups])) for i in range(S)]
    res = {}
    for name, sched, seed0 in (("reference", B.SCHEDULE_REFERENCE, 10), ("parallel", B.SCHEDULE_PARALLEL, 50)):
        kw = dict(nchain=C, thin=thin, nparameter=2 * D, schedule=sched)
        phi_b, sub_b = E.run_hier(fx.ct, trials, pp, hp, E.Tuning(nmc=201, pop_migration_prob=0.05, sub_migration_prob=0.05,
                                                                  seeds=[seed0 + r for r in range(R)], **kw), phi_st, sub_st)
        phi2 = E.PopState(phi_b.theta[:, -1], phi_b.lp[:, -1], phi_b.ll[:, -1])
        sub2 = [E.PopState(o.theta[:, -1], o.lp[:, -1], o.ll[:, -1]) for o in sub_b]
        phi_o, sub_o = E.run_hier(fx.ct, trials, pp, hp, E.Tuning(nmc=nmc, seeds=[seed0 + 500 + r for r in range(R)], **kw), phi2, sub2)
        res[name] = (phi_o.theta[:, 1:], np.stack([o.theta[:, 1:] for o in sub_o], axis=1))  # [R,n,C,2D], [R,S,n,C,D]
    for which, idx in (("phi", 0), ("subject0", 1)):
        a = res["reference"][idx] if idx == 0 else res["reference"][idx][:, 0]
        b = res["parallel"][idx] if idx == 0 else res["parallel"][idx][:, 0]
        sa = np.stack([summaries(a[r]) for r in range(R)])
        sb = np.stack([summaries(b[r]) for r in range(R)])
        z = np.abs(sa.mean(0) - sb.mean(0)) / np.sqrt(sa.var(0, ddof=1) / R + sb.var(0, ddof=1) / R)
        assert np.mean(z <= 2.0) >= 0.8 and z.max() < 4.5, (which, z)
        assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
