"""Experiment (GPU; output of round 2 in profiles/r02_ddm_hier.txt):
the reference README's second example as a hierarchical DDM recovery study -- S subjects (default 32) x 256 trials,
free a, sz, t0, v, z (start-point variability ON: the midpoint-rule path), truncated-normal population
distribution -- through the resident engine: DE-MCMC iterations/s, trial-likelihoods/s, R-hat and recovered population
means.  Usage: python tools/exp_ddm_hier.py [n_subject] [n_iter]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ggdmc_b200 import engine as E, workloads as W
from ggdmc_b200.model import PriorTable


def build(S, rng):
    ct, p_vector, pop_mean, pop_scale = W.ddm_readme_model()
    D = ct.npar
    lower = np.array([0.0, 0.0, 0.0, -10.0, 0.0])  # README.md:283
    pp = PriorTable(D, pop_mean.copy(), pop_scale.copy(), lower, np.full(D, np.inf), np.full(D, 1, np.int32), np.ones(D, np.uint8), ct.pnames)
    hlo = np.concatenate([np.maximum(pop_mean - 2.0, lower), np.full(D, 1e-3)])
    hhi = np.concatenate([pop_mean + 2.0, np.full(D, 2.0)])
    hp = PriorTable(2 * D, hlo, hhi, np.zeros(2 * D), np.zeros(2 * D), np.full(2 * D, 6, np.int32), np.ones(2 * D, np.uint8),
                    [f"loc_{n}" for n in ct.pnames] + [f"sca_{n}" for n in ct.pnames])
    truths = np.maximum(pop_mean + pop_scale * rng.standard_normal((S, D)), lower + 1e-3)
    trials = [simulate_accuracy_coded(dict(zip(ct.pnames, t)), 128, rng) for t in truths]
    return ct, pp, hp, pop_mean, pop_scale, truths, trials


def simulate_accuracy_coded(p, n_per_stim, rng):
    """W.ddm_simulate draws with one drift per stimulus and r2 as the upper boundary; the README model has ONE drift
    towards the matching response, which is the upper boundary of its cell: same paths, cells s1.r1 / s1.r2 swapped."""
    from ggdmc_b200.model import Trials
    full = dict(a=p["a"], st0=0.0, sv=0.0, sz=p["sz"], t0=p["t0"], z=p["z"])
    full["v.s1"] = full["v.s2"] = p["v"]
    tr = W.ddm_simulate(np.array([full[k] for k in W.DDM_PNAMES]), n_per_stim, rng)
    cell = np.where(tr.cell < 2, 1 - tr.cell, tr.cell).astype(np.uint16)
    order = np.argsort(cell, kind="stable")
    return Trials(tr.rt[order], cell[order])


def starts(ct, pp, hp, pop_mean, pop_scale, truths, trials, C, rng):
    S, D = truths.shape
    phi0 = np.concatenate([pop_mean, pop_scale])[None, None, :] * (1.0 + 0.05 * rng.standard_normal((1, C, 2 * D)))
    subj0 = truths[:, None, None, :] * (1.0 + 0.02 * rng.standard_normal((S, 1, C, D)))
    return phi0, subj0


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    rng = np.random.default_rng(20260106)
    ct, pp, hp, pop_mean, pop_scale, truths, trials = build(S, rng)
    D, C = ct.npar, 3 * 2 * ct.npar
    phi0, subj0 = starts(ct, pp, hp, pop_mean, pop_scale, truths, trials, C, rng)
    ll = E.sumloglike(ct, trials, subj0.reshape(S, C, D)).reshape(S, 1, C)
    ph = np.broadcast_to(phi0.reshape(1, C, 2 * D), (S, C, 2 * D)).reshape(S * C, 2 * D)
    lp = E.sumlogprior(pp, subj0.reshape(S * C, D), np.ascontiguousarray(ph[:, :D]), np.ascontiguousarray(ph[:, D:])).reshape(S, 1, C)
    phi_lp = E.sumlogprior(hp, phi0.reshape(C, 2 * D)).reshape(1, C)
    thin, nmc = 4, n_iter // 4 + 1
    tun = E.Tuning(nmc=nmc, nchain=C, thin=thin, nparameter=2 * D, pop_migration_prob=0.05, sub_migration_prob=0.05, seeds=[9032])
    eng = E.Engine(ct, trials, pp, hp, tun, E.PopState(phi0, phi_lp, lp.sum(axis=0)), [E.PopState(subj0[s], lp[s], ll[s]) for s in range(S)])
    eng.iterate(20)
    eng.counters()
    ms = eng.iterate(n_iter - 20)
    nlik, _, _ = eng.counters()
    st = eng.state()
    eng.close()
    print(f"{S} subjects x {sum(len(t.rt) for t in trials) / S:.0f} trials, {C} chains: {(n_iter - 20) / (ms * 1e-3):.0f} DE-MCMC iterations/s, "
          f"{nlik / (ms * 1e-3):.3e} trial-likelihoods/s")
    print("population means:", dict(zip(ct.pnames, pop_mean)), "\nphi location state mean:", st["phi_theta"][0, :, :D].mean(0).round(3))
