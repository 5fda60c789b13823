"""Experiment (GPU): do the PARALLEL and REFERENCE schedules give the same posterior?"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ggdmc_b200 import _lib as B, engine as E
from oracle import binding as ob
from helpers import load_fixture, sane_starts
from test_gpu_sampler import hier_setup
from test_gpu_posterior import summaries  # noqa

def rhat(x):
    n = x.shape[0]; cm = x.mean(0); W = x.var(0, ddof=1).mean(0); Bn = cm.var(0, ddof=1)
    return np.sqrt((n - 1) / n + Bn / W)

def single(burn_nmc, nmc, thin, R=8):
    fx = load_fixture(6); tr, od = fx.trials("sub"), fx.odata("sub")
    prior, oprior = fx.prior("sub_prior"), fx.oprior("sub_prior")
    D, C = fx.ct.npar, 3 * fx.ct.npar
    rng = np.random.default_rng(2026); starts = []
    for _ in range(R):
        th = sane_starts(fx, C, rng, jitter=0.1)
        starts.append((th, np.array([ob.sumlogprior(oprior, t) for t in th]), np.array([ob.sumloglike(fx.om, od, t) for t in th])))
    st = E.PopState(np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts]), np.stack([s[2] for s in starts]))
    arms = {}
    for name, sched, seed0 in (("reference", 0, 100), ("parallel", 1, 200), ("simultaneous", 2, 300)):
        t0 = time.time()
        b = E.run_subject(fx.ct, tr, prior, E.Tuning(nmc=burn_nmc, nchain=C, thin=thin, nparameter=D, sub_migration_prob=0.06, schedule=sched, seeds=[seed0 + r for r in range(R)]), st)
        o = E.run_subject(fx.ct, tr, prior, E.Tuning(nmc=nmc, nchain=C, thin=thin, nparameter=D, schedule=sched, seeds=[seed0 + 50 + r for r in range(R)]),
                          E.PopState(b.theta[:, -1], b.lp[:, -1], b.ll[:, -1]))
        arms[name] = o.theta[:, 1:]
        print(name, "time %.1fs" % (time.time() - t0), "rhat max", max(rhat(arms[name][r]).max() for r in range(R)).round(4), flush=True)
    stat = {k: np.stack([summaries(v[r]) for r in range(R)]) for k, v in arms.items()}
    for a, b in (("reference", "parallel"), ("reference", "simultaneous")):
        z = np.abs(stat[a].mean(0) - stat[b].mean(0)) / np.sqrt(stat[a].var(0, ddof=1) / R + stat[b].var(0, ddof=1) / R)
        print("single", a, "vs", b, "frac z<=2: %.2f max z %.2f" % (np.mean(z <= 2), z.max()), "mean-row z", z[0].round(1))
    sd = {k: v.reshape(-1, D).std(0) for k, v in arms.items()}
    print("sd ratio ref/par", (sd["reference"] / sd["parallel"]).round(3))

def hier(burn_nmc, nmc, thin, R=8):
    fx = load_fixture(2); S, D = fx.n_pop, fx.ct.npar; C = 6 * D
    trials = [fx.trials(f"pop{s}") for s in range(S)]; pp, hp = fx.prior("p_prior"), fx.prior("h_prior")
    rng = np.random.default_rng(7); setups = [hier_setup(fx, S, C, rng) for _ in range(R)]
    phi_st = E.PopState(np.stack([s[0][0] for s in setups]), np.stack([s[0][1] for s in setups]), np.stack([s[0][2] for s in setups]))
    sub_st = [E.PopState(np.stack([s[1][i][0] for s in setups]), np.stack([s[1][i][1] for s in setups]), np.stack([s[1][i][2] for s in setups])) for i in range(S)]
    res = {}
    for name, sched, seed0 in (("reference", 0, 10), ("parallel", 1, 50), ("simultaneous", 2, 90)):
        t0 = time.time(); kw = dict(nchain=C, thin=thin, nparameter=2 * D, schedule=sched)
        pb, sb = E.run_hier(fx.ct, trials, pp, hp, E.Tuning(nmc=burn_nmc, pop_migration_prob=0.05, sub_migration_prob=0.05, seeds=[seed0 + r for r in range(R)], **kw), phi_st, sub_st)
        po, so = E.run_hier(fx.ct, trials, pp, hp, E.Tuning(nmc=nmc, seeds=[seed0 + 500 + r for r in range(R)], **kw),
                            E.PopState(pb.theta[:, -1], pb.lp[:, -1], pb.ll[:, -1]), [E.PopState(o.theta[:, -1], o.lp[:, -1], o.ll[:, -1]) for o in sb])
        res[name] = (po.theta[:, 1:], so[0].theta[:, 1:])
        print(name, "time %.1fs" % (time.time() - t0), "phi rhat max", max(rhat(res[name][0][r]).max() for r in range(R)).round(4), flush=True)
    for a, b in (("reference", "parallel"), ("reference", "simultaneous")):
        for idx, nm in ((0, "phi"), (1, "subj0")):
            sa = np.stack([summaries(res[a][idx][r]) for r in range(R)]); sb_ = np.stack([summaries(res[b][idx][r]) for r in range(R)])
            z = np.abs(sa.mean(0) - sb_.mean(0)) / np.sqrt(sa.var(0, ddof=1) / R + sb_.var(0, ddof=1) / R)
            print("hier", nm, a, "vs", b, "frac z<=2: %.2f max z %.2f" % (np.mean(z <= 2), z.max()), "mean-row z", z[0].round(1))
            if idx == 0:
                print("   means", sa.mean(0)[0].round(3)); print("        ", sb_.mean(0)[0].round(3))

if __name__ == "__main__":
    single(501, 1001, 8)
    hier(1001, 2001, 8)
