"""Experiment (GPU): README hierarchical recovery study (32 subjects x 768 trials, 78 chains):
do the three schedules give the same phi / subject posteriors once R-hat < 1.05?"""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ggdmc_b200 import engine as E, workloads as W
from test_gpu_posterior import summaries

def rhat(x):
    n = x.shape[0]; cm = x.mean(0); Wv = x.var(0, ddof=1).mean(0); Bn = cm.var(0, ddof=1)
    return np.sqrt((n - 1) / n + Bn / Wv)

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
burn_nmc, nmc, thin = int(sys.argv[2]) if len(sys.argv) > 2 else 501, int(sys.argv[3]) if len(sys.argv) > 3 else 1001, 8
w = W.hierarchical("c2", 6, 32, 768, n_replicate=R)
D = w.spec.ct.npar
res = {}
for name, sched, seed0 in (("reference", 0, 10), ("parallel", 1, 50), ("simultaneous", 2, 90)):
    t0 = time.time()
    tun = W.tuning_for(w, nmc=burn_nmc, thin=thin, seeds=[seed0 + r for r in range(R)], schedule=sched, pop_migration_prob=0.05, sub_migration_prob=0.05)
    pb, sb = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
    tun2 = W.tuning_for(w, nmc=nmc, thin=thin, seeds=[seed0 + 500 + r for r in range(R)], schedule=sched, pop_migration_prob=0.0, sub_migration_prob=0.0)
    po, so = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun2, E.PopState(pb.theta[:, -1], pb.lp[:, -1], pb.ll[:, -1]),
                        [E.PopState(o.theta[:, -1], o.lp[:, -1], o.ll[:, -1]) for o in sb])
    res[name] = (po.theta[:, 1:], so[0].theta[:, 1:], so[5].theta[:, 1:])
    print(name, "time %.1fs" % (time.time() - t0), "phi rhat max", max(rhat(res[name][0][r]).max() for r in range(R)).round(4),
          "subj0 rhat max", max(rhat(res[name][1][r]).max() for r in range(R)).round(4), flush=True)
truth = np.concatenate([w.spec.pop_mean, w.spec.pop_scale])
for a, b in (("reference", "parallel"), ("reference", "simultaneous"), ("parallel", "simultaneous")):
    for idx, nm in ((0, "phi"), (1, "subj0"), (2, "subj5")):
        sa = np.stack([summaries(res[a][idx][r]) for r in range(R)]); sb_ = np.stack([summaries(res[b][idx][r]) for r in range(R)])
        z = np.abs(sa.mean(0) - sb_.mean(0)) / np.sqrt(sa.var(0, ddof=1) / R + sb_.var(0, ddof=1) / R)
        print(nm, a, "vs", b, "frac z<=2: %.2f max z %.2f" % (np.mean(z <= 2), z.max()), "mean-row z", z[0].round(1))
print("truth   ", truth.round(3))
for k in res:
    print("%-12s" % k, np.stack([summaries(res[k][0][r]) for r in range(R)]).mean(0)[0].round(3))
