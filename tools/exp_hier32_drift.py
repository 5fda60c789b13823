"""Experiment (GPU): drift of phi means over a long run, per schedule (README hierarchy, 32 subjects)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ggdmc_b200 import engine as E, workloads as W
R, thin = 4, 8
w = W.hierarchical("c2", 6, 32, 768, n_replicate=R)
for name, sched, seed0, nseg, seg_nmc in (("parallel", 1, 50, 12, 1001), ("simultaneous", 2, 90, 12, 1001), ("reference", 0, 10, 6, 501)):
    t0 = time.time()
    phi, subj = w.phi_start, w.subj_start
    rows = []
    for seg in range(nseg):
        mig = 0.05 if seg == 0 else 0.0
        tun = W.tuning_for(w, nmc=seg_nmc, thin=thin, seeds=[seed0 + 100 * seg + r for r in range(R)], schedule=sched, pop_migration_prob=mig, sub_migration_prob=mig)
        po, so = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, phi, subj)
        phi = E.PopState(po.theta[:, -1], po.lp[:, -1], po.ll[:, -1]); subj = [E.PopState(o.theta[:, -1], o.lp[:, -1], o.ll[:, -1]) for o in so]
        m = po.theta[:, 1:].mean(axis=(1, 2))  # [R, 2D]
        rows.append(m)
        print(name, "seg", seg, "its", (seg + 1) * (seg_nmc - 1) * thin, "loc[1,8,10] per rep:", m[:, 1].round(3), m[:, 8].round(3), "sca[1]:", m[:, 14].round(3), "t=%.0fs" % (time.time() - t0), flush=True)
