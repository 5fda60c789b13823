"""Shared by tests/test_gpu_posterior.py and tools/exp_c2_recovery.py: the README's staged hierarchical fit through the
engine, the package's own R-hat, and replicate-level z-scores."""
import numpy as np

from ggdmc_b200 import engine as E
from ggdmc_b200 import workloads as W

# README.md:181-196: (nmc, thin, sub_migration_prob, pop_migration_prob) of StartSampling and the two RestartSampling calls
README_STAGES = [(500, 8, 0.06, 0.0), (1000, 8, 0.0, 0.05), (1000, 8, 0.0, 0.01)]


def gelman_pkg(theta):
    """Point estimate of the potential scale reduction factor per parameter exactly as the package computes it
    (.gelman_diag / .compute_psrf_components, R/model-class.R:1559-1690; note its var_w = row variance of the mean
    within-chain covariance matrix / nchain).  theta [n, nchain, npar]."""
    n, m, D = theta.shape
    Wm = np.mean([np.cov(theta[:, k, :].T) for k in range(m)], axis=0)
    Bm = n * np.cov(theta.mean(0).T)
    w, b = np.diag(Wm), np.diag(Bm)
    var_w = Wm.var(axis=1, ddof=1) / m
    var_b = 2.0 * b ** 2 / (m - 1)
    V = (n - 1) / n * w + (1 + 1 / m) / n * b
    var_V = ((n - 1) ** 2 * var_w + (1 + 1 / m) ** 2 * var_b) / n ** 2
    df_V = 2.0 * V ** 2 / var_V
    return np.sqrt((df_V + 3) / (df_V + 1) * ((n - 1) / n + (1 + 1 / m) / n * (b / w)))


def summaries(x):
    flat = x.reshape(-1, x.shape[-1])
    return np.stack([flat.mean(0), np.quantile(flat, 0.05, axis=0), np.quantile(flat, 0.5, axis=0), np.quantile(flat, 0.975, axis=0)])


def zscores(a, b):
    """a, b [R, n, C, D] -> |difference| of every summary between two arms in units of its replicate-level standard error."""
    sa = np.stack([summaries(a[r]) for r in range(a.shape[0])])
    sb = np.stack([summaries(b[r]) for r in range(b.shape[0])])
    return np.abs(sa.mean(0) - sb.mean(0)) / np.sqrt(sa.var(0, ddof=1) / a.shape[0] + sb.var(0, ddof=1) / b.shape[0])


def run_stages(w, schedule, seeds, stages=README_STAGES, start=None, keep_subjects=(0, 15, 31)):
    """A staged fit from `start` = (phi state, subject states) (default: the workload's start values); returns
    ([per stage (phi samples [R, nmc - 1, C, 2 D], [subject samples [R, nmc - 1, C, D] ...])], final state)."""
    phi, subj = start if start is not None else (w.phi_start, w.subj_start)
    out = []
    for i, (nmc, thin, sub_mig, pop_mig) in enumerate(stages):
        tun = W.tuning_for(w, nmc=nmc, thin=thin, seeds=[s + 100 * i for s in seeds], schedule=schedule, pop_migration_prob=pop_mig,
                           sub_migration_prob=sub_mig)
        po, so = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, phi, subj)
        phi = E.PopState(po.theta[:, -1], po.lp[:, -1], po.ll[:, -1])
        subj = [E.PopState(o.theta[:, -1], o.lp[:, -1], o.ll[:, -1]) for o in so]
        out.append((po.theta[:, 1:].copy(), [so[k].theta[:, 1:].copy() for k in keep_subjects if k < len(so)]))
    return out, (phi, subj)
