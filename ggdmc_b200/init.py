"""Start-point generation on the GPU: `initialise_theta` / `initialise_phi` (reference: R/phi.R:141-332).

The reference draws one candidate at a time in R (`ggdmcPrior::rprior`), scores it with one R -> C++
likelihood call and keeps the first candidate per chain whose log prior and log likelihood are finite --
up to `max_init_attempts` sequential round trips per chain, per subject.  Here all chains (and all
subjects) are scored together: candidates are drawn in batches on the host, their log priors and log
likelihoods come from two C-ABI calls (`ggdmc_b200_sumlogprior`, `ggdmc_b200_sumloglike_init`, the latter
with the R-side `.sumlog` rule: densities <= 0 are floored at .Machine$double.eps, R/phi.R:3-13), and
every chain takes the first valid candidate of its own candidate sequence.  The random stream is numpy's,
not R's, so start values are reproducible for a seed but are not the reference's draws.
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np

from . import engine as E
from .api import Posterior, _flatten_dmi, _scalar, _strs
from .model import (DIST_BETA_LU, DIST_CAUCHY, DIST_GAMMA_L, DIST_LNORM_L, DIST_NORM, DIST_TNORM, DIST_UNIF, PriorTable,
                    flatten_prior, slot)


def rprior(table: PriorTable, n: int, rng: np.random.Generator) -> np.ndarray:
    """n draws from the joint prior -> [n, npar] (ggdmcPrior::rprior; families of @hdr/prior.h:186)."""
    out = np.empty((n, table.npar))
    for i in range(table.npar):
        p0, p1, lo, up, dist = table.p0[i], table.p1[i], table.lower[i], table.upper[i], int(table.dist[i])
        if dist == DIST_TNORM:
            x = rng.normal(p0, p1, size=n)
            bad = (x < lo) | (x > up)
            guard = 0
            while bad.any() and guard < 10000:
                x[bad] = rng.normal(p0, p1, size=int(bad.sum()))
                bad = (x < lo) | (x > up)
                guard += 1
        elif dist == DIST_UNIF:
            x = rng.uniform(p0, p1, size=n)
        elif dist == DIST_NORM:
            x = rng.normal(p0, p1, size=n)
        elif dist == DIST_BETA_LU:
            x = lo + (up - lo) * rng.beta(p0, p1, size=n)
        elif dist == DIST_GAMMA_L:
            x = (lo if np.isfinite(lo) else 0.0) + rng.gamma(p0, p1, size=n)
        elif dist == DIST_LNORM_L:
            x = (lo if np.isfinite(lo) else 0.0) + rng.lognormal(p0, p1, size=n)
        elif dist == DIST_CAUCHY:
            flo = 0.5 + np.arctan((lo - p0) / p1) / np.pi if np.isfinite(lo) else 0.0
            fup = 0.5 + np.arctan((up - p0) / p1) / np.pi if np.isfinite(up) else 1.0
            u = rng.uniform(flo, fup, size=n)
            x = p0 + p1 * np.tan(np.pi * (u - 0.5))
        else:
            raise ValueError(f"unknown dist_id {dist}")
        out[:, i] = x
    return out


def _first_valid(valid: np.ndarray) -> np.ndarray:
    """valid [n_chain, n_try] -> index of the first True per chain, -1 if none."""
    idx = np.argmax(valid, axis=1)
    idx[~valid.any(axis=1)] = -1
    return idx


def _new_samples(theta_input, npar: int, pnames: List[str]) -> Posterior:
    """set_up_new_samples (R/phi.R:72-88): NaN thetas, -Inf log prior / likelihood."""
    nmc, nchain, thin = int(_scalar(slot(theta_input, "nmc"))), int(_scalar(slot(theta_input, "nchain"))), int(_scalar(slot(theta_input, "thin")))
    return Posterior(theta=np.full((npar, nchain, nmc), np.nan), summed_log_prior=np.full((nchain, nmc), -np.inf),
                     log_likelihoods=np.full((nchain, nmc), -np.inf), start=1, npar=npar, pnames=list(pnames), nmc=nmc, thin=thin,
                     nchain=nchain)


def initialise_thetas(theta_input, priors, dmis: Sequence[Any], seed: Optional[int] = None, batch: int = 8) -> List[Posterior]:
    """`initialise_theta` for many subjects at once (what initialise_phi loops over, R/phi.R:286-296)."""
    flat = [_flatten_dmi(d) for d in dmis]
    ct = flat[0][0]
    trials = [t for _, t in flat]
    p_prior = flatten_prior(slot(priors, "p_prior"))
    nchain = int(_scalar(slot(theta_input, "nchain")))
    max_attempts = int(_scalar(slot(theta_input, "max_init_attempts")))
    S, D = len(dmis), ct.npar
    rng = np.random.default_rng(seed)
    theta = np.full((S, nchain, D), np.nan)
    lp = np.full((S, nchain), -np.inf)
    ll = np.full((S, nchain), -np.inf)
    todo = np.ones((S, nchain), dtype=bool)
    attempts = 0
    while todo.any():
        if attempts >= max_attempts:
            s, c = np.argwhere(todo)[0]
            raise RuntimeError(f"Chain {c + 1}: Failed to find valid theta after {max_attempts} attempts")  # R/phi.R:199-201
        n_try = min(batch, max_attempts - attempts)
        cand = rprior(p_prior, S * nchain * n_try, rng).reshape(S, nchain * n_try, D)
        slp = E.sumlogprior(p_prior, cand.reshape(-1, D)).reshape(S, nchain, n_try)
        sll = E.sumloglike(ct, trials, cand, init_rule=True).reshape(S, nchain, n_try)
        valid = np.isfinite(slp) & np.isfinite(sll)
        cand = cand.reshape(S, nchain, n_try, D)
        for s in range(S):
            first = _first_valid(valid[s])
            take = todo[s] & (first >= 0)
            cs = np.nonzero(take)[0]
            theta[s, cs] = cand[s, cs, first[cs]]
            lp[s, cs] = slp[s, cs, first[cs]]
            ll[s, cs] = sll[s, cs, first[cs]]
            todo[s, cs] = False
        attempts += n_try
    out = []
    for s in range(S):
        post = _new_samples(theta_input, D, ct.pnames)
        post.theta[:, :, 0] = theta[s].T
        post.summed_log_prior[:, 0] = lp[s]
        post.log_likelihoods[:, 0] = ll[s]
        out.append(post)
    return out


def initialise_theta(theta_input, priors, dmi, seed: Optional[int] = None) -> Posterior:
    """R/phi.R:141-204."""
    if isinstance(dmi, (list, tuple)):
        dmi = dmi[0]  # "Use the first instance in the dmi list."
    return initialise_thetas(theta_input, priors, [dmi], seed)[0]


def initialise_phi(theta_input, priors, dmis: Sequence[Any], seed: Optional[int] = None, batch: int = 8):
    """R/phi.R:256-332: subject start values from p_prior, then phi chains from h_prior scored with the
    hyper-likelihood of the subjects' chain-k thetas (`.sumloghlike`, R/phi.R:37-48).  Returns
    {"phi": posterior, "subject_theta": [posterior, ...]}."""
    if slot(priors, "h_prior") is None:
        raise ValueError("hyper prior is NULL")
    p_prior, h_prior = flatten_prior(slot(priors, "p_prior")), flatten_prior(slot(priors, "h_prior"))
    subj = initialise_thetas(theta_input, priors, dmis, seed, batch)
    nchain = int(_scalar(slot(theta_input, "nchain")))
    max_attempts = int(_scalar(slot(theta_input, "max_init_attempts")))
    S, D = len(dmis), p_prior.npar
    rng = np.random.default_rng(None if seed is None else seed + 7919)
    th_s = np.stack([p.theta[:, :, 0].T for p in subj])  # [S, nchain, D]
    phi = _new_samples(theta_input, 2 * D, h_prior.pnames)
    todo = np.ones(nchain, dtype=bool)
    attempts = 0
    while todo.any():
        if attempts >= max_attempts:
            raise RuntimeError(f"Chain {np.nonzero(todo)[0][0] + 1}: Failed to find valid phi after {max_attempts} attempts")
        n_try = min(batch, max_attempts - attempts)
        cand = rprior(h_prior, nchain * n_try, rng).reshape(nchain, n_try, 2 * D)
        slp = E.sumlogprior(h_prior, cand.reshape(-1, 2 * D)).reshape(nchain, n_try)
        # hyper-likelihood of candidate (k, j): sum over subjects of log p(theta_{s,k} | candidate)
        x = np.broadcast_to(th_s[:, :, None, :], (S, nchain, n_try, D)).reshape(-1, D)
        ph = np.broadcast_to(cand[None], (S, nchain, n_try, 2 * D)).reshape(-1, 2 * D)
        sll = E.sumlogprior(p_prior, x, np.ascontiguousarray(ph[:, :D]), np.ascontiguousarray(ph[:, D:])).reshape(S, nchain, n_try).sum(0)
        valid = np.isfinite(slp) & np.isfinite(sll)
        first = _first_valid(valid)
        cs = np.nonzero(todo & (first >= 0))[0]
        phi.theta[:, cs, 0] = cand[cs, first[cs]].T
        phi.summed_log_prior[cs, 0] = slp[cs, first[cs]]
        phi.log_likelihoods[cs, 0] = sll[cs, first[cs]]
        todo[cs] = False
        attempts += n_try
    return {"phi": phi, "subject_theta": subj}
