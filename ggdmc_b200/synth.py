"""Synthetic LBA data and start states for benchmarks, smoke tests and recovery tests.

Host-side helper (numpy): simulates the standard LBA race the way `lbaModel::simulate` is used in
the reference's recovery scripts (README.md:94-95, tests/testthat/Group1/data/6_lba_Bv_model.r):
start point k ~ U(0, A), drift v ~ N(mean_v, sd_v) truncated at 0 for positive-drift accumulators,
finishing time (b - k) / v + t0, the fastest accumulator responds.  Nothing here runs on the
sampling hot path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .model import CellTable, Trials


def cell_params(ct: CellTable, theta: np.ndarray, cell: int) -> np.ndarray:
    """6 x n_acc matrix (rows A, b, mean_v, sd_v, st0, t0) of one cell -- SURVEY.md A.1."""
    src = ct.param_src[cell]
    P = np.where(src >= 0, theta[np.maximum(src, 0)], ct.const_val[np.maximum(-1 - src, 0)])
    P = P.astype(np.float64)
    P[1] += P[0]
    return P


def condition_groups(ct: CellTable) -> List[List[int]]:
    """Cells that share a stimulus condition (cell name minus its last component, the response)."""
    groups = {}
    for c, name in enumerate(ct.cell_names):
        groups.setdefault(name.rsplit(".", 1)[0], []).append(c)
    return list(groups.values())


def simulate_subject(ct: CellTable, node_1_index: np.ndarray, theta: np.ndarray, n_trial: int,
                     rng: np.random.Generator) -> Trials:
    """Balanced design: n_trial split evenly over the stimulus conditions."""
    groups = condition_groups(ct)
    per = np.full(len(groups), n_trial // len(groups))
    per[: n_trial - per.sum()] += 1
    rts, cells = [], []
    for g, n in zip(groups, per):
        if n == 0:
            continue
        c0 = g[0]
        P = cell_params(ct, theta, c0)  # column j belongs to accumulator node_1_index[c0, j]
        na = ct.n_acc
        t = np.empty((n, na))
        for j in range(na):
            A, b, mv, sv, st0, t0 = P[:, j]
            k = rng.uniform(0.0, max(A, 0.0), size=n)
            v = rng.normal(mv, sv, size=n)
            if ct.posdrift[node_1_index[c0, j]]:
                bad = v <= 0
                while bad.any():
                    v[bad] = rng.normal(mv, sv, size=int(bad.sum()))
                    bad = v <= 0
            else:
                v = np.where(v <= 0, np.nan, v)
            t[:, j] = (b - k) / v + t0 + st0 * rng.uniform(size=n)
        t = np.where(np.isnan(t), np.inf, t)
        win = np.argmin(t, axis=1)
        acc_of_col = node_1_index[c0]
        cell_of_acc = {int(node_1_index[c, 0]): c for c in g}
        rts.append(t[np.arange(n), win])
        cells.append(np.array([cell_of_acc[int(acc_of_col[w])] for w in win], dtype=np.uint16))
    rt = np.concatenate(rts)
    cell = np.concatenate(cells)
    order = np.argsort(cell, kind="stable")
    return Trials(rt[order], cell[order])


def rtnorm(mean: np.ndarray, sd: np.ndarray, lower: float, rng: np.random.Generator, size=None) -> np.ndarray:
    x = rng.normal(mean, sd, size=size)
    bad = x < lower
    while np.any(bad):
        x = np.where(bad, rng.normal(mean, sd, size=size), x)
        bad = x < lower
    return x


def simulate_population(ct: CellTable, node_1_index: np.ndarray, pop_mean: np.ndarray, pop_scale: np.ndarray, n_subject: int,
                        n_trial: int, seed: int):
    """theta_s ~ tnorm(pop_mean, pop_scale, lower 0); returns (true thetas [S, npar], [Trials])."""
    rng = np.random.default_rng(seed)
    thetas = np.stack([rtnorm(pop_mean, pop_scale, 0.0, rng) for _ in range(n_subject)])
    trials = [simulate_subject(ct, node_1_index, thetas[s], n_trial, rng) for s in range(n_subject)]
    return thetas, trials
