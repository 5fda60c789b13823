"""The BASELINE.json workload shapes, built from synthetic data.

  C1  single-subject LBA B x v model (13 par, 24 cells, 2 acc, 768 trials), 39 chains
  C2  hierarchical recovery study: 32 subjects x 768 trials, 78 chains
  C3  likelihood-only sweep: 5-par 2-acc model, 15 chains, 1e3 .. 1e7 trials
  C4  scaled hierarchy: 1024 subjects x 768 trials, 78 chains (subjects sharded over GPUs)
  C5  4-accumulator model, 96 cells, 17 par, 102 chains, 256 subjects x 2048 trials

The model tables (cell table, generating population values, prior layout) come from the committed
fixtures tests/golden/lba_data{6,5,2}.npz, which were extracted from the reference's own fixture
files; the data are simulated here.  Start states are the generating values with jitter and their
log prior / log likelihood are evaluated ON THE GPU through the C ABI (the job of
initialise_theta / initialise_phi, R/phi.R:141-332, minus the rejection sampling).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import engine as E
from . import synth
from .model import CellTable, PriorTable, Trials

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


@dataclass
class ModelSpec:
    ct: CellTable
    node_1_index: np.ndarray
    pop_mean: np.ndarray
    pop_scale: np.ndarray
    p_prior: PriorTable  # hyper-likelihood layout (tnorm, lower 0)
    h_prior: PriorTable  # prior on phi
    sub_prior: PriorTable  # single-subject prior


def _prior(g, prefix) -> PriorTable:
    return PriorTable(len(g[f"{prefix}_p0"]), g[f"{prefix}_p0"].copy(), g[f"{prefix}_p1"].copy(), g[f"{prefix}_lower"].copy(),
                      g[f"{prefix}_upper"].copy(), g[f"{prefix}_dist"].astype(np.int32), g[f"{prefix}_log_p"].astype(np.uint8),
                      [str(s) for s in g[f"{prefix}_names"]])


def load_model(k: int) -> ModelSpec:
    g = np.load(os.path.join(GOLDEN, f"lba_data{k}.npz"))
    ct = CellTable(int(g["param_src"].shape[2]), int(g["param_src"].shape[0]), len(g["pnames"]), g["param_src"].astype(np.int32),
                   g["const_val"].astype(np.float64), g["posdrift"].astype(np.uint8), [str(s) for s in g["pnames"]],
                   [str(s) for s in g["cell_names"]])
    return ModelSpec(ct, g["node_1_index"].astype(np.int64), g["pop_mean"].copy(), g["pop_scale"].copy(), _prior(g, "p_prior"),
                     _prior(g, "h_prior"), _prior(g, "sub_prior"))


def sweep_model():
    """BASELINE config 3's model: the 5-parameter, 2-accumulator LBA of the reference's first test model
    (tests/testthat/Group1/data/0_lba_model0.r:22-30: A, B, mean_v.false, mean_v.true, t0 free; sd_v = 1, st0 = 0),
    2 stimuli x 2 responses = 4 cells.  Returns (cell table, node_1_index, p_vector, uniform prior)."""
    pnames = ["A", "B", "mean_v.false", "mean_v.true", "t0"]
    const_val = np.array([1.0, 0.0])  # sd_v, st0
    n_cell, n_acc = 4, 2
    param_src = np.zeros((n_cell, 6, n_acc), dtype=np.int32)
    node_1 = np.zeros((n_cell, n_acc), dtype=np.int64)
    names = []
    for s_ in range(2):
        for r in range(2):
            c = 2 * s_ + r
            names.append(f"s{s_ + 1}.r{r + 1}")
            node_1[c] = (r, 1 - r)  # responder first
            for j, acc in enumerate((r, 1 - r)):
                param_src[c, :, j] = (0, 1, 3 if acc == s_ else 2, -1, -2, 4)  # rows A, B, mean_v, sd_v, st0, t0
    ct = CellTable(n_acc, n_cell, 5, param_src, const_val, np.array([1, 1], dtype=np.uint8), pnames, names)
    p_vector = np.array([0.75, 1.25, 1.5, 2.5, 0.15])
    prior = PriorTable(5, np.zeros(5), np.full(5, 10.0), np.zeros(5), np.full(5, 10.0), np.full(5, 6, dtype=np.int32),
                       np.ones(5, dtype=np.uint8), pnames)
    return ct, node_1, p_vector, prior


DDM_PNAMES = ["a", "st0", "sv", "sz", "t0", "v.s1", "v.s2", "z"]


def ddm_model(precision: float = 3.0, s: float = 1.0, fixed=()):
    """A 4-cell DDM design (model type "fastdm"): stimulus S (s1, s2) x response R (r1, r2), drift rate by stimulus,
    r2 = upper boundary; free parameters DDM_PNAMES minus `fixed` (names held at the constant 0, the usual way to switch a
    variability off), further constants d = 0, precision, s.  Rows follow model.DDM_CORE.
    Returns (cell table, p_vector, uniform prior)."""
    from .model import DDM_CORE
    pnames = [n for n in DDM_PNAMES if n not in fixed]
    cnames = ["d", "precision", "s"] + [n for n in DDM_PNAMES if n in fixed]
    src = np.zeros((4, len(DDM_CORE), 2), dtype=np.int32)
    for c in range(4):
        for r, core in enumerate(DDM_CORE):
            name = f"v.s{c // 2 + 1}" if core == "v" else core
            src[c, r, :] = pnames.index(name) if name in pnames else -1 - cnames.index(name)
    const = np.array([0.0, precision, s] + [0.0] * (len(cnames) - 3))
    ct = CellTable(2, 4, len(pnames), src, const, np.array([0, 1, 0, 1], dtype=np.uint8), list(pnames),
                   ["s1.r1", "s1.r2", "s2.r1", "s2.r2"], "fastdm")
    full = dict(zip(DDM_PNAMES, [1.2, 0.1, 0.6, 0.2, 0.2, -1.8, 1.8, 0.6]))
    lo = dict(zip(DDM_PNAMES, [0.2, 0.0, 0.0, 0.0, 0.0, -6.0, -6.0, 0.05]))
    hi = dict(zip(DDM_PNAMES, [4.0, 0.5, 3.0, 1.0, 0.6, 6.0, 6.0, 3.5]))
    n = len(pnames)
    prior = PriorTable(n, np.array([lo[k] for k in pnames]), np.array([hi[k] for k in pnames]), np.zeros(n), np.zeros(n),
                       np.full(n, 6, dtype=np.int32), np.ones(n, dtype=np.uint8), list(pnames))
    return ct, np.array([full[k] for k in pnames]), prior


def ddm_readme_model():
    """The DDM of the reference's second README example (README.md:247-300): p_map all "1", factors S = (s1, s2),
    match_map s1 -> r1, s2 -> r2, constants d = 0, s = 1, st0 = 0, sv = 0, precision = 3, free parameters
    a, sz, t0, v, z; the matching response of a cell is the upper boundary (dmi@is_positive_drift per cell).
    Returns (cell table, p_vector of README.md:297, population mean / scale of README.md:278-279)."""
    from .model import DDM_CORE
    pnames = ["a", "sz", "t0", "v", "z"]
    cnames = ["d", "s", "st0", "sv", "precision"]
    const = np.array([0.0, 1.0, 0.0, 0.0, 3.0])
    src = np.zeros((4, len(DDM_CORE), 2), dtype=np.int32)
    for r, core in enumerate(DDM_CORE):
        src[:, r, :] = pnames.index(core) if core in pnames else -1 - cnames.index(core)
    upper = np.array([1, 0, 0, 1], dtype=np.uint8)  # cells s1.r1, s1.r2, s2.r1, s2.r2: the match is the upper boundary
    ct = CellTable(2, 4, 5, src, const, upper, pnames, ["s1.r1", "s1.r2", "s2.r1", "s2.r2"], "fastdm")
    p_vector = np.array([1.0, 0.25, 0.15, 2.5, 0.38])
    pop_mean, pop_scale = p_vector.copy(), np.array([0.05, 0.01, 0.02, 0.5, 0.01])
    return ct, p_vector, pop_mean, pop_scale


def ddm_simulate(theta: np.ndarray, n_per_stim: int, rng: np.random.Generator, dt: float = 1e-3, s: float = 1.0, pnames=None) -> Trials:
    """Euler-Maruyama simulation of ddm_model()'s design (synthetic inputs only): trials grouped by cell.
    `pnames` names the entries of theta (default DDM_PNAMES); parameters not named are 0."""
    p = dict.fromkeys(DDM_PNAMES, 0.0)
    p.update(zip(pnames or DDM_PNAMES, theta))
    rts, cells = [], []
    for stim in range(2):
        v = p[f"v.s{stim + 1}"] + p["sv"] * rng.standard_normal(n_per_stim)
        x = p["z"] + p["sz"] * (rng.uniform(size=n_per_stim) - 0.5)
        t = np.zeros(n_per_stim)
        done = np.zeros(n_per_stim, dtype=bool)
        resp = np.zeros(n_per_stim, dtype=int)
        for _ in range(20000):
            live = ~done
            if not live.any():
                break
            x[live] += v[live] * dt + s * np.sqrt(dt) * rng.standard_normal(live.sum())
            t[live] += dt
            up, lo = live & (x >= p["a"]), live & (x <= 0)
            resp[up] = 1
            done |= up | lo
        rt = t + p["t0"] + p["st0"] * rng.uniform(size=n_per_stim)
        rts.append(rt[done])
        cells.append((2 * stim + resp[done]).astype(np.uint16))
    rt, cell = np.concatenate(rts), np.concatenate(cells)
    order = np.argsort(cell, kind="stable")
    return Trials(rt[order], cell[order])


@dataclass
class HierWorkload:
    name: str
    spec: ModelSpec
    true_theta: np.ndarray  # [S, npar]
    trials: object  # E.TrialsStack (list-like of Trials)
    nchain: int
    phi_start: E.PopState
    subj_start: object  # E.PopStateStack (list-like of PopState)

    @property
    def n_trial_total(self) -> int:
        return int(sum(len(t.rt) for t in self.trials))


def shard_bounds(n_subject: int, rank: int, world: int):
    """Subjects [begin, end) owned by `rank`: contiguous, sizes differ by at most one."""
    return rank * n_subject // world, (rank + 1) * n_subject // world


@dataclass
class PopulationShard:
    """CPU-side description of subjects [subject_begin, subject_end) of a hierarchical problem."""

    spec: ModelSpec
    subject_begin: int
    subject_end: int
    true_theta: np.ndarray  # [S_local, npar]
    trials: List[Trials]
    nchain: int
    phi0: np.ndarray  # [R, C, 2 npar]  (identical on every shard)
    subj0: np.ndarray  # [S_local, R, C, npar]


def build_population(model_k: int, n_subject: int, n_trial: int, n_replicate: int = 1, data_seed: int = 20260101,
                     subject_begin: int = 0, subject_end: Optional[int] = None, start_seed: int = 1234) -> PopulationShard:
    """Synthetic data + start values (no GPU).  Everything about a subject depends only on
    (seed, GLOBAL subject index) and the phi start only on the seed, so the shards built by the
    ranks of a multi-GPU run are exactly the slices of the single-GPU problem."""
    spec = load_model(model_k)
    ct = spec.ct
    subject_end = n_subject if subject_end is None else subject_end
    D, C = ct.npar, 3 * 2 * ct.npar  # nchain = 3 x (number of phi parameters), R/sampling.R:1-10
    R = n_replicate
    thetas, trials = [], []
    for s in range(subject_begin, subject_end):
        rng = np.random.default_rng([data_seed, s])
        th = synth.rtnorm(spec.pop_mean, spec.pop_scale, 0.0, rng)
        thetas.append(th)
        trials.append(synth.simulate_subject(ct, spec.node_1_index, th, n_trial, rng))
    thetas = np.stack(thetas) if thetas else np.zeros((0, D))
    prng = np.random.default_rng([start_seed, 0xF1])
    center = np.concatenate([spec.pop_mean, spec.pop_scale])
    phi0 = np.abs(center[None, None, :] * (1.0 + 0.05 * prng.standard_normal((R, C, 2 * D))))
    subj0 = np.empty((len(trials), R, C, D))
    for i, s in enumerate(range(subject_begin, subject_end)):
        srng = np.random.default_rng([start_seed, s])
        subj0[i] = np.abs(thetas[i][None, None, :] * (1.0 + 0.05 * srng.standard_normal((R, C, D))))
    return PopulationShard(spec, subject_begin, subject_end, thetas, trials, C, phi0, subj0)


def hierarchical(name: str, model_k: int, n_subject: int, n_trial: int, n_replicate: int = 1, data_seed: int = 20260101,
                 subject_begin: int = 0, subject_end: Optional[int] = None, start_seed: int = 1234) -> HierWorkload:
    """build_population + log prior / log likelihood of the start states, evaluated on the GPU."""
    sh = build_population(model_k, n_subject, n_trial, n_replicate, data_seed, subject_begin, subject_end, start_seed)
    spec, ct, trials, phi0, subj0, C = sh.spec, sh.spec.ct, sh.trials, sh.phi0, sh.subj0, sh.nchain
    R, D, S = n_replicate, ct.npar, len(sh.trials)
    ll = E.sumloglike(ct, trials, subj0.reshape(S, R * C, D)).reshape(S, R, C)
    ph = np.broadcast_to(phi0.reshape(1, R * C, 2 * D), (S, R * C, 2 * D)).reshape(S * R * C, 2 * D)
    lp = E.sumlogprior(spec.p_prior, subj0.reshape(S * R * C, D), np.ascontiguousarray(ph[:, :D]),
                       np.ascontiguousarray(ph[:, D:])).reshape(S, R, C)
    phi_lp = E.sumlogprior(spec.h_prior, phi0.reshape(R * C, 2 * D)).reshape(R, C)
    phi_ll = lp.sum(axis=0)  # local subjects only; refreshed (and all-reduced) by the first phi step anyway
    return HierWorkload(name, spec, sh.true_theta, E.TrialsStack(trials), C, E.PopState(phi0, phi_lp, phi_ll),
                        E.PopStateStack(subj0, lp, ll))


def tuning_for(w: HierWorkload, nmc: int, thin: int, seeds, schedule=None, pop_migration_prob=0.05, sub_migration_prob=0.05,
               subject_begin=0, n_subject_total=0, device=-1) -> E.Tuning:
    from . import _lib as B
    return E.Tuning(nmc=nmc, nchain=w.nchain, thin=thin, nparameter=2 * w.spec.ct.npar, pop_migration_prob=pop_migration_prob,
                    sub_migration_prob=sub_migration_prob, schedule=B.SCHEDULE_PARALLEL if schedule is None else schedule,
                    seeds=list(seeds), subject_begin=subject_begin, n_subject_total=n_subject_total, device=device)
