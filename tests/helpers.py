"""Shared test helpers: golden fixtures -> engine inputs and oracle inputs."""
import os
from dataclasses import dataclass
from typing import List

import numpy as np

from ggdmc_b200 import api

from ggdmc_b200.model import CellTable, PriorTable, Trials
from oracle import binding as ob

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRIOR_FIELDS = ("p0", "p1", "lower", "upper", "dist", "log_p")


@dataclass
class Fixture:
    k: int
    g: dict
    ct: CellTable
    om: "ob.OModel"

    def prior(self, prefix) -> PriorTable:
        g = self.g
        return PriorTable(len(g[f"{prefix}_p0"]), g[f"{prefix}_p0"].copy(), g[f"{prefix}_p1"].copy(), g[f"{prefix}_lower"].copy(),
                          g[f"{prefix}_upper"].copy(), g[f"{prefix}_dist"].astype(np.int32), g[f"{prefix}_log_p"].astype(np.uint8),
                          [str(s) for s in g[f"{prefix}_names"]])

    def oprior(self, prefix) -> "ob.OPrior":
        return ob.OPrior(*[self.g[f"{prefix}_{f}"] for f in PRIOR_FIELDS])

    def trials(self, which) -> Trials:
        return Trials(self.g[f"{which}_rt"].copy(), self.g[f"{which}_cell"].astype(np.uint16))

    def odata(self, which) -> "ob.OData":
        return ob.OData(self.g[f"{which}_rt"], self.g[f"{which}_cell"])

    @property
    def n_pop(self) -> int:
        return int(self.g["n_pop"])


_cache = {}


def load_fixture(k: int) -> Fixture:
    if k not in _cache:
        g = dict(np.load(os.path.join(GOLDEN, f"lba_data{k}.npz")))
        ct = CellTable(int(g["param_src"].shape[2]), int(g["param_src"].shape[0]), len(g["pnames"]), g["param_src"].astype(np.int32),
                       g["const_val"].astype(np.float64), g["posdrift"].astype(np.uint8), [str(s) for s in g["pnames"]],
                       [str(s) for s in g["cell_names"]])
        om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar)
        _cache[k] = Fixture(k, g, ct, om)
    return _cache[k]


def sane_starts(fx: Fixture, nchain: int, rng: np.random.Generator, center=None, jitter=0.05) -> np.ndarray:
    """Start vectors near the generating values (README.md:44-66) with multiplicative jitter."""
    c = fx.g["p_vector"] if center is None else center
    return c[None, :] * (1.0 + jitter * rng.standard_normal((nchain, len(c))))


def cond_mask_tolerance(logd_ref: np.ndarray, rel=1e-10):
    """Tolerance for per-trial log densities.

    1e-10 relative on the log density (the north-star bound) wherever the density is well
    conditioned.  Where the reference's own arithmetic cancels catastrophically -- densities near
    the 1e-10 floor, or (1 - cdf) a few ulps above 0 -- two correct FP64 implementations differ by
    kappa * eps; there the bound widens with 1/density (absolute error ~1e-15 on the density).
    """
    d = np.exp(logd_ref)
    tol = rel * np.maximum(np.abs(logd_ref), 1.0) + 4e-15 / np.maximum(d, 1e-300)
    return tol


def fixture_objects(k):
    """Rebuild the reference's model / dmi / prior objects of fixture k from the committed golden file."""
    fx = load_fixture(k)
    g = fx.g
    model = api.Model(parameter_x_condition_names=[str(s) for s in g["pxc_names"]], pnames=fx.ct.pnames, cell_names=fx.ct.cell_names,
                      constants=api.NamedVector(g["const_val"], [str(s) for s in g["const_names"]]), model_boolean=g["model_boolean"],
                      type="lba", npar=fx.ct.npar)

    def dmi_of(which):
        tr = fx.trials(which)
        data = api.NamedList({fx.ct.cell_names[c]: tr.rt[tr.cell == c] for c in np.unique(tr.cell)})
        return api.DMI(model=model, data=data, node_1_index=g["node_1_index"], is_positive_drift=g["is_positive_drift"])

    return fx, model, dmi_of


# ---- DDM ("fastdm") test model: stimulus S (s1, s2) x response R (r1, r2); drift rate by stimulus; r2 = upper boundary ----
DDM_PNAMES = ["a", "st0", "sv", "sz", "t0", "v.s1", "v.s2", "z"]
DDM_CONST = {"d": 0.0, "precision": 3.0, "s": 1.0}


def ddm_model(precision=3.0, s=1.0):
    """(CellTable, OModel) of a 4-cell DDM: rows a, d, precision, s, st0, sv, sz, t0, v, z (ggdmc_b200.model.DDM_CORE);
    free parameters DDM_PNAMES (alphabetical like the reference's pnames), constants d, precision, s."""
    from ggdmc_b200.model import DDM_CORE
    cells = ["s1.r1", "s1.r2", "s2.r1", "s2.r2"]
    cnames = list(DDM_CONST)
    src = np.zeros((4, 10, 2), dtype=np.int32)
    for c in range(4):
        for r, core in enumerate(DDM_CORE):
            name = f"v.s{c // 2 + 1}" if core == "v" else core
            src[c, r, :] = DDM_PNAMES.index(name) if name in DDM_PNAMES else -1 - cnames.index(name)
    const = np.array([0.0, precision, s])
    upper = np.array([0, 1, 0, 1], dtype=np.uint8)
    ct = CellTable(2, 4, len(DDM_PNAMES), src, const, upper, list(DDM_PNAMES), cells, "fastdm")
    return ct, ob.OModel(src, const, upper, ct.npar, type=ob.MODEL_DDM)


def ddm_theta(rng: np.random.Generator, kind: int) -> np.ndarray:
    """A plausible parameter vector in DDM_PNAMES order.  kind 0: no variability; 1: + sv; 2: + sz; 3: + st0."""
    a = rng.uniform(0.6, 2.2)
    th = dict(a=a, st0=0.0, sv=0.0, sz=0.0, t0=rng.uniform(0.1, 0.3), z=a * rng.uniform(0.35, 0.65))
    th["v.s1"], th["v.s2"] = rng.normal(-1.5, 1.0), rng.normal(1.5, 1.0)
    if kind >= 1:
        th["sv"] = rng.uniform(0.05, 1.5)
    if kind >= 2:
        th["sz"] = a * rng.uniform(0.02, 0.3)
    if kind >= 3:
        th["st0"] = rng.uniform(0.02, 0.25)
    return np.array([th[n] for n in DDM_PNAMES])


def ddm_simulate(theta: np.ndarray, n_per_stim: int, rng: np.random.Generator, dt=1e-3, s=1.0):
    """Euler-Maruyama simulation of the 4-cell model above (test data only): returns (rt, cell) grouped by cell."""
    p = dict(zip(DDM_PNAMES, theta))
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
        keep = done
        rts.append(rt[keep])
        cells.append((2 * stim + resp[keep]).astype(np.uint16))
    rt, cell = np.concatenate(rts), np.concatenate(cells)
    order = np.argsort(cell, kind="stable")
    return rt[order], cell[order]


def ddm_prior(kind="sub"):
    """Uniform priors over DDM_PNAMES wide enough for the generators above: (PriorTable, OPrior)."""
    lo = np.array([0.2, 0.0, 0.0, 0.0, 0.0, -6.0, -6.0, 0.05])
    hi = np.array([4.0, 0.5, 3.0, 1.0, 0.6, 6.0, 6.0, 3.5])
    n = len(lo)
    dist, logp = np.full(n, 6, np.int32), np.ones(n, np.uint8)
    return (PriorTable(n, lo.copy(), hi.copy(), np.zeros(n), np.zeros(n), dist, logp, list(DDM_PNAMES)),
            ob.OPrior(lo, hi, np.zeros(n), np.zeros(n), dist, logp))


def ddm_objects(rt=None, cell=None):
    """The reference-style `model` / `dmi` objects (ggdmc_b200.api mirrors) of the 4-cell DDM design above:
    model_boolean [ncell x n_pxc x n_acc] with one TRUE per (cell, core parameter, accumulator), type "fastdm",
    dmi@is_positive_drift per CELL (upper-boundary response)."""
    pxc = ["a", "d", "precision", "s", "st0", "sv", "sz", "t0", "v.s1", "v.s2", "z"]
    cells = ["s1.r1", "s1.r2", "s2.r1", "s2.r2"]
    mb = np.zeros((4, len(pxc), 2), dtype=bool)
    for c in range(4):
        for k, name in enumerate(pxc):
            mb[c, k, :] = (name == f"v.s{c // 2 + 1}") if name.startswith("v.") else True
    model = api.Model(parameter_x_condition_names=pxc, pnames=list(DDM_PNAMES), cell_names=cells,
                      constants=api.NamedVector(list(DDM_CONST.values()), list(DDM_CONST)), model_boolean=mb, type="fastdm",
                      npar=len(DDM_PNAMES))
    node_1 = np.array([[0, 1], [1, 0], [0, 1], [1, 0]])
    data = None
    if rt is not None:
        data = api.NamedList({cells[c]: np.asarray(rt)[np.asarray(cell) == c] for c in np.unique(cell)})
    return model, api.DMI(model=model, data=data, node_1_index=node_1, is_positive_drift=np.array([False, True, False, True]))
