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
