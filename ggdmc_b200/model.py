"""Host-side flattening of the reference's S4 `model` / `dmi` / `prior` objects.

The reference marshals these S4 objects into C++ classes at the `.Call` boundary
(src/de2R.cpp:8-171 via ggdmcHeaders' `new_likelihood`, `new_prior`, ...).  The
B200 engine takes plain arrays instead; this module is that marshaller.  It
accepts either objects decoded from `.rda` files (:mod:`ggdmc_b200.rda`) or the
plain-Python mirrors in :mod:`ggdmc_b200.api` -- anything exposing the
reference's slot names through :func:`slot`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Sequence

import numpy as np

# core LBA parameter rows, alphabetical (design_class::set_parameter_values,
# @hdr/design_light.h:314-344)
LBA_CORE = ("A", "B", "mean_v", "sd_v", "st0", "t0")
# core DDM parameter rows, alphabetical like every core table; ddm_class::set_parameters reads them by
# row index (@hdr/ddm.h:194-214: a = 0, d = 1, precision = 2, s = 3, st0 = 4, sv = 5, sz = 6, t0 = 7, v = 8, z = 9)
DDM_CORE = ("a", "d", "precision", "s", "st0", "sv", "sz", "t0", "v", "z")
CORE_OF_TYPE = {"lba": LBA_CORE, "fastdm": DDM_CORE}

# prior::DistributionType (@hdr/prior.h:186)
DIST_TNORM, DIST_BETA_LU, DIST_GAMMA_L, DIST_LNORM_L, DIST_CAUCHY, DIST_UNIF, DIST_NORM = 1, 2, 3, 4, 5, 6, 7


def slot(obj: Any, name: str) -> Any:
    """Read slot `name` from an RS4, a mapping, or a plain object."""
    if hasattr(obj, "attrs") and not isinstance(obj, np.ndarray) and name in getattr(obj, "attrs", {}):
        return obj.attrs[name]
    if isinstance(obj, dict):
        return obj[name]
    return getattr(obj, name)


def _has_slot(obj: Any, name: str) -> bool:
    try:
        slot(obj, name)
        return True
    except (AttributeError, KeyError):
        return False


def _scalar(x) -> Any:
    a = np.asarray(x)
    return a.ravel()[0] if a.size else None


def _strings(x) -> List[str]:
    if isinstance(x, str):
        return [x]
    return [str(s) for s in x]


@dataclass
class CellTable:
    """Flattened model: the per-cell rows x n_acc parameter source table.

    ``param_src[c, r, j] >= 0`` is an index into theta; ``< 0`` refers to
    ``const_val[-1 - k]``.  Column j = 0 is the responding accumulator
    (``dmi@node_1_index``), rows follow :data:`LBA_CORE` (type "lba") or
    :data:`DDM_CORE` (type "fastdm").
    """

    n_acc: int
    n_cell: int
    npar: int
    param_src: np.ndarray  # int32 [n_cell, rows, n_acc]
    const_val: np.ndarray  # float64
    posdrift: np.ndarray  # uint8; "lba": [n_acc] is_positive_drift, "fastdm": [n_cell] upper-boundary response
    pnames: List[str]
    cell_names: List[str]
    type: str = "lba"


def build_cell_table(model: Any, node_1_index: Any, is_positive_drift: Any) -> CellTable:
    """model_boolean + node_1_index + constants + pnames -> param_src."""
    mtype = _strings(slot(model, "type"))[0] if _has_slot(model, "type") else "lba"
    if mtype not in CORE_OF_TYPE:
        raise ValueError("Undefined model type")  # @hdr/likelihood.h:312
    core_rows = CORE_OF_TYPE[mtype]
    pxc = _strings(slot(model, "parameter_x_condition_names"))
    pnames = _strings(slot(model, "pnames"))
    cell_names = _strings(slot(model, "cell_names"))
    constants = slot(model, "constants")
    cnames = list(getattr(constants, "names", None) or getattr(constants, "attrs", {}).get("names") or [])
    cvals = np.asarray(constants, dtype=np.float64).ravel()
    mb = np.asarray(slot(model, "model_boolean")).astype(bool)
    n1 = np.asarray(node_1_index).astype(np.int64)
    n_cell, n_pxc, n_acc = mb.shape
    if n1.shape != (n_cell, n_acc):
        raise ValueError(f"node_1_index shape {n1.shape} != ({n_cell}, {n_acc})")
    if len(pxc) != n_pxc or len(cell_names) != n_cell:
        raise ValueError("model_boolean dimensions disagree with names")
    core_of = [name.split(".", 1)[0] for name in pxc]
    for nm in core_of:
        if nm not in core_rows:
            raise ValueError(f"unknown core parameter {nm!r} for model type {mtype!r}")
    src = np.zeros((n_cell, len(core_rows), n_acc), dtype=np.int32)
    for c in range(n_cell):
        for j in range(n_acc):
            acc = int(n1[c, j])
            for r, core in enumerate(core_rows):
                ks = [k for k in range(n_pxc) if core_of[k] == core and mb[c, k, acc]]
                if len(ks) != 1:
                    raise ValueError(f"cell {cell_names[c]} acc {acc}: {len(ks)} sources for {core}")
                name = pxc[ks[0]]
                if name in pnames:
                    src[c, r, j] = pnames.index(name)
                elif name in cnames:
                    src[c, r, j] = -1 - cnames.index(name)
                else:
                    raise ValueError(f"{name} is neither a free parameter nor a constant")
    pd = np.asarray(is_positive_drift).astype(np.uint8).ravel()
    if mtype == "fastdm":  # ddm_likelihood indexes it by cell (@hdr/likelihood.h:142): TRUE = upper-boundary response
        if pd.size != n_cell:
            raise ValueError("is_positive_drift length != number of cells (model type 'fastdm')")
    elif pd.size != n_acc:
        raise ValueError("is_positive_drift length != number of accumulators")
    return CellTable(n_acc, n_cell, len(pnames), src, cvals.copy(), pd, pnames, cell_names, mtype)


@dataclass
class Trials:
    """One subject's data, grouped by ascending model cell index (stable)."""

    rt: np.ndarray  # float64 [n]
    cell: np.ndarray  # uint16 [n]


def flatten_data(data: Any, cell_names: Sequence[str]) -> Trials:
    """`dmi@data` (named list cell_name -> RT vector, empty cells omitted) -> arrays."""
    names = list(getattr(data, "names", None) or [])
    lookup = {n: i for i, n in enumerate(cell_names)}
    rts, cells = [], []
    order = sorted(range(len(names)), key=lambda i: lookup[names[i]])
    for i in order:
        v = np.asarray(data[i], dtype=np.float64).ravel()
        rts.append(v)
        cells.append(np.full(v.size, lookup[names[i]], dtype=np.uint16))
    if not rts:
        return Trials(np.zeros(0), np.zeros(0, dtype=np.uint16))
    return Trials(np.concatenate(rts), np.concatenate(cells))


@dataclass
class PriorTable:
    npar: int
    p0: np.ndarray
    p1: np.ndarray
    lower: np.ndarray
    upper: np.ndarray
    dist: np.ndarray  # int32
    log_p: np.ndarray  # uint8
    pnames: List[str]


def flatten_prior(plist: Any) -> PriorTable:
    """`prior@p_prior` / `@h_prior`: named list of list(p0,p1,lower,upper,dist_id,log_p)."""
    names = list(getattr(plist, "names", None) or [])
    n = len(plist)
    out = PriorTable(n, *(np.zeros(n) for _ in range(4)), np.zeros(n, np.int32), np.zeros(n, np.uint8), names)
    for i in range(n):
        e = plist[i]
        out.p0[i] = _scalar(e["p0"])
        out.p1[i] = _scalar(e["p1"])
        out.lower[i] = _scalar(e["lower"])
        out.upper[i] = _scalar(e["upper"])
        out.dist[i] = int(_scalar(e["dist_id"]))
        out.log_p[i] = 1 if bool(_scalar(e["log_p"])) else 0
    return out
