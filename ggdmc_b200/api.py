"""Python mirror of the reference's `.Call` interface (R/RcppExports.R:7-83, src/de2R.cpp:8-171).

`run_subject(config_r, dmi, samples)`, `run_hyper(config_r, dmi, samples)` and
`run(config_r, dmis, samples)` take objects that expose the reference's S4 slot names -- either
objects decoded from `.rda` files (:func:`ggdmc_b200.rda.read_rda`) or the plain dataclasses below
(`Model`, `DMI`, `Prior`, `ThetaInput`, `DEInput`, `Config`, `Posterior`, same slot names and
meaning as R/model-class.R:36-57, 238-254, 1288-1313, 1467-1485) -- flatten them exactly like the R
glue does (ggdmc_b200/r/ggdmc_b200_glue.cpp) and call the C ABI.  Same argument meaning, same
errors ("Require three or more chains.", "Undefined model type"), same `posterior` slots back.
Nothing here computes a density or a proposal.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib as B
from . import engine as E
from .model import PriorTable, Trials, build_cell_table, flatten_data, flatten_prior, slot


# ---- plain mirrors of the S4 classes -----------------------------------------------------------
class NamedList(list):
    """R named list: `.names` + item access by name (what `prior@p_prior`, `dmi@data` are)."""

    def __init__(self, items: Dict[str, Any]):
        super().__init__(items.values())
        self.names = list(items.keys())

    def __getitem__(self, k):
        if isinstance(k, str):
            return super().__getitem__(self.names.index(k))
        return super().__getitem__(k)


class NamedVector(np.ndarray):
    """R named numeric vector (`model@constants`)."""

    def __new__(cls, values, names):
        o = np.asarray(values, dtype=np.float64).view(cls)
        o.names = list(names)
        return o

    def __array_finalize__(self, obj):
        self.names = getattr(obj, "names", None)


@dataclass
class Model:  # S4 `model` (ggdmcModel::BuildModel)
    parameter_x_condition_names: List[str]
    pnames: List[str]
    cell_names: List[str]
    constants: NamedVector
    model_boolean: np.ndarray  # logical [ncell, n_pxc, n_acc]
    type: str = "lba"
    npar: int = 0


@dataclass
class DMI:  # S4 `dmi` (ggdmcModel::BuildDMI)
    model: Model
    data: Any  # LBA: NamedList cell_name -> RT vector; hyper: matrix nsubject x npar
    node_1_index: Optional[np.ndarray] = None
    is_positive_drift: Optional[np.ndarray] = None


@dataclass
class Prior:  # S4 `prior` (ggdmcPrior::set_priors)
    nparameter: int
    pnames: List[str]
    p_prior: Any
    h_prior: Any = None


@dataclass
class ThetaInput:  # R/model-class.R:36-57
    nmc: int = 500
    nchain: int = 3
    thin: int = 1
    nparameter: int = 0
    pnames: List[str] = field(default_factory=list)
    report_length: int = 100
    max_init_attempts: int = 1000
    is_print: bool = False


@dataclass
class DEInput:  # R/model-class.R:1288-1313
    pop_migration_prob: float = 0.0
    sub_migration_prob: float = 0.0
    gamma_precursor: float = 2.38
    rp: float = 0.001
    is_hblocked: bool = False
    is_pblocked: bool = False
    nparameter: int = 0
    nchain: int = 3
    pop_debug: bool = False
    sub_debug: bool = False


@dataclass
class Config:  # R/model-class.R:1467-1485
    prior: Prior
    theta_input: ThetaInput
    de_input: DEInput
    seed: int = 1
    main_seed: int = 1
    core_id: int = 1


@dataclass
class Posterior:  # R/model-class.R:238-254; arrays shaped like R's: theta [npar, nchain, nmc]
    theta: np.ndarray
    summed_log_prior: np.ndarray  # [nchain, nmc]
    log_likelihoods: np.ndarray  # [nchain, nmc]
    start: int
    npar: int
    pnames: List[str]
    nmc: int
    thin: int
    nchain: int


def prior_list(table: PriorTable) -> NamedList:
    """PriorTable -> the reference's list-of-lists form."""
    return NamedList({n: NamedList({"p0": table.p0[i], "p1": table.p1[i], "lower": table.lower[i], "upper": table.upper[i],
                                    "dist_id": float(table.dist[i]), "log_p": bool(table.log_p[i])})
                      for i, n in enumerate(table.pnames)})


# ---- helpers -------------------------------------------------------------------------------------
def _scalar(x):
    a = np.asarray(x)
    return a.ravel()[0]


def _strs(x) -> List[str]:
    return [str(s) for s in x] if not isinstance(x, str) else [x]


def _tuning(config_r, schedule, seeds, extra_nchain=None) -> E.Tuning:
    ti, de = slot(config_r, "theta_input"), slot(config_r, "de_input")
    seeds = list(seeds) if seeds is not None else [int(_scalar(slot(config_r, "seed")))]
    is_print = bool(_scalar(slot(ti, "is_print")))
    return E.Tuning(nmc=int(_scalar(slot(ti, "nmc"))), nchain=int(_scalar(slot(ti, "nchain"))), thin=int(_scalar(slot(ti, "thin"))),
                    nparameter=int(_scalar(slot(de, "nparameter"))), pop_migration_prob=float(_scalar(slot(de, "pop_migration_prob"))),
                    sub_migration_prob=float(_scalar(slot(de, "sub_migration_prob"))),
                    gamma_precursor=float(_scalar(slot(de, "gamma_precursor"))), rp=float(_scalar(slot(de, "rp"))),
                    is_hblocked=bool(_scalar(slot(de, "is_hblocked"))), is_pblocked=bool(_scalar(slot(de, "is_pblocked"))),
                    report_length=int(_scalar(slot(ti, "report_length"))) if is_print else 0, schedule=schedule, seeds=seeds)


def _last_valid_slice(samples) -> int:
    """Slice to continue from: the last one whose thetas are all finite (a fresh initialise_* object
    has only slice 1 filled; a finished fit has all of them) -- R/phi.R:77-80, R/sampling.R:426-429."""
    th = np.asarray(slot(samples, "theta"))
    ok = np.all(np.isfinite(th), axis=(0, 1))
    idx = np.nonzero(ok)[0]
    if idx.size == 0:
        raise ValueError("samples has no slice with finite thetas")
    return int(idx[-1])


def _start(samples_list: Sequence[Any]) -> E.PopState:
    """posterior objects of the replicates -> PopState [R, C, D]."""
    th, lp, ll = [], [], []
    for s in samples_list:
        k = _last_valid_slice(s)
        th.append(np.ascontiguousarray(np.asarray(slot(s, "theta"))[:, :, k].T))
        lp.append(np.asarray(slot(s, "summed_log_prior"))[:, k].copy())
        ll.append(np.asarray(slot(s, "log_likelihoods"))[:, k].copy())
    return E.PopState(np.stack(th), np.stack(lp), np.stack(ll))


def _posterior(out: E.PopSamples, r: int, pnames: List[str], thin: int) -> Posterior:
    nmc, nchain, npar = out.theta.shape[1:]
    return Posterior(theta=np.transpose(out.theta[r], (2, 1, 0)), summed_log_prior=out.lp[r].T, log_likelihoods=out.ll[r].T, start=1,
                     npar=npar, pnames=list(pnames), nmc=nmc, thin=thin, nchain=nchain)


def _flatten_dmi(dmi):
    model = slot(dmi, "model")
    mtype = _strs(slot(model, "type"))[0]
    if mtype not in B.MODEL_TYPES:  # "lba" -> lba_likelihood, "fastdm" -> ddm_likelihood (@hdr/likelihood.h:279-305)
        raise B.GgdmcError(B.ERR_ARG, "Undefined model type")  # @hdr/likelihood.h:312
    ct = build_cell_table(model, slot(dmi, "node_1_index"), slot(dmi, "is_positive_drift"))
    return ct, flatten_data(slot(dmi, "data"), ct.cell_names)


def _as_list(x) -> list:
    return list(x) if isinstance(x, (list, tuple)) and not hasattr(x, "attrs") else [x]


# ---- the three entry points ---------------------------------------------------------------------
def run_subject(config_r, dmi, samples, schedule: int = B.SCHEDULE_PARALLEL, progress=None) -> Union[Posterior, List[Posterior]]:
    """`run_subject` (src/de2R.cpp:8-23).  `config_r` / `samples` may be lists of equal length: the
    replicates the reference would fork (R/sampling.R:26-55) run as one batched call."""
    configs, samp = _as_list(config_r), _as_list(samples)
    if len(configs) != len(samp):
        raise ValueError("one samples object per config")
    ct, trials = _flatten_dmi(dmi)
    prior = slot(configs[0], "prior")
    tun = _tuning(configs[0], schedule, [int(_scalar(slot(c, "seed"))) for c in configs])
    out = E.run_subject(ct, trials, flatten_prior(slot(prior, "p_prior")), tun, _start(samp), progress)
    res = [_posterior(out, r, ct.pnames, tun.thin) for r in range(len(configs))]
    return res if isinstance(config_r, (list, tuple)) else res[0]


def run_hyper(config_r, dmi, samples, schedule: int = B.SCHEDULE_PARALLEL, progress=None) -> Union[Posterior, List[Posterior]]:
    """`run_hyper` (src/de2R.cpp:30-47): dmi@data is the nsubject x npar matrix of subject thetas."""
    configs, samp = _as_list(config_r), _as_list(samples)
    prior = slot(configs[0], "prior")
    tun = _tuning(configs[0], schedule, [int(_scalar(slot(c, "seed"))) for c in configs])
    data = np.asarray(slot(dmi, "data"), dtype=np.float64)
    out = E.run_hyper(flatten_prior(slot(prior, "p_prior")), flatten_prior(slot(prior, "h_prior")), data, tun, _start(samp), progress)
    pnames = _strs(slot(slot(configs[0], "theta_input"), "pnames"))
    res = [_posterior(out, r, pnames, tun.thin) for r in range(len(configs))]
    return res if isinstance(config_r, (list, tuple)) else res[0]


def run(config_r, dmis, samples, schedule: int = B.SCHEDULE_PARALLEL, progress=None):
    """`run` (src/de2R.cpp:123-171): returns {"phi": posterior, "subject_theta": [posterior, ...]}
    (a list of such dicts when `config_r` / `samples` are lists of replicates)."""
    configs, samp = _as_list(config_r), (list(samples) if isinstance(samples, (list, tuple)) else [samples])
    if len(configs) != len(samp):
        raise ValueError("one samples object per config")
    flat = [_flatten_dmi(d) for d in dmis]
    ct = flat[0][0]
    for c, _ in flat[1:]:
        if not np.array_equal(c.param_src, ct.param_src):
            raise B.GgdmcError(B.ERR_ARG, "all subjects of one call must share one model")
    prior = slot(configs[0], "prior")
    tun = _tuning(configs[0], schedule, [int(_scalar(slot(c, "seed"))) for c in configs])
    S = len(dmis)
    phi_start = _start([s["phi"] for s in samp])
    subj_start = [_start([s["subject_theta"][i] for s in samp]) for i in range(S)]
    phi_out, subj_out = E.run_hier(ct, [t for _, t in flat], flatten_prior(slot(prior, "p_prior")), flatten_prior(slot(prior, "h_prior")),
                                   tun, phi_start, subj_start, progress)
    phi_names = _strs(slot(slot(configs[0], "theta_input"), "pnames"))
    res = [{"phi": _posterior(phi_out, r, phi_names, tun.thin),
            "subject_theta": [_posterior(subj_out[i], r, ct.pnames, tun.thin) for i in range(S)]} for r in range(len(configs))]
    return res if isinstance(config_r, (list, tuple)) else res[0]
