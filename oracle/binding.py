"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/libggdmc_oracle.so (the C restatement)
and oracle/_ref/libggdmc_ref.so (the reference's own object code, when it was built)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_up = C.POINTER(C.c_uint)
c_u8p = C.POINTER(C.c_ubyte)
c_u16p = C.POINTER(C.c_ushort)


class Addr(C.Structure):
    _fields_ = [("pop", C.c_uint), ("iter", C.c_uint), ("sweep", C.c_uint), ("chain", C.c_uint),
                ("purpose", C.c_uint), ("slot", C.c_uint)]


class Rng(C.Structure):
    _fields_ = [("mode", C.c_int), ("u", c_dp), ("n", C.c_long), ("pos", C.c_long), ("seed", C.c_ulonglong),
                ("burn_static_ctor", C.c_int), ("first_like_done", C.c_int), ("rec", c_dp), ("rec_n", C.c_long), ("rec_cap", C.c_long)]


class Model(C.Structure):
    _fields_ = [("n_acc", C.c_int), ("n_cell", C.c_int), ("npar", C.c_int), ("param_src", c_ip),
                ("const_val", c_dp), ("posdrift", c_u8p), ("type", C.c_int)]


MODEL_LBA, MODEL_DDM = 0, 1


class Data(C.Structure):
    _fields_ = [("n_trial", C.c_int), ("rt", c_dp), ("cell", c_u16p)]


class Prior(C.Structure):
    _fields_ = [("npar", C.c_int), ("p0", c_dp), ("p1", c_dp), ("lower", c_dp), ("upper", c_dp), ("dist", c_ip),
                ("log_p", c_u8p)]


class DE(C.Structure):
    _fields_ = [("pop_migration_prob", C.c_double), ("sub_migration_prob", C.c_double), ("gamma_precursor", C.c_double),
                ("rp", C.c_double), ("is_hblocked", C.c_int), ("is_pblocked", C.c_int), ("nparameter", C.c_int),
                ("nchain", C.c_int), ("jacobi", C.c_int)]


class Pop(C.Structure):
    _fields_ = [("npar", C.c_int), ("nchain", C.c_int), ("nmc", C.c_int), ("thin", C.c_int), ("theta", c_dp),
                ("lp", c_dp), ("ll", c_dp), ("out_theta", c_dp), ("out_lp", c_dp), ("out_ll", c_dp), ("store_i", C.c_int)]


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is mounted)."""
    so = os.path.join(HERE, "libggdmc_oracle.so")
    srcs = [os.path.join(HERE, f) for f in ("ggdmc_oracle.c", "rmath_port.c", "ggdmc_oracle.h", "rmath_port.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", HERE, "libggdmc_oracle.so"], check=True, capture_output=True)
    ref = os.path.join(HERE, "_ref", "libggdmc_ref.so")
    ref_srcs = [os.path.join(HERE, f) for f in ("ref_harness.cpp", "ref_harness2.cpp", "ref_shim.c", "rmath_port.c")]
    ref_stale = force or not os.path.exists(ref) or any(os.path.getmtime(s) > os.path.getmtime(ref) for s in ref_srcs)
    if os.path.exists("/root/reference/src/de.o") and ref_stale:
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "libggdmc_oracle.so"))
        L.orc_uniform.restype = C.c_double
        L.orc_sumloglike.restype = C.c_double
        L.orc_sumloglike_rinit.restype = C.c_double
        L.orc_tnorm_d.restype = C.c_double
        L.orc_tnorm_d.argtypes = [C.c_double] * 5 + [C.c_int]
        L.orc_sumlogprior.restype = C.c_double
        L.orc_sumloghlike.restype = C.c_double
        L.orc_time_sumloglike.restype = C.c_double
        L.orc_get_subchains.restype = C.c_int
        L.orc_pnorm5.restype = C.c_double
        L.orc_pnorm5.argtypes = [C.c_double] * 3 + [C.c_int] * 2
        L.orc_dnorm4.restype = C.c_double
        L.orc_dnorm4.argtypes = [C.c_double] * 3 + [C.c_int]
        _lib = L
    return _lib


_ref = None


def ref_lib() -> Optional[C.CDLL]:
    """The reference's own object code (src/de.o) behind C wrappers, or None if never built."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libggdmc_ref.so")
        if os.path.exists("/root/reference/src/de.o"):
            build()  # (re)link the harness where the reference is mounted; elsewhere the prebuilt file is used as it is
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_get_subchains.restype = C.c_uint
        L.ref_tnorm_d.restype = C.c_double
        L.ref_tnorm_d.argtypes = [C.c_double] * 5 + [C.c_int]
        L.ref_uniform_stream_pos.restype = C.c_long
        _ref = L
    return _ref


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a: np.ndarray, t=c_dp):
    return a.ctypes.data_as(t)


class OModel:
    """Keeps the numpy buffers alive next to the C struct."""

    def __init__(self, param_src, const_val, posdrift, npar, type=MODEL_LBA):
        self.param_src = np.ascontiguousarray(param_src, dtype=np.int32)
        self.const_val = f64(const_val if len(const_val) else [0.0])
        self.posdrift = np.ascontiguousarray(posdrift, dtype=np.uint8)
        n_cell, rows, n_acc = self.param_src.shape
        assert rows == (10 if type == MODEL_DDM else 6)
        assert self.posdrift.size == (n_cell if type == MODEL_DDM else n_acc)
        self.c = Model(n_acc, n_cell, int(npar), ptr(self.param_src, c_ip), ptr(self.const_val), ptr(self.posdrift, c_u8p), int(type))
        self.n_acc, self.n_cell, self.npar, self.type = n_acc, n_cell, int(npar), int(type)


class OData:
    def __init__(self, rt, cell):
        order = np.argsort(np.asarray(cell), kind="stable")
        self.order = order
        self.rt = f64(np.asarray(rt)[order])
        self.cell = np.ascontiguousarray(np.asarray(cell)[order], dtype=np.uint16)
        self.c = Data(len(self.rt), ptr(self.rt), ptr(self.cell, c_u16p))


class OPrior:
    def __init__(self, p0, p1, lower, upper, dist, log_p):
        self.p0, self.p1, self.lower, self.upper = f64(p0), f64(p1), f64(lower), f64(upper)
        self.dist = np.ascontiguousarray(dist, dtype=np.int32)
        self.log_p = np.ascontiguousarray(log_p, dtype=np.uint8)
        self.c = Prior(len(self.p0), ptr(self.p0), ptr(self.p1), ptr(self.lower), ptr(self.upper), ptr(self.dist, c_ip),
                       ptr(self.log_p, c_u8p))


class OPop:
    """A chain population with sample storage, laid out like the reference's arma objects."""

    def __init__(self, theta0, lp0, ll0, nmc, thin):
        theta0 = f64(theta0)  # [nchain, npar]
        self.nchain, self.npar = theta0.shape
        self.theta = theta0.copy()
        self.lp, self.ll = f64(lp0).copy(), f64(ll0).copy()
        self.out_theta = np.full((nmc, self.nchain, self.npar), np.nan)
        self.out_lp = np.full((nmc, self.nchain), -np.inf)
        self.out_ll = np.full((nmc, self.nchain), -np.inf)
        self.out_theta[0], self.out_lp[0], self.out_ll[0] = self.theta, self.lp, self.ll
        self.c = Pop(self.npar, self.nchain, nmc, thin, ptr(self.theta), ptr(self.lp), ptr(self.ll), ptr(self.out_theta),
                     ptr(self.out_lp), ptr(self.out_ll), 0)


def make_rng(seed: Optional[int] = None, stream: Optional[np.ndarray] = None, burn: bool = False, record: int = 0):
    """record > 0: keep the first `record` uniforms handed out, in order (r.recorded())."""
    r = Rng()
    if record:
        buf = np.zeros(record)
        r.rec, r.rec_n, r.rec_cap = ptr(buf), 0, record
        r._rec = buf
    if stream is not None:
        s = f64(stream)
        r.mode, r.u, r.n, r.pos = 0, ptr(s), len(s), 0
        r._keep = s
    else:
        r.mode, r.seed = 1, int(seed)
    r.burn_static_ctor = 1 if burn else 0
    r.first_like_done = 0
    return r


def recorded(r: Rng) -> np.ndarray:
    return r._rec[: r.rec_n].copy()


def make_de(nparameter, nchain, pop_migration_prob=0.0, sub_migration_prob=0.0, gamma_precursor=2.38, rp=0.001,
            is_hblocked=False, is_pblocked=False, jacobi=0) -> DE:
    return DE(pop_migration_prob, sub_migration_prob, gamma_precursor, rp, int(is_hblocked), int(is_pblocked),
              int(nparameter), int(nchain), int(jacobi))


def trial_logdens(m: OModel, d: OData, theta) -> np.ndarray:
    """Per-trial log densities, returned in the CALLER's original trial order."""
    th = f64(theta)
    out = np.zeros(len(d.rt))
    lib().orc_trial_logdens(C.byref(m.c), C.byref(d.c), ptr(th), ptr(out))
    back = np.empty_like(out)
    back[d.order] = out
    return back


def sumloglike(m: OModel, d: OData, theta) -> float:
    th = f64(theta)
    return lib().orc_sumloglike(C.byref(m.c), C.byref(d.c), ptr(th), None, None)


def sumloglike_rinit(m: OModel, d: OData, theta) -> float:
    th = f64(theta)
    return lib().orc_sumloglike_rinit(C.byref(m.c), C.byref(d.c), ptr(th))


def sumlogprior(p: OPrior, x, p0=None, p1=None) -> float:
    xx = f64(x)
    a = f64(p0) if p0 is not None else None
    b = f64(p1) if p1 is not None else None
    return lib().orc_sumlogprior(C.byref(p.c), ptr(a) if a is not None else None, ptr(b) if b is not None else None, ptr(xx))


def run_subject(de: DE, pop: OPop, prior: OPrior, m: OModel, d: OData, rng: Rng, pop_id: int, n_iter: int) -> None:
    lib().orc_run_subject(C.byref(de), C.byref(pop.c), C.byref(prior.c), C.byref(m.c), C.byref(d.c), C.byref(rng), C.c_uint(pop_id),
                          C.c_uint(n_iter))


def run_hyper(de: DE, phi: OPop, p_prior: OPrior, h_prior: OPrior, data_theta, rng: Rng, n_iter: int) -> None:
    x = f64(data_theta)
    lib().orc_run_hyper(C.byref(de), C.byref(phi.c), C.byref(p_prior.c), C.byref(h_prior.c), ptr(x), int(x.shape[0]), C.byref(rng),
                        C.c_uint(n_iter))


def run_hier(de: DE, phi: OPop, subj: Sequence[OPop], p_prior: OPrior, h_prior: OPrior, m: OModel, datas: Sequence[OData],
             rng: Rng, n_iter: int, first_subject_id: int = 0) -> None:
    S = len(subj)
    pops = (Pop * S)(*[s.c for s in subj])
    ds = (Data * S)(*[d.c for d in datas])
    lib().orc_run_hier(C.byref(de), C.byref(phi.c), pops, S, C.byref(p_prior.c), C.byref(h_prior.c), C.byref(m.c), ds,
                       C.byref(rng), C.c_uint(n_iter), C.c_uint(first_subject_id))
    for i, s in enumerate(subj):  # store_i lives in the struct copies
        s.c.store_i = pops[i].store_i


def time_sumloglike(m: OModel, d: OData, thetas, reps: int) -> float:
    th = f64(thetas)
    return lib().orc_time_sumloglike(C.byref(m.c), C.byref(d.c), ptr(th), int(th.shape[0]), int(reps))


# ---- tier-2: the reference's own sampler object code (oracle/ref_harness2.cpp) ---------------------
def _model_args(m: OModel):
    ref_lib().ref2_set_model_type(int(m.type))  # "lba" or "fastdm" objects (oracle/ref_harness2.cpp)
    return [m.n_acc, m.n_cell, ptr(m.param_src, c_ip), ptr(m.const_val), ptr(m.posdrift, c_u8p)]


def _prior_args(p: OPrior):
    return [ptr(p.p0), ptr(p.p1), ptr(p.lower), ptr(p.upper), ptr(p.dist, c_ip), ptr(p.log_p, c_u8p)]


def ddm_cell(P, is_upper: bool, rt) -> tuple:
    """orc_ddm_cell: densities of one DDM cell (P = a, d, precision, s, st0, sv, sz, t0, v, z); returns (valid, dens)."""
    P, rt = f64(P), f64(rt)
    out = np.zeros(len(rt))
    ok = lib().orc_ddm_cell(ptr(P), int(bool(is_upper)), ptr(rt), len(rt), ptr(out))
    return bool(ok), out


def ref2_ddm_density(m: OModel, d: OData, theta) -> np.ndarray:
    """likelihood_class::ddm_likelihood of src/de.o: density of every trial of d (cell-grouped order = d's order)."""
    th = f64(theta)
    out = np.zeros(len(d.rt))
    ref_lib().ref2_ddm_density(m.n_acc, m.n_cell, ptr(m.param_src, c_ip), ptr(m.const_val), ptr(m.posdrift, c_u8p), ptr(d.rt),
                               ptr(d.cell, c_u16p), len(d.rt), ptr(th), len(th), ptr(out))
    return out


def ref2_prime() -> None:
    ref_lib().ref2_prime()


def ref2_set_stream(u: np.ndarray) -> np.ndarray:
    u = f64(u)
    ref_lib().ref_set_uniform_stream(ptr(u), len(u))
    return u


def ref2_sweep_subject(kind: int, para_idx: int, nparameter: int, m: OModel, d: OData, prior: OPrior, theta, lp, ll, gamma_precursor=2.38,
                       rp=0.001):
    """de_class::crossover (kind 0) / migration (kind 1) of src/de.o on chains theta [nchain, npar]."""
    th, a, b = f64(theta).copy(), f64(lp).copy(), f64(ll).copy()
    nchain, npar = th.shape
    ref_lib().ref2_sweep_subject(kind, para_idx, nchain, nparameter, C.c_double(gamma_precursor), C.c_double(rp), *_model_args(m),
                                 ptr(d.rt), ptr(d.cell, c_u16p), len(d.rt), npar, *_prior_args(prior), ptr(th), ptr(a), ptr(b))
    return th, a, b


def ref2_run_chains(nparameter: int, m: OModel, d: OData, prior: OPrior, theta0, lp0, ll0, nmc: int, thin: int, sub_migration_prob=0.0,
                    is_pblocked=False, gamma_precursor=2.38, rp=0.001):
    """de_class::run_chains of src/de.o; returns (theta [nmc, nchain, npar], lp, ll [nmc, nchain])."""
    th0 = f64(theta0)
    nchain, npar = th0.shape
    ot, olp, oll = np.zeros((nmc, nchain, npar)), np.zeros((nmc, nchain)), np.zeros((nmc, nchain))
    ref_lib().ref2_run_chains(nchain, nparameter, C.c_double(sub_migration_prob), C.c_double(gamma_precursor), C.c_double(rp),
                              int(is_pblocked), nmc, thin, *_model_args(m), ptr(d.rt), ptr(d.cell, c_u16p), len(d.rt), npar,
                              *_prior_args(prior), ptr(th0), ptr(f64(lp0)), ptr(f64(ll0)), ptr(ot), ptr(olp), ptr(oll))
    return ot, olp, oll


def ref2_run_hchains(nparameter: int, m: OModel, datas: Sequence[OData], p_prior: OPrior, h_prior: OPrior, phi_start, subj_starts, nmc: int,
                     thin: int, pop_migration_prob=0.0, sub_migration_prob=0.0, is_hblocked=False, is_pblocked=False,
                     gamma_precursor=2.38, rp=0.001):
    """de_class::run_hchains of src/de.o; returns (phi (theta, lp, ll), [subject (theta, lp, ll)])."""
    S = len(datas)
    off = np.zeros(S + 1, dtype=np.int64)
    for s in range(S):
        off[s + 1] = off[s] + len(datas[s].rt)
    rt = f64(np.concatenate([x.rt for x in datas]))
    cell = np.ascontiguousarray(np.concatenate([x.cell for x in datas]), dtype=np.uint16)
    phi0, plp0, pll0 = f64(phi_start[0]), f64(phi_start[1]), f64(phi_start[2])
    nchain, npar = phi0.shape[0], phi0.shape[1] // 2
    st0 = f64(np.stack([s[0] for s in subj_starts]))
    slp0, sll0 = f64(np.stack([s[1] for s in subj_starts])), f64(np.stack([s[2] for s in subj_starts]))
    pot, polp, poll = np.zeros((nmc, nchain, 2 * npar)), np.zeros((nmc, nchain)), np.zeros((nmc, nchain))
    sot, solp, soll = np.zeros((S, nmc, nchain, npar)), np.zeros((S, nmc, nchain)), np.zeros((S, nmc, nchain))
    ref_lib().ref2_run_hchains(nchain, nparameter, C.c_double(pop_migration_prob), C.c_double(sub_migration_prob),
                               C.c_double(gamma_precursor), C.c_double(rp), int(is_hblocked), int(is_pblocked), nmc, thin, S,
                               *_model_args(m), off.ctypes.data_as(C.POINTER(C.c_longlong)), ptr(rt), ptr(cell, c_u16p), npar,
                               *_prior_args(p_prior), *_prior_args(h_prior), ptr(phi0), ptr(plp0), ptr(pll0), ptr(st0), ptr(slp0),
                               ptr(sll0), ptr(pot), ptr(polp), ptr(poll), ptr(sot), ptr(solp), ptr(soll))
    return (pot, polp, poll), [(sot[s], solp[s], soll[s]) for s in range(S)]
