"""Array-level Python binding of the C ABI (include/ggdmc_b200.h).

Everything here is a thin marshaller: numpy arrays in, one C call, numpy arrays out.  The
computation happens in libggdmc_b200.so on the GPU; nothing here computes densities or proposals.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as B
from .model import CellTable, PriorTable, Trials


class _Keep:
    """C struct + the numpy buffers it points into."""

    def __init__(self, c, *bufs):
        self.c = c
        self.bufs = bufs


def _model(ct: CellTable) -> _Keep:
    ps = np.ascontiguousarray(ct.param_src, dtype=np.int32)
    cv = B.f64(ct.const_val if len(ct.const_val) else [0.0])
    pd = np.ascontiguousarray(ct.posdrift, dtype=np.uint8)
    mtype = B.MODEL_TYPES[getattr(ct, "type", "lba")]
    m = B.ModelT(ct.n_acc, ct.n_cell, ct.npar, len(ct.const_val), B.ptr(ps, B.c_i32p), B.ptr(cv), B.ptr(pd, B.c_u8p), mtype)
    return _Keep(m, ps, cv, pd)


class TrialsStack:
    """All subjects' trials already concatenated (rt, cell, offsets): skips S small concatenations per call."""

    def __init__(self, subjects: Sequence[Trials]):
        self.n = len(subjects)
        self.off = np.zeros(self.n + 1, dtype=np.int64)
        self.off[1:] = np.cumsum([len(t.rt) for t in subjects])
        self.rt = B.f64(np.concatenate([t.rt for t in subjects]))
        self.cell = np.ascontiguousarray(np.concatenate([t.cell for t in subjects]), dtype=np.uint16)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return Trials(self.rt[self.off[i]:self.off[i + 1]], self.cell[self.off[i]:self.off[i + 1]])

    def __iter__(self):
        return (self[i] for i in range(self.n))


def _trials(subjects) -> _Keep:
    if isinstance(subjects, TrialsStack):
        t = B.TrialsT(subjects.n, B.ptr(subjects.off, B.c_i64p), B.ptr(subjects.rt), B.ptr(subjects.cell, B.c_u16p))
        return _Keep(t, subjects)
    off = np.zeros(len(subjects) + 1, dtype=np.int64)
    for i, t in enumerate(subjects):
        off[i + 1] = off[i] + len(t.rt)
    rt = B.f64(np.concatenate([t.rt for t in subjects]) if len(subjects) else np.zeros(0))
    cell = np.ascontiguousarray(np.concatenate([t.cell for t in subjects]), dtype=np.uint16)
    if rt.size == 0:
        rt, cell = np.zeros(1), np.zeros(1, dtype=np.uint16)
    t = B.TrialsT(len(subjects), B.ptr(off, B.c_i64p), B.ptr(rt), B.ptr(cell, B.c_u16p))
    return _Keep(t, off, rt, cell)


def _prior(p: PriorTable) -> _Keep:
    a = [B.f64(p.p0), B.f64(p.p1), B.f64(p.lower), B.f64(p.upper)]
    dist = np.ascontiguousarray(p.dist, dtype=np.int32)
    lg = np.ascontiguousarray(p.log_p, dtype=np.uint8)
    c = B.PriorT(p.npar, B.ptr(a[0]), B.ptr(a[1]), B.ptr(a[2]), B.ptr(a[3]), B.ptr(dist, B.c_i32p), B.ptr(lg, B.c_u8p))
    return _Keep(c, *a, dist, lg)


@dataclass
class Tuning:
    """theta_input + de_input + seeds, i.e. everything `config` carries besides the prior."""

    nmc: int
    nchain: int
    thin: int = 1
    nparameter: int = 0  # de_input@nparameter (sets gamma)
    pop_migration_prob: float = 0.0
    sub_migration_prob: float = 0.0
    gamma_precursor: float = 2.38
    rp: float = 0.001
    is_hblocked: bool = False
    is_pblocked: bool = False
    report_length: int = 0
    schedule: int = B.SCHEDULE_PARALLEL
    seeds: Sequence[int] = (1,)
    device: int = -1
    subject_begin: int = 0
    n_subject_total: int = 0


def _config(t: Tuning) -> _Keep:
    seeds = np.ascontiguousarray(np.asarray(t.seeds, dtype=np.uint64))
    c = B.ConfigT(int(t.nmc), int(t.nchain), int(t.thin), int(t.report_length), float(t.pop_migration_prob),
                  float(t.sub_migration_prob), float(t.gamma_precursor), float(t.rp), int(bool(t.is_hblocked)),
                  int(bool(t.is_pblocked)), int(t.nparameter), int(t.schedule), len(seeds), int(t.device),
                  B.ptr(seeds, B.c_u64p), int(t.subject_begin), int(t.n_subject_total))
    return _Keep(c, seeds)


@dataclass
class PopState:
    """Start state of one population for R replicates: theta [R, C, D], lp / ll [R, C]."""

    theta: np.ndarray
    lp: np.ndarray
    ll: np.ndarray

    def __post_init__(self):
        self.theta = B.f64(self.theta)
        self.lp = B.f64(self.lp)
        self.ll = B.f64(self.ll)
        if self.theta.ndim == 2:
            self.theta, self.lp, self.ll = self.theta[None], self.lp[None], self.ll[None]

    def c(self) -> B.StartT:
        return B.StartT(B.ptr(self.theta), B.ptr(self.lp), B.ptr(self.ll))


@dataclass
class PopSamples:
    """Samples of one population: theta [R, nmc, C, D], lp / ll [R, nmc, C] (slot 0 = start)."""

    theta: np.ndarray
    lp: np.ndarray
    ll: np.ndarray

    @staticmethod
    def empty(R, nmc, C_, D) -> "PopSamples":
        return PopSamples(np.empty((R, nmc, C_, D)), np.empty((R, nmc, C_)), np.empty((R, nmc, C_)))

    def c(self) -> B.SamplesT:
        R, nmc, C_, D = self.theta.shape
        return B.SamplesT(D, C_, nmc, B.ptr(self.theta), B.ptr(self.lp), B.ptr(self.ll))


def _addr(a: np.ndarray) -> int:
    return a.__array_interface__["data"][0]


class PopStateStack:
    """Start states of S populations held in three stacked arrays (theta [S, R, C, D], lp / ll [S, R, C]).
    Behaves like a list of PopState; the engine gets all S address triples without touching S objects."""

    def __init__(self, theta, lp, ll):
        self.theta, self.lp, self.ll = B.f64(theta), B.f64(lp), B.f64(ll)
        assert self.theta.ndim == 4 and self.lp.shape == self.theta.shape[:3] == self.ll.shape

    def __len__(self):
        return self.theta.shape[0]

    def __getitem__(self, i):
        return PopState(self.theta[i], self.lp[i], self.ll[i])

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class HierOutputs(list):
    """List of per-subject PopSamples that are views of three big arrays (kept in .big)."""

    big = None


def _start_array(starts):
    """[S] ggdmc_start_t as one numpy block of addresses (ctypes struct construction is ~10 us each)."""
    S = len(starts)
    tab = np.empty((S, 3), dtype=np.uint64)
    if isinstance(starts, PopStateStack):
        idx = np.arange(S, dtype=np.uint64)
        for j, a in enumerate((starts.theta, starts.lp, starts.ll)):
            tab[:, j] = np.uint64(_addr(a)) + idx * np.uint64(a.strides[0])
    else:
        for i, s in enumerate(starts):
            tab[i, 0], tab[i, 1], tab[i, 2] = _addr(s.theta), _addr(s.lp), _addr(s.ll)
    return (tab, starts), tab.ctypes.data_as(C.POINTER(B.StartT))


_SAMPLES_DT = np.dtype([("npar", "<i4"), ("nchain", "<i4"), ("nmc", "<i4"), ("_pad", "<i4"), ("theta", "<u8"), ("lp", "<u8"), ("ll", "<u8")])
assert _SAMPLES_DT.itemsize == C.sizeof(B.SamplesT)


def _samples_array(outs):
    S = len(outs)
    tab = np.zeros(S, dtype=_SAMPLES_DT)
    big = getattr(outs, "big", None)
    if big is not None:
        _, _, nmc, nchain, npar = big[0].shape
        idx = np.arange(S, dtype=np.uint64)
        tab["npar"], tab["nchain"], tab["nmc"] = npar, nchain, nmc
        for name, a in zip(("theta", "lp", "ll"), big):
            tab[name] = np.uint64(_addr(a)) + idx * np.uint64(a.strides[0])
    else:
        for i, o in enumerate(outs):
            _, nmc, nchain, npar = o.theta.shape
            tab[i] = (npar, nchain, nmc, 0, _addr(o.theta), _addr(o.lp), _addr(o.ll))
    return (tab, outs), tab.ctypes.data_as(C.POINTER(B.SamplesT))


def _progress(cb):
    if cb is None:
        return C.cast(None, B.PROGRESS_FN)
    return B.PROGRESS_FN(lambda i, _u: cb(int(i)))


def run_subject(ct: CellTable, trials: Trials, p_prior: PriorTable, tuning: Tuning, start: PopState, progress=None) -> PopSamples:
    m, t, p, cfg = _model(ct), _trials([trials]), _prior(p_prior), _config(tuning)
    R = len(tuning.seeds)
    out = PopSamples.empty(R, tuning.nmc, tuning.nchain, ct.npar)
    oc, sc, err, cb = out.c(), start.c(), B.errbuf(), _progress(progress)
    rc = B.lib().ggdmc_b200_run_subject(C.byref(m.c), C.byref(t.c), C.byref(p.c), C.byref(cfg.c), C.byref(sc), C.byref(oc),
                                        cb, None, err)
    B.check(rc, err)
    return out


def run_hyper(p_prior: PriorTable, h_prior: PriorTable, data_theta: np.ndarray, tuning: Tuning, start: PopState,
              progress=None) -> PopSamples:
    p, h, cfg = _prior(p_prior), _prior(h_prior), _config(tuning)
    x = B.f64(data_theta)
    R = len(tuning.seeds)
    out = PopSamples.empty(R, tuning.nmc, tuning.nchain, h_prior.npar)
    oc, sc, err, cb = out.c(), start.c(), B.errbuf(), _progress(progress)
    rc = B.lib().ggdmc_b200_run_hyper(C.byref(p.c), C.byref(h.c), B.ptr(x), int(x.shape[0]), C.byref(cfg.c), C.byref(sc),
                                      C.byref(oc), cb, None, err)
    B.check(rc, err)
    return out


def alloc_hier_outputs(S: int, R: int, nmc: int, nchain: int, npar: int, touch: bool = False):
    """Output arrays for :func:`run_hier`: (phi PopSamples, [subject PopSamples]).  All subjects live in
    one allocation, so the engine returns them with a single device->host copy per array."""
    mk = np.zeros if touch else np.empty
    phi_out = PopSamples(mk((R, nmc, nchain, 2 * npar)), mk((R, nmc, nchain)), mk((R, nmc, nchain)))
    big_t, big_lp, big_ll = mk((S, R, nmc, nchain, npar)), mk((S, R, nmc, nchain)), mk((S, R, nmc, nchain))
    if touch:
        for a in (big_t, big_lp, big_ll):
            a.fill(0.0)
    subj = HierOutputs(PopSamples(big_t[s], big_lp[s], big_ll[s]) for s in range(S))
    subj.big = (big_t, big_lp, big_ll)
    return phi_out, subj


def run_hier(ct: CellTable, trials: Sequence[Trials], p_prior: PriorTable, h_prior: PriorTable, tuning: Tuning,
             phi_start: PopState, subj_start: Sequence[PopState], progress=None, out=None):
    """`run` of the reference: returns (phi PopSamples, [subject PopSamples]).  `out` may carry
    preallocated outputs from :func:`alloc_hier_outputs`."""
    import os, time
    t0 = time.perf_counter()
    m, t, p, h, cfg = _model(ct), _trials(trials), _prior(p_prior), _prior(h_prior), _config(tuning)
    R, S = len(tuning.seeds), len(trials)
    phi_out, subj_out = out if out is not None else alloc_hier_outputs(S, R, tuning.nmc, tuning.nchain, ct.npar)
    _keep_s, starts = _start_array(subj_start)
    _keep_o, outs = _samples_array(subj_out)
    pc, poc, err, cb = phi_start.c(), phi_out.c(), B.errbuf(), _progress(progress)
    t1 = time.perf_counter()
    rc = B.lib().ggdmc_b200_run(C.byref(m.c), C.byref(t.c), C.byref(p.c), C.byref(h.c), C.byref(cfg.c), C.byref(pc), starts,
                                C.byref(poc), outs, cb, None, err)
    B.check(rc, err)
    if os.environ.get("GGDMC_B200_TIMING"):
        print(f"[ggdmc_b200] python marshal {1e3 * (t1 - t0):.3f} ms, C call {1e3 * (time.perf_counter() - t1):.3f} ms", file=__import__("sys").stderr)
    return phi_out, subj_out


def trial_logdens(ct: CellTable, trials: Trials, theta: np.ndarray) -> np.ndarray:
    """log n1PDF of every trial of one subject for each row of theta -> [n_theta, n_trial]."""
    m, t = _model(ct), _trials([trials])
    th = B.f64(np.atleast_2d(theta))
    out = np.empty((th.shape[0], len(trials.rt)))
    err = B.errbuf()
    B.check(B.lib().ggdmc_b200_trial_logdens(C.byref(m.c), C.byref(t.c), B.ptr(th), th.shape[0], B.ptr(out), err), err)
    return out


def trial_logdens_hot(ct: CellTable, trials: Trials, theta: np.ndarray, seed: int = 0, pop: int = 0, iteration: int = 0):
    """Per-trial log n1PDF through the SAMPLER's trial loops (parity probe of the production path) -> ([n_theta, n_trial],
    [n_theta] sums).  The `t0 + st0 U` draws are those of (seed, pop, iteration, sweep 0, chain = row of theta)."""
    m, t = _model(ct), _trials([trials])
    th = B.f64(np.atleast_2d(theta))
    out = np.empty((th.shape[0], len(trials.rt)))
    sums = np.empty(th.shape[0])
    err = B.errbuf()
    B.check(B.lib().ggdmc_b200_trial_logdens_hot(C.byref(m.c), C.byref(t.c), B.ptr(th), th.shape[0], C.c_uint64(seed), C.c_uint32(pop),
                                                 C.c_uint32(iteration), B.ptr(out), B.ptr(sums), err), err)
    return out, sums


def sumloglike(ct: CellTable, trials: Sequence[Trials], theta: np.ndarray, init_rule: bool = False) -> np.ndarray:
    """theta [S, n_theta, npar] -> summed log-likelihoods [S, n_theta].  init_rule: densities <= 0 are floored at
    .Machine$double.eps like the R-side initialisation path does (R/phi.R:3-13)."""
    m, t = _model(ct), _trials(trials)
    th = B.f64(theta)
    assert th.ndim == 3 and th.shape[0] == len(trials) and th.shape[2] == ct.npar
    out = np.empty(th.shape[:2])
    err = B.errbuf()
    f = B.lib().ggdmc_b200_sumloglike_init if init_rule else B.lib().ggdmc_b200_sumloglike
    B.check(f(C.byref(m.c), C.byref(t.c), B.ptr(th), th.shape[1], B.ptr(out), err), err)
    return out


def sumlogprior(prior: PriorTable, x: np.ndarray, p0: Optional[np.ndarray] = None, p1: Optional[np.ndarray] = None) -> np.ndarray:
    p = _prior(prior)
    xx = B.f64(np.atleast_2d(x))
    a = B.f64(np.atleast_2d(p0)) if p0 is not None else None
    b = B.f64(np.atleast_2d(p1)) if p1 is not None else None
    out = np.empty(xx.shape[0])
    err = B.errbuf()
    B.check(B.lib().ggdmc_b200_sumlogprior(C.byref(p.c), B.ptr(xx), B.ptr(a), B.ptr(b), xx.shape[0], B.ptr(out), err), err)
    return out


def select_chains(nchain: int, k: Optional[np.ndarray] = None, u_partner: Optional[np.ndarray] = None,
                  u_mig: Optional[np.ndarray] = None):
    """Device evaluation of get_chains / get_subchains from explicit uniforms."""
    n = len(u_partner) if u_partner is not None else len(u_mig)
    kk = np.ascontiguousarray(k, dtype=np.int32) if k is not None else None
    up = B.f64(u_partner) if u_partner is not None else None
    um = B.f64(u_mig) if u_mig is not None else None
    op = np.zeros((n, 2), dtype=np.int32) if up is not None else None
    om = np.zeros((n, nchain), dtype=np.int32) if um is not None else None
    on = np.zeros(n, dtype=np.int32) if um is not None else None
    err = B.errbuf()
    B.check(B.lib().ggdmc_b200_select_chains(nchain, n, B.ptr(kk, B.c_i32p), B.ptr(up), B.ptr(op, B.c_i32p), B.ptr(um),
                                             B.ptr(om, B.c_i32p), B.ptr(on, B.c_i32p), err), err)
    return op, om, on


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    B.lib().ggdmc_b200_philox(c, k, o)
    return list(o)


def measure_fp64_tflops(device: int = -1) -> float:
    err = B.errbuf()
    v = B.lib().ggdmc_b200_measure_fp64_tflops(device, err)
    if v < 0:
        raise B.GgdmcError(B.ERR_CUDA, err.value.decode())
    return v


def device_count() -> int:
    return int(B.lib().ggdmc_b200_device_count())


class Engine:
    """Resident engine: data, state and sample storage stay in HBM between calls."""

    def __init__(self, ct: CellTable, trials: Sequence[Trials], p_prior: PriorTable, h_prior: Optional[PriorTable],
                 tuning: Tuning, phi_start: Optional[PopState], subj_start: Sequence[PopState]):
        self._keep = [_model(ct), _trials(trials), _prior(p_prior), _prior(h_prior) if h_prior is not None else None,
                      _config(tuning)]
        m, t, p, h, cfg = self._keep
        S = len(trials)
        self.R, self.S, self.C, self.D = len(tuning.seeds), S, tuning.nchain, ct.npar
        self.hier = h_prior is not None
        keep_s, starts = _start_array(subj_start)
        self._starts = (keep_s, starts, subj_start, phi_start)
        pc = phi_start.c() if phi_start is not None else None
        self.h = C.c_void_p()
        err = B.errbuf()
        rc = B.lib().ggdmc_b200_engine_create(C.byref(m.c), C.byref(t.c), C.byref(p.c), C.byref(h.c) if h else None,
                                              C.byref(cfg.c), C.byref(pc) if pc is not None else None, starts,
                                              C.byref(self.h), err)
        B.check(rc, err)

    def iterate(self, n_iter: int) -> float:
        """Advance n_iter iterations; returns the CUDA-event time in ms."""
        ms = C.c_float(0)
        err = B.errbuf()
        B.check(B.lib().ggdmc_b200_engine_iterate(self.h, int(n_iter), C.byref(ms), err), err)
        return float(ms.value)

    def iterate_flushed(self, n_iter: int, flush_bytes: int = 256 << 20) -> float:
        """Like iterate(), with an L2 flush before every iteration (outside the timed brackets)."""
        ms = C.c_float(0)
        err = B.errbuf()
        B.check(B.lib().ggdmc_b200_engine_iterate_flushed(self.h, int(n_iter), C.c_int64(flush_bytes), C.byref(ms), err), err)
        return float(ms.value)

    def time_likelihood(self, reps: int = 10):
        """(mean ms per launch of the likelihood kernel, trial-likelihoods per launch)."""
        ms, n = C.c_float(0), C.c_int64(0)
        err = B.errbuf()
        B.check(B.lib().ggdmc_b200_engine_time_likelihood(self.h, int(reps), C.byref(ms), C.byref(n), err), err)
        return float(ms.value), int(n.value)

    def state(self):
        R, S, C_, D = self.R, self.S, self.C, self.D
        st = np.empty((S, R, C_, D)); slp = np.empty((S, R, C_)); sll = np.empty((S, R, C_))
        if self.hier:
            pt = np.empty((R, C_, 2 * D)); plp = np.empty((R, C_)); pll = np.empty((R, C_))
        else:
            pt = plp = pll = None
        err = B.errbuf()
        B.check(B.lib().ggdmc_b200_engine_state(self.h, B.ptr(pt), B.ptr(plp), B.ptr(pll), B.ptr(st), B.ptr(slp), B.ptr(sll),
                                                err), err)
        return dict(phi_theta=pt, phi_lp=plp, phi_ll=pll, theta=st, lp=slp, ll=sll)

    def profile(self, enable: bool = True) -> None:
        err = B.errbuf()
        B.check(B.lib().ggdmc_b200_engine_profile(self.h, int(enable), err), err)

    def counters(self):
        """(trial-likelihoods evaluated, summed likelihood-kernel ms, likelihood launches) since the last call."""
        n, ms, k = C.c_int64(0), C.c_double(0), C.c_int64(0)
        err = B.errbuf()
        B.check(B.lib().ggdmc_b200_engine_counters(self.h, C.byref(n), C.byref(ms), C.byref(k), err), err)
        return int(n.value), float(ms.value), int(k.value)

    @property
    def persistent(self) -> bool:
        """True when the iterations run inside the persistent sampler kernel (whole iterations per launch)."""
        return bool(B.lib().ggdmc_b200_engine_is_persistent(self.h))

    @property
    def launch_count(self) -> int:
        return int(B.lib().ggdmc_b200_engine_launch_count(self.h))

    def close(self):
        if self.h:
            B.lib().ggdmc_b200_engine_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    err = B.errbuf()
    B.check(B.lib().ggdmc_b200_comm_unique_id(buf, err), err)
    return bytes(buf)


def comm_init(n_rank: int, rank: int, uid: bytes, device: int) -> None:
    buf = (C.c_uint8 * 128)(*uid)
    err = B.errbuf()
    B.check(B.lib().ggdmc_b200_comm_init(n_rank, rank, buf, device, err), err)


def comm_finalize() -> None:
    B.lib().ggdmc_b200_comm_finalize()
