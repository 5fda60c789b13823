"""Test infrastructure: drives ggdmc_b200/r/ggdmc_b200_glue.cpp (the Rcpp glue a ggdmc maintainer drops into src/)
compiled against a stand-in for Rcpp (tests/host/mock_rcpp/Rcpp.h).  R objects are assembled here from the Python
mirrors of the S4 classes (ggdmc_b200.api) through the harness's small C API and handed to the glue's
run_subject / run_hyper / run exactly as `.Call` would hand them over."""
import ctypes as C
import os
import subprocess

import numpy as np

from ggdmc_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    from ggdmc_b200 import _lib as B
    B.build()
    out = os.path.join(ROOT, "tests", "host", "libglue_harness.so")
    src = [os.path.join(ROOT, "tests", "host", "glue_harness.cpp"), os.path.join(ROOT, "tests", "host", "mock_rcpp", "Rcpp.h"),
           os.path.join(ROOT, "ggdmc_b200", "r", "ggdmc_b200_glue.cpp"), os.path.join(ROOT, "include", "ggdmc_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in src):
        so_dir = os.path.join(ROOT, "ggdmc_b200")
        subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-fPIC", "-shared", "-I", os.path.join(ROOT, "tests", "host", "mock_rcpp"),
                        "-I", os.path.join(ROOT, "include"), "-o", out, src[0], "-L", so_dir, "-lggdmc_b200", f"-Wl,-rpath,{so_dir}"],
                       check=True)
    L = C.CDLL(out)
    vp = C.c_void_p
    for name, res, args in [("gh_real", vp, [C.POINTER(C.c_double), C.c_long]), ("gh_int", vp, [C.POINTER(C.c_int), C.c_long]),
                            ("gh_lgl", vp, [C.POINTER(C.c_int), C.c_long]), ("gh_str", vp, [C.POINTER(C.c_char_p), C.c_long]),
                            ("gh_list", vp, [C.c_long]), ("gh_s4", vp, [C.c_char_p]), ("gh_set_attr", None, [vp, C.c_char_p, vp]),
                            ("gh_list_set", None, [vp, C.c_long, vp]), ("gh_get_attr", vp, [vp, C.c_char_p]), ("gh_list_get", vp, [vp, C.c_long]),
                            ("gh_length", C.c_long, [vp]), ("gh_kind", C.c_int, [vp]), ("gh_real_ptr", C.POINTER(C.c_double), [vp]),
                            ("gh_int_ptr", C.POINTER(C.c_int), [vp]), ("gh_str_at", C.c_char_p, [vp, C.c_long]), ("gh_last_error", C.c_char_p, []),
                            ("gh_set_option", None, [C.c_char_p, C.c_char_p]), ("gh_schedule_option", C.c_int, []),
                            ("gh_run_subject", vp, [vp, vp, vp]), ("gh_run_hyper", vp, [vp, vp, vp]), ("gh_run", vp, [vp, vp, vp]),
                            ("gh_run_subject_batch", vp, [vp, vp, vp]), ("gh_run_batch", vp, [vp, vp, vp]),
                            ("gh_sumloglike_init_batch", vp, [vp, vp]), ("gh_sumlogprior_batch", vp, [vp, vp, vp, vp]),
                            ("gh_flatten_model", C.c_int, [vp, C.POINTER(C.c_int), C.c_long, C.POINTER(C.c_int)]),
                            ("gh_flatten_trials", C.c_long, [vp, C.POINTER(C.c_double), C.POINTER(C.c_ushort), C.c_long]),
                            ("gh_start_slice", C.c_long, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long])]:
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


# ---- Python values -> mock R objects -------------------------------------------------------------
def r_real(x, dim=None, names=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel(order="F") if np.ndim(x) > 1 else np.atleast_1d(np.asarray(x, dtype=np.float64)))
    o = lib().gh_real(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
    _decorate(o, x, dim, names)
    return o


def _ints(x, maker, dim=None):
    a = np.ascontiguousarray((np.asarray(x).ravel(order="F") if np.ndim(x) > 1 else np.atleast_1d(np.asarray(x))).astype(np.int32))
    o = maker(a.ctypes.data_as(C.POINTER(C.c_int)), a.size)
    _decorate(o, x, dim, None)
    return o


def r_int(x, dim=None):
    return _ints(x, lib().gh_int, dim)


def r_lgl(x, dim=None):
    return _ints(np.asarray(x).astype(bool), lib().gh_lgl, dim)


def r_str(x):
    s = [x] if isinstance(x, str) else list(x)
    arr = (C.c_char_p * len(s))(*[str(v).encode() for v in s])
    return lib().gh_str(arr, len(s))


def _decorate(o, x, dim, names):
    if dim is None and np.ndim(x) > 1:
        dim = np.shape(x)
    if dim is not None:
        lib().gh_set_attr(o, b"dim", r_int(list(dim)))
    if names is not None:
        lib().gh_set_attr(o, b"names", r_str(names))


def r_list(items, names=None):
    o = lib().gh_list(len(items))
    for i, it in enumerate(items):
        lib().gh_list_set(o, i, it)
    if names is not None:
        lib().gh_set_attr(o, b"names", r_str(names))
    return o


def r_s4(klass, **slots):
    o = lib().gh_s4(klass.encode())
    for k, v in slots.items():
        lib().gh_set_attr(o, k.encode(), v)
    return o


def r_model(m: api.Model):
    return r_s4("model", parameter_x_condition_names=r_str(m.parameter_x_condition_names), pnames=r_str(m.pnames),
                cell_names=r_str(m.cell_names), constants=r_real(np.asarray(m.constants), names=m.constants.names),
                model_boolean=r_lgl(m.model_boolean), type=r_str(m.type), npar=r_int(m.npar))


def r_dmi(d: api.DMI):
    if isinstance(d.data, api.NamedList):
        data = r_list([r_real(v) for v in d.data], names=d.data.names)
    else:
        data = r_real(np.asarray(d.data))  # hyper: nsubject x npar matrix
    slots = dict(model=r_model(d.model), data=data)
    if d.node_1_index is not None:
        slots["node_1_index"] = r_int(d.node_1_index)
        slots["is_positive_drift"] = r_lgl(d.is_positive_drift)
    return r_s4("dmi", **slots)


def r_prior_list(pl: api.NamedList):
    return r_list([r_list([r_real(e["p0"]), r_real(e["p1"]), r_real(e["lower"]), r_real(e["upper"]), r_real(e["dist_id"]), r_lgl([e["log_p"]])],
                          names=["p0", "p1", "lower", "upper", "dist_id", "log_p"]) for e in pl], names=pl.names)


def r_config(c: api.Config):
    pr = c.prior
    slots = dict(nparameter=r_int(pr.nparameter), pnames=r_str(pr.pnames), p_prior=r_prior_list(pr.p_prior))
    if pr.h_prior is not None:
        slots["h_prior"] = r_prior_list(pr.h_prior)
    ti, de = c.theta_input, c.de_input
    return r_s4("config", prior=r_s4("prior", **slots),
                theta_input=r_s4("theta_input", nmc=r_int(ti.nmc), nchain=r_int(ti.nchain), thin=r_int(ti.thin), nparameter=r_int(ti.nparameter),
                                 pnames=r_str(ti.pnames), report_length=r_int(ti.report_length),
                                 max_init_attempts=r_int(ti.max_init_attempts), is_print=r_lgl([ti.is_print])),
                de_input=r_s4("de_input", pop_migration_prob=r_real(de.pop_migration_prob), sub_migration_prob=r_real(de.sub_migration_prob),
                              gamma_precursor=r_real(de.gamma_precursor), rp=r_real(de.rp), is_hblocked=r_lgl([de.is_hblocked]),
                              is_pblocked=r_lgl([de.is_pblocked]), nparameter=r_int(de.nparameter), nchain=r_int(de.nchain),
                              pop_debug=r_lgl([de.pop_debug]), sub_debug=r_lgl([de.sub_debug])),
                seed=r_real(float(c.seed)), main_seed=r_real(float(c.main_seed)), core_id=r_int(c.core_id))


def r_posterior(p: api.Posterior):
    return r_s4("posterior", theta=r_real(p.theta), summed_log_prior=r_real(p.summed_log_prior), log_likelihoods=r_real(p.log_likelihoods),
                start=r_int(p.start), npar=r_int(p.npar), pnames=r_str(p.pnames), nmc=r_int(p.nmc), thin=r_int(p.thin), nchain=r_int(p.nchain))


# ---- mock R objects -> Python ----------------------------------------------------------------------
def _real(o):
    n = lib().gh_length(o)
    a = np.ctypeslib.as_array(lib().gh_real_ptr(o), shape=(n,)).copy()
    d = lib().gh_get_attr(o, b"dim")
    if d:
        dims = np.ctypeslib.as_array(lib().gh_int_ptr(d), shape=(lib().gh_length(d),)).copy()
        a = a.reshape(tuple(int(v) for v in dims), order="F")
    return a


def _int1(o):
    return int(lib().gh_int_ptr(o)[0])


def py_posterior(o) -> api.Posterior:
    g = lambda name: lib().gh_get_attr(o, name.encode())
    pn = g("pnames")
    return api.Posterior(_real(g("theta")), _real(g("summed_log_prior")), _real(g("log_likelihoods")), _int1(g("start")), _int1(g("npar")),
                         [lib().gh_str_at(pn, i).decode() for i in range(lib().gh_length(pn))], _int1(g("nmc")), _int1(g("thin")),
                         _int1(g("nchain")))


def _check(o):
    if not o:
        raise RuntimeError(lib().gh_last_error().decode())
    return o


def run_subject(config: api.Config, dmi: api.DMI, samples: api.Posterior) -> api.Posterior:
    return py_posterior(_check(lib().gh_run_subject(r_config(config), r_dmi(dmi), r_posterior(samples))))


def run_hyper(config: api.Config, dmi: api.DMI, samples: api.Posterior) -> api.Posterior:
    return py_posterior(_check(lib().gh_run_hyper(r_config(config), r_dmi(dmi), r_posterior(samples))))


def run(config: api.Config, dmis, samples):
    return _py_hier(_check(lib().gh_run(r_config(config), r_list([r_dmi(d) for d in dmis]), _r_hier(samples))))


def _py_hier(o):
    subj = lib().gh_list_get(o, 1)
    return {"phi": py_posterior(lib().gh_list_get(o, 0)),
            "subject_theta": [py_posterior(lib().gh_list_get(subj, i)) for i in range(lib().gh_length(subj))]}


def _r_hier(samples):
    return r_list([r_posterior(samples["phi"]), r_list([r_posterior(p) for p in samples["subject_theta"]])], names=["phi", "subject_theta"])


def run_subject_batch(configs, dmi: api.DMI, samples):
    o = _check(lib().gh_run_subject_batch(r_list([r_config(c) for c in configs]), r_dmi(dmi), r_list([r_posterior(p) for p in samples])))
    return [py_posterior(lib().gh_list_get(o, r)) for r in range(lib().gh_length(o))]


def run_batch(configs, dmis, samples):
    o = _check(lib().gh_run_batch(r_list([r_config(c) for c in configs]), r_list([r_dmi(d) for d in dmis]), r_list([_r_hier(s) for s in samples])))
    return [_py_hier(lib().gh_list_get(o, r)) for r in range(lib().gh_length(o))]


def sumloglike_init_batch(dmis, theta):
    """theta [npar, n_candidate, n_subject] (R layout) -> [n_candidate, n_subject]."""
    return _real(_check(lib().gh_sumloglike_init_batch(r_list([r_dmi(d) for d in dmis]), r_real(theta))))


def sumlogprior_batch(prior_list: api.NamedList, x, p0=None, p1=None):
    """x [npar, n] (R layout) -> [n]; p0 / p1 [npar, n] or None."""
    empty = np.zeros(0)
    return _real(_check(lib().gh_sumlogprior_batch(r_prior_list(prior_list), r_real(x), r_real(empty if p0 is None else p0),
                                                   r_real(empty if p1 is None else p1))))

