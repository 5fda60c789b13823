"""The Rcpp glue (ggdmc_b200/r/ggdmc_b200_glue.cpp) compiled against the Rcpp stand-in of tests/host/mock_rcpp: its
flattening rules (model_boolean + node_1_index + constants -> param_src, dmi@data -> trials, posterior -> start slice)
against the committed fixtures and against the Python mirror.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from ggdmc_b200 import api
from ggdmc_b200 import engine as E
import glue_mock as G
from helpers import fixture_objects


@pytest.mark.parametrize("k", [2, 3, 5, 6])
def test_glue_flattens_model_and_data_like_the_fixtures(k):
    fx, model, dmi_of = fixture_objects(k)
    dmi = dmi_of("sub")
    r = G.r_dmi(dmi)
    n = fx.ct.param_src.size
    buf, dims = (C.c_int * n)(), (C.c_int * 4)()
    assert G.lib().gh_flatten_model(r, buf, n, dims) == 0, G.lib().gh_last_error()
    assert list(dims)[:3] == [fx.ct.n_acc, fx.ct.n_cell, fx.ct.npar]
    assert np.array_equal(np.array(buf[:]).reshape(fx.ct.param_src.shape), fx.ct.param_src)
    tr = fx.trials("sub")
    rt, cell = (C.c_double * len(tr.rt))(), (C.c_ushort * len(tr.rt))()
    assert G.lib().gh_flatten_trials(r, rt, cell, len(tr.rt)) == len(tr.rt)
    assert np.array_equal(np.array(rt[:]), tr.rt) and np.array_equal(np.array(cell[:]), tr.cell)


def test_glue_start_slice_rule():
    """Continue from the last slice whose thetas are all finite (fresh initialise_* object: slice 1; finished fit: the last)."""
    rng = np.random.default_rng(1)
    D, Cn, nmc = 4, 5, 6
    th = rng.normal(size=(D, Cn, nmc))
    lp, ll = rng.normal(size=(Cn, nmc)), rng.normal(size=(Cn, nmc))
    for last in (0, 3, nmc - 1):
        t = th.copy()
        t[:, :, last + 1:] = np.nan
        post = api.Posterior(t, lp, ll, 1, D, [f"p{i}" for i in range(D)], nmc, 1, Cn)
        out_t, out_lp, out_ll = (C.c_double * (D * Cn))(), (C.c_double * Cn)(), (C.c_double * Cn)()
        assert G.lib().gh_start_slice(G.r_posterior(post), out_t, out_lp, out_ll, D * Cn) == Cn
        assert np.array_equal(np.array(out_t[:]).reshape(Cn, D), th[:, :, last].T)
        assert np.array_equal(np.array(out_lp[:]), lp[:, last]) and np.array_equal(np.array(out_ll[:]), ll[:, last])
        assert api._last_valid_slice(post) == last  # the Python mirror applies the same rule


def test_glue_surfaces_library_errors_as_r_errors():
    fx, model, dmi_of = fixture_objects(2)
    D = fx.ct.npar
    prior = api.Prior(nparameter=D, pnames=fx.ct.pnames, p_prior=api.prior_list(fx.prior("sub_prior")))
    st = api.Posterior(np.ones((D, 2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), 1, D, fx.ct.pnames, 3, 1, 2)
    cfg = api.Config(prior=prior, theta_input=api.ThetaInput(nmc=3, nchain=2, thin=1, nparameter=D, pnames=fx.ct.pnames),
                     de_input=api.DEInput(nparameter=D, nchain=2), seed=1)
    bad = dmi_of("sub")
    bad.model = api.Model(model.parameter_x_condition_names, model.pnames, model.cell_names, model.constants, model.model_boolean, type="ddm2")
    with pytest.raises(RuntimeError, match="Undefined model type"):  # raised by the glue itself, like @hdr/likelihood.h:312
        G.run_subject(cfg, bad, st)
    with pytest.raises(RuntimeError, match="Require three or more chains."):  # src/de.cpp:7-10, checked before any device work
        G.run_subject(cfg, dmi_of("sub"), st)
    if E.device_count() == 0:  # and a valid request without a GPU fails loudly instead of falling back to a CPU path
        st3 = api.Posterior(np.ones((D, 3, 3)), np.zeros((3, 3)), np.zeros((3, 3)), 1, D, fx.ct.pnames, 3, 1, 3)
        cfg3 = api.Config(prior=prior, theta_input=api.ThetaInput(nmc=3, nchain=3, thin=1, nparameter=D, pnames=fx.ct.pnames),
                          de_input=api.DEInput(nparameter=D, nchain=3), seed=1)
        with pytest.raises(RuntimeError, match="(?i)cuda|device"):
            G.run_subject(cfg3, dmi_of("sub"), st3)


def test_glue_refuses_malformed_and_mixed_models():
    """The advisor's findings on flatten_model: a parameter_x_condition column that is neither a free parameter nor a
    constant, two sources for one core parameter, and subjects whose models differ must all stop with an R error
    (api.run / model.build_cell_table raise in the same cases) instead of sampling a wrong likelihood."""
    fx, model, dmi_of = fixture_objects(2)
    n = fx.ct.param_src.size
    buf, dims = (C.c_int * n)(), (C.c_int * 4)()
    # (1) unknown name: rename a constant so that its column has no source
    bad = dmi_of("sub")
    cn = list(model.constants.names)
    bad.model = api.Model(model.parameter_x_condition_names, model.pnames, model.cell_names,
                          api.NamedVector(np.asarray(model.constants), ["nobody_" + c for c in cn]), model.model_boolean, type="lba")
    assert G.lib().gh_flatten_model(G.r_dmi(bad), buf, n, dims) == -1
    assert b"neither a free parameter" in G.lib().gh_last_error()
    # (2) two sources for one (cell, accumulator, core parameter)
    mb = np.array(model.model_boolean, copy=True)
    pxc = list(model.parameter_x_condition_names)
    a_cols = [k for k, nm in enumerate(pxc) if nm.split(".")[0] == "B"]
    if len(a_cols) >= 2:
        mb[0, a_cols[0], 0] = True
        mb[0, a_cols[1], 0] = True
        two = dmi_of("sub")
        two.model = api.Model(pxc, model.pnames, model.cell_names, model.constants, mb, type="lba")
        assert G.lib().gh_flatten_model(G.r_dmi(two), buf, n, dims) == -1
        assert b"more than one source" in G.lib().gh_last_error()
    # (3) mixed models in one hierarchical call
    fx6, model6, dmi6 = fixture_objects(6)
    D = fx.ct.npar
    prior = api.Prior(nparameter=D, pnames=fx.ct.pnames, p_prior=api.prior_list(fx.prior("p_prior")), h_prior=api.prior_list(fx.prior("h_prior")))
    st = api.Posterior(np.ones((D, 3, 2)), np.zeros((3, 2)), np.zeros((3, 2)), 1, D, fx.ct.pnames, 2, 1, 3)
    phi = api.Posterior(np.ones((2 * D, 3, 2)), np.zeros((3, 2)), np.zeros((3, 2)), 1, 2 * D, fx.ct.pnames * 2, 2, 1, 3)
    cfg = api.Config(prior=prior, theta_input=api.ThetaInput(nmc=2, nchain=3, thin=1, nparameter=2 * D, pnames=fx.ct.pnames * 2),
                     de_input=api.DEInput(nparameter=2 * D, nchain=3), seed=1)
    with pytest.raises(RuntimeError, match="must share one model"):
        G.run(cfg, [dmi_of("pop0"), dmi6("pop0")], {"phi": phi, "subject_theta": [st, st]})


def test_glue_start_slice_needs_a_finite_slice():
    """No slice with finite thetas: an R error, not an out-of-bounds read (api._last_valid_slice raises as well)."""
    D, Cn, nmc = 3, 4, 5
    post = api.Posterior(np.full((D, Cn, nmc), np.nan), np.zeros((Cn, nmc)), np.zeros((Cn, nmc)), 1, D, ["a", "b", "c"], nmc, 1, Cn)
    out_t, out_lp, out_ll = (C.c_double * (D * Cn))(), (C.c_double * Cn)(), (C.c_double * Cn)()
    assert G.lib().gh_start_slice(G.r_posterior(post), out_t, out_lp, out_ll, D * Cn) == -1
    assert b"no slice with finite thetas" in G.lib().gh_last_error()
    with pytest.raises(ValueError):
        api._last_valid_slice(post)


def test_glue_schedule_option():
    """options(ggdmc.schedule = ...) selects the chain-update schedule; unset = the two-half parallel schedule."""
    from ggdmc_b200 import _lib as B
    L = G.lib()
    L.gh_set_option(b"ggdmc.schedule", None)
    assert L.gh_schedule_option() == B.SCHEDULE_PARALLEL
    for name, val in ((b"reference", B.SCHEDULE_REFERENCE), (b"parallel", B.SCHEDULE_PARALLEL), (b"simultaneous", B.SCHEDULE_SIMULTANEOUS)):
        L.gh_set_option(b"ggdmc.schedule", name)
        assert L.gh_schedule_option() == val
    L.gh_set_option(b"ggdmc.schedule", b"fastest")
    assert L.gh_schedule_option() == -1 and b"ggdmc.schedule" in L.gh_last_error()
    L.gh_set_option(b"ggdmc.schedule", None)


def test_restart_with_changed_nmc_and_thin_takes_the_last_slice():
    """RestartSampling with new nmc / thin (R/sampling.R:370-421) hands the previous fit over as `samples` and a config
    with the NEW theta_input: the glue's start state is the previous fit's last finite slice whatever its nmc / thin were,
    and the new fit's arrays are sized by the new nmc (checked here on the flattened objects; the GPU run of the same
    scenario is tests/test_gpu_api.py::test_restart_with_changed_nmc_and_thin)."""
    rng = np.random.default_rng(3)
    D, Cn = 4, 6
    for old_nmc, filled in ((10, 10), (7, 4)):
        th = rng.normal(size=(D, Cn, old_nmc))
        th[:, :, filled:] = np.nan  # an interrupted fit: only the first `filled` slices were written
        lp, ll = rng.normal(size=(Cn, old_nmc)), rng.normal(size=(Cn, old_nmc))
        prev = api.Posterior(th, lp, ll, 1, D, [f"p{i}" for i in range(D)], old_nmc, 8, Cn)
        out_t, out_lp, out_ll = (C.c_double * (D * Cn))(), (C.c_double * Cn)(), (C.c_double * Cn)()
        assert G.lib().gh_start_slice(G.r_posterior(prev), out_t, out_lp, out_ll, D * Cn) == Cn
        assert np.array_equal(np.array(out_t[:]).reshape(Cn, D), th[:, :, filled - 1].T)
        assert np.array_equal(np.array(out_lp[:]), lp[:, filled - 1])


def test_patch_applies(tmp_path):
    """ggdmc_b200/r/patch/apply.sh on a tree with the reference's layout: the sampler sources are replaced by the glue,
    the wrappers of the batch routines are registered, parallel_lapply is the batched one -- exactly once."""
    import os
    import shutil
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("the reference checkout is not mounted here")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = tmp_path / "ggdmc"
    pkg.mkdir()
    for d in ("R", "src"):
        shutil.copytree(os.path.join(ref, d), pkg / d)
    for f in ("DESCRIPTION", "NAMESPACE"):
        shutil.copy(os.path.join(ref, f), pkg / f)
    r = subprocess.run(["sh", os.path.join(root, "ggdmc_b200", "r", "patch", "apply.sh"), str(pkg), root], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = sorted(os.listdir(pkg / "src"))
    assert src == ["Makevars", "RcppExports.cpp", "ggdmc_b200_glue.cpp"], src
    exports = (pkg / "src" / "RcppExports.cpp").read_text()
    for fn in ("run_subject", "run_hyper", "run", "run_subject_batch", "run_batch", "sumloglike_init_batch", "sumlogprior_batch"):
        assert exports.count(f'{{"_ggdmc_{fn}", (DL_FUNC) &_ggdmc_{fn},') == 1, fn
    assert "RcppArmadillo" not in exports and exports.index("_ggdmc_run_batch(SEXP") < exports.index("CallEntries[]")
    sampling = (pkg / "R" / "sampling.R").read_text()
    assert sampling.count("parallel_lapply <- function(") == 1 and "parallel::mclapply" not in sampling and "run_subject_batch(config_list" in sampling
    assert sampling.count("{") == sampling.count("}")
    assert "StartSampling <- function" in sampling  # the callers are untouched
    assert root in (pkg / "src" / "Makevars").read_text()
    desc = (pkg / "DESCRIPTION").read_text()
    assert "ggdmcHeaders" not in desc and "RcppArmadillo" not in desc and "Imports:" in desc
    r2 = subprocess.run(["sh", os.path.join(root, "ggdmc_b200", "r", "patch", "apply.sh"), str(pkg), root], capture_output=True, text=True)
    assert r2.returncode != 0 and "already applied" in r2.stderr
