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
