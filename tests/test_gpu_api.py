"""GPU tests (-m gpu) of the reference-level Python mirror: ggdmc_b200.api (run_subject / run_hyper / run on
objects with the reference's S4 slot names) and ggdmc_b200.init (initialise_theta / initialise_phi)."""
import numpy as np
import pytest

from ggdmc_b200 import _lib as B
from ggdmc_b200 import api, init
from ggdmc_b200 import engine as E
from oracle import binding as ob
from helpers import fixture_objects, load_fixture

pytestmark = pytest.mark.gpu


def test_flattening_of_reference_objects_matches_goldens():
    for k in (2, 3, 5, 6):
        fx, model, dmi_of = fixture_objects(k)
        from ggdmc_b200.model import build_cell_table, flatten_data
        ct = build_cell_table(model, fx.g["node_1_index"], fx.g["is_positive_drift"])
        assert np.array_equal(ct.param_src, fx.ct.param_src) and np.array_equal(ct.const_val, fx.ct.const_val)
        tr = flatten_data(dmi_of("sub").data, ct.cell_names)
        assert np.array_equal(tr.rt, fx.trials("sub").rt) and np.array_equal(tr.cell, fx.trials("sub").cell)


def test_run_subject_through_reference_interface():
    """StartSampling_subject's inner call: config + dmi + fresh start samples -> posterior (R/sampling.R:446-502)."""
    fx, model, dmi_of = fixture_objects(6)
    dmi = dmi_of("sub")
    D, nchain, nmc, thin = fx.ct.npar, 3 * fx.ct.npar, 6, 2
    prior = api.Prior(nparameter=D, pnames=fx.ct.pnames, p_prior=api.prior_list(fx.prior("sub_prior")))
    ti = api.ThetaInput(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, pnames=fx.ct.pnames, report_length=2, is_print=True)
    de = api.DEInput(sub_migration_prob=0.06, nparameter=D, nchain=nchain)
    configs = [api.Config(prior=prior, theta_input=ti, de_input=de, seed=s) for s in (101, 202, 303)]  # ncore = 3 replicates
    starts = [init.initialise_theta(ti, prior, dmi, seed=s) for s in (1, 2, 3)]
    for st in starts:  # a fresh start object: slice 1 valid, the rest NaN / -Inf (R/phi.R:72-88)
        assert np.all(np.isfinite(st.theta[:, :, 0])) and np.all(np.isnan(st.theta[:, :, 1:]))
        assert np.all(np.isfinite(st.log_likelihoods[:, 0])) and np.all(np.isfinite(st.summed_log_prior[:, 0]))
        # the stored scores are the oracle's R-init-path values
        od, op = fx.odata("sub"), fx.oprior("sub_prior")
        for c in (0, 7, nchain - 1):
            ref = ob.sumloglike_rinit(fx.om, od, st.theta[:, c, 0])
            assert abs(st.log_likelihoods[c, 0] - ref) <= 1e-9 * abs(ref)
            assert abs(st.summed_log_prior[c, 0] - ob.sumlogprior(op, st.theta[:, c, 0])) <= 1e-9
    seen = []
    fits = api.run_subject(configs, dmi, starts, progress=seen.append)
    assert len(fits) == 3 and seen  # progress callback fired (report_length = 2)
    for st, fit in zip(starts, fits):
        assert fit.theta.shape == (D, nchain, nmc) and fit.summed_log_prior.shape == (nchain, nmc) and fit.nmc == nmc
        assert fit.pnames == fx.ct.pnames and fit.thin == thin and fit.start == 1 and fit.npar == D and fit.nchain == nchain
        assert np.array_equal(fit.theta[:, :, 0], st.theta[:, :, 0]) and np.array_equal(fit.log_likelihoods[:, 0], st.log_likelihoods[:, 0])
        assert np.all(np.isfinite(fit.theta)) and not np.array_equal(fit.theta[:, :, 0], fit.theta[:, :, -1])
    # RestartSampling_subject: feed the fit back as `samples`; sampling continues from its last slice
    again = api.run_subject(configs[0], dmi, fits[0])
    assert np.array_equal(again.theta[:, :, 0], fits[0].theta[:, :, -1])
    # same seed, same start -> same fit (reproducibility); other seed -> other fit
    rep = api.run_subject(configs[0], dmi, starts[0])
    assert np.array_equal(rep.theta, fits[0].theta) and not np.array_equal(fits[0].theta, fits[1].theta)


def test_run_and_run_hyper_through_reference_interface():
    """StartSampling / StartSampling_hyper inner calls (R/sampling.R:247-304, 615-677)."""
    fx, model, dmi_of = fixture_objects(2)
    S, D = fx.n_pop, fx.ct.npar
    dmis = [dmi_of(f"pop{s}") for s in range(S)]
    nchain, nmc, thin = 6 * D, 5, 2
    pp, hp = fx.prior("p_prior"), fx.prior("h_prior")
    prior = api.Prior(nparameter=2 * D, pnames=hp.pnames, p_prior=api.prior_list(pp), h_prior=api.prior_list(hp))
    ti = api.ThetaInput(nmc=nmc, nchain=nchain, thin=thin, nparameter=2 * D, pnames=hp.pnames)
    de = api.DEInput(pop_migration_prob=0.05, sub_migration_prob=0.05, nparameter=2 * D, nchain=nchain)
    cfg = api.Config(prior=prior, theta_input=ti, de_input=de, seed=9032)
    start = init.initialise_phi(ti, prior, dmis, seed=5)
    assert start["phi"].theta.shape == (2 * D, nchain, nmc) and len(start["subject_theta"]) == S
    # phi start scores = h_prior density and hyper-likelihood of the subjects' chain-k thetas (oracle)
    opp, ohp = fx.oprior("p_prior"), fx.oprior("h_prior")
    for c in (0, nchain - 1):
        phi_c = start["phi"].theta[:, c, 0]
        hl = sum(ob.sumlogprior(opp, s.theta[:, c, 0], phi_c[:D], phi_c[D:]) for s in start["subject_theta"])
        assert abs(start["phi"].log_likelihoods[c, 0] - hl) <= 1e-9 * abs(hl)
        assert abs(start["phi"].summed_log_prior[c, 0] - ob.sumlogprior(ohp, phi_c)) <= 1e-9
    fit = api.run(cfg, dmis, start)
    assert set(fit) == {"phi", "subject_theta"} and len(fit["subject_theta"]) == S
    assert fit["phi"].theta.shape == (2 * D, nchain, nmc) and fit["subject_theta"][0].theta.shape == (D, nchain, nmc)
    assert fit["phi"].pnames == hp.pnames and fit["subject_theta"][0].pnames == fx.ct.pnames
    assert np.all(np.isfinite(fit["phi"].theta)) and np.array_equal(fit["phi"].theta[:, :, 0], start["phi"].theta[:, :, 0])
    refit = api.run(cfg, dmis, fit)  # RestartSampling
    assert np.array_equal(refit["phi"].theta[:, :, 0], fit["phi"].theta[:, :, -1])
    # hyper-only fit on the matrix of "true" subject thetas
    hyper_dmi = api.DMI(model=api.Model([], [], [], api.NamedVector([], []), np.zeros((0, 0, 0), bool), type="hyper"), data=fx.g["hyper_data"])
    phi_start = start["phi"]
    hfit = api.run_hyper(cfg, hyper_dmi, phi_start)
    assert hfit.theta.shape == (2 * D, nchain, nmc) and np.all(np.isfinite(hfit.theta[:, :, 0]))


def test_reference_errors_surface():
    fx, model, dmi_of = fixture_objects(2)
    D = fx.ct.npar
    prior = api.Prior(nparameter=D, pnames=fx.ct.pnames, p_prior=api.prior_list(fx.prior("sub_prior")))
    ti = api.ThetaInput(nmc=3, nchain=2, thin=1, nparameter=D, pnames=fx.ct.pnames)
    cfg = api.Config(prior=prior, theta_input=ti, de_input=api.DEInput(nparameter=D, nchain=2), seed=1)
    st = api.Posterior(np.ones((D, 2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), 1, D, fx.ct.pnames, 3, 1, 2)
    with pytest.raises(B.GgdmcError, match="three or more chains"):
        api.run_subject(cfg, dmi_of("sub"), st)
    bad = dmi_of("sub")
    bad.model = api.Model(model.parameter_x_condition_names, model.pnames, model.cell_names, model.constants, model.model_boolean, type="ddm2")
    with pytest.raises(B.GgdmcError, match="Undefined model type"):
        api.run_subject(cfg, bad, st)


def test_rcpp_glue_entry_points_equal_the_python_mirror():
    """ggdmc_b200/r/ggdmc_b200_glue.cpp -- the file a ggdmc maintainer drops into src/ -- compiled against the Rcpp
    stand-in (tests/host/mock_rcpp) and driven with R-like objects: run_subject, run_hyper and run return posterior
    objects with the same slots and the same numbers as the Python mirror of the interface (same seed = same Philox key)."""
    import glue_mock as G
    fx, model, dmi_of = fixture_objects(2)
    S, D = 4, fx.ct.npar
    nchain, nmc, thin = 6 * D, 4, 2
    pp, hp = fx.prior("p_prior"), fx.prior("h_prior")
    prior = api.Prior(nparameter=2 * D, pnames=hp.pnames, p_prior=api.prior_list(pp), h_prior=api.prior_list(hp))
    ti = api.ThetaInput(nmc=nmc, nchain=nchain, thin=thin, nparameter=2 * D, pnames=hp.pnames, report_length=2, is_print=True)
    de = api.DEInput(pop_migration_prob=0.2, sub_migration_prob=0.2, nparameter=2 * D, nchain=nchain)
    cfg = api.Config(prior=prior, theta_input=ti, de_input=de, seed=4242)
    dmis = [dmi_of(f"pop{s}") for s in range(S)]
    start = init.initialise_phi(ti, prior, dmis, seed=3)

    def same(a: api.Posterior, b: api.Posterior):
        assert np.array_equal(a.theta, b.theta) and np.array_equal(a.summed_log_prior, b.summed_log_prior)
        assert np.array_equal(a.log_likelihoods, b.log_likelihoods)
        assert (a.start, a.npar, a.pnames, a.nmc, a.thin, a.nchain) == (b.start, b.npar, b.pnames, b.nmc, b.thin, b.nchain)

    # run (hierarchical)
    ref = api.run(cfg, dmis, start)
    got = G.run(cfg, dmis, start)
    same(got["phi"], ref["phi"])
    assert len(got["subject_theta"]) == S
    for a, b in zip(got["subject_theta"], ref["subject_theta"]):
        same(a, b)
    # run_subject (first subject, its own prior), continuing from the hierarchical fit like RestartSampling does
    sub_prior = api.Prior(nparameter=D, pnames=fx.ct.pnames, p_prior=api.prior_list(fx.prior("sub_prior")))
    cfg1 = api.Config(prior=sub_prior, theta_input=api.ThetaInput(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, pnames=fx.ct.pnames),
                      de_input=api.DEInput(sub_migration_prob=0.1, nparameter=D, nchain=nchain), seed=7)
    st1 = ref["subject_theta"][0]
    same(G.run_subject(cfg1, dmis[0], st1), api.run_subject(cfg1, dmis[0], st1))
    # run_hyper on a matrix of subject-level estimates
    hyper_dmi = api.DMI(model=api.Model([], [], [], api.NamedVector([], []), np.zeros((0, 0, 0), bool), type="hyper"), data=fx.g["hyper_data"])
    same(G.run_hyper(cfg, hyper_dmi, start["phi"]), api.run_hyper(cfg, hyper_dmi, start["phi"]))
    # the ncore replicates as one call (what parallel_lapply becomes): equal to one call per replicate
    cfgs = [api.Config(prior=prior, theta_input=ti, de_input=de, seed=sd) for sd in (4242, 99)]
    starts = [start, init.initialise_phi(ti, prior, dmis, seed=8)]
    batch = G.run_batch(cfgs, dmis, starts)
    same(batch[0]["phi"], ref["phi"])
    second = api.run(cfgs[1], dmis, starts[1])
    same(batch[1]["phi"], second["phi"])
    for a, b in zip(batch[1]["subject_theta"], second["subject_theta"]):
        same(a, b)
    cfgs1 = [api.Config(prior=sub_prior, theta_input=cfg1.theta_input, de_input=cfg1.de_input, seed=sd) for sd in (7, 8, 9)]
    sts = [ref["subject_theta"][0], ref["subject_theta"][1], ref["subject_theta"][2]]
    fits = G.run_subject_batch(cfgs1, dmis[0], sts)
    assert len(fits) == 3
    for c1, s1, f in zip(cfgs1, sts, fits):
        same(f, api.run_subject(c1, dmis[0], s1))


def test_rcpp_glue_scores_start_value_candidates_in_bulk():
    """sumloglike_init_batch / sumlogprior_batch of the glue: what an R-side initialise_theta / initialise_phi would call
    instead of one likelihood round trip per candidate (R/phi.R:166-201, 300-326)."""
    import glue_mock as G
    fx, model, dmi_of = fixture_objects(6)
    S, D, n_cand = 3, fx.ct.npar, 11
    dmis = [dmi_of(f"pop{s}") for s in range(S)]
    rng = np.random.default_rng(2)
    pp = fx.prior("p_prior")
    cand = init.rprior(pp, S * n_cand, rng).reshape(S, n_cand, D)  # [subject][candidate][par]
    cand[0, 0, :] = 50.0  # a hopeless candidate: densities underflow to 0 -> the eps floor of .sumlog applies
    got = G.sumloglike_init_batch(dmis, np.transpose(cand, (2, 1, 0)))  # R array npar x n_candidate x n_subject
    ref = E.sumloglike(fx.ct, [fx.trials(f"pop{s}") for s in range(S)], cand, init_rule=True)
    assert got.shape == (n_cand, S) and np.array_equal(got.T, ref) and np.all(np.isfinite(got))
    od = fx.odata("pop1")
    for j in (0, 5):
        o = ob.sumloglike_rinit(fx.om, od, cand[1, j])
        assert abs(got[j, 1] - o) <= 1e-9 * abs(o)
    x = cand.reshape(-1, D)
    lp = G.sumlogprior_batch(api.prior_list(pp), x.T)
    assert np.array_equal(lp, E.sumlogprior(pp, x))
    phi = np.abs(rng.normal(1.0, 0.2, size=(x.shape[0], 2 * D)))
    lp2 = G.sumlogprior_batch(api.prior_list(pp), x.T, phi[:, :D].T, phi[:, D:].T)
    assert np.array_equal(lp2, E.sumlogprior(pp, x, np.ascontiguousarray(phi[:, :D]), np.ascontiguousarray(phi[:, D:])))



def test_restart_with_changed_nmc_and_thin():
    """RestartSampling_subject with a different nmc / thin (R/sampling.R:370-421 builds a new theta_input and passes the
    previous fit as `samples`): through the Rcpp glue and through the Python mirror the continuation starts at the
    previous fit's last slice, has the NEW nmc / thin, and is what the oracle produces from that state with the same seed
    (the reference's theta_class constructor is not in de.o; that it starts from the last stored sample is what
    R/sampling.R:426-429 relies on when it binds old and new samples together)."""
    import glue_mock as G
    from test_gpu_sampler import compare
    fx, model, dmi_of = fixture_objects(2)
    dmi = dmi_of("sub")
    D, nchain = fx.ct.npar, 3 * fx.ct.npar
    prior = api.Prior(nparameter=D, pnames=fx.ct.pnames, p_prior=api.prior_list(fx.prior("sub_prior")))
    de = api.DEInput(sub_migration_prob=0.1, nparameter=D, nchain=nchain)
    ti1 = api.ThetaInput(nmc=5, nchain=nchain, thin=2, nparameter=D, pnames=fx.ct.pnames)
    first = api.run_subject(api.Config(prior=prior, theta_input=ti1, de_input=de, seed=11), dmi, init.initialise_theta(ti1, prior, dmi, seed=4))
    ti2 = api.ThetaInput(nmc=8, nchain=nchain, thin=3, nparameter=D, pnames=fx.ct.pnames)
    cfg2 = api.Config(prior=prior, theta_input=ti2, de_input=de, seed=12)
    again_py = api.run_subject(cfg2, dmi, first)
    again_r = G.run_subject(cfg2, dmi, first)
    for again in (again_py, again_r):
        assert again.theta.shape == (D, nchain, 8) and again.nmc == 8 and again.thin == 3
        assert np.array_equal(again.theta[:, :, 0], first.theta[:, :, -1])
        assert np.array_equal(again.log_likelihoods[:, 0], first.log_likelihoods[:, -1])
    assert np.array_equal(again_py.theta, again_r.theta)
    # the oracle from the same state: 7 x 3 iterations of run_chains in the engine's default (two-half) order
    pop = ob.OPop(np.ascontiguousarray(first.theta[:, :, -1].T), first.summed_log_prior[:, -1].copy(), first.log_likelihoods[:, -1].copy(), 8, 3)
    de_o = ob.make_de(D, nchain, sub_migration_prob=0.1, jacobi=1)
    ob.run_subject(de_o, pop, fx.oprior("sub_prior"), fx.om, fx.odata("sub"), ob.make_rng(seed=12), 0, 7 * 3)
    gpu = E.PopSamples(np.ascontiguousarray(again_py.theta.transpose(2, 1, 0))[None], again_py.summed_log_prior.T[None].copy(), again_py.log_likelihoods.T[None].copy())
    compare(gpu, 0, pop, "restarted fit")
