"""GPU sampler tests (-m gpu): whole DE-MCMC trajectories from the CUDA engine against the oracle.

Both sides draw every uniform from the same counter-addressed Philox source, and the proposal
arithmetic is done with the same IEEE operations, so theta trajectories must agree EXACTLY (a
difference could only come from an accept decision whose MH ratio lies within ~1e-13 of its
uniform); log prior / log likelihood values agree to FP64 rounding."""
import numpy as np
import pytest

from ggdmc_b200 import _lib as B
from ggdmc_b200 import engine as E
from ggdmc_b200.model import Trials
from oracle import binding as ob
from helpers import load_fixture, sane_starts

pytestmark = pytest.mark.gpu

SCHEDULES = [(B.SCHEDULE_PARALLEL, 1), (B.SCHEDULE_REFERENCE, 0), (B.SCHEDULE_SIMULTANEOUS, 2)]  # (engine schedule, oracle schedule)


@pytest.fixture(autouse=True, params=["default", "persistent", "launches"])
def kernel_path(request, monkeypatch):
    """Every test of this module runs three times: with the engine's own choice between the persistent sampler kernel and
    the launch sequence (persistent for fits without a phi level), with the persistent kernel forced wherever it applies
    (the PARALLEL schedule of an LBA fit without per-parameter sweeps), and with the launch sequence forced."""
    monkeypatch.delenv("GGDMC_B200_PERSIST", raising=False)
    monkeypatch.delenv("GGDMC_B200_NO_PERSIST", raising=False)
    if request.param == "persistent":
        monkeypatch.setenv("GGDMC_B200_PERSIST", "1")
    elif request.param == "launches":
        monkeypatch.setenv("GGDMC_B200_NO_PERSIST", "1")
    return request.param


def subject_start(fx, od, oprior, nchain, rng, center=None, phi=None):
    th = sane_starts(fx, nchain, rng, center=center)
    D = th.shape[1]
    lp = np.array([ob.sumlogprior(oprior, th[c], None if phi is None else phi[c, :D], None if phi is None else phi[c, D:])
                   for c in range(nchain)])
    ll = np.array([ob.sumloglike(fx.om, od, th[c]) for c in range(nchain)])
    return th, lp, ll


def compare(gpu: E.PopSamples, r: int, pop: ob.OPop, what: str):
    th, lp, ll = gpu.theta[r], gpu.lp[r], gpu.ll[r]
    same = np.array_equal(th, pop.out_theta)
    if not same:
        bad = np.argwhere(th != pop.out_theta)
        raise AssertionError(f"{what}: theta trajectories differ first at (slot, chain, par) = {bad[0]}, {len(bad)} entries")
    for a, b, nm in ((lp, pop.out_lp, "lp"), (ll, pop.out_ll, "ll")):
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), f"{what}: {nm} finiteness differs"
        assert np.all(np.abs(a[fin] - b[fin]) <= 1e-9 * np.maximum(1.0, np.abs(b[fin]))), f"{what}: {nm} differs"


@pytest.mark.parametrize("schedule,jacobi", SCHEDULES)
@pytest.mark.parametrize("pblocked", [False, True])
def test_run_subject_trajectory(schedule, jacobi, pblocked):
    """run_subject (src/de2R.cpp:8-23): crossover + migration sweeps, thinning, storage."""
    fx = load_fixture(2)
    rng = np.random.default_rng(7)
    tr, od = fx.trials("sub"), fx.odata("sub")
    prior, oprior = fx.prior("sub_prior"), fx.oprior("sub_prior")
    D, nchain, nmc, thin = fx.ct.npar, 3 * fx.ct.npar, 5, 2
    seeds = [9032, 77]
    starts = [subject_start(fx, od, oprior, nchain, rng) for _ in seeds]
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, sub_migration_prob=0.35, is_pblocked=pblocked, schedule=schedule,
                   seeds=seeds)
    st = E.PopState(np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts]), np.stack([s[2] for s in starts]))
    out = E.run_subject(fx.ct, tr, prior, tun, st)
    for r, seed in enumerate(seeds):
        pop = ob.OPop(*starts[r], nmc, thin)
        de = ob.make_de(D, nchain, sub_migration_prob=0.35, is_pblocked=pblocked, jacobi=jacobi)
        ob.run_subject(de, pop, oprior, fx.om, od, ob.make_rng(seed=seed), 0, (nmc - 1) * thin)
        compare(out, r, pop, f"run_subject seed {seed}")
    assert not np.array_equal(out.theta[0], out.theta[1])
    assert not np.array_equal(out.theta[0, 0], out.theta[0, -1])  # chains moved


@pytest.mark.parametrize("schedule,jacobi", SCHEDULES)
def test_run_hyper_trajectory(schedule, jacobi):
    """run_hyper (src/de2R.cpp:30-47): phi chains against a fixed matrix of subject thetas."""
    fx = load_fixture(2)
    rng = np.random.default_rng(8)
    pp, hp, opp, ohp = fx.prior("p_prior"), fx.prior("h_prior"), fx.oprior("p_prior"), fx.oprior("h_prior")
    data = fx.g["hyper_data"]
    D = pp.npar
    nchain, nmc, thin = 3 * 2 * D, 6, 2
    center = np.concatenate([fx.g["pop_mean"], fx.g["pop_scale"]])
    phi0 = center[None, :] * (1.0 + 0.1 * rng.standard_normal((nchain, 2 * D)))
    lp0 = np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(nchain)])
    ll0 = np.array([sum(ob.sumlogprior(opp, data[s], phi0[c, :D], phi0[c, D:]) for s in range(data.shape[0])) for c in range(nchain)])
    seed = 4242
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=2 * D, sub_migration_prob=0.3, schedule=schedule, seeds=[seed])
    out = E.run_hyper(pp, hp, data, tun, E.PopState(phi0, lp0, ll0))
    pop = ob.OPop(phi0, lp0, ll0, nmc, thin)
    de = ob.make_de(2 * D, nchain, sub_migration_prob=0.3, jacobi=jacobi)
    ob.run_hyper(de, pop, opp, ohp, data, ob.make_rng(seed=seed), (nmc - 1) * thin)
    compare(out, 0, pop, "run_hyper")
    assert not np.array_equal(out.theta[0, 0], out.theta[0, -1])


def hier_setup(fx, S, nchain, rng):
    D = fx.ct.npar
    opp, ohp = fx.oprior("p_prior"), fx.oprior("h_prior")
    center = np.concatenate([fx.g["pop_mean"], fx.g["pop_scale"]])
    phi0 = center[None, :] * (1.0 + 0.1 * rng.standard_normal((nchain, 2 * D)))
    subj = []
    for s in range(S):
        od = fx.odata(f"pop{s}")
        subj.append(subject_start(fx, od, opp, nchain, rng, center=fx.g["ps"][s], phi=phi0))
    lp0 = np.array([ob.sumlogprior(ohp, phi0[c]) for c in range(nchain)])
    ll0 = np.array([sum(ob.sumlogprior(opp, subj[s][0][c], phi0[c, :D], phi0[c, D:]) for s in range(S)) for c in range(nchain)])
    return (phi0, lp0, ll0), subj


@pytest.mark.parametrize("schedule,jacobi", SCHEDULES)
@pytest.mark.parametrize("blocked", [False, True])
def test_run_hierarchical_trajectory(schedule, jacobi, blocked):
    """run (src/de2R.cpp:123-171 -> run_hchains): phi step, subject steps with phi-driven priors,
    stale-lp and refreshed-hyper-ll quirks, migration at both levels, per-parameter blocking."""
    fx = load_fixture(2)
    rng = np.random.default_rng(21)
    S = fx.n_pop
    D = fx.ct.npar
    nchain, nmc, thin = 3 * 2 * D, 4, 2
    if blocked:
        nmc = 3
    phi_s, subj_s = hier_setup(fx, S, nchain, rng)
    seed = 31337
    kw = dict(pop_migration_prob=0.3, sub_migration_prob=0.3, is_hblocked=blocked, is_pblocked=blocked)
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=2 * D, schedule=schedule, seeds=[seed], **kw)
    phi_out, subj_out = E.run_hier(fx.ct, [fx.trials(f"pop{s}") for s in range(S)], fx.prior("p_prior"), fx.prior("h_prior"), tun,
                                   E.PopState(*phi_s), [E.PopState(*s) for s in subj_s])
    phi = ob.OPop(*phi_s, nmc, thin)
    pops = [ob.OPop(*s, nmc, thin) for s in subj_s]
    de = ob.make_de(2 * D, nchain, jacobi=jacobi, **kw)
    ob.run_hier(de, phi, pops, fx.oprior("p_prior"), fx.oprior("h_prior"), fx.om, [fx.odata(f"pop{s}") for s in range(S)],
                ob.make_rng(seed=seed), (nmc - 1) * thin)
    compare(phi_out, 0, phi, "phi")
    for s in range(S):
        compare(subj_out[s], 0, pops[s], f"subject {s}")
    assert not np.array_equal(phi_out.theta[0, 0], phi_out.theta[0, -1])


@pytest.mark.skipif(ob.ref_lib() is None, reason="oracle/_ref (reference object code) was never built")
def test_gpu_reference_schedule_vs_reference_object_code():
    """The GPU engine (REFERENCE schedule) against the reference's OWN machine code: the oracle replays the
    engine's counter-addressed draws and records them in the order the reference consumes uniforms; that
    recording is injected behind Rf_runif and de_class::run_hchains of src/de.o runs on it.  theta of every
    stored sample must be identical; log prior / log likelihood agree to FP64 rounding."""
    ob.ref2_prime()
    fx = load_fixture(2)
    rng = np.random.default_rng(77)
    S, D = fx.n_pop, fx.ct.npar
    nchain, nmc, thin, seed = 6 * D, 4, 2, 424242
    phi_s, subj_s = hier_setup(fx, S, nchain, rng)
    kw = dict(pop_migration_prob=0.3, sub_migration_prob=0.3)
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=2 * D, schedule=B.SCHEDULE_REFERENCE, seeds=[seed], **kw)
    phi_out, subj_out = E.run_hier(fx.ct, [fx.trials(f"pop{s}") for s in range(S)], fx.prior("p_prior"), fx.prior("h_prior"), tun,
                                   E.PopState(*phi_s), [E.PopState(*s) for s in subj_s])
    # the draw sequence of that run, in the reference's consumption order
    datas = [fx.odata(f"pop{s}") for s in range(S)]
    opp, ohp = fx.oprior("p_prior"), fx.oprior("h_prior")
    rec = ob.make_rng(seed=seed, record=4000000)
    ob.run_hier(ob.make_de(2 * D, nchain, **kw), ob.OPop(*phi_s, nmc, thin), [ob.OPop(*s, nmc, thin) for s in subj_s], opp, ohp, fx.om,
                datas, rec, (nmc - 1) * thin)
    stream = ob.recorded(rec)
    ob.ref2_set_stream(stream)
    (pt, plp, pll), subs = ob.ref2_run_hchains(2 * D, fx.om, datas, opp, ohp, phi_s, subj_s, nmc, thin, **kw)
    assert ob.ref_lib().ref_uniform_stream_pos() == len(stream)
    assert np.array_equal(phi_out.theta[0], pt)
    assert np.allclose(phi_out.ll[0], pll, rtol=1e-9, atol=0) and np.allclose(phi_out.lp[0], plp, rtol=1e-9, atol=0)
    for s in range(S):
        assert np.array_equal(subj_out[s].theta[0], subs[s][0]), f"subject {s}"
        fin = np.isfinite(subs[s][2])
        assert np.allclose(subj_out[s].ll[0][fin], subs[s][2][fin], rtol=1e-9, atol=0)
    assert not np.array_equal(pt[0], pt[-1])


def test_replicates_batch_equals_separate_runs():
    fx = load_fixture(2)
    rng = np.random.default_rng(3)
    tr, od = fx.trials("sub"), fx.odata("sub")
    prior, oprior = fx.prior("sub_prior"), fx.oprior("sub_prior")
    D, nchain = fx.ct.npar, 3 * fx.ct.npar
    seeds = [5, 6, 7]
    starts = [subject_start(fx, od, oprior, nchain, rng) for _ in seeds]
    st = E.PopState(np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts]), np.stack([s[2] for s in starts]))
    tun = E.Tuning(nmc=4, nchain=nchain, thin=3, nparameter=D, sub_migration_prob=0.2, seeds=seeds)
    both = E.run_subject(fx.ct, tr, prior, tun, st)
    for r, seed in enumerate(seeds):
        one = E.run_subject(fx.ct, tr, prior, E.Tuning(nmc=4, nchain=nchain, thin=3, nparameter=D, sub_migration_prob=0.2, seeds=[seed]),
                            E.PopState(*starts[r]))
        assert np.array_equal(one.theta[0], both.theta[r]) and np.array_equal(one.ll[0], both.ll[r])


def test_subject_shard_equals_slice_of_full_run():
    """Independent-subject engine: running subjects [2, 4) with subject_begin = 2 reproduces the
    same trajectories as the full run (draw addresses use GLOBAL subject ids)."""
    fx = load_fixture(2)
    rng = np.random.default_rng(13)
    S, D = fx.n_pop, fx.ct.npar
    nchain = 3 * D
    prior, oprior = fx.prior("sub_prior"), fx.oprior("sub_prior")
    trials = [fx.trials(f"pop{s}") for s in range(S)]
    starts = [subject_start(fx, fx.odata(f"pop{s}"), oprior, nchain, rng, center=fx.g["ps"][s]) for s in range(S)]
    tun = E.Tuning(nmc=2, nchain=nchain, thin=1, nparameter=D, sub_migration_prob=0.2, seeds=[99])
    full = E.Engine(fx.ct, trials, prior, None, tun, None, [E.PopState(*s) for s in starts])
    full.iterate(6)
    a = full.state()
    tun2 = E.Tuning(nmc=2, nchain=nchain, thin=1, nparameter=D, sub_migration_prob=0.2, seeds=[99], subject_begin=2, n_subject_total=S)
    part = E.Engine(fx.ct, trials[2:], prior, None, tun2, None, [E.PopState(*s) for s in starts[2:]])
    part.iterate(6)
    b = part.state()
    assert np.array_equal(a["theta"][2:], b["theta"]) and np.array_equal(a["ll"][2:], b["ll"])
    assert full.launch_count > 0
    full.close()
    part.close()


def test_error_behaviour():
    """"Require three or more chains." (src/de.cpp:7-10) and argument errors surface as exceptions."""
    fx = load_fixture(2)
    tr = fx.trials("sub")
    prior = fx.prior("sub_prior")
    D = fx.ct.npar
    st = E.PopState(np.ones((2, D)), np.zeros(2), np.zeros(2))
    with pytest.raises(B.GgdmcError) as ei:
        E.run_subject(fx.ct, tr, prior, E.Tuning(nmc=3, nchain=2, nparameter=D), st)
    assert ei.value.code == B.ERR_CHAINS and "three or more chains" in str(ei.value)
    bad = Trials(tr.rt, np.full(len(tr.rt), 9999, np.uint16))
    st = E.PopState(np.ones((3, D)), np.zeros(3), np.zeros(3))
    with pytest.raises(B.GgdmcError) as ei:
        E.run_subject(fx.ct, bad, prior, E.Tuning(nmc=3, nchain=3, nparameter=D), st)
    assert ei.value.code == B.ERR_ARG


def test_results_do_not_depend_on_host_layout():
    """The engine has a no-staging path for inputs that already have the device layout (trials grouped by cell with
    every subject a multiple of 8 trials; start and output arrays of the subjects adjacent in memory) and a general
    path (counting sort + padding, per-subject copies).  Both must give the same bits."""
    from ggdmc_b200 import workloads as W
    w = W.hierarchical("layout", 6, 5, 64, n_replicate=2)
    D, C_ = w.spec.ct.npar, w.nchain
    tun = W.tuning_for(w, nmc=4, thin=3, seeds=[11, 12], pop_migration_prob=0.3, sub_migration_prob=0.3)
    stacked_trials = w.trials
    assert all(len(t.rt) % 8 == 0 and np.all(np.diff(t.cell.astype(int)) >= 0) for t in stacked_trials)  # direct path applies
    phi_a, subj_a = E.run_hier(w.spec.ct, stacked_trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)

    # general path: every subject's trials interleaved round-robin over its cells (a stable grouping by cell restores
    # exactly the original order), separately allocated start states and separately allocated output arrays
    rng = np.random.default_rng(3)
    mixed = []
    for t in stacked_trials:
        cells = [np.nonzero(t.cell == c)[0] for c in np.unique(t.cell)]
        order = [idx[i] for i in range(max(len(c) for c in cells)) for idx in cells if i < len(idx)]
        assert sorted(order) == list(range(len(t.rt))) and not np.all(np.diff(t.cell[order].astype(int)) >= 0)
        mixed.append(Trials(t.rt[order].copy(), t.cell[order].copy()))
    pad = [np.empty(rng.integers(1, 50)) for _ in range(3 * len(mixed))]  # keeps the allocations apart
    starts = [E.PopState(s.theta.copy(), s.lp.copy(), s.ll.copy()) for s in w.subj_start]
    R, nmc = 2, 4
    outs = [E.PopSamples.empty(R, nmc, C_, D) for _ in mixed]
    phi_out = E.PopSamples.empty(R, nmc, C_, 2 * D)
    phi_b, subj_b = E.run_hier(w.spec.ct, mixed, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, starts, out=(phi_out, outs))
    del pad
    assert np.array_equal(phi_a.theta, phi_b.theta) and np.array_equal(phi_a.ll, phi_b.ll) and np.array_equal(phi_a.lp, phi_b.lp)
    for a, b in zip(subj_a, subj_b):
        assert np.array_equal(a.theta, b.theta) and np.array_equal(a.lp, b.lp) and np.array_equal(a.ll, b.ll)
    assert np.all(np.isfinite(phi_a.theta)) and not np.array_equal(phi_a.theta[:, 0], phi_a.theta[:, -1])

    # a ragged subject (61 trials: padding, odd tail) changes only that subject's likelihoods, not the machinery
    ragged = [Trials(t.rt[:61].copy(), t.cell[:61].copy()) if i == 2 else t for i, t in enumerate(stacked_trials)]
    ll2 = E.sumloglike(w.spec.ct, ragged, np.stack([s.theta[0] for s in w.subj_start]))
    ll1 = E.sumloglike(w.spec.ct, stacked_trials, np.stack([s.theta[0] for s in w.subj_start]))
    assert np.array_equal(np.delete(ll1, 2, axis=0), np.delete(ll2, 2, axis=0)) and not np.array_equal(ll1[2], ll2[2])


def test_hierarchical_replicates_batch_equals_separate_runs():
    """The ncore replicates of StartSampling are a batch dimension of one call (R/sampling.R:30-55 forks instead):
    replicate r of a batched hierarchical run must be the run with seed r alone, bit for bit."""
    from ggdmc_b200 import workloads as W
    w = W.hierarchical("reps", 6, 5, 64, n_replicate=2)
    seeds = [21, 22]
    kw = dict(nmc=4, thin=2, pop_migration_prob=0.3, sub_migration_prob=0.3)
    phi_b, subj_b = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, W.tuning_for(w, seeds=seeds, **kw), w.phi_start,
                               w.subj_start)
    for r, seed in enumerate(seeds):
        phi0 = E.PopState(w.phi_start.theta[r:r + 1].copy(), w.phi_start.lp[r:r + 1].copy(), w.phi_start.ll[r:r + 1].copy())
        subj0 = [E.PopState(s.theta[r:r + 1].copy(), s.lp[r:r + 1].copy(), s.ll[r:r + 1].copy()) for s in w.subj_start]
        phi_1, subj_1 = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, W.tuning_for(w, seeds=[seed], **kw), phi0, subj0)
        assert np.array_equal(phi_1.theta[0], phi_b.theta[r]) and np.array_equal(phi_1.ll[0], phi_b.ll[r])
        for a, b in zip(subj_1, subj_b):
            assert np.array_equal(a.theta[0], b.theta[r]) and np.array_equal(a.ll[0], b.ll[r]) and np.array_equal(a.lp[0], b.lp[r])
    assert not np.array_equal(phi_b.theta[0], phi_b.theta[1])


def test_profiling_can_be_switched_mid_fit():
    """Per-launch profiling (bench.py's roofline pass) runs an iteration as plain launches on one stream, with the sweep
    decisions drawn in line instead of on the side stream; a resident fit may switch it on and off at any iteration and
    must end in the same state as an undisturbed one (the groups' own iteration counters stay in step in both orders)."""
    from ggdmc_b200 import workloads as W
    w = W.hierarchical("paths", 6, 7, 64, n_replicate=1)
    tun = W.tuning_for(w, nmc=2, thin=1 << 30, seeds=[77], pop_migration_prob=0.3, sub_migration_prob=0.3)

    def engine():
        return E.Engine(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)

    a = engine()
    a.iterate(60)
    ref = a.state()
    a.close()
    b = engine()
    b.iterate(3)
    b.profile(True)
    b.iterate(2)
    b.profile(False)
    b.iterate(50)  # long enough for the iteration graph to be captured after the switch
    b.profile(True)
    b.iterate(1)
    b.profile(False)
    b.iterate(4)
    got = b.state()
    b.close()
    for k in ("phi_theta", "phi_lp", "phi_ll", "theta", "lp", "ll"):
        assert np.array_equal(ref[k], got[k]), k


def test_execution_paths_agree_bit_for_bit(monkeypatch):
    """The engine's execution strategies are scheduling only: the captured iteration graph vs plain launches, the phi
    sweep on its side stream vs in line, the fused phi half-sweep launch vs its four separate kernels, 1 / 2 / 3 subject
    groups, per-launch priorities on or off, the register-capped build of the short kernels -- every combination must
    produce the same samples, bit for bit."""
    from ggdmc_b200 import workloads as W
    w = W.hierarchical("paths", 6, 7, 64, n_replicate=2)

    def fit(tun):
        phi, subj = E.run_hier(w.spec.ct, w.trials, w.spec.p_prior, w.spec.h_prior, tun, w.phi_start, w.subj_start)
        return [phi.theta.copy(), phi.lp.copy(), phi.ll.copy()] + [a.copy() for s in subj for a in (s.theta, s.lp, s.ll)]

    # a one-shot call of 9 iterations runs plain launches whatever the switch says; 51 iterations are long enough for the
    # iteration graph to be captured (the sweep decisions drawn one iteration ahead, and on the side stream beside the
    # last MH tests, are part of it)
    short = W.tuning_for(w, nmc=4, thin=3, seeds=[31, 32], pop_migration_prob=0.3, sub_migration_prob=0.3)
    long_ = W.tuning_for(w, nmc=18, thin=3, seeds=[31, 32], pop_migration_prob=0.3, sub_migration_prob=0.3)
    envs_short = ({"GGDMC_B200_NO_GRAPH": "1"}, {"GGDMC_B200_NO_OVERLAP": "1"}, {"GGDMC_B200_NO_FUSED_PHI": "1"},
                  {"GGDMC_B200_NO_HI_SMALL": "1"}, {"GGDMC_B200_GROUPS": "1"}, {"GGDMC_B200_GROUPS": "3"},
                  {"GGDMC_B200_NO_SB_ASIDE": "1"}, {"GGDMC_B200_NO_SWEEP_AHEAD": "1"}, {"GGDMC_B200_SHORT_WAVE_WARPS": "0"},
                  {"GGDMC_B200_NO_GRAPH": "1", "GGDMC_B200_NO_FUSED_PHI": "1", "GGDMC_B200_GROUPS": "1", "GGDMC_B200_NO_OVERLAP": "1"})
    envs_long = ({"GGDMC_B200_NO_GRAPH": "1"}, {"GGDMC_B200_NO_SB_ASIDE": "1"}, {"GGDMC_B200_SHORT_WAVE_WARPS": "0"}, {"GGDMC_B200_NO_SWEEP_AHEAD": "1", "GGDMC_B200_GROUPS": "3"},
                 {"GGDMC_B200_NO_OVERLAP": "1", "GGDMC_B200_GROUPS": "2"})
    for tun, envs in ((short, envs_short), (long_, envs_long)):
        base = fit(tun)
        assert np.all(np.isfinite(base[0])) and not np.array_equal(base[0][:, 0], base[0][:, -1])
        for env in envs:
            with monkeypatch.context() as m:
                for k, v in env.items():
                    m.setenv(k, v)
                other = fit(tun)
            for a, b in zip(base, other):
                assert np.array_equal(a, b), env
