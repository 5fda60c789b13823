"""GPU parity tests (-m gpu) of the LBA edge cases the reference's fixtures leave out: the CUDA
path, through the C ABI, against known answers computed by the reference's own object code
(tests/golden/lba_edge_ref.npz), against the oracle with the same addressed `t0 + st0 U` draws, trial by trial through
the SAMPLER's trial loops (n1pdf_fast2 / n1pdf_fast, hot / cold split), and as whole trajectories with st0 free."""
import ctypes as C
import os

import numpy as np
import pytest

from ggdmc_b200 import _lib as B
from ggdmc_b200 import engine as E
from ggdmc_b200.model import CellTable, PriorTable, Trials
from oracle import binding as ob
from helpers import GOLDEN, load_fixture, sane_starts
import lba_edge
from test_lba_edge_cpu import LOOSE, _close
from test_gpu_sampler import SCHEDULES, compare, kernel_path  # noqa: F401  (kernel_path: autouse fixture, both kernel paths)

pytestmark = pytest.mark.gpu
G = dict(np.load(os.path.join(GOLDEN, "lba_edge_ref.npz")))


def _grid_trials(ncell):
    n = len(lba_edge.RT_GRID)
    return Trials(np.tile(lba_edge.RT_GRID, ncell), np.repeat(np.arange(ncell, dtype=np.uint16), n))


@pytest.mark.parametrize("na", [2, 4])
def test_edge_trial_logdens_vs_reference_known_answers(na):
    """ggdmc_b200_trial_logdens on every edge case against lba_class::dlba of de.o: st0 > 0 (the addressed U_ST0 draws),
    every validate_parameters rule, A < 1e-10, sd_v = 0, NaN / inf parameters, rt below / at / just above t0."""
    n = len(lba_edge.RT_GRID)
    n_checked = 0
    for g, (ct, theta, lst, pd) in enumerate(lba_edge.model_for(na)):
        key = f"na{na}_g{g}"
        got = E.trial_logdens(ct, _grid_trials(ct.n_cell), theta[None])[0].reshape(ct.n_cell, n)
        for c, (name, _) in enumerate(lst):
            try:
                _close(np.exp(got[c]), G[f"{key}_dens"][c], LOOSE.get(name, 0.0))
            except AssertionError as e:
                raise AssertionError(f"{key} {name}: {e}")
            n_checked += n
    assert n_checked == 52 * n


def _oracle_cells(ct, theta, pd, seed, pop, iteration, chain):
    """densities of the all-free edge model through the oracle with the engine's addressed st0 draws"""
    L = ob.lib()
    na, rt = ct.n_acc, ob.f64(lba_edge.RT_GRID)
    u = lba_edge.philox_u_st0(L, ob, seed, pop, iteration, chain, ct.n_cell * na)
    out = np.zeros((ct.n_cell, len(rt)))
    for c in range(ct.n_cell):
        P = theta[c * 6 * na:(c + 1) * 6 * na].reshape(6, na).copy()
        P[1] = P[0] + P[1]
        L.orc_lba_cell(ob.ptr(ob.f64(P)), na, ob.ptr(pd, ob.c_u8p), ob.ptr(ob.f64(u[c * na:(c + 1) * na])), ob.ptr(rt), len(rt), ob.ptr(out[c]))
    return out


@pytest.mark.parametrize("na", [2, 4])
def test_edge_cases_through_the_sampler_trial_loops(na):
    """The same cases through the production likelihood code (like_eval: table build with addressed draws, hot loop, cold
    loop, running product): per-trial values and the sums the sampler would see, for three chains (three draw addresses)."""
    n = len(lba_edge.RT_GRID)
    seed, pop, it = 9032, 5, 17
    for g, (ct, theta, lst, pd) in enumerate(lba_edge.model_for(na)):
        th3 = np.stack([theta, theta, theta])
        got, sums = E.trial_logdens_hot(ct, _grid_trials(ct.n_cell), th3, seed=seed, pop=pop, iteration=it)
        for k in range(3):
            ref = _oracle_cells(ct, theta, pd, seed, pop, it, k)
            gk = got[k].reshape(ct.n_cell, n)
            for c, (name, _) in enumerate(lst):
                try:
                    _close(np.exp(gk[c]), ref[c], LOOSE.get(name, 0.0))
                except AssertionError as e:
                    raise AssertionError(f"group {g} chain {k} {name}: {e}")
            # the sum: -inf when a density is exactly 0, else the sum of the per-trial logs (well-conditioned cases to 1e-10)
            lr = np.log(ref)
            tot = lr.sum()
            assert np.isfinite(sums[k]) == np.isfinite(tot)
            if np.isfinite(tot):
                assert abs(sums[k] - gk.sum()) <= 1e-10 * abs(tot)
        if any(nm == "st0_all" for nm, _ in lst):
            c = [nm for nm, _ in lst].index("st0_all")
            a, b = got[0].reshape(ct.n_cell, n)[c], got[1].reshape(ct.n_cell, n)[c]
            assert not np.array_equal(a, b)  # different chains draw different t0 + st0 U


@pytest.mark.parametrize("k", [2, 6, 3, 5])
def test_sampler_trial_loops_per_trial_vs_oracle_on_fixtures(k):
    """North-star check 1 for the function the timed loop runs (n1pdf_fast2 via norm_pairs_stepmajor<4> for 2 accumulators,
    n1pdf_fast for 4), not only for its sums: per-trial log densities <= 1e-10 relative against the oracle."""
    fx = load_fixture(k)
    rng = np.random.default_rng(300 + k)
    n_strict = 0
    for s in range(2):
        tr, od = fx.trials(f"pop{s}"), fx.odata(f"pop{s}")
        thetas = np.concatenate([fx.g["pop_theta_all"][s][:20], sane_starts(fx, 30, rng, center=fx.g["ps"][s])])
        got, sums = E.trial_logdens_hot(fx.ct, tr, thetas)
        generic = E.trial_logdens(fx.ct, tr, thetas)
        for i, th in enumerate(thetas):
            ref = ob.trial_logdens(fx.om, od, th)
            fin = np.isfinite(ref)
            assert np.all(got[i][~fin] < np.log(1e-12))
            strict = fin & (ref > np.log(1e-4))
            err = np.abs(got[i][strict] - ref[strict])
            assert np.all(err <= 1e-10 * np.maximum(np.abs(ref[strict]), 1.0)), (k, s, i, err.max())
            n_strict += int(strict.sum())
            # hot loop vs the generic entry point: same arithmetic up to the lock-step evaluation order
            both = np.isfinite(got[i]) & np.isfinite(generic[i]) & strict
            assert np.all(np.abs(got[i][both] - generic[i][both]) <= 1e-11 * np.maximum(np.abs(ref[both]), 1.0))
            if np.all(fin) and np.all(ref > np.log(1e-12)):
                assert abs(sums[i] - ref.sum()) <= 1e-10 * abs(ref.sum())
    assert n_strict > 5000


def _st0_model(fx):
    """fixture model with the non-decision variability st0 made a free parameter (one more entry of theta)"""
    src = fx.ct.param_src.copy()
    D = fx.ct.npar
    src[:, 4, :] = D
    ct = CellTable(fx.ct.n_acc, fx.ct.n_cell, D + 1, src, fx.ct.const_val, fx.ct.posdrift, list(fx.ct.pnames) + ["st0"], fx.ct.cell_names)
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar)
    p = fx.prior("sub_prior")
    prior = PriorTable(D + 1, np.append(p.p0, 0.0), np.append(p.p1, 0.4), np.append(p.lower, 0.0), np.append(p.upper, 0.0),
                       np.append(p.dist, 6).astype(np.int32), np.append(p.log_p, 1).astype(np.uint8), list(p.pnames) + ["st0"])
    oprior = ob.OPrior(prior.p0, prior.p1, prior.lower, prior.upper, prior.dist, prior.log_p)
    return ct, om, prior, oprior


@pytest.mark.parametrize("schedule,jacobi", SCHEDULES)
def test_run_subject_trajectory_with_st0_free(schedule, jacobi):
    """`t0 + st0 * U` (@hdr/lba.h:117) inside the sampler: with st0 free every likelihood call consumes n_acc draws per
    cell; the engine addresses them (U_ST0, slot = cell * n_acc + accumulator, chain = proposing chain), the oracle
    replays the same addresses -> identical theta trajectories in every schedule."""
    fx = load_fixture(2)
    ct, om, prior, oprior = _st0_model(fx)
    rng = np.random.default_rng(70)
    tr, od = fx.trials("sub"), fx.odata("sub")
    D, nchain, nmc, thin = ct.npar, 3 * ct.npar, 5, 2
    seeds = [9032, 78]
    starts = []
    for seed in seeds:
        th = np.column_stack([sane_starts(fx, nchain, rng), rng.uniform(0.02, 0.12, nchain)])
        lp = np.array([ob.sumlogprior(oprior, th[c]) for c in range(nchain)])
        # start log-likelihoods: any finite values do (the first MH test of a chain compares against them); use st0-free ones
        ll = np.array([ob.sumloglike(om, od, np.append(th[c, :-1], 0.0)) for c in range(nchain)])
        starts.append((th, lp, ll))
    tun = E.Tuning(nmc=nmc, nchain=nchain, thin=thin, nparameter=D, sub_migration_prob=0.3, schedule=schedule, seeds=seeds)
    st = E.PopState(np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts]), np.stack([s[2] for s in starts]))
    out = E.run_subject(ct, tr, prior, tun, st)
    for r, seed in enumerate(seeds):
        pop = ob.OPop(*starts[r], nmc, thin)
        de = ob.make_de(D, nchain, sub_migration_prob=0.3, jacobi=jacobi)
        ob.run_subject(de, pop, oprior, om, od, ob.make_rng(seed=seed), 0, (nmc - 1) * thin)
        compare(out, r, pop, f"run_subject st0 free, seed {seed}")
    assert not np.array_equal(out.theta[0, 0], out.theta[0, -1])
    assert np.ptp(out.theta[0, -1][:, -1]) > 0  # st0 itself moved
