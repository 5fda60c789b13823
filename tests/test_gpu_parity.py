"""GPU parity tests (run with -m gpu on a B200): the CUDA engine, called through the C ABI,
against the CPU oracle on the same seeded inputs and against the reference's golden vectors."""
import ctypes as C

import numpy as np
import pytest

from ggdmc_b200 import _lib as B
from ggdmc_b200 import engine as E
from ggdmc_b200.model import PriorTable, Trials
from oracle import binding as ob
from helpers import cond_mask_tolerance, load_fixture, sane_starts

pytestmark = pytest.mark.gpu


def test_philox_device_known_answers():
    assert E.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert E.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert E.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---- check 1: per-trial log densities on fixed theta arrays ---------------------------------
@pytest.mark.parametrize("k", [2, 3, 5, 6])
def test_trial_logdens_vs_oracle(k):
    fx = load_fixture(k)
    rng = np.random.default_rng(100 + k)
    tr, od = fx.trials("pop0"), fx.odata("pop0")
    thetas = np.concatenate([fx.g["pop_theta_all"][0][:30], sane_starts(fx, 30, rng), sane_starts(fx, 10, rng, jitter=0.3)])
    got = E.trial_logdens(fx.ct, tr, thetas)
    n_strict = 0
    for i, th in enumerate(thetas):
        ref = ob.trial_logdens(fx.om, od, th)
        fin = np.isfinite(ref)
        # exact zeros of the reference ((1 - cdf) clamped to 0) must be zero-or-tiny here as well
        assert np.all(got[i][~fin] < np.log(1e-12))
        tol = cond_mask_tolerance(ref[fin])
        err = np.abs(got[i][fin] - ref[fin])
        bad = err > tol
        assert not bad.any(), (k, i, np.where(bad)[0][:5], err[bad][:5], ref[fin][bad][:5])
        strict = ref[fin] > np.log(1e-4)
        n_strict += int(strict.sum())
        assert np.all(err[strict] <= 1e-10 * np.maximum(np.abs(ref[fin][strict]), 1.0))
    assert n_strict > 1000


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
def test_sumloglike_vs_oracle_and_goldens(k):
    fx = load_fixture(k)
    S = fx.n_pop
    trials = [fx.trials(f"pop{s}") for s in range(S)]
    theta = fx.g["pop_theta_all"][:S]
    got = E.sumloglike(fx.ct, trials, theta)
    n_clean = 0
    for s in range(S):
        od = fx.odata(f"pop{s}")
        for c in range(theta.shape[1]):
            ld = ob.trial_logdens(fx.om, od, theta[s, c])
            ref = ob.sumloglike(fx.om, od, theta[s, c])
            if np.all(ld > np.log(1e-12)):
                assert abs(got[s, c] - ref) <= 1e-10 * abs(ref), (s, c, got[s, c], ref)
                gold = fx.g[f"pop{s}_ll"][c]
                assert abs(got[s, c] - gold) <= 1e-10 * abs(gold)  # the reference's own known answer
                n_clean += 1
            elif np.isfinite(ref):
                assert abs(got[s, c] - ref) <= 0.05 * abs(ref)
    assert n_clean > 0.7 * S * theta.shape[1]


def test_sumloglike_sane_region_strict():
    """In the region the sampler actually visits (near the generating values) the sums agree to 1e-12."""
    fx = load_fixture(6)
    rng = np.random.default_rng(5)
    S = 8
    trials = [fx.trials(f"pop{s}") for s in range(S)]
    theta = np.stack([sane_starts(fx, 39, rng, center=fx.g["ps"][s]) for s in range(S)])
    got = E.sumloglike(fx.ct, trials, theta)
    for s in range(S):
        od = fx.odata(f"pop{s}")
        for c in range(39):
            ref = ob.sumloglike(fx.om, od, theta[s, c])
            assert abs(got[s, c] - ref) <= 1e-12 * abs(ref)


def test_sumloglike_properties_full_size():
    """BASELINE config-4 shapes (768 trials x 78 chains, 256 of the 1024 subjects here): additivity over
    trials, invariance to trial order, chain-order equivariance."""
    fx = load_fixture(6)
    rng = np.random.default_rng(9)
    S, Cn = 256, 78
    base = [fx.trials(f"pop{s % 32}") for s in range(S)]
    theta = np.stack([sane_starts(fx, Cn, rng, center=fx.g["ps"][s % 32], jitter=0.08) for s in range(S)])
    full = E.sumloglike(fx.ct, base, theta)
    fin = np.isfinite(full)  # -inf is legitimate: a survivor CDF clamped to 1 makes a density exactly 0
    assert fin.mean() > 0.5 and not np.isnan(full).any()

    def close(a, b, tol):
        both = np.isfinite(a) & np.isfinite(b)
        # sums of ~768 log densities of size ~1..20 that may cancel to ~0: absolute + relative bound
        return np.array_equal(np.isfinite(a), np.isfinite(b)) and np.all(np.abs(a[both] - b[both]) <= 1e-10 + tol * np.abs(b[both]))

    # subjects 32.. repeat the data of subjects 0..31 with other thetas; same theta -> same value
    again = E.sumloglike(fx.ct, base[:32], theta[32:64])
    assert np.array_equal(again, full[32:64])
    # additivity: split every subject's trials in two halves
    h1 = [Trials(t.rt[::2], t.cell[::2]) for t in base]
    h2 = [Trials(t.rt[1::2], t.cell[1::2]) for t in base]
    s12 = E.sumloglike(fx.ct, h1, theta) + E.sumloglike(fx.ct, h2, theta)
    assert close(s12, full, 1e-13)
    # trial order does not matter beyond rounding
    perm = [rng.permutation(len(t.rt)) for t in base]
    shuf = [Trials(t.rt[p], t.cell[p]) for t, p in zip(base, perm)]
    assert close(E.sumloglike(fx.ct, shuf, theta), full, 1e-13)
    # chain order equivariance (bit exact: each chain is an independent block)
    rev = E.sumloglike(fx.ct, base, theta[:, ::-1])
    assert np.array_equal(rev[:, ::-1], full)


# ---- priors ------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 6])
def test_prior_and_hyperlikelihood_vs_goldens(k):
    fx = load_fixture(k)
    pp, hp = fx.prior("p_prior"), fx.prior("h_prior")
    phi = fx.g["phi_theta"]
    npar = phi.shape[1] // 2
    allth = fx.g["pop_theta_all"]
    # h_prior on phi
    got = E.sumlogprior(hp, phi)
    assert np.max(np.abs(got - fx.g["phi_lp"])) <= 1e-12 * np.max(np.abs(fx.g["phi_lp"]))
    # subject prior goldens
    for s in range(fx.n_pop):
        got = E.sumlogprior(pp, allth[s])
        assert np.max(np.abs(got - fx.g[f"pop{s}_lp"])) <= 1e-12 * np.max(np.abs(got))
    # hyper-likelihood of chain c = sum over subjects of the phi-driven tnorm log density
    tot = np.zeros(phi.shape[0])
    for s in range(allth.shape[0]):
        tot += E.sumlogprior(pp, allth[s], phi[:, :npar], phi[:, npar:])
    assert np.max(np.abs(tot - fx.g["phi_ll"]) / np.abs(fx.g["phi_ll"])) <= 1e-12


def test_all_prior_families_vs_oracle():
    rng = np.random.default_rng(0)
    n = 4000
    for dist in (1, 2, 3, 4, 5, 6, 7):
        for lg in (0, 1):
            lo, up = (0.0, 10.0) if dist == 2 else (float(rng.choice([0.0, -np.inf])), float(rng.choice([np.inf, 10.0])))
            pt = PriorTable(1, np.array([1.3]), np.array([0.8]), np.array([lo]), np.array([up]), np.array([dist], np.int32),
                            np.array([lg], np.uint8), ["x"])
            x = rng.uniform(-1, 11, size=(n, 1))
            p0, p1 = rng.uniform(0.2, 4, size=(n, 1)), rng.uniform(-0.1, 3, size=(n, 1))
            got = E.sumlogprior(pt, x, p0, p1)
            op = ob.OPrior([1.3], [0.8], [lo], [up], [dist], [lg])
            for i in range(0, n, 7):
                ref = ob.sumlogprior(op, x[i], p0[i], p1[i])
                assert (np.isnan(ref) and np.isnan(got[i])) or ref == got[i] or abs(ref - got[i]) <= 1e-11 * max(1.0, abs(ref)), \
                    (dist, lg, x[i], p0[i], p1[i], ref, got[i])


# ---- check 2: chain selection is bit exact given the same uniforms ----------------------------
def test_chain_selection_bit_exact():
    rng = np.random.default_rng(42)
    L = ob.lib()
    R = ob.ref_lib()
    for nchain in (3, 4, 15, 39, 78, 102, 300):
        n = 200
        k = rng.integers(0, nchain, size=n).astype(np.int32)
        up = rng.uniform(size=(n, nchain - 1))
        um = rng.uniform(size=(n, nchain + 1))
        um[:5, 0] = [1e-9, 0.999999, 0.5, 1.0 / nchain, 2.0 / nchain]
        op, om, on = E.select_chains(nchain, k, up, um)
        for i in range(n):
            r = ob.make_rng(stream=up[i])
            o = (C.c_uint * 2)()
            L.orc_get_chains(nchain, int(k[i]), 2, C.byref(r), C.byref(ob.Addr()), o)
            assert list(o) == list(op[i])
            r = ob.make_rng(stream=um[i])
            s = (C.c_uint * nchain)()
            nn = L.orc_get_subchains(nchain, C.byref(r), C.byref(ob.Addr()), s)
            assert nn == on[i] and list(s[:nn]) == list(om[i][:nn]) and np.all(om[i][nn:] == -1)
            if R is not None and i < 40:  # the reference's own machine code
                u = ob.f64(up[i])
                R.ref_set_uniform_stream(ob.ptr(u), len(u))
                o2 = (C.c_uint * 2)()
                R.ref_get_chains(nchain, int(k[i]), 2, o2)
                assert list(o2) == list(op[i])
                u = ob.f64(um[i])
                R.ref_set_uniform_stream(ob.ptr(u), len(u))
                s2 = (C.c_uint * nchain)()
                n2 = R.ref_get_subchains(nchain, s2)
                assert n2 == on[i] and list(s2[:n2]) == list(om[i][:n2])


def test_likelihood_sweep_config3_shapes():
    """BASELINE config 3 (likelihood-only sweep: 5-parameter 2-accumulator model, 15 chains, 1e3 .. 1e7 trials of ONE
    subject): the trial-chunked grid (a subject with > 8192 trials is split over blocks) against the oracle at the
    sizes the oracle finishes in seconds, and at 1e7 trials through additivity / tiling."""
    from ggdmc_b200 import synth
    from ggdmc_b200 import workloads as W
    ct, node_1, p_vector, _ = W.sweep_model()
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar)
    rng = np.random.default_rng(20260103)
    theta = p_vector * (1.0 + 0.05 * rng.uniform(-1, 1, size=(15, 5)))  # p_vector +- 5 % per chain
    base = synth.simulate_subject(ct, node_1, p_vector, 100_000, rng)
    for n in (1_000, 10_000, 100_000):
        sub = Trials(base.rt[:: 100_000 // n][:n].copy(), base.cell[:: 100_000 // n][:n].copy())
        got = E.sumloglike(ct, [sub], theta[None])[0]
        od = ob.OData(sub.rt, sub.cell)
        ref = np.array([ob.sumloglike(om, od, th) for th in theta])
        assert np.all(np.isfinite(ref)) and np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-12, (n, got, ref)
        # per-trial log densities of the first chain at this size
        ld = E.trial_logdens(ct, sub, theta[:1])[0]
        assert abs(ld.sum() - ref[0]) <= 1e-10 * abs(ref[0])
    ll_1e5 = E.sumloglike(ct, [base], theta[None])[0]
    # 1e6 and 1e7 trials = the 1e5 block tiled: the sum must scale (each block keeps its own fixed-order partial sums)
    for reps in (10, 100):
        big = Trials(np.tile(base.rt, reps), np.tile(base.cell, reps))
        got = E.sumloglike(ct, [big], theta[None])[0]
        assert np.max(np.abs(got - reps * ll_1e5) / np.abs(reps * ll_1e5)) <= 1e-11, reps
    # ragged split of 1e6 trials into 7 subjects: additivity across subjects and chunk boundaries
    big = Trials(np.tile(base.rt, 10), np.tile(base.cell, 10))
    cuts = np.sort(rng.choice(np.arange(1, 1_000_000), size=6, replace=False))
    parts = [Trials(r.copy(), c.copy()) for r, c in zip(np.split(big.rt, cuts), np.split(big.cell, cuts))]
    per = E.sumloglike(ct, parts, np.broadcast_to(theta, (7, 15, 5)).copy())
    assert np.max(np.abs(per.sum(axis=0) - 10 * ll_1e5) / np.abs(10 * ll_1e5)) <= 1e-11
