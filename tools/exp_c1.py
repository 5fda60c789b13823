"""BASELINE config 1 (README single-subject LBA B x v model: 13 parameters, 768 trials, 39 chains, nmc 500, thin 8,
sub_migration_prob 0.06, 3 replicates): the whole StartSampling_subject job through ggdmc_b200_run_subject (replicates
batched) next to de_class::run_chains of the reference's own object code on one host core per replicate."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggdmc_b200 import engine as E, synth, workloads as W
from oracle import binding as ob

spec = W.load_model(6)
ct = spec.ct
D, C, nmc, thin, R = ct.npar, 3 * ct.npar, 500, 8, 3
rng = np.random.default_rng(20260101)
theta_true = synth.rtnorm(spec.pop_mean, spec.pop_scale, 0.0, rng)
tr = synth.simulate_subject(ct, spec.node_1_index, theta_true, 768, rng)
x0 = np.abs(theta_true[None, None, :] * (1.0 + 0.05 * rng.standard_normal((R, C, D))))
ll0 = E.sumloglike(ct, [tr], x0.reshape(1, R * C, D)).reshape(R, C)
lp0 = E.sumlogprior(spec.sub_prior, x0.reshape(R * C, D)).reshape(R, C)
tun = E.Tuning(nmc=nmc, nchain=C, thin=thin, nparameter=D, sub_migration_prob=0.06, seeds=[9032, 9033, 9034])
E.run_subject(ct, tr, spec.sub_prior, E.Tuning(nmc=3, nchain=C, thin=2, nparameter=D, sub_migration_prob=0.06, seeds=[1, 2, 3]),
              E.PopState(x0, lp0, ll0))  # warm-up: context, module load, memory pool
t0 = time.perf_counter()
out = E.run_subject(ct, tr, spec.sub_prior, tun, E.PopState(x0, lp0, ll0))
dt = time.perf_counter() - t0
n_iter = (nmc - 1) * thin
n_lik = R * n_iter * C * 768 * (0.94 + 0.06 * 0.5)
print(f"ggdmc_b200_run_subject: {R} replicates x {n_iter} iterations x {C} chains x 768 trials in {dt*1e3:.0f} ms "
      f"({R * n_iter / dt:.0f} replicate-iterations/s, {n_lik / dt:.3e} trial-likelihoods/s, host buffers in and out)")
assert np.all(np.isfinite(out.theta))
if ob.ref_lib() is not None:
    om = ob.OModel(ct.param_src, ct.const_val, ct.posdrift, ct.npar)
    od = ob.OData(tr.rt, tr.cell)
    pr = spec.sub_prior
    op = ob.OPrior(pr.p0, pr.p1, pr.lower, pr.upper, pr.dist, pr.log_p)
    ob.ref2_prime()
    ob.ref_lib().ref_set_uniform_stream(None, 0)
    it_ref = 60
    t0 = time.perf_counter()
    ob.ref2_run_chains(D, om, od, op, x0[0], lp0[0], ll0[0], it_ref + 1, 1, sub_migration_prob=0.06)
    dtr = time.perf_counter() - t0
    print(f"reference object code (src/de.o, run_chains, 1 core): {it_ref / dtr:.1f} iterations/s per replicate -> the same job "
          f"({n_iter} iterations, one core per replicate) would take {n_iter / (it_ref / dtr):.0f} s; speed-up {n_iter / (it_ref / dtr) / dt:.0f}x")
