"""BASELINE config 2 -- the README's hierarchical recovery study (32 subjects x 768 trials, B x v model, 78 chains) with the
README's three stages (README.md:181-196: StartSampling thin 8, sub migration 0.06; RestartSampling nmc 1000, pop migration
0.05; RestartSampling nmc 1000, pop migration 0.01) in the REFERENCE schedule (= the reference's own trajectories, see
tests/test_gpu_sampler.py) and in the default PARALLEL schedule, 3 replicates each (the README's ncore = 3).  Prints what
tests/test_gpu_posterior.py::test_readme_recovery_study_c2 asserts: R-hat by the package's own definition
(R/model-class.R:1559-1690), agreement of the two arms in units of the replicate-level Monte-Carlo standard error, and
the recovery of the generating population means.

    python tools/exp_c2_recovery.py [n_extra_stages]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from ggdmc_b200 import _lib as B, engine as E, workloads as W  # noqa: E402
from recovery import gelman_pkg, run_stages, zscores  # noqa: E402

R = 3
extra = int(sys.argv[1]) if len(sys.argv) > 1 else 10
w = W.hierarchical("c2", 6, 32, 768, n_replicate=R)
D = w.spec.ct.npar


def report(name, stages):
    for i, (phi, subj) in enumerate(stages):
        half = phi.shape[1] // 2
        rh = np.array([gelman_pkg(phi[r, half:]) for r in range(R)])
        print(f"{name} stage {i}: phi R-hat (second half of the stage) max {rh.max():.3f} median {np.median(rh):.3f}; "
              f"location means {np.round(phi[:, half:, :, :D].mean((0, 1, 2)), 3)}", flush=True)


# 1. the README's stages in both schedules, from the same start values
for name, sched, seed0 in (("parallel", B.SCHEDULE_PARALLEL, 9032), ("reference", B.SCHEDULE_REFERENCE, 5000)):
    t0 = time.perf_counter()
    st, state = run_stages(w, sched, [seed0 + r for r in range(R)])
    print(f"{name}: README stages in {time.perf_counter() - t0:.1f} s")
    report(name + " README", st)
    if name == "parallel":
        readme_state = state
print("generating population means", np.round(w.spec.pop_mean, 3))
# 2. keep going in the fast schedule: one long stage (default 6000 x 8 iterations)
R4 = R
st, conv = run_stages(w, B.SCHEDULE_PARALLEL, [77 + r for r in range(R)], stages=[(1000 * extra // 2 + 1, 8, 0.0, 0.01)], start=readme_state)
phi_long = st[0][0]
for lo in range(0, phi_long.shape[1], 1000):
    rh = np.array([gelman_pkg(phi_long[r, lo:lo + 1000]) for r in range(R)])
    print(f"parallel, continued, samples {lo}-{lo + 1000}: R-hat max {rh.max():.3f} median {np.median(rh):.3f}", flush=True)
# 3. from that state: both schedules, fresh seeds
t0 = time.perf_counter()
pa, _ = run_stages(w, B.SCHEDULE_PARALLEL, [300 + r for r in range(R)], stages=[(4001, 8, 0.0, 0.01)], start=conv)
ra, _ = run_stages(w, B.SCHEDULE_REFERENCE, [600 + r for r in range(R)], stages=[(1001, 8, 0.0, 0.01)], start=conv)
print(f"comparison arms in {time.perf_counter() - t0:.1f} s")
pp, ps = pa[0][0], pa[0][1]
for n_use in (500, 1000, 2000, 4000):
    rh = np.array([gelman_pkg(pp[r, :n_use]) for r in range(R)])
    print(f"parallel arm, first {n_use} samples: R-hat max {rh.max():.3f} (parameter {np.unravel_index(rh.argmax(), rh.shape)[1]}) median {np.median(rh):.3f}; "
          f"subject 0: {max(gelman_pkg(ps[0][r, :n_use]).max() for r in range(R)):.3f}")
rh = np.array([gelman_pkg(ra[0][0][r]) for r in range(R)])
print(f"reference arm, 1000 samples: R-hat max {rh.max():.3f} median {np.median(rh):.3f}; subject 0: {max(gelman_pkg(ra[0][1][0][r]).max() for r in range(R)):.3f}")
for nm, a, b in (("phi", pp, ra[0][0]), ("subject 0", ps[0], ra[0][1][0]), ("subject 15", ps[1], ra[0][1][1]), ("subject 31", ps[2], ra[0][1][2])):
    z = zscores(a, b)
    print(f"parallel vs reference, {nm}: fraction of summaries within 2 MCSE {np.mean(z <= 2):.2f}, max z {z.max():.2f}, sorted top {np.round(np.sort(z.ravel())[-4:], 2)}")
flat = pp.reshape(-1, 2 * D)
print("phi location posterior mean", np.round(flat.mean(0)[:D], 3), "sd", np.round(flat.std(0)[:D], 3))
print("generating - mean, in sd:", np.round((w.spec.pop_mean - flat.mean(0)[:D]) / flat.std(0)[:D], 2))
