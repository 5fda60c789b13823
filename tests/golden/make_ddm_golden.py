"""Generate tests/golden/ddm_ref.npz: DDM ("fastdm") trial densities computed by the REFERENCE'S OWN OBJECT CODE
(likelihood_class::ddm_likelihood of /root/reference/src/de.o, driven through oracle/ref_harness2.cpp) on seeded
inputs.  Run HERE, where /root/reference is mounted; everywhere else (the GPU box) the tests only read the .npz.

The reference ships no DDM fixture (its Group5 scripts load a ddm_data0.rda that is not in the repository), so these
vectors are the known answers for the DDM path: the oracle must reproduce them bit for bit
(tests/test_ddm_cpu.py::test_ddm_golden_vectors), the CUDA kernel within its tolerance
(tests/test_gpu_ddm.py::test_ddm_trial_logdens_vs_reference_golden_vectors).

Inputs are exactly those of tests/test_gpu_ddm.py::test_ddm_trial_logdens_vs_oracle: per (precision, s) setting a
360-trial data set (short / typical / very long response times over the four cells) and 48 parameter vectors cycling
through the four variability combinations.
Usage: python tests/golden/make_ddm_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import binding as ob  # noqa: E402
from helpers import ddm_model, ddm_theta  # noqa: E402
from test_ddm_cpu import _edge_thetas, _grid_data  # noqa: E402

SETTINGS = [(3.0, 1.0), (2.5, 1.0), (3.0, 0.1)]


def inputs(precision, s):
    rng = np.random.default_rng(int(10 * precision + 100 * s))
    od = _grid_data(rng, 360)
    thetas = []
    for it in range(48):
        th = ddm_theta(rng, it % 4)
        if s != 1.0:
            for i in (0, 2, 5, 6):
                th[i] *= s
            th[7] = th[0] / s * 0.5
            th[3] = min(th[3], 0.2 * th[0] / s)
        thetas.append(th)
    return od, np.stack(thetas)


def main():
    assert ob.ref_lib() is not None, "needs /root/reference/src/de.o"
    out = {"settings": np.array(SETTINGS)}
    for k, (precision, s) in enumerate(SETTINGS):
        ct, om = ddm_model(precision, s)
        od, thetas = inputs(precision, s)
        if k == 0:  # the edge cases ride along with the first setting
            thetas = np.concatenate([thetas, np.stack(_edge_thetas(np.random.default_rng(3)))])
        dens = np.stack([ob.ref2_ddm_density(om, od, th) for th in thetas])
        out[f"rt{k}"], out[f"cell{k}"], out[f"theta{k}"], out[f"dens{k}"] = od.rt, od.cell, thetas, dens
    np.savez_compressed(os.path.join(HERE, "ddm_ref.npz"), **out)
    print("wrote ddm_ref.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
