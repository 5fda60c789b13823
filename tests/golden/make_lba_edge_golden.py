"""Generate tests/golden/lba_edge_ref.npz: LBA node-1 densities of the edge cases in tests/lba_edge.py computed by the
REFERENCE'S OWN OBJECT CODE -- lba_class::set_parameters, validate_parameters and dlba of /root/reference/src/de.o,
driven through oracle/ref_harness.cpp::ref_lba_cell, plus the invalid-cell rule of @hdr/likelihood.h:105 -- with the
uniforms of `t0 + st0 * Rf_runif(0, 1)` (@hdr/lba.h:117) injected.  Run HERE, where /root/reference is mounted;
everywhere else (the GPU box) the tests only read the .npz.

The injected uniforms are the counter-addressed draws the CUDA engine itself uses for a cell table built at address
(seed 0, population 0, iteration 0, sweep 0, chain 0), slot = cell * n_acc + accumulator, so the golden densities can
be compared with ggdmc_b200_trial_logdens directly.  The fixtures of the reference all have st0 = 0, strictly positive
parameters and A >> 1e-10; these vectors pin what they leave out.
Usage: python tests/golden/make_lba_edge_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import binding as ob  # noqa: E402
import lba_edge  # noqa: E402


def main():
    R, L = ob.ref_lib(), ob.lib()
    assert R is not None, "needs /root/reference/src/de.o"
    out = {"rt": lba_edge.RT_GRID}
    for na in (2, 4):
        for g, (ct, theta, lst, pd) in enumerate(lba_edge.model_for(na)):
            u_all = lba_edge.philox_u_st0(L, ob, 0, 0, 0, 0, ct.n_cell * na)
            dens, valid = [], []
            for c, (name, P) in enumerate(lst):
                Pb = P.copy()
                Pb[1] = Pb[0] + Pb[1]  # design_light.h:336-340
                u = u_all[c * na:(c + 1) * na]
                stream = ob.f64(np.concatenate([[0.5, 0.5], u]))  # 2 draws of the constructor's dummy parameters, then n_acc
                R.ref_set_uniform_stream(ob.ptr(stream), len(stream))
                o = np.zeros(len(lba_edge.RT_GRID))
                v = R.ref_lba_cell(ob.ptr(ob.f64(Pb)), na, ob.ptr(pd, ob.c_u8p), ob.ptr(ob.f64(lba_edge.RT_GRID)), len(o), ob.ptr(o))
                assert R.ref_uniform_stream_pos() == 2 + na
                dens.append(o)
                valid.append(v)
            key = f"na{na}_g{g}"
            out[f"{key}_names"] = np.array([n for n, _ in lst])
            out[f"{key}_P"] = np.stack([P for _, P in lst])
            out[f"{key}_posdrift"] = pd
            out[f"{key}_u"] = u_all
            out[f"{key}_valid"] = np.array(valid, np.int32)
            out[f"{key}_dens"] = np.stack(dens)
    np.savez_compressed(os.path.join(HERE, "lba_edge_ref.npz"), **out)
    print("wrote lba_edge_ref.npz:", {k: v.shape for k, v in out.items()})
    for k in out:
        if k.endswith("_valid"):
            print(k, int(out[k].sum()), "valid of", len(out[k]))


if __name__ == "__main__":
    main()
