"""Generate the committed golden fixtures under tests/golden/ from the reference's own
known-answer data (run HERE, where /root/reference is mounted; the GPU box only reads the .npz).

Source: /root/reference/tests/testthat/Group1/data/lba_data{2..6}.rda (lba_data0/1 are stale,
SURVEY.md 8c).  Each .rda holds start samples produced by the reference's init path
(R/phi.R:141-204, 256-332), whose `log_likelihoods[,1]` / `summed_log_prior[,1]` are known answers
of the hot-path densities at `theta[,,1]`.

For every fixture we store the flattened model (cell table), the priors, and
  * sub:   the single-subject dmi + its nchain start thetas with ll / lp goldens,
  * pop_k: the first N_POP subjects of the hierarchical start (data, thetas, ll, lp goldens),
  * phi:   phi start thetas with hyper-ll / hyper-prior goldens -- these sum over ALL 32
           subjects, so chain-k thetas of all 32 subjects are stored too (theta only).
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ggdmc_b200.model import build_cell_table, flatten_data, flatten_prior  # noqa: E402
from ggdmc_b200.rda import read_rda  # noqa: E402

SRC = "/root/reference/tests/testthat/Group1/data"
N_POP = {2: 4, 3: 6, 4: 2, 5: 3, 6: 32}


def prior_arrays(prefix, pt, out):
    for f in ("p0", "p1", "lower", "upper", "dist", "log_p"):
        out[f"{prefix}_{f}"] = getattr(pt, f)
    out[f"{prefix}_names"] = np.array(pt.pnames)


def main():
    for k in (2, 3, 4, 5, 6):
        d = read_rda(f"{SRC}/lba_data{k}.rda")
        out = {}
        dmi = d["sub_dmis"][0]
        ct = build_cell_table(dmi["model"], dmi["node_1_index"], dmi["is_positive_drift"])
        out.update(param_src=ct.param_src, const_val=ct.const_val, posdrift=ct.posdrift, pnames=np.array(ct.pnames),
                   cell_names=np.array(ct.cell_names))
        # raw model slots, so the S4 -> cell-table flattening itself can be tested where /root/reference is absent
        mdl = dmi["model"]
        out["model_boolean"] = np.asarray(mdl["model_boolean"]).astype(bool)
        out["pxc_names"] = np.array([str(s) for s in mdl["parameter_x_condition_names"]])
        out["const_names"] = np.array([str(s) for s in mdl["constants"].attrs["names"]])
        out["is_positive_drift"] = np.asarray(dmi["is_positive_drift"]).astype(bool)
        out["node_1_index"] = np.asarray(dmi["node_1_index"]).astype(np.int32)
        out["accumulators"] = np.array([str(a) for a in dmi["model"]["accumulators"]])
        # single subject
        tr = flatten_data(dmi["data"], ct.cell_names)
        ss = d["sub_samples"]
        out.update(sub_rt=tr.rt, sub_cell=tr.cell,
                   sub_theta=np.ascontiguousarray(np.asarray(ss["theta"])[:, :, 0].T),
                   sub_ll=np.asarray(ss["log_likelihoods"])[:, 0], sub_lp=np.asarray(ss["summed_log_prior"])[:, 0])
        prior_arrays("sub_prior", flatten_prior(d["sub_priors"]["p_prior"]), out)
        # hierarchical
        pp = d["pop_priors"]
        prior_arrays("p_prior", flatten_prior(pp["p_prior"]), out)
        prior_arrays("h_prior", flatten_prior(pp["h_prior"]), out)
        ps = d["pop_samples"]
        phi = ps["phi"]
        out.update(phi_theta=np.ascontiguousarray(np.asarray(phi["theta"])[:, :, 0].T),
                   phi_ll=np.asarray(phi["log_likelihoods"])[:, 0], phi_lp=np.asarray(phi["summed_log_prior"])[:, 0])
        subj = ps["subject_theta"]
        all_theta = np.stack([np.asarray(s["theta"])[:, :, 0].T for s in subj])  # [nsubj, nchain, npar]
        out["pop_theta_all"] = all_theta
        npop = N_POP[k]
        for s in range(npop):
            dm = d["pop_dmis"][s]
            ct_s = build_cell_table(dm["model"], dm["node_1_index"], dm["is_positive_drift"])
            assert np.array_equal(ct_s.param_src, ct.param_src), "subjects do not share one model"
            trs = flatten_data(dm["data"], ct.cell_names)
            out[f"pop{s}_rt"] = trs.rt
            out[f"pop{s}_cell"] = trs.cell
            out[f"pop{s}_ll"] = np.asarray(subj[s]["log_likelihoods"])[:, 0]
            out[f"pop{s}_lp"] = np.asarray(subj[s]["summed_log_prior"])[:, 0]
        out["n_pop"] = np.array(npop)
        # the hyper_dmi data matrix (nsubject x npar "true" thetas) for run_hyper
        out["hyper_data"] = np.asarray(d["hyper_dmi"]["data"], dtype=np.float64)
        # generating values of the recovery study (README.md:44-66): handy as sane start points
        def by_pnames(v):  # named numeric vector -> model pnames order
            names = list(v.names)
            return np.array([float(np.asarray(v)[names.index(n)]) for n in ct.pnames])
        out["p_vector"] = by_pnames(d["p_vector"])
        out["pop_mean"] = by_pnames(d["pop_mean"])
        out["pop_scale"] = by_pnames(d["pop_scale"])
        out["ps"] = np.asarray(d["ps"], dtype=np.float64)
        path = os.path.join(ROOT, "tests", "golden", f"lba_data{k}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB", "cells", ct.n_cell, "acc", ct.n_acc, "npar", ct.npar)


if __name__ == "__main__":
    main()
