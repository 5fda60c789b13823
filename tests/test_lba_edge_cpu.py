"""CPU tests of the LBA edge cases: the oracle against known answers computed by the reference's
own object code (tests/golden/lba_edge_ref.npz, made by tests/golden/make_lba_edge_golden.py), and the engine's device
headers, built for the host, against both -- including the trial functions of the sampler's hot loop, trial by trial."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from helpers import GOLDEN, load_fixture, sane_starts
import lba_edge

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = dict(np.load(os.path.join(GOLDEN, "lba_edge_ref.npz")))


def groups():
    for na in (2, 4):
        for g in range(3):
            key = f"na{na}_g{g}"
            yield na, key, G[f"{key}_names"], G[f"{key}_P"], G[f"{key}_posdrift"], G[f"{key}_u"], G[f"{key}_valid"], G[f"{key}_dens"]


def test_cases_match_the_golden_file():
    """The committed vectors belong to the committed case list (names, parameters, uniforms)."""
    L = ob.lib()
    for na in (2, 4):
        for g, (ct, theta, lst, pd) in enumerate(lba_edge.model_for(na)):
            key = f"na{na}_g{g}"
            assert [n for n, _ in lst] == [str(s) for s in G[f"{key}_names"]]
            assert np.array_equal(np.stack([P for _, P in lst]), G[f"{key}_P"], equal_nan=True)
            assert np.array_equal(pd, G[f"{key}_posdrift"])
            assert np.array_equal(lba_edge.philox_u_st0(L, ob, 0, 0, 0, 0, ct.n_cell * na), G[f"{key}_u"])
    assert np.array_equal(G["rt"], lba_edge.RT_GRID)


def test_oracle_reproduces_reference_edge_densities_bitwise():
    """orc_lba_cell == lba_class::{set_parameters, validate_parameters, dlba} of de.o on every edge case, bit for bit,
    without the reference mounted: st0 > 0 with injected uniforms, each validity rule, point-mass start points,
    sd_v = 0, NaN / inf parameters, rt below / at / just above t0."""
    L = ob.lib()
    rt = ob.f64(G["rt"])
    n_cases = n_nonfloor = 0
    for na, key, names, Ps, pd, u, valid, dens in groups():
        for c in range(len(names)):
            Pb = Ps[c].copy()
            Pb[1] = Pb[0] + Pb[1]
            o = np.zeros(len(rt))
            v = L.orc_lba_cell(ob.ptr(ob.f64(Pb)), na, ob.ptr(pd, ob.c_u8p), ob.ptr(ob.f64(u[c * na:(c + 1) * na])), ob.ptr(rt), len(rt),
                               ob.ptr(o))
            assert v == valid[c], (key, names[c])
            assert np.array_equal(o, dens[c]), (key, names[c], o, dens[c])
            n_cases += 1
            n_nonfloor += int(np.sum(dens[c] != 1e-10))
    assert n_cases == 104 and n_nonfloor > 500
    # the cases do what their names say
    by = {(na, str(n)): (v, d) for na, key, names, Ps, pd, u, valid, dens in groups() for n, v, d in zip(names, valid, dens)}
    for na in (2, 4):
        assert all(by[(na, f"invalid_{r}_{j}")][0] == 0 for r in ("A_neg", "b_neg", "b_lt_A", "sdv_neg", "st0_neg", "t0_neg") for j in (0, na - 1))
        assert all(np.all(by[(na, f"invalid_{r}_0")][1] == 1e-10) for r in ("A_neg", "t0_neg"))
        assert by[(na, "b_equals_A")][0] == 1 and by[(na, "sdv_zero_winner")][0] == 1 and by[(na, "nan_A_0")][0] == 1
        assert not np.array_equal(by[(na, "st0_all")][1], by[(na, "regular")][1])  # the draws moved t0
        assert abs(by[(na, "regular")][1][0] - 1e-10) < 1e-19  # rt < t0: the floor, times (1 - floor) per survivor


@pytest.fixture(scope="module")
def hostmath():
    out = os.path.join(ROOT, "tests", "host", "libhostmath_edge.so")
    src = os.path.join(ROOT, "tests", "host", "host_math_harness.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-x", "c++", src, "-o", out], check=True)
    return C.CDLL(out)


# Cases in which the reference's own formula is ill-conditioned: with a start-point range A of 1e-10 (1e-8) the general
# branch divides Phi(z1) - Phi(z2), two numbers that agree to ~10 (8) digits, by A: two correct FP64 implementations of
# Phi differ by eps * b / A and more where Phi >> phi (@hdr/lba.h:235-244, 331-337).  Everything else is held to 1e-10.
LOOSE = {"A_at_threshold": 2e-2, "A_small": 2e-4}


def _close(got, ref, extra_rel=0.0):
    """density-level agreement: relative 1e-10 on the log density where it is well conditioned, absolute ~1e-15 on the
    density where the reference's own 1 - cdf cancels (helpers.cond_mask_tolerance), identical zeros / floors"""
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    z = ref == 0.0
    assert np.all(got[z] < 1e-12)
    f = ref == 1e-10
    assert np.all(np.abs(got[f] - 1e-10) <= 1e-22) or np.all(np.abs(np.log(got[f]) - np.log(1e-10)) < 1e-9)
    ok = ~z & ~f & np.isfinite(ref)
    lr, lg = np.log(ref[ok]), np.log(got[ok])
    tol = 1e-10 * np.maximum(np.abs(lr), 1.0) + 4e-15 / np.maximum(ref[ok], 1e-300) + extra_rel
    assert np.all(np.abs(lg - lr) <= tol), (got[ok], ref[ok])


def test_device_headers_on_host_reproduce_edge_densities(hostmath):
    """gg_lba.cuh (cellacc_build, cell_class_update, n1pdf_any: the table build and every branch of the generic path)
    against the reference's known answers."""
    H = hostmath
    rt = ob.f64(G["rt"])
    for na, key, names, Ps, pd, u, valid, dens in groups():
        for c in range(len(names)):
            o = np.zeros(len(rt))
            cls = H.hm_lba_cell_u(ob.ptr(ob.f64(Ps[c])), na, ob.ptr(pd, ob.c_u8p), ob.ptr(ob.f64(u[c * na:(c + 1) * na])), ob.ptr(rt), len(rt),
                                  ob.ptr(o))
            assert (cls != 1) == bool(valid[c]), (key, names[c], cls)
            try:
                _close(o, dens[c], LOOSE.get(str(names[c]), 0.0))
            except AssertionError as e:
                raise AssertionError(f"{key} {names[c]}: {e}")


@pytest.mark.parametrize("k", [2, 6])
def test_hot_loop_trial_functions_per_trial_vs_oracle(hostmath, k):
    """n1pdf_fast2<2> -- what the sampler's 2-accumulator trial loop runs -- trial by trial against the oracle on the
    fixture data: <= 1e-10 relative on the log density (north-star check 1 for the production function, not a sum)."""
    H, L = hostmath, ob.lib()
    fx = load_fixture(k)
    assert fx.ct.n_acc == 2
    rng = np.random.default_rng(k)
    n_fast2 = 0
    worst = 0.0
    for s in range(3):
        rt, cell = fx.g[f"pop{s}_rt"], fx.g[f"pop{s}_cell"]
        for th in list(fx.g["pop_theta_all"][s][::13]) + list(sane_starts(fx, 8, rng, center=fx.g["ps"][s])):
            th = ob.f64(th)
            for cc in np.unique(cell):
                P = np.zeros((6, 2))
                L.orc_cell_params(C.byref(fx.om.c), ob.ptr(th), int(cc), ob.ptr(P))
                r = ob.f64(rt[cell == cc])
                ref, got = np.zeros_like(r), np.zeros_like(r)
                how = np.zeros(len(r), np.int32)
                L.orc_lba_cell(ob.ptr(P), 2, ob.ptr(fx.om.posdrift, ob.c_u8p), None, ob.ptr(r), len(r), ob.ptr(ref))
                P2 = P.copy()
                P2[1] -= P2[0]
                if not np.array_equal(P2[0] + P2[1], P[1]):
                    continue  # A + (b - A) != b in floating point: the two sides would see different thresholds
                H.hm_lba_cell_hot2(ob.ptr(ob.f64(P2)), ob.ptr(fx.om.posdrift, ob.c_u8p), ob.ptr(r), len(r), ob.ptr(got), how.ctypes.data_as(C.POINTER(C.c_int)))
                f2 = how == 2
                n_fast2 += int(f2.sum())
                _close(got, ref)
                strict = f2 & (ref > 1e-4)
                if strict.any():
                    rel = np.abs(np.log(got[strict]) - np.log(ref[strict])) / np.maximum(np.abs(np.log(ref[strict])), 1.0)
                    worst = max(worst, rel.max())
    assert n_fast2 > 5000 and worst <= 1e-10, (n_fast2, worst)


@pytest.mark.parametrize("k", [3, 5])
def test_hot_loop_one_trial_path_vs_oracle(hostmath, k):
    """n1pdf_fast<NACC> (the 4-accumulator trial loop) trial by trial against the oracle."""
    H, L = hostmath, ob.lib()
    fx = load_fixture(k)
    na = fx.ct.n_acc
    rng = np.random.default_rng(k)
    n = 0
    for s in range(2):
        rt, cell = fx.g[f"pop{s}_rt"], fx.g[f"pop{s}_cell"]
        for th in list(fx.g["pop_theta_all"][s][::13]) + list(sane_starts(fx, 6, rng, center=fx.g["ps"][s])):
            th = ob.f64(th)
            for cc in np.unique(cell):
                P = np.zeros((6, na))
                L.orc_cell_params(C.byref(fx.om.c), ob.ptr(th), int(cc), ob.ptr(P))
                r = ob.f64(rt[cell == cc])
                ref, got = np.zeros_like(r), np.zeros_like(r)
                L.orc_lba_cell(ob.ptr(P), na, ob.ptr(fx.om.posdrift, ob.c_u8p), None, ob.ptr(r), len(r), ob.ptr(ref))
                P2 = P.copy()
                P2[1] -= P2[0]
                if not np.array_equal(P2[0] + P2[1], P[1]):
                    continue
                H.hm_lba_cell_hot1(ob.ptr(ob.f64(P2)), na, ob.ptr(fx.om.posdrift, ob.c_u8p), ob.ptr(r), len(r), ob.ptr(got))
                _close(got, ref)
                n += len(r)
    assert n > 5000
