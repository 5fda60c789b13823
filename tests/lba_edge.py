"""LBA edge cases shared by tests/golden/make_lba_edge_golden.py (which runs them through the reference's own object
code) and by the CPU / GPU parity tests.

Every case is one CELL of an all-free model: rows A, B, mean_v, sd_v, st0, t0 x n_acc accumulators (the reference's
design_class::set_parameter_values turns row B into b = A + B, @hdr/design_light.h:336-340).  The cases walk through
lba_class::set_parameters (@hdr/lba.h:88-119: `t0 + st0 U` with U > 0 and the drift denominator and its 1e-10 floor),
every rule of validate_parameters (:121-146), the A < 1e-10 point-mass branches of d() and p() (:221-227, 315-320),
sd_v = 0, NaN and infinite parameters, and response times below, at and just above t0.
"""
import numpy as np

ROWS = ("A", "B", "mean_v", "sd_v", "st0", "t0")
RT_GRID = np.array([0.05, 0.2, 0.2 + 1e-12, 0.2000001, 0.21, 0.26, 0.3, 0.31, 0.45, 0.7, 1.1, 1.9, 4.0, 25.0])


def _base(na):
    P = np.zeros((6, na))
    P[0] = [0.75, 0.6, 0.9, 0.5][:na]      # A
    P[1] = [0.85, 1.1, 0.7, 1.3][:na]      # B
    P[2] = [2.5, 1.1, 0.4, 1.8][:na]       # mean_v
    P[3] = [1.0, 1.2, 0.8, 1.5][:na]       # sd_v
    P[4] = 0.0                             # st0
    P[5] = 0.2                             # t0
    return P


def cases(na):
    """-> list of (name, P [6, na] with row 1 = B (not yet b), posdrift [na])"""
    out = []

    def add(name, P, pd=None):
        out.append((name, P, np.ones(na, np.uint8) if pd is None else np.asarray(pd, np.uint8)))

    add("regular", _base(na))
    P = _base(na); P[4] = 0.1; add("st0_all", P)
    P = _base(na); P[4, 0] = 0.25; add("st0_winner_only", P)
    P = _base(na); P[4, na - 1] = 0.3; add("st0_last_only", P)
    P = _base(na); P[4] = [0.05, 0.4, 0.15, 0.2][:na]; P[5] = [0.1, 0.25, 0.2, 0.05][:na]; add("st0_t0_per_accumulator", P)
    # validate_parameters, one rule at a time, in the first and in the last accumulator
    for j in (0, na - 1):
        P = _base(na); P[0, j] = -0.1; add(f"invalid_A_neg_{j}", P)
        P = _base(na); P[0, j] = 0.2; P[1, j] = -0.5; add(f"invalid_b_neg_{j}", P)
        P = _base(na); P[0, j] = 1.0; P[1, j] = -0.25; add(f"invalid_b_lt_A_{j}", P)
        P = _base(na); P[3, j] = -0.3; add(f"invalid_sdv_neg_{j}", P)
        P = _base(na); P[4, j] = -0.01; add(f"invalid_st0_neg_{j}", P)
        P = _base(na); P[5, j] = -0.05; add(f"invalid_t0_neg_{j}", P)
    P = _base(na); P[1, 0] = 0.0; add("b_equals_A", P)  # b < A is false: valid
    # point-mass start point
    P = _base(na); P[0, 0] = 0.0; add("A_zero_winner", P)
    P = _base(na); P[0, na - 1] = 5e-11; add("A_tiny_survivor", P)
    P = _base(na); P[0] = 0.0; add("A_zero_all", P)
    P = _base(na); P[0, 0] = 1e-10; add("A_at_threshold", P)  # A < 1e-10 is false: the general branch
    P = _base(na); P[0] = 0.0; P[4] = 0.2; add("A_zero_st0", P)
    # sd_v = 0 passes validate_parameters
    P = _base(na); P[3, 0] = 0.0; add("sdv_zero_winner", P)
    P = _base(na); P[3, na - 1] = 0.0; add("sdv_zero_survivor", P)
    # NaN / inf in every row
    for r, nm in enumerate(ROWS):
        for j in (0, na - 1):
            P = _base(na); P[r, j] = np.nan; add(f"nan_{nm}_{j}", P)
    P = _base(na); P[2, 0] = np.inf; add("inf_meanv_winner", P)
    P = _base(na); P[2, na - 1] = -np.inf; add("neginf_meanv_survivor", P)
    P = _base(na); P[1, 0] = np.inf; add("inf_B", P)
    P = _base(na); P[3, na - 1] = np.inf; add("inf_sdv", P)
    # drift denominator: floor at 1e-10, no truncation, mixed
    P = _base(na); P[2] = -45.0; add("denom_floor", P)
    P = _base(na); P[2, 0] = -7.0; add("denom_small", P)
    add("no_posdrift", _base(na), np.zeros(na))
    add("mixed_posdrift", _base(na), [1, 0, 0, 1][:na])
    P = _base(na); P[2] = [-0.5, -1.5, 0.2, -2.0][:na]; add("negative_drifts", P, np.zeros(na))
    # extreme but regular
    P = _base(na); P[3] = 1e-3; add("sdv_small", P)
    P = _base(na); P[3] = 40.0; add("sdv_large", P)
    P = _base(na); P[0] = 1e-8; add("A_small", P)
    P = _base(na); P[1] = 30.0; add("B_large", P)
    P = _base(na); P[5] = 0.7; add("t0_large", P)  # most of the grid is below t0
    P = _base(na); P[5] = 0.0; add("t0_zero", P)
    return out


def philox_u_st0(L, ob, seed, pop, iteration, chain, n_slots):
    """The addressed draws the CUDA engine uses for `t0 + st0 U` (purpose U_ST0, slot = cell * n_acc + accumulator)."""
    import ctypes as C
    r = ob.make_rng(seed=seed)
    L.orc_uniform.restype = C.c_double
    u = np.empty(n_slots)
    for k in range(n_slots):
        a = ob.Addr(pop, iteration, 0, chain, 3, k)
        u[k] = L.orc_uniform(C.byref(r), C.byref(a))
    return u


def model_for(na):
    """(CellTable, theta, list of (name, P, posdrift groups)) -- one all-free model per posdrift pattern, because
    is_positive_drift belongs to the model, not to the cell."""
    from ggdmc_b200.model import CellTable
    groups = {}
    for name, P, pd in cases(na):
        groups.setdefault(tuple(int(x) for x in pd), []).append((name, P))
    out = []
    for pd, lst in groups.items():
        ncell = len(lst)
        src = np.arange(ncell * 6 * na, dtype=np.int32).reshape(ncell, 6, na)
        theta = np.concatenate([P.reshape(-1) for _, P in lst])
        ct = CellTable(na, ncell, len(theta), src, np.zeros(1), np.array(pd, np.uint8), [f"p{i}" for i in range(len(theta))],
                       [n for n, _ in lst])
        out.append((ct, theta, lst, np.array(pd, np.uint8)))
    return out
