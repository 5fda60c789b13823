// ggdmc_b200 -- LBA node-1 density (n1PDF): winner's defective PDF x survivors' (1 - CDF).
//
// Replaces lba_class::{set_parameters, validate_parameters, d, p, dlba} (@hdr/lba.h:88-146,
// 213-345, 560-572) and design_class::set_parameter_values (@hdr/design_light.h:314-344).
// The reference rebuilds a 6 x n_acc parameter matrix per cell per likelihood call and walks the
// trials of that cell; here the per-(cell, accumulator) quantities that do not depend on the trial
// -- the parameters themselves, max(Phi(mean_v/sd_v), 1e-10), and the reciprocals the trial loop
// multiplies by -- are computed once per (chain, cell, accumulator) into a shared-memory table
// (CellAcc) and every trial only gathers its cell's row.  Branch structure, floors (1e-10), the
// clamp of the survivor CDF to [1e-10, 1] and the NaN -> 1e-10 rules are the reference's.
#pragma once
#include "gg_math.cuh"
#include "gg_fastmath.cuh"
#include <stdint.h>

namespace gg {

struct CellAcc {
    double b, A, mean_v, sd_v, t0a, inv_sdv, inv_A, inv_denom;
};

// The table holds every DISTINCT (cell, accumulator) row once (the model says which entries draw the same six
// parameters, gg_host.cuh ModelDev); a trial reaches its cell's rows through the cell's row indices.
struct RowRef {
    const CellAcc *rows;
    const uint16_t *idx; // [n_acc] row of accumulator j
    GG_HD const CellAcc &operator[](int j) const { return rows[idx[j]]; }
};

// 1 / x for the cell table: hardware seed + two Newton steps (<= 1 ulp) when x is a comfortable positive normal
// number -- true of A, sd_v and the drift denominator of every regular cell --, IEEE division otherwise, so the
// generic path keeps the reference's behaviour for zeros, negatives, infinities and NaN
GG_HD bool rcp_safe(double x) { return x > 1e-290 && x < 1e290; }
GG_HD double rcp_table(double x) { return rcp_safe(x) ? fm::rcp_pos(x) : 1.0 / x; }

// Entry (cell, column j) of the table from the raw core parameters of that column (A, B, mean_v, sd_v, st0, t0);
// u_st0 is the uniform of `t0 + st0 * U` (@hdr/lba.h:117).
GG_HD void cellacc_build(CellAcc &e, double A, double B, double mean_v, double sd_v, double st0, double t0,
                         bool posdrift, double u_st0)
{
    double b = A + B; // design_light.h:336-340: row B += row A
    const double inv_sdv = rcp_table(sd_v);
    const double zv = rcp_safe(sd_v) ? mean_v * inv_sdv : mean_v / sd_v;
    double denom = posdrift ? fmax(fm::norm_cdf_lowlatency(zv), kFloor) : 1.0; // lba.h:112-115
    e.b = b;
    e.A = A;
    e.mean_v = mean_v;
    e.sd_v = sd_v;
    // lba.h:117: t0 + st0 * U as a separately rounded product and sum like the reference's scalar code (no FMA contraction:
    // the value is a threshold -- rt > t0 -- and feeds every z of the cell); t0 + 0 * U == t0 + 0
#ifdef __CUDA_ARCH__
    e.t0a = __dadd_rn(t0, __dmul_rn(st0, (st0 != 0.0) ? u_st0 : 0.0));
#else
    e.t0a = t0 + st0 * ((st0 != 0.0) ? u_st0 : 0.0);
#endif
    e.inv_sdv = inv_sdv;
    e.inv_A = rcp_table(A);
    e.inv_denom = posdrift ? rcp_table(denom) : 1.0;
}

typedef fm::Pair PhiPair;

// Phi(z) and phi(z) of the standard normal (Rf_pnorm5(z,0,1,1,0), Rf_dnorm4(z,0,1,0)): one shared
// exponential, coefficients in constant memory, no branches (gg_fastmath.cuh).
GG_HD PhiPair norm_both(double z) { return fm::norm_pair(z); }

// 1 / dt for the trial loop: dt == 0 must give +inf like the reference's x / (sd_v * 0)
GG_HD double rcp_time(double dt)
{
    double r = fm::rcp_pos(dt);
    return dt == 0.0 ? INFINITY : r;
}

// Cell classes decided once per (chain, cell) when the table is built.
enum : uint8_t {
    kCellRegular = 0,  // all parameters finite, A >= 1e-10 and sd_v > 0 for every accumulator: fast path
    kCellInvalid = 1,  // validate_parameters() fails: every trial has density 1e-10 (@hdr/likelihood.h:105)
    kCellGeneric = 2   // anything else (A < 1e-10 point-mass branch, sd_v == 0, NaN / inf parameters)
};

// density of one trial whose cell's table row is e[0 .. n_acc): the reference's arithmetic with all of
// its branches and NaN rules (lba_class::d / p, @hdr/lba.h:213-248, 286-345)
// `e` is anything indexable by accumulator: a `const CellAcc *` (rows of one cell side by side) or a RowRef (rows of a
// de-duplicated table picked through the cell's row indices)
template <int NACC, class E>
GG_HD double n1pdf_generic_body(double rt, const E &e, int n_acc_rt)
{
    const int n_acc = NACC > 0 ? NACC : n_acc_rt;
    double t0a = e[0].t0a;
    double dt = rt - t0a;
    double rdt = rcp_time(dt);
    double pdf;
    {
        const double b = e[0].b, A = e[0].A, mv = e[0].mean_v, sv = e[0].sd_v;
        if (0.0 > dt) { // lba.h:217-219
            pdf = kFloor;
        } else if (A < kFloor) { // lba.h:221-227
            pdf = fmax(b / (dt * dt) * (dnorm4(b / dt, mv, sv, false) * e[0].inv_denom), kFloor); // the reference divides the density by denom first
        } else { // lba.h:231-244
            double rts = e[0].inv_sdv * rdt, tv = mv * dt;
            PhiPair n1 = norm_both((b - tv) * rts);
            PhiPair n2 = norm_both(((b - A) - tv) * rts);
            double t1 = mv * (n1.cdf - n2.cdf);
            double t2 = sv * (n2.pdf - n1.pdf);
            pdf = fmax((t1 + t2) * (e[0].inv_A * e[0].inv_denom), kFloor);
        }
        if (isnan(pdf)) pdf = kFloor; // lba.h:247
    }
    for (int j = 1; j < n_acc; ++j) { // p(), lba.h:286-345
        const double b = e[j].b, A = e[j].A, mv = e[j].mean_v, sv = e[j].sd_v;
        if (e[j].t0a != t0a) {
            t0a = e[j].t0a;
            dt = rt - t0a;
            rdt = rcp_time(dt);
        }
        double cdf;
        if (0.0 > dt) { // :310-312
            cdf = kFloor;
        } else if (A < kFloor) { // :315-320
            cdf = pnorm5(b / dt, mv, sv, false) * e[j].inv_denom;
            cdf = cdf < kFloor ? kFloor : (1.0 < cdf ? 1.0 : cdf);
        } else { // :324-338
            double ts = sv * dt, rts = e[j].inv_sdv * rdt, tv = mv * dt;
            double x1 = b - tv, x2 = x1 - A;
            PhiPair n1 = norm_both(x1 * rts);
            PhiPair n2 = norm_both(x2 * rts);
            double s = x2 * n2.cdf - x1 * n1.cdf + ts * (n2.pdf - n1.pdf);
            cdf = (1.0 + s * e[j].inv_A) * e[j].inv_denom;
            cdf = cdf < kFloor ? kFloor : (1.0 < cdf ? 1.0 : cdf); // std::clamp (NaN passes through)
        }
        pdf = pdf * (1.0 - cdf);      // :341
        if (isnan(pdf)) pdf = kFloor; // :342
    }
    return pdf;
}

// Can the fast path take this trial?  Needs rt > t0 for every accumulator (then every z is finite).
template <int NACC, class E>
GG_HD bool n1pdf_fast_ok(double rt, const E &e, int n_acc_rt)
{
    const int n_acc = NACC > 0 ? NACC : n_acc_rt;
    bool ok = (rt - e[0].t0a) > 0.0;
#pragma unroll
    for (int j = 1; j < n_acc; ++j) ok = ok && ((rt - e[j].t0a) > 0.0);
    return ok;
}

// Fast path for a trial of a REGULAR cell with n1pdf_fast_ok(): same formulas as the generic body,
// but with finite parameters, A >= 1e-10, sd_v > 0 and rt > t0 the point-mass branches and every NaN
// rule are dead code and every z is finite.
template <int NACC, class E>
GG_HD double n1pdf_fast(double rt, const E &e, int n_acc_rt)
{
    const int n_acc = NACC > 0 ? NACC : n_acc_rt;
    double t0a = e[0].t0a;
    double dt = rt - t0a;
    double rdt = fm::rcp_pos(dt);
    double pdf;
    {
        const double b = e[0].b, A = e[0].A, mv = e[0].mean_v, sv = e[0].sd_v;
        const double rts = e[0].inv_sdv * rdt, tv = mv * dt;
        const double z[2] = {(b - tv) * rts, ((b - A) - tv) * rts};
        double cdf[2], phi[2];
        fm::norm_pairs_finite<2>(z, cdf, phi);
        const double t1 = mv * (cdf[0] - cdf[1]);
        const double t2 = sv * (phi[1] - phi[0]);
        pdf = fmax((t1 + t2) * (e[0].inv_A * e[0].inv_denom), kFloor);
    }
#pragma unroll
    for (int j = 1; j < n_acc; ++j) {
        const double b = e[j].b, A = e[j].A, mv = e[j].mean_v, sv = e[j].sd_v;
        if (e[j].t0a != t0a) {
            t0a = e[j].t0a;
            dt = rt - t0a;
            rdt = fm::rcp_pos(dt);
        }
        const double ts = sv * dt, rts = e[j].inv_sdv * rdt, tv = mv * dt;
        const double x1 = b - tv, x2 = x1 - A;
        const double z[2] = {x1 * rts, x2 * rts};
        double cdf2[2], phi[2];
        fm::norm_pairs_finite<2>(z, cdf2, phi);
        const double s = x2 * cdf2[1] - x1 * cdf2[0] + ts * (phi[1] - phi[0]);
        double cdf = (1.0 + s * e[j].inv_A) * e[j].inv_denom;
        cdf = cdf < kFloor ? kFloor : (1.0 < cdf ? 1.0 : cdf);
        pdf = pdf * (1.0 - cdf);
    }
    return pdf;
}

// Two fast-path trials of one thread advanced together: per accumulator the four (Phi, phi) pairs of the two
// trials form one lock-step batch, and everything around them comes in two independent copies.  A warp whose
// consecutive FP64 instructions depend on each other cannot fill the FP64 pipe even with every scheduler slot
// taken (measured on B200, tools/cuda_probe/dfma_lat.cu: 3.0 / 2.5 / 2.2 cycles per DFMA per scheduler with
// 1 / 2 / 4 independent chains per warp at 6 warps per scheduler).  Per trial the arithmetic is n1pdf_fast's.
template <int NACC, class E>
GG_HD void n1pdf_fast2(double rtA, const E &eA, double rtB, const E &eB, int n_acc_rt, double &outA, double &outB)
{
    const int n_acc = NACC > 0 ? NACC : n_acc_rt;
    double t0A = eA[0].t0a, t0B = eB[0].t0a;
    double dtA = rtA - t0A, dtB = rtB - t0B;
    double rdtA = fm::rcp_pos(dtA), rdtB = fm::rcp_pos(dtB);
    double pdfA, pdfB;
    {
        const double rtsA = eA[0].inv_sdv * rdtA, tvA = eA[0].mean_v * dtA;
        const double rtsB = eB[0].inv_sdv * rdtB, tvB = eB[0].mean_v * dtB;
        const double z[4] = {(eA[0].b - tvA) * rtsA, ((eA[0].b - eA[0].A) - tvA) * rtsA, (eB[0].b - tvB) * rtsB,
                             ((eB[0].b - eB[0].A) - tvB) * rtsB};
        double cdf[4], phi[4];
        fm::norm_pairs_stepmajor<4>(z, cdf, phi);
        const double a1 = eA[0].mean_v * (cdf[0] - cdf[1]), a2 = eA[0].sd_v * (phi[1] - phi[0]);
        const double b1 = eB[0].mean_v * (cdf[2] - cdf[3]), b2 = eB[0].sd_v * (phi[3] - phi[2]);
        pdfA = fmax((a1 + a2) * (eA[0].inv_A * eA[0].inv_denom), kFloor);
        pdfB = fmax((b1 + b2) * (eB[0].inv_A * eB[0].inv_denom), kFloor);
    }
#pragma unroll
    for (int j = 1; j < n_acc; ++j) {
        if (eA[j].t0a != t0A) {
            t0A = eA[j].t0a;
            dtA = rtA - t0A;
            rdtA = fm::rcp_pos(dtA);
        }
        if (eB[j].t0a != t0B) {
            t0B = eB[j].t0a;
            dtB = rtB - t0B;
            rdtB = fm::rcp_pos(dtB);
        }
        const double tsA = eA[j].sd_v * dtA, rtsA = eA[j].inv_sdv * rdtA, tvA = eA[j].mean_v * dtA;
        const double tsB = eB[j].sd_v * dtB, rtsB = eB[j].inv_sdv * rdtB, tvB = eB[j].mean_v * dtB;
        const double x1A = eA[j].b - tvA, x2A = x1A - eA[j].A;
        const double x1B = eB[j].b - tvB, x2B = x1B - eB[j].A;
        const double z[4] = {x1A * rtsA, x2A * rtsA, x1B * rtsB, x2B * rtsB};
        double c[4], phi[4];
        fm::norm_pairs_stepmajor<4>(z, c, phi);
        const double sA = x2A * c[1] - x1A * c[0] + tsA * (phi[1] - phi[0]);
        const double sB = x2B * c[3] - x1B * c[2] + tsB * (phi[3] - phi[2]);
        double cdfA = (1.0 + sA * eA[j].inv_A) * eA[j].inv_denom;
        double cdfB = (1.0 + sB * eB[j].inv_A) * eB[j].inv_denom;
        cdfA = cdfA < kFloor ? kFloor : (1.0 < cdfA ? 1.0 : cdfA);
        cdfB = cdfB < kFloor ? kFloor : (1.0 < cdfB ? 1.0 : cdfB);
        pdfA = pdfA * (1.0 - cdfA);
        pdfB = pdfB * (1.0 - cdfB);
    }
    outA = pdfA;
    outB = pdfB;
}

// density of any trial of any cell class (used outside the hot loop)
template <int NACC, class E>
GG_HD double n1pdf_any(uint8_t cls, double rt, const E &e, int n_acc)
{
    if (cls == kCellInvalid) return kFloor;
    if (cls == kCellRegular && n1pdf_fast_ok<NACC>(rt, e, n_acc)) return n1pdf_fast<NACC>(rt, e, n_acc);
    return n1pdf_generic_body<NACC>(rt, e, n_acc);
}

// class of a cell from the raw parameters of all its accumulators (v = A, B, mean_v, sd_v, st0, t0)
GG_HD uint8_t cell_class_update(uint8_t cls, double A, double B, double mean_v, double sd_v, double st0, double t0)
{
    const double b = A + B;
    if ((A < 0.0) || (b < 0.0) || (b < A) || (sd_v < 0.0) || (st0 < 0.0) || (t0 < 0.0)) return kCellInvalid;
    if (cls == kCellInvalid) return cls;
    const bool regular = (A >= kFloor) && (sd_v > 0.0) && isfinite(A) && isfinite(b) && isfinite(mean_v) && isfinite(sd_v) &&
                         isfinite(st0) && isfinite(t0);
    return regular ? cls : (uint8_t)kCellGeneric;
}

} // namespace gg
