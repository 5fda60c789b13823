// ggdmc_b200 -- DDM ("fastdm") first-passage density of one trial.
//
// Replaces ddm::ddm_class::{set_parameters, set_precision, validate_parameters, g -> integral_t0 ->
// integral_z -> integral_v -> g_no_var, get_N, compute_g_factor, compute_g_series} (@hdr/ddm.h:120-138,
// 187-309, 344-549; anonymous namespace, compiled into src/de.o, decoded from its object code) and the
// per-cell body of likelihood_class::ddm_likelihood (@hdr/likelihood.h:140-158).
//
// The reference keeps one ddm_class object, resets its 32 members for every cell of every likelihood
// call and walks the cell's trials.  Here everything that does not depend on the trial -- the scaled
// parameters, their squares, the integration step widths and thresholds derived from `precision`, and
// the verdict of validate_parameters -- is computed once per (chain, cell) into a shared-memory table
// (DdmCell); a trial gathers its cell's row and evaluates
//     g(rt) = integral over the non-decision range st0 ( integral over the start-point range sz (
//             closed-form integral over drift variability sv of the lower-boundary density ) )
// with the reference's midpoint rules and its choice between the small-time and the large-time series
// (Navarro & Fuss 2009; term counts from get_N).  The term counts and the integration grids -- everything that decides
// WHICH numbers are added -- follow the object code; the terms themselves are produced with fewer special-function calls
// (reciprocals instead of divisions, recurrences for the large-time series), which moves results by a few ulp
// (tests: <= 1e-10 relative on the log density against the oracle, itself bit-identical to the reference's machine code).
//
// The header also compiles as plain C++ (GG_HD empty) so tests can check it on the host against the oracle.
#pragma once
#include "gg_math.cuh"
#include <stdint.h>

// the one-abscissa evaluation and the start-point rule are each called from two places of the non-decision rule: one
// out-of-line copy of each
#ifdef __CUDACC__
#define GG_DDM_FN __device__ __noinline__
#else
#define GG_DDM_FN inline
#endif

namespace gg {

constexpr int kDdmRows = 10; // a, d, precision, s, st0, sv, sz, t0, v, z (alphabetical core names, like the LBA's six)
constexpr int kLbaRows = 6;

struct DdmCell {
    double a, v, sv, st0, zr, szr, t_offset; // m_a, m_v, m_sv, m_st0, m_zr, m_szr, m_t_offset
    double int_t0, int_z, var_eps;           // TUNE_INT_T0, TUNE_INT_Z, TUNE_SZ_EPSILON == TUNE_ST0_EPSILON
    double a2, v2, sv2;                      // @hdr/ddm.h:223-225
    double inv_a2, ln_inv_a2;                // 1 / a^2 (t / a^2 and the factor's division become multiplications) and its log
    int valid, no_var;                       // validate_parameters(); sv == 0
};
static_assert(sizeof(DdmCell) == 128, "DdmCell is two 64-byte lines");

constexpr double kDdmEpsilon = 1e-6;            // ddm::EPSILON, de.o .rodata+0x1b8
constexpr double kPi = 3.141592653589793;       // .rodata+0xec8
constexpr double kTwoPi = 6.283185307179586;    // .rodata+0xec0
constexpr double kPiSq = 9.869604401089358;     // m_pi2, @hdr/ddm.h:141
constexpr double kInvPi = 0.3183098861837907;
constexpr double kLnPi = 1.1447298858494002, kLn2Pi = 1.8378770664093453, kLn2 = 0.6931471805599453;
constexpr double kLnEpsilon = -13.815510557964274; // log(1e-6)

// the object code converts with cvttsd2si: NaN and out-of-range values give INT_MIN
GG_HD int ddm_trunc(double x) { return (x >= -2147483648.0 && x < 2147483648.0) ? (int)x : (-2147483647 - 1); }

// set_parameters(matrix, is_lower) @hdr/ddm.h:187-226 + set_precision :120-138 + validate_parameters :229-309.
// P = column 0 of the ten rows; is_upper = dmi@is_positive_drift of the cell (@hdr/likelihood.h:142 passes its negation).
GG_HD void ddmcell_build(DdmCell &q, const double *P, bool is_upper)
{
    const double s = P[3], scale = (1.0 != s) ? 1.0 / s : 1.0, prec = P[2]; // :194-195
    q.st0 = P[4];
    q.a = P[0] * scale;                               // :203
    q.sv = P[5] * scale;                              // :204
    q.v = is_upper ? (-P[8]) * scale : P[8] * scale;  // :205-206
    q.t_offset = 0.5 * P[4] + P[7];                   // :208
    const double zr0 = P[9] / q.a;                    // :216
    q.zr = is_upper ? 1.0 - zr0 : zr0;                // :217
    q.szr = P[6] / q.a;                               // :218
    q.int_t0 = exp(prec * -1.03758) * 0.089045;       // :130
    q.int_z = exp(prec * -1.022373) * 0.508061;       // :131
    q.var_eps = pow(10.0, -(2.0 + prec));             // :136-137
    q.a2 = q.a * q.a;
    q.v2 = q.v * q.v;
    q.sv2 = q.sv * q.sv;
    q.inv_a2 = 1.0 / q.a2;
    q.ln_inv_a2 = log(q.inv_a2);
    q.no_var = q.sv == 0;
    bool ok = true; // comparisons are false on NaN, like the reference's
    if (q.a <= 0) ok = false;                         // :232
    if (q.szr < 0 || q.szr > 1.0) ok = false;         // :240
    if (q.st0 < 0) ok = false;                        // :250
    if (q.sv < 0) ok = false;                         // :259
    if (q.t_offset < 0) ok = false;                   // :268
    if (q.zr - 0.5 * q.szr <= 0) ok = false;          // :278
    if (q.zr + 0.5 * q.szr >= 1.0) ok = false;        // :288
    if (s <= 0) ok = false;                           // :298
    q.valid = ok;
}

// 1 / sqrt(x) and sin / cos of pi x: the device has cheaper dedicated forms (MUFU.RSQ64H seed + Newton; exact argument
// reduction for multiples of pi); the host build used by the CPU tests spells them with libm
GG_HD double ddm_rsqrt(double x)
{
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
GG_HD void ddm_sincospi(double x, double *s, double *c)
{
#ifdef __CUDA_ARCH__
    sincospi(x, s, c);
#else
    *s = sin(kPi * x);
    *c = cos(kPi * x);
#endif
}

// ddm_trunc(ceil(sqrt(x))) -- the form every term count of get_N has -- without the FP64 square root: a float root
// gives the candidate, two FP64 comparisons make it the smallest k with k^2 >= x.  Outside [0, 1e12) (term counts no
// sampler state produces) the reference's own expression is evaluated, so NaN / overflow keep their INT_MIN meaning.
GG_HD int ddm_ceil_sqrt(double x)
{
    if (!(x >= 0.0 && x < 1e12)) return ddm_trunc(ceil(sqrt(x)));
    int k = (int)ceilf(sqrtf((float)x));
    const double kd = (double)k;
    if (k > 0 && (kd - 1.0) * (kd - 1.0) >= x) --k;
    else if (kd * kd < x) ++k;
    return k;
}

// compute_g_series, @hdr/ddm.h:344-379.  Same sums as the reference, fewer special-function calls:
//  * small-time series: the division by 2 t/a^2 of every term becomes one reciprocal;
//  * large-time series sum_i i exp(-(pi i)^2 ta / 2) sin(pi i zr): exp(.)_i = q^(i^2) with q = exp(-pi^2 ta / 2) is
//    advanced by two multiplications per term (ratio q^(2i+1)), sin(pi i zr) by the three-term recurrence
//    s_(i+1) = 2 cos(pi zr) s_i - s_(i-1) -- one exp and one sincospi per evaluation instead of one exp and one sin per
//    term.  The recurrences' error grows like i^2 ulp; beyond 48 terms (never chosen for a t the small-time series
//    handles in a handful) the terms are evaluated one by one like the reference does.
GG_HD double ddm_series(double ta, double zr, bool use_small, int N)
{
    double sum = 0.0;
    if (use_small) {
        const double norm = ddm_rsqrt((ta * ta * ta) * kTwoPi);
        const double m_inv_two_ta = -0.5 / ta;
        const int hi = N / 2, lo = -(N / 2);
        for (int i = lo; i <= hi; ++i) {
            const double d = ((double)i + (double)i) + zr;
            sum = exp((d * d) * m_inv_two_ta) * d + sum;
        }
        return sum * norm;
    }
    if (N > 48) {
        for (int i = 1; i <= N; ++i) {
            const double d = kPi * (double)i;
            sum = (double)i * (exp(-0.5 * d * d * ta) * sin(d * zr)) + sum;
        }
        return kPi * sum;
    }
    double s, c;
    ddm_sincospi(zr, &s, &c);
    const double q = exp((-0.5 * kPiSq) * ta), q2 = q * q, two_c = c + c;
    double e = q, r = q2 * q, s_prev = 0.0;
    for (int i = 1; i <= N; ++i) {
        sum = ((double)i * e) * s + sum;
        e *= r;
        r *= q2;
        const double s_next = two_c * s - s_prev;
        s_prev = s;
        s = s_next;
    }
    return kPi * sum;
}

// compute_g_factor, @hdr/ddm.h:383-405
GG_HD double ddm_factor(const DdmCell &q, double t, double zr)
{
    double f;
    if (q.no_var) {
        f = exp((-q.a * zr) * q.v - (0.5 * q.v2) * t) * q.inv_a2;
    } else {
        const double denom = 1.0 + q.sv2 * t;
        const double e = (-0.5 * ((q.v2 * t + (q.a * (q.v + q.v)) * zr) - ((q.a2 * zr) * zr) * q.sv2)) / denom;
        f = exp(e) * (q.inv_a2 * ddm_rsqrt(denom));
    }
    return isfinite(f) ? f : 0.0;
}

// integral_v (@hdr/ddm.h:457-485) and g_no_var (:433-454): the two share every step but the factor
GG_DDM_FN double ddm_integral_v(const DdmCell &q, double t, double zr)
{
    if (0 >= t) return 0.0;
    const double ta = t * q.inv_a2;
    const double factor = ddm_factor(q, t, zr);
    if (factor == 0) return 0.0;
    const double eps = kDdmEpsilon / factor;
    // get_N, :408-430: nl = max(ceil(1 / (pi sqrt t)), ceil(sqrt(-2 log(pi ta eps) / (pi^2 ta)))) terms of the large-time
    // series, ns = ceil(max(sqrt ta + 1, sqrt(-2 ta log(2 eps sqrt(2 pi ta))) + 2)) of the small-time one (ceil and max
    // commute, and ceil(x + integer) = ceil(x) + integer)
    int nl = ddm_trunc(ceil(ddm_rsqrt(t) * kInvPi));
    const double pe = (kPi * ta) * eps;
    if (1.0 > pe) {
        const int k = ddm_ceil_sqrt((log(pe) * -2.0) / (kPiSq * ta));
        if (nl < k) nl = k;
    }
    int ns = 2;
    const double rt2 = sqrt(ta * kTwoPi);
    if (1.0 > (rt2 + rt2) * eps) {
        const double x1 = (-2.0 * ta) * log(rt2 * (eps + eps));
        if (ta > 1e-30 && ta < 1e12 && x1 > 1e-30 && x1 < 1e12) {
            const int k1 = ddm_ceil_sqrt(x1) + 2, k2 = ddm_ceil_sqrt(ta) + 1;
            ns = k2 < k1 ? k1 : k2;
        } else { // roots that vanish next to the added integer, overflow, NaN: the reference's expression as it stands
            const double t1 = sqrt(x1) + 2.0, t2 = sqrt(ta) + 1.0;
            ns = ddm_trunc(ceil(t2 < t1 ? t1 : t2));
        }
    }
    const bool use_small = ns < nl;
    return ddm_series(ta, zr, use_small, use_small ? ns : nl) * factor;
}

// The start-point rule.  integrate_v_over_zr (@hdr/ddm.h:488-505) with integral_v (:457-485), compute_g_factor (:383-405)
// and get_N (:408-430) folded into one function so that what depends on t alone is computed once per t and not once per
// start-point abscissa:
//  * t / a^2, ceil(1 / (pi sqrt t)), sqrt(2 pi ta), ceil(sqrt ta) + 1, the factor's zr-independent multiplier
//    c = 1 / (a^2 sqrt(1 + sv^2 t));
//  * the LOGARITHMS.  get_N needs log(pi ta eps) and log(2 eps sqrt(2 pi ta)) with eps = 1e-6 / factor and
//    factor = exp(e) c, so log eps = log 1e-6 - e - log c: with log(ta) and log(c) taken once per t, no logarithm is left
//    inside the abscissa loop (the conditions `1 > ...` are still tested on the products themselves, like the reference's;
//    a sum that rounding pushes across 0 falls back to the reference's expression).  This moves the argument of a term
//    count's ceil by ~1e-14 relative instead of ~1e-16.
// The reference's midpoint rule is kept exactly: max(4, trunc(width / step)) abscissae, x accumulated by += step,
// `upper > x` as the loop test.  (Measured: 1.50e8 -> 2.11e8 trial-likelihoods/s with all variabilities on; the single
// abscissa of a model without start-point variability keeps the plain ddm_integral_v, which needs fewer registers and
// is 10-16 % faster there.)
GG_DDM_FN double ddm_integrate_v_over_zr(const DdmCell &q, double t)
{
    const double lower = q.zr - 0.5 * q.szr, upper = 0.5 * q.szr + q.zr, width = upper - lower;
    int n = ddm_trunc(width / q.int_z);
    if (n < 4) n = 4;
    const double step = width / (double)n, weight = step, divisor = q.szr;
    double x = 0.5 * step + lower;
    // every abscissa contributes 0 when 0 >= t (:459); a NaN t ends as 0 through the factor's isfinite test (:389, :403)
    if (!(t > 0)) return 0.0 / divisor;
    const double ta = t * q.inv_a2;
    const double lt = log(ta);
    const double log_pi_ta = lt + kLnPi;
    const double rt2 = sqrt(ta * kTwoPi);
    const double log_2rt2 = 0.5 * (lt + kLn2Pi) + kLn2;
    const int nl0 = ddm_trunc(ceil(ddm_rsqrt(t) * kInvPi));
    const bool ta_ok = ta > 1e-30 && ta < 1e12;
    const int k2p1 = ta_ok ? ddm_ceil_sqrt(ta) + 1 : 0;
    double denom = 1.0, c = q.inv_a2, log_c = q.ln_inv_a2;
    if (!q.no_var) {
        denom = 1.0 + q.sv2 * t;
        c = q.inv_a2 * ddm_rsqrt(denom);
        log_c = q.ln_inv_a2 - 0.5 * log(denom);
    }
    const double v2t = q.v2 * t, two_av = q.a * (q.v + q.v);

    double sum = 0.0;
    for (; upper > x; x += step) {
        const double e_arg = q.no_var ? (-q.a * x) * q.v - (0.5 * q.v2) * t
                                      : (-0.5 * ((v2t + two_av * x) - ((q.a2 * x) * x) * q.sv2)) / denom;
        double factor = exp(e_arg) * c;
        if (!isfinite(factor)) factor = 0.0;
        double val = 0.0;
        if (factor != 0) {
            const double eps = kDdmEpsilon / factor;
            const double le = kLnEpsilon - (e_arg + log_c);
            int nl = nl0;
            if (1.0 > (kPi * ta) * eps) {
                const double L = log_pi_ta + le;
                const int k = L < 0 ? ddm_ceil_sqrt((L * -2.0) / (kPiSq * ta)) : 1;
                if (nl < k) nl = k;
            }
            int ns = 2;
            if (1.0 > (rt2 + rt2) * eps) {
                const double x1 = (-2.0 * ta) * (log_2rt2 + le);
                if (ta_ok && x1 > 1e-30 && x1 < 1e12) {
                    const int k1 = ddm_ceil_sqrt(x1) + 2;
                    ns = k2p1 < k1 ? k1 : k2p1;
                } else { // vanishing / overflowing roots, a sum rounded across 0, NaN: the reference's expression as it stands
                    const double t1 = sqrt((-2.0 * ta) * log(rt2 * (eps + eps))) + 2.0, t2 = sqrt(ta) + 1.0;
                    ns = ddm_trunc(ceil(t2 < t1 ? t1 : t2));
                }
            }
            const bool use_small = ns < nl;
            val = ddm_series(ta, x, use_small, use_small ? ns : nl) * factor;
        }
        sum = val * weight + sum;
    }
    return sum / divisor;
}

// integral_z, @hdr/ddm.h:508-514
GG_HD double ddm_integral_z(const DdmCell &q, double t)
{
    return q.var_eps > q.szr ? ddm_integral_v(q, t, q.zr) : ddm_integrate_v_over_zr(q, t);
}

// g (@hdr/ddm.h:545-549) -> integral_t0 (:537-542) with integrate_z_over_t (:517-534)
GG_HD double ddm_g(const DdmCell &q, double rt)
{
    const double t = rt - q.t_offset;
    if (q.var_eps > q.st0) return ddm_integral_z(q, t);
    const double lower = t - q.st0 * 0.5, upper = 0.5 * q.st0 + t, width = upper - lower;
    int n = ddm_trunc(width / q.int_t0);
    if (n < 4) n = 4;
    const double step = width / (double)n;
    double sum = 0.0;
    for (double x = 0.5 * step + lower; upper > x; x += step) sum = ddm_integral_z(q, x) * step + sum;
    return sum / q.st0;
}

// density of one trial as likelihood_class::ddm_likelihood stores it: 1e-10 for every trial of an invalid cell
// (@hdr/likelihood.h:158), else dddm (@hdr/ddm.h:552-560)
GG_HD double ddm_density(const DdmCell &q, double rt) { return q.valid ? ddm_g(q, rt) : kFloor; }

// what sumloglike takes the log of: std::max(density, DBL_MIN) (@hdr/likelihood.h:303); NaN stays NaN
GG_HD double ddm_floor(double x) { return x < DBL_MIN ? DBL_MIN : x; }

} // namespace gg
