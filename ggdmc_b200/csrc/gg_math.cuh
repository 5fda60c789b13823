// ggdmc_b200 -- FP64 special functions of the LBA / prior densities (device code).
//
// What the reference gets from libR's nmath (Rf_pnorm5, Rf_dnorm4, Rf_dunif, ...; call sites
// @hdr/lba.h:112-117,213-345, @hdr/tnorm.h:59-127, @hdr/prior.h:331-412) is implemented here for
// sm_100a.  Same special-case behaviour (NaN propagation, sigma <= 0, infinities); the regular
// path is 0.5*erfc(-z/sqrt2) and a compensated exp(-z^2/2), accurate to a few ulp.
//
// The header also compiles as plain C++ (GG_HD empty) so tests can check the arithmetic on the
// host against mpmath; the product only ever runs it on the GPU.
#pragma once
#include <math.h>
#include <float.h>

#ifdef __CUDACC__
#define GG_HD __device__ __forceinline__
#define GG_COLD __device__ __noinline__ // rare or long code that must not be copied into every caller (instruction cache)
#else
#define GG_HD inline
#define GG_COLD inline
#endif

namespace gg {

constexpr double kFloor = 1e-10;              // de.o .rodata+0xee0
constexpr double kInvSqrt2 = 0.70710678118654752440;
constexpr double kInvSqrt2Pi = 0.398942280401432677939946059934;
constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;

// exp(-z^2/2) with the rounding error of z*z compensated (z*z = s + r exactly).
GG_HD double exp_mhalf_sq(double z)
{
    double s = z * z;
    double r = fma(z, z, -s);
    double e = exp(-0.5 * s);
    return fma(e, -0.5 * r, e);
}

// Rf_pnorm5(z, 0, 1, lower_tail = TRUE, log_p = FALSE)
GG_HD double pnorm_std(double z) { return 0.5 * erfc(-z * kInvSqrt2); }
// Rf_pnorm5(z, 0, 1, lower_tail = FALSE, log_p = FALSE)
GG_HD double pnorm_std_upper(double z) { return 0.5 * erfc(z * kInvSqrt2); }
// Rf_dnorm4(z, 0, 1, log = FALSE)
GG_HD double dnorm_std(double z)
{
    double az = fabs(z);
    if (!(az < 1e154)) return (az != az) ? z : 0.0; // NaN propagates; inf / huge -> 0
    return kInvSqrt2Pi * exp_mhalf_sq(az);
}

// Rf_pnorm5(x, mu, sigma, lower_tail, log_p = FALSE), nmath/pnorm.c special cases included
GG_HD double pnorm5(double x, double mu, double sigma, bool lower)
{
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (isinf(x) && mu == x) return NAN;
    if (sigma <= 0.0) {
        if (sigma < 0.0) return NAN;
        double p = (x < mu) ? 0.0 : 1.0;
        return lower ? p : 1.0 - p;
    }
    double z = (x - mu) / sigma;
    if (isinf(z)) {
        double p = (x < mu) ? 0.0 : 1.0;
        return lower ? p : 1.0 - p;
    }
    return lower ? pnorm_std(z) : pnorm_std_upper(z);
}

// Rf_dnorm4(x, mu, sigma, give_log), nmath/dnorm.c special cases included
GG_HD double dnorm4(double x, double mu, double sigma, bool give_log)
{
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (sigma < 0.0) return NAN;
    if (isinf(sigma)) return give_log ? -INFINITY : 0.0;
    if (isinf(x) && mu == x) return NAN;
    if (sigma == 0.0) return (x == mu) ? INFINITY : (give_log ? -INFINITY : 0.0);
    double z = (x - mu) / sigma;
    if (isinf(z)) return give_log ? -INFINITY : 0.0;
    z = fabs(z);
    if (z >= 2.0 * 1.3407807929942596e154) return give_log ? -INFINITY : 0.0;
    if (give_log) return -(kLnSqrt2Pi + 0.5 * z * z + log(sigma));
    return kInvSqrt2Pi * exp_mhalf_sq(z) / sigma;
}

// dnorm4(x, mu, sigma, true) with log(sigma) supplied by the caller (the hyper-likelihood evaluates thousands of x per sigma)
GG_HD double dnorm4_log_pre(double x, double mu, double sigma, double log_sigma)
{
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (sigma < 0.0) return NAN;
    if (isinf(sigma)) return -INFINITY;
    if (isinf(x) && mu == x) return NAN;
    if (sigma == 0.0) return (x == mu) ? INFINITY : -INFINITY;
    double z = (x - mu) / sigma;
    if (isinf(z)) return -INFINITY;
    z = fabs(z);
    if (z >= 2.0 * 1.3407807929942596e154) return -INFINITY;
    return -(kLnSqrt2Pi + 0.5 * z * z + log_sigma);
}

// Rf_dunif
GG_HD double dunif(double x, double a, double b, bool give_log)
{
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (b <= a) return NAN;
    if (a <= x && x <= b) return give_log ? -log(b - a) : 1.0 / (b - a);
    return give_log ? -INFINITY : 0.0;
}

GG_HD double dlnorm(double x, double meanlog, double sdlog, bool give_log)
{
    if (isnan(x) || isnan(meanlog) || isnan(sdlog)) return x + meanlog + sdlog;
    if (sdlog < 0.0) return NAN;
    if (isinf(x) && log(x) == meanlog) return NAN;
    if (sdlog == 0.0) return (log(x) == meanlog) ? INFINITY : (give_log ? -INFINITY : 0.0);
    if (x <= 0.0) return give_log ? -INFINITY : 0.0;
    double y = (log(x) - meanlog) / sdlog;
    return give_log ? -(kLnSqrt2Pi + 0.5 * y * y + log(x * sdlog)) : kInvSqrt2Pi * exp(-0.5 * y * y) / (x * sdlog);
}

GG_HD double dcauchy(double x, double loc, double scale, bool give_log)
{
    if (isnan(x) || isnan(loc) || isnan(scale)) return x + loc + scale;
    if (scale <= 0.0) return NAN;
    double y = (x - loc) / scale;
    double d = 3.14159265358979323846 * scale * (1.0 + y * y);
    return give_log ? -log(d) : 1.0 / d;
}

GG_HD double pcauchy(double x, double loc, double scale)
{
    if (isnan(x) || isnan(loc) || isnan(scale)) return x + loc + scale;
    if (scale <= 0.0) return NAN;
    x = (x - loc) / scale;
    if (isnan(x)) return NAN;
    if (isinf(x)) return (x < 0.0) ? 0.0 : 1.0;
    if (fabs(x) > 1.0) {
        double y = atan(1.0 / x) / 3.14159265358979323846;
        return (x > 0.0) ? (1.0 - y) : -y;
    }
    return 0.5 + atan(x) / 3.14159265358979323846;
}

GG_HD double dgamma(double x, double shape, double scale, bool give_log)
{
    if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
    if (shape < 0.0 || scale <= 0.0) return NAN;
    if (x < 0.0) return give_log ? -INFINITY : 0.0;
    if (shape == 0.0) return (x == 0.0) ? INFINITY : (give_log ? -INFINITY : 0.0);
    if (x == 0.0) {
        if (shape < 1.0) return INFINITY;
        if (shape > 1.0) return give_log ? -INFINITY : 0.0;
        return give_log ? -log(scale) : 1.0 / scale;
    }
    double lg = (shape - 1.0) * log(x / scale) - x / scale - lgamma(shape) - log(scale);
    return give_log ? lg : exp(lg);
}

GG_HD double dbeta(double x, double a, double b, bool give_log)
{
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (a < 0.0 || b < 0.0) return NAN;
    if (x < 0.0 || x > 1.0) return give_log ? -INFINITY : 0.0;
    if (x == 0.0 || x == 1.0) {
        double e = (x == 0.0) ? a : b, o = (x == 0.0) ? b : a;
        if (e < 1.0) return INFINITY;
        if (e > 1.0) return give_log ? -INFINITY : 0.0;
        return give_log ? log(o) : o;
    }
    double lg = (a - 1.0) * log(x) + (b - 1.0) * log1p(-x) + lgamma(a + b) - lgamma(a) - lgamma(b);
    return give_log ? lg : exp(lg);
}

// tnorm_class::set_parameters + d  (@hdr/tnorm.h:59-67, 118-127)
GG_HD double tnorm_d(double x, double mean, double sd, double lower, double upper, bool log_p)
{
    double denom = pnorm5(upper, mean, sd, true) - pnorm5(lower, mean, sd, true);
    if (x < lower || x > upper) return log_p ? -INFINITY : kFloor;
    return log_p ? dnorm4(x, mean, sd, true) - log(denom) : dnorm4(x, mean, sd, false) / denom;
}

// prior_class::dcauchy_trunc (@hdr/prior.h:25-58)
GG_HD double dcauchy_trunc(double x, double loc, double scale, double lower, double upper, bool log_p)
{
    if (0.0 >= scale || !(lower < upper)) return NAN;
    double den = dcauchy(x, loc, scale, false);
    double Fu = pcauchy(upper, loc, scale), Fl = pcauchy(lower, loc, scale);
    double out = 0.0;
    if (x >= lower && upper >= x) out = den / (Fu - Fl);
    if (log_p) return out > 0.0 ? log(out) : -INFINITY;
    return out;
}

// log of the probability mass a normal(mean, sd) puts on [lower, upper]: tnorm_class::set_parameters (@hdr/tnorm.h:59-67)
GG_COLD double tnorm_logmass(double lower, double upper, double mean, double sd)
{
    return log(pnorm5(upper, mean, sd, true) - pnorm5(lower, mean, sd, true));
}

// one element of prior_class::dprior (@hdr/prior.h:342-407)
GG_COLD double dprior1(int dist, double x, double p0, double p1, double lower, double upper, bool log_p)
{
    switch (dist) {
    case 1: return tnorm_d(x, p0, p1, lower, upper, log_p);
    case 2: {
        double range = upper - lower, xs = (x - lower) / range, den = -INFINITY;
        if (p0 >= 0.0 && p1 >= 0.0) den = dbeta(xs, p0, p1, log_p);
        return log_p ? den - log(range) : den / range;
    }
    case 3: return dgamma(isfinite(lower) ? x - lower : x, p0, p1, log_p);
    case 4: return dlnorm(isfinite(lower) ? x - lower : x, p0, p1, log_p);
    case 5: return dcauchy_trunc(x, p0, p1, lower, upper, log_p);
    case 6: {
        double v = dunif(x, p0, p1, log_p);
        return isnan(v) ? -1e10 : v;
    }
    case 7: return dnorm4(x, p0, p1, log_p);
    default: return NAN;
    }
}

} // namespace gg
