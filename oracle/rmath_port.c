/* TEST INFRASTRUCTURE ONLY -- part of the CPU oracle; never linked into the product library.
 *
 * Restatement of the handful of R "nmath" routines the reference's hot path calls through
 * libR (the reference links them as Rf_pnorm5 / Rf_dnorm4 / Rf_dunif / ...; see
 * `nm -C /root/reference/src/de.o | grep ' U '`).  libR is a third-party dependency that is NOT
 * under /root/reference (DESCRIPTION:18-19 pins only "R >= 3.5.0"), so its published algorithms are
 * restated here:
 *   - pnorm: W. J. Cody, "Rational Chebyshev approximations for the error function",
 *     Math. Comp. 23 (1969) 631-637, in the form of R's nmath/pnorm.c (pnorm_both, the
 *     non-log branches; R 4.x keeps the coefficients unchanged since R 1.x).
 *   - dnorm: nmath/dnorm.c (R >= 3.1: plain formula for |x| < 5, split-argument otherwise).
 *   - dunif, dlnorm, dcauchy, pcauchy: closed forms of nmath/{dunif,dlnorm,dcauchy,pcauchy}.c.
 *   - dgamma, dbeta: closed forms via lgamma (nmath uses saddle-point forms that agree with
 *     these to ~1e-14 relative; they are not on the LBA fixtures' path).
 * Used by oracle/ggdmc_oracle.c and by the R-API shim that lets the reference's own object code
 * (src/de.o) run without R (oracle/ref_shim.c).
 */
#include <math.h>
#include <float.h>
#include "rmath_port.h"

#define M_SQRT_32 5.656854249492380195206754896838
#define M_1_SQRT_2PI 0.398942280401432677939946059934
#define M_LN_SQRT_2PI 0.918938533204672741780329736406

/* nmath/pnorm.c: pnorm_both(), i_tail: 0 = lower only, 1 = upper only, 2 = both; log_p = 0 */
void orc_pnorm_both(double x, double *cum, double *ccum, int i_tail)
{
    static const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                                18154.981253343561249, 0.065682337918207449113};
    static const double b[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                                45507.789335026729956};
    static const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                                597.27027639480026226, 2494.5375852903726711, 6848.1904505362823326,
                                11602.651437647350124, 9842.7148383839780218, 1.0765576773720192317e-8};
    static const double d[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                                6485.558298266760755, 18615.571640885098091, 34900.952721145977266,
                                38912.003286093271411, 19685.429676859990727};
    static const double p[6] = {0.21589853405795699, 0.1274011611602473639, 0.022235277870649807,
                                0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
    static const double q[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                                0.00378239633202758244, 7.29751555083966205e-5};
    double xden, xnum, temp, del, eps, xsq, y;
    int i, lower, upper;

    if (isnan(x)) { *cum = *ccum = x; return; }
    eps = DBL_EPSILON * 0.5;
    lower = i_tail != 1;
    upper = i_tail != 0;
    y = fabs(x);
    if (y <= 0.67448975) { /* qnorm(3/4) */
        if (y > eps) {
            xsq = x * x;
            xnum = a[4] * xsq;
            xden = xsq;
            for (i = 0; i < 3; ++i) {
                xnum = (xnum + a[i]) * xsq;
                xden = (xden + b[i]) * xsq;
            }
        } else
            xnum = xden = 0.0;
        temp = x * (xnum + a[3]) / (xden + b[3]);
        if (lower) *cum = 0.5 + temp;
        if (upper) *ccum = 0.5 - temp;
    } else if (y <= M_SQRT_32) {
        xnum = c[8] * y;
        xden = y;
        for (i = 0; i < 7; ++i) {
            xnum = (xnum + c[i]) * y;
            xden = (xden + d[i]) * y;
        }
        temp = (xnum + c[7]) / (xden + d[7]);
        xsq = trunc(y * 16) / 16;
        del = (y - xsq) * (y + xsq);
        *cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
        *ccum = 1.0 - *cum;
        if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
    } else if ((lower && -37.5193 < x && x < 8.2924) || (upper && -8.2924 < x && x < 37.5193)) {
        xsq = 1.0 / (x * x);
        xnum = p[5] * xsq;
        xden = xsq;
        for (i = 0; i < 4; ++i) {
            xnum = (xnum + p[i]) * xsq;
            xden = (xden + q[i]) * xsq;
        }
        temp = xsq * (xnum + p[4]) / (xden + q[4]);
        temp = (M_1_SQRT_2PI - temp) / y;
        xsq = trunc(x * 16) / 16;
        del = (x - xsq) * (x + xsq);
        *cum = exp(-xsq * xsq * 0.5) * exp(-del * 0.5) * temp;
        *ccum = 1.0 - *cum;
        if (x > 0.) { temp = *cum; if (lower) *cum = *ccum; *ccum = temp; }
    } else {
        if (x > 0) { *cum = 1.; *ccum = 0.; }
        else { *cum = 0.; *ccum = 1.; }
    }
}

/* nmath/pnorm.c: pnorm5(x, mu, sigma, lower_tail, log_p) */
double orc_pnorm5(double x, double mu, double sigma, int lower_tail, int log_p)
{
    double p, cp;
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma <= 0) {
        if (sigma < 0) return NAN;
        p = (x < mu) ? 0. : 1.;
        p = lower_tail ? p : 1. - p;
        return log_p ? log(p) : p;
    }
    p = (x - mu) / sigma;
    if (!isfinite(p)) {
        p = (x < mu) ? 0. : 1.;
        p = lower_tail ? p : 1. - p;
        return log_p ? log(p) : p;
    }
    x = p;
    orc_pnorm_both(x, &p, &cp, lower_tail ? 0 : 1);
    p = lower_tail ? p : cp;
    return log_p ? log(p) : p; /* (the reference's hot path never asks for log_p) */
}

/* nmath/dnorm.c: dnorm4(x, mu, sigma, give_log) */
double orc_dnorm4(double x, double mu, double sigma, int give_log)
{
    if (isnan(x) || isnan(mu) || isnan(sigma)) return x + mu + sigma;
    if (sigma < 0) return NAN;
    if (!isfinite(sigma)) return give_log ? -INFINITY : 0.;
    if (!isfinite(x) && mu == x) return NAN;
    if (sigma == 0) return (x == mu) ? INFINITY : (give_log ? -INFINITY : 0.);
    x = (x - mu) / sigma;
    if (!isfinite(x)) return give_log ? -INFINITY : 0.;
    x = fabs(x);
    if (x >= 2 * sqrt(DBL_MAX)) return give_log ? -INFINITY : 0.;
    if (give_log) return -(M_LN_SQRT_2PI + 0.5 * x * x + log(sigma));
    if (x < 5) return M_1_SQRT_2PI * exp(-0.5 * x * x) / sigma;
    if (x > sqrt(-2 * M_LN2 * (DBL_MIN_EXP + 1 - DBL_MANT_DIG))) return 0.;
    {
        double x1 = ldexp(nearbyint(ldexp(x, 16)), -16);
        double x2 = x - x1;
        return M_1_SQRT_2PI / sigma * (exp(-0.5 * x1 * x1) * exp((-0.5 * x2 - x1) * x2));
    }
}

/* nmath/dunif.c */
double orc_dunif(double x, double a, double b, int give_log)
{
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (b <= a) return NAN;
    if (a <= x && x <= b) return give_log ? -log(b - a) : 1. / (b - a);
    return give_log ? -INFINITY : 0.;
}

/* nmath/dlnorm.c */
double orc_dlnorm(double x, double meanlog, double sdlog, int give_log)
{
    double y;
    if (isnan(x) || isnan(meanlog) || isnan(sdlog)) return x + meanlog + sdlog;
    if (sdlog < 0) return NAN;
    if (!isfinite(x) && log(x) == meanlog) return NAN;
    if (sdlog == 0) return (log(x) == meanlog) ? INFINITY : (give_log ? -INFINITY : 0.);
    if (x <= 0) return give_log ? -INFINITY : 0.;
    y = (log(x) - meanlog) / sdlog;
    return give_log ? -(M_LN_SQRT_2PI + 0.5 * y * y + log(x * sdlog))
                    : M_1_SQRT_2PI * exp(-0.5 * y * y) / (x * sdlog);
}

/* nmath/dcauchy.c */
double orc_dcauchy(double x, double location, double scale, int give_log)
{
    double y;
    if (isnan(x) || isnan(location) || isnan(scale)) return x + location + scale;
    if (scale <= 0) return NAN;
    y = (x - location) / scale;
    return give_log ? -log(M_PI * scale * (1. + y * y)) : 1. / (M_PI * scale * (1. + y * y));
}

/* nmath/pcauchy.c */
double orc_pcauchy(double x, double location, double scale, int lower_tail, int log_p)
{
    double v;
    if (isnan(x) || isnan(location) || isnan(scale)) return x + location + scale;
    if (scale <= 0) return NAN;
    x = (x - location) / scale;
    if (isnan(x)) return NAN;
    if (!isfinite(x)) {
        v = (x < 0) ? 0. : 1.;
        v = lower_tail ? v : 1. - v;
        return log_p ? log(v) : v;
    }
    if (!lower_tail) x = -x;
    if (fabs(x) > 1) {
        double y = atan(1 / x) / M_PI;
        v = (x > 0) ? (1. - y) : -y; /* R: R_D_Clog(y) / R_D_val(-y) */
    } else
        v = 0.5 + atan(x) / M_PI;
    return log_p ? log(v) : v;
}

/* dgamma(x, shape, scale): closed form (nmath/dgamma.c uses dpois_raw; same value) */
double orc_dgamma(double x, double shape, double scale, int give_log)
{
    double lg;
    if (isnan(x) || isnan(shape) || isnan(scale)) return x + shape + scale;
    if (shape < 0 || scale <= 0) return NAN;
    if (x < 0) return give_log ? -INFINITY : 0.;
    if (shape == 0) return (x == 0) ? INFINITY : (give_log ? -INFINITY : 0.);
    if (x == 0) {
        if (shape < 1) return INFINITY;
        if (shape > 1) return give_log ? -INFINITY : 0.;
        return give_log ? -log(scale) : 1 / scale;
    }
    lg = (shape - 1) * log(x / scale) - x / scale - lgamma(shape) - log(scale);
    return give_log ? lg : exp(lg);
}

/* dbeta(x, a, b): closed form (nmath/dbeta.c uses dbinom_raw; same value) */
double orc_dbeta(double x, double a, double b, int give_log)
{
    double lg;
    if (isnan(x) || isnan(a) || isnan(b)) return x + a + b;
    if (a < 0 || b < 0) return NAN;
    if (x < 0 || x > 1) return give_log ? -INFINITY : 0.;
    if (x == 0 || x == 1) {
        double e = (x == 0) ? a : b;
        if (e < 1) return INFINITY;
        if (e > 1) return give_log ? -INFINITY : 0.;
        return give_log ? log((x == 0) ? b : a) : ((x == 0) ? b : a);
    }
    lg = (a - 1) * log(x) + (b - 1) * log1p(-x) + lgamma(a + b) - lgamma(a) - lgamma(b);
    return give_log ? lg : exp(lg);
}
