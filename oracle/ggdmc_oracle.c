/* TEST INFRASTRUCTURE ONLY -- CPU oracle (see ggdmc_oracle.h for scope and parity status).
 *
 * Every function cites the reference lines it restates.  `@hdr/x.h:NN` is line NN of
 * ggdmcHeaders/x.h, a header-only dependency that is not under /root/reference; its code is
 * present as object code in /root/reference/src/de.o (line tables in DWARF), decoded in
 * SURVEY.md Appendix A/B.  ggdmcHeaders has no pinned version (DESCRIPTION:20-23 LinkingTo only).
 */
#include "ggdmc_oracle.h"
#include "rmath_port.h"
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FLOOR_ 1e-10 /* de.o .rodata+0xee0 */

/* ------------------------------------------------------------------------------------------- */
/* uniform source                                                                               */
/* ------------------------------------------------------------------------------------------- */

/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Not from the reference (which uses R's
 * global Mersenne-Twister stream, src/RcppExports.cpp:19); this is the engine's counter-based
 * replacement, restated here independently so trajectories can be compared draw for draw. */
void orc_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4])
{
    unsigned c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        unsigned long long p0 = 0xD2511F53ULL * c0, p1 = 0xCD9E8D57ULL * c2;
        unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
        unsigned n1 = (unsigned)p1;
        unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
        unsigned n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

double orc_uniform(orc_rng *r, const orc_addr *a)
{
    double u;
    if (r->mode == 0) {
        if (r->pos >= r->n) {
            fprintf(stderr, "orc_uniform: stream exhausted at %ld\n", r->pos);
            abort();
        }
        u = r->u[r->pos++];
    } else {
        unsigned ctr[4], key[2], w[4];
        ctr[0] = a->slot >> 2;
        ctr[1] = (a->purpose << 28) | ((a->sweep & 0xFFFu) << 16) | (a->chain & 0xFFFFu);
        ctr[2] = a->pop;
        ctr[3] = a->iter;
        key[0] = (unsigned)(r->seed & 0xFFFFFFFFu);
        key[1] = (unsigned)(r->seed >> 32);
        orc_philox4x32_10(ctr, key, w);
        u = ((double)w[a->slot & 3u] + 0.5) * (1.0 / 4294967296.0);
    }
    if (r->rec) {
        if (r->rec_n >= r->rec_cap) {
            fprintf(stderr, "orc_uniform: record buffer full\n");
            abort();
        }
        r->rec[r->rec_n++] = u;
    }
    return u;
}

/* nmath runif(a, b): a == b returns a WITHOUT consuming a draw */
static double runif_(orc_rng *r, const orc_addr *a, double lo, double hi)
{
    if (lo == hi) return lo;
    return lo + (hi - lo) * orc_uniform(r, a);
}

/* ------------------------------------------------------------------------------------------- */
/* LBA node-1 density                                                                           */
/* ------------------------------------------------------------------------------------------- */

/* design_class::set_parameter_values, @hdr/design_light.h:314-344: rows
 * A, B, mean_v, sd_v, st0, t0; then row B += row A so the density sees b = A + B. */
void orc_cell_params(const orc_model *m, const double *theta, int cell, double *P)
{
    int na = m->n_acc, rows = m->type == ORC_MODEL_DDM ? ORC_DDM_ROWS : ORC_LBA_ROWS;
    const int *src = m->param_src + (size_t)cell * rows * na;
    for (int r = 0; r < rows; ++r)
        for (int j = 0; j < na; ++j) {
            int s = src[r * na + j];
            P[r * na + j] = s >= 0 ? theta[s] : m->const_val[-1 - s];
        }
    if (m->type != ORC_MODEL_DDM) /* :336-340 only the row NAMED "B" gets += the row above it; the DDM has no such row */
        for (int j = 0; j < na; ++j) P[1 * na + j] += P[0 * na + j];
}

/* lba_class::{set_parameters :88-119, validate_parameters :121-146, dlba :560-572, d :213-248,
 * p :286-345} of @hdr/lba.h, and the invalid-cell rule of @hdr/likelihood.h:105.
 * P rows: A, b, mean_v, sd_v, st0, t0.  u_st0[j] is the uniform drawn for t0 + st0*U (:117). */
int orc_lba_cell(const double *P, int na, const unsigned char *posdrift, const double *u_st0, const double *rt, int n,
                 double *out)
{
    const double *A = P, *b = P + na, *mv = P + 2 * na, *sv = P + 3 * na, *st0 = P + 4 * na, *t0 = P + 5 * na;
    double denom[16], t0a[16];
    int valid = 1;
    if (na > 16) abort();
    for (int j = 0; j < na; ++j) {
        denom[j] = posdrift[j] ? fmax(orc_pnorm5(mv[j] / sv[j], 0.0, 1.0, 1, 0), FLOOR_) : 1.0; /* :112-115 */
        t0a[j] = t0[j] + st0[j] * (u_st0 ? u_st0[j] : 0.0);                                        /* :117 */
    }
    for (int j = 0; j < na; ++j) /* :121-146; comparisons are false on NaN */
        if (A[j] < 0 || b[j] < 0 || b[j] < A[j] || sv[j] < 0 || st0[j] < 0 || t0[j] < 0) valid = 0;
    if (!valid) {
        for (int i = 0; i < n; ++i) out[i] = FLOOR_; /* likelihood.h:105 */
        return 0;
    }
    for (int i = 0; i < n; ++i) {
        double pdf, dt = rt[i] - t0a[0];
        if (0 > dt) { /* :217-219 */
            pdf = FLOOR_;
        } else if (A[0] < FLOOR_) { /* :221-227 */
            /* object code (de.o, lba_class::d +0xf6..+0x198): the normal density is divided by the drift denominator
             * first (:224-225), THEN multiplied by b / (dt * dt) (:226) -- one rounding apart from the left-to-right product */
            double term = orc_dnorm4(b[0] / dt, mv[0], sv[0], 0) / denom[0];
            pdf = fmax(b[0] / (dt * dt) * term, FLOOR_);
        } else { /* :231-244 */
            double ts = sv[0] * dt, tv = mv[0] * dt;
            double t1 = mv[0] * (orc_pnorm5((b[0] - tv) / ts, 0.0, 1.0, 1, 0) - orc_pnorm5((b[0] - A[0] - tv) / ts, 0.0, 1.0, 1, 0));
            double t2 = sv[0] * (orc_dnorm4((b[0] - A[0] - tv) / ts, 0.0, 1.0, 0) - orc_dnorm4((b[0] - tv) / ts, 0.0, 1.0, 0));
            pdf = fmax((t1 + t2) / (A[0] * denom[0]), FLOOR_);
        }
        if (isnan(pdf)) pdf = FLOOR_; /* :247 */
        for (int j = 1; j < na; ++j) { /* p(), :286-345 */
            double cdf;
            dt = rt[i] - t0a[j];
            if (0 > dt) {
                cdf = FLOOR_; /* :310-312 */
            } else if (A[j] < FLOOR_) { /* :315-320 */
                cdf = orc_pnorm5(b[j] / dt, mv[j], sv[j], 0, 0) / denom[j];
                cdf = cdf < FLOOR_ ? FLOOR_ : (1.0 < cdf ? 1.0 : cdf); /* std::clamp */
            } else { /* :324-338 */
                double ts = sv[j] * dt, tv = mv[j] * dt, x1 = b[j] - tv, x2 = x1 - A[j];
                double z1 = x1 / ts, z2 = x2 / ts;
                cdf = (1.0 + (x2 * orc_pnorm5(z2, 0.0, 1.0, 1, 0) - x1 * orc_pnorm5(z1, 0.0, 1.0, 1, 0) +
                              ts * (orc_dnorm4(z2, 0.0, 1.0, 0) - orc_dnorm4(z1, 0.0, 1.0, 0))) / A[j]) / denom[j];
                cdf = cdf < FLOOR_ ? FLOOR_ : (1.0 < cdf ? 1.0 : cdf); /* std::clamp: NaN passes through */
            }
            pdf = pdf * (1.0 - cdf); /* :341 */
            if (isnan(pdf)) pdf = FLOOR_; /* :342 */
        }
        out[i] = pdf;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------- */
/* DDM ("fastdm") first-passage density                                                         */
/* ------------------------------------------------------------------------------------------- */
/* ddm::ddm_class of @hdr/ddm.h (anonymous namespace; compiled into src/de.o at -O0 with line
 * tables, decoded with objdump -d -C -l; member offsets from its DWARF).  The arithmetic is the
 * fast-dm density of Voss & Voss (2007) / Navarro & Fuss (2009): small- and large-time series for the
 * lower-boundary density, closed-form integral over the drift variability sv, midpoint rules over the
 * start-point range sz and the non-decision range st0.  Operation ORDER below follows the object code
 * so that the restatement is bit-identical to it (checked in tests/test_oracle_cpu.py through
 * likelihood_class::ddm_likelihood of de.o). */
typedef struct {
    double a, v, sv, st0, zr, szr, t_offset; /* m_a @0, m_v @8, m_sv @0x28, m_st0 @0x30, m_zr @0x38, m_szr @0x20, m_t_offset @0x58 */
    double s;                                /* m_s @0x40 */
    double int_t0, int_z, sz_eps, st0_eps;   /* TUNE_INT_T0 @0xa0, TUNE_INT_Z @0xa8, TUNE_SZ_EPSILON @0xb8, TUNE_ST0_EPSILON @0xc0 */
    double a2, v2, sv2;                      /* @0xe8, @0xf0, @0xf8 */
} ddm_par;

#define DDM_EPSILON 1e-6                 /* ddm::EPSILON, .rodata+0x1b8 */
#define DDM_PI 3.141592653589793         /* .rodata+0xec8 */
#define DDM_2PI 6.283185307179586        /* .rodata+0xec0 */
#define DDM_PI2 9.869604401089358        /* m_pi2 @0xe0, set by the constructor (@hdr/ddm.h:141) */

/* x86 cvttsd2si: NaN and out-of-range values give INT_MIN (the object code converts with it) */
static int cvt_trunc(double x) { return (x >= -2147483648.0 && x < 2147483648.0) ? (int)x : (-2147483647 - 1); }

/* set_parameters(matrix, is_lower), @hdr/ddm.h:187-226, with set_precision :120-138.
 * P = column 0 of rows a, d, precision, s, st0, sv, sz, t0, v, z. */
static void ddm_set(ddm_par *q, const double *P, int is_lower)
{
    double s = P[3], scale = (1.0 != s) ? 1.0 / s : 1.0, prec = P[2], tmp;   /* :194-195 */
    q->s = s;
    q->st0 = P[4];                                                            /* :201 */
    q->a = P[0] * scale;                                                      /* :203 */
    q->sv = P[5] * scale;                                                     /* :204 */
    q->v = is_lower ? P[8] * scale : (-P[8]) * scale;                         /* :205-206 */
    q->t_offset = 0.5 * P[4] + P[7];                                          /* :208 */
    tmp = P[9] / q->a;                                                        /* :216 */
    q->zr = is_lower ? tmp : 1.0 - tmp;                                       /* :217 */
    q->szr = P[6] / q->a;                                                     /* :218 */
    q->int_t0 = exp(prec * -1.03758) * 0.089045;                              /* :130 */
    q->int_z = exp(prec * -1.022373) * 0.508061;                              /* :131 */
    q->sz_eps = pow(10.0, -(2.0 + prec));                                     /* :136 */
    q->st0_eps = pow(10.0, -(2.0 + prec));                                    /* :137 */
    q->a2 = q->a * q->a; q->v2 = q->v * q->v; q->sv2 = q->sv * q->sv;         /* :223-225 */
}

/* validate_parameters, @hdr/ddm.h:229-309 (comparisons false on NaN) */
static int ddm_valid(const ddm_par *q)
{
    int ok = 1;
    if (q->a <= 0) ok = 0;                          /* :232 */
    if (q->szr < 0 || q->szr > 1.0) ok = 0;         /* :240 */
    if (q->st0 < 0) ok = 0;                         /* :250 */
    if (q->sv < 0) ok = 0;                          /* :259 */
    if (q->t_offset < 0) ok = 0;                    /* :268 */
    if (q->zr - 0.5 * q->szr <= 0) ok = 0;          /* :278 */
    if (q->zr + 0.5 * q->szr >= 1.0) ok = 0;        /* :288 */
    if (q->s <= 0) ok = 0;                          /* :298 */
    return ok;
}

/* compute_g_series, @hdr/ddm.h:344-379 */
static double ddm_series(double ta, double zr, int use_small, int N)
{
    double sum = 0.0;
    if (use_small) {
        double t3 = ta * ta * ta;                   /* :356 */
        double norm = 1.0 / sqrt(t3 * DDM_2PI);     /* :357 */
        int lo = -(N / 2), hi = N / 2;              /* :359-360 */
        for (int i = lo; i <= hi; ++i) {
            double d = ((double)i + (double)i) + zr;            /* :364 */
            sum = exp((-d * d) / (ta + ta)) * d + sum;          /* :365 */
        }
        return sum * norm;                          /* :367 */
    }
    for (int i = 1; i <= N; ++i) {                  /* :372 */
        double d = DDM_PI * (double)i;              /* :374 */
        sum = (double)i * (exp(-0.5 * d * d * ta) * sin(d * zr)) + sum; /* :375 */
    }
    return DDM_PI * sum;                            /* :377 */
}

/* compute_g_factor, @hdr/ddm.h:383-405 */
static double ddm_factor(const ddm_par *q, double t, double zr, int no_var)
{
    double f;
    if (no_var) {
        f = exp((-q->a * zr) * q->v - (0.5 * q->v2) * t) / q->a2;                                       /* :388 */
    } else {
        double denom = 1.0 + q->sv2 * t;                                                                 /* :396 */
        double e = (-0.5 * ((q->v2 * t + (q->a * (q->v + q->v)) * zr) - ((q->a2 * zr) * zr) * q->sv2)) / denom; /* :397-398 */
        f = exp(e) / (q->a2 * sqrt(denom));                                                              /* :401 */
    }
    return isfinite(f) ? f : 0.0;                                                                        /* :389, :403 */
}

/* get_N, @hdr/ddm.h:408-430: number of terms of the small-time and of the large-time series */
static void ddm_get_n(double t, double ta, double eps, int *n_small, int *n_large)
{
    int nl = cvt_trunc(ceil(1.0 / (DDM_PI * sqrt(t)))), ns;                     /* :409 (t, not t / a^2) */
    if (1.0 > (DDM_PI * ta) * eps) {                                            /* :410 */
        double x = (log((DDM_PI * ta) * eps) * -2.0) / (DDM_PI2 * ta);          /* :412 */
        int k = cvt_trunc(ceil(sqrt(x)));                                       /* :413 */
        if (nl < k) nl = k;                                                     /* :414 std::max */
    }
    if (1.0 > (sqrt(ta * DDM_2PI) + sqrt(ta * DDM_2PI)) * eps) {                /* :418 */
        double lg = log(sqrt(ta * DDM_2PI) * (eps + eps));                      /* :420 */
        double t1 = sqrt((-2.0 * ta) * lg) + 2.0;                               /* :421 */
        double t2 = sqrt(ta) + 1.0;                                             /* :422 */
        ns = cvt_trunc(ceil(t2 < t1 ? t1 : t2));                                /* :423 std::max(t2, t1) */
    } else
        ns = 2;                                                                 /* :427 */
    *n_small = ns;
    *n_large = nl;
}

/* work counters for the measurement tools (tools/exp_ddm.py turns them into algorithmic flops per trial):
 * series evaluations that got past the t > 0 and factor != 0 tests, and the terms they summed */
static long long g_ddm_evals, g_ddm_small_terms, g_ddm_large_terms;
void orc_ddm_counters(long long out[3], int reset)
{
    out[0] = g_ddm_evals; out[1] = g_ddm_small_terms; out[2] = g_ddm_large_terms;
    if (reset) g_ddm_evals = g_ddm_small_terms = g_ddm_large_terms = 0;
}

/* integral_v, @hdr/ddm.h:457-485, and g_no_var :433-454 (the sv == 0 branch; same steps, other factor) */
static double ddm_integral_v(const ddm_par *q, double t, double zr)
{
    int no_var, ns, nl, use_small;
    double ta, factor, eps;
    if (0 >= t) return 0.0;                          /* :459 / :434 */
    no_var = q->sv == 0;                             /* :463 */
    ta = t / q->a2;                                  /* :469 / :439 */
    factor = ddm_factor(q, t, zr, no_var);           /* :472 / :440 */
    if (factor == 0) return 0.0;                     /* :474 / :444 */
    eps = DDM_EPSILON / factor;                      /* :479 / :449 */
    ddm_get_n(t, ta, eps, &ns, &nl);                 /* :480 / :450 */
    use_small = ns < nl;                             /* :481 / :451 */
    ++g_ddm_evals;
    if (use_small) g_ddm_small_terms += ns > 0 ? 2 * (ns / 2) + 1 : 0; else g_ddm_large_terms += nl > 0 ? nl : 0;
    return ddm_series(ta, zr, use_small, use_small ? ns : nl) * factor; /* :484 / :453 */
}

/* integral_z :508-514 and integrate_v_over_zr :488-505 (midpoint rule over the start-point range) */
static double ddm_integral_z(const ddm_par *q, double t)
{
    double lower, upper, width, step, sum = 0.0;
    int n;
    if (q->sz_eps > q->szr) return ddm_integral_v(q, t, q->zr);  /* :512 */
    lower = q->zr - 0.5 * q->szr;                                /* :489 */
    upper = 0.5 * q->szr + q->zr;                                /* :490 */
    width = upper - lower;                                       /* :491 */
    n = cvt_trunc(width / q->int_z);                             /* :492-494 */
    if (n < 4) n = 4;
    step = width / (double)n;                                    /* :495 */
    for (double x = 0.5 * step + lower; upper > x; x += step)    /* :498 */
        sum = ddm_integral_v(q, t, x) * step + sum;              /* :501 */
    return sum / q->szr;                                         /* :504 */
}

/* integral_t0 :537-542 and integrate_z_over_t :517-534 (midpoint rule over the non-decision range) */
static double ddm_integral_t0(const ddm_par *q, double t)
{
    double lower, upper, width, step, sum = 0.0;
    int n;
    if (q->st0_eps > q->st0) return ddm_integral_z(q, t);        /* :541 */
    lower = t - q->st0 * 0.5;                                    /* :518 */
    upper = 0.5 * q->st0 + t;                                    /* :519 */
    width = upper - lower;                                       /* :521 */
    n = cvt_trunc(width / q->int_t0);                            /* :522-523 */
    if (n < 4) n = 4;
    step = width / (double)n;                                    /* :524 */
    for (double x = 0.5 * step + lower; upper > x; x += step)    /* :528 */
        sum = ddm_integral_z(q, x) * step + sum;                 /* :530 */
    return sum / q->st0;                                         /* :533 */
}

/* ddm_likelihood's per-cell body, @hdr/likelihood.h:140-158: set_parameters(matrix, !is_positive_drift[cell]),
 * validate, then dddm (@hdr/ddm.h:552-560: g(rt) = integral_t0(rt - m_t_offset), :545-549) or 1e-10 for every trial. */
int orc_ddm_cell(const double *P, int is_upper, const double *rt, int n, double *out)
{
    ddm_par q;
    ddm_set(&q, P, !is_upper);
    if (!ddm_valid(&q)) {
        for (int i = 0; i < n; ++i) out[i] = FLOOR_; /* likelihood.h:158 */
        return 0;
    }
    for (int i = 0; i < n; ++i) out[i] = ddm_integral_t0(&q, rt[i] - q.t_offset);
    return 1;
}

/* densities of one cell's trials for either model family; returns validity */
static int cell_density(const orc_model *m, const double *theta, int c, const double *u_st0, const double *rt, int n, double *out)
{
    double P[ORC_DDM_ROWS * 16];
    orc_cell_params(m, theta, c, P);
    if (m->type == ORC_MODEL_DDM) {
        double col0[ORC_DDM_ROWS];
        for (int r = 0; r < ORC_DDM_ROWS; ++r) col0[r] = P[r * m->n_acc]; /* the DDM reads column 0 only (@hdr/ddm.h:194-214) */
        return orc_ddm_cell(col0, m->posdrift[c] != 0, rt, n, out);
    }
    return orc_lba_cell(P, m->n_acc, m->posdrift, u_st0, rt, n, out);
}

/* the log the sampler takes of one density: LBA plain (@hdr/likelihood.h:288), DDM floored at DBL_MIN (:303) */
static double log_density(const orc_model *m, double x)
{
    if (m->type == ORC_MODEL_DDM) return log(x < DBL_MIN ? DBL_MIN : x); /* std::max(x, DBL_MIN): NaN stays NaN */
    return log(x);
}

static void check_grouped(const orc_data *d)
{
    for (int i = 1; i < d->n_trial; ++i)
        if (d->cell[i] < d->cell[i - 1]) {
            fprintf(stderr, "oracle: trials must be grouped by ascending cell index\n");
            abort();
        }
}

/* per-trial log density; likelihood_class::lba_likelihood, @hdr/likelihood.h:73-108, without
 * the st0 draws (all fixtures and benchmark models have st0 = 0 => t0 + 0*U = t0 exactly) */
void orc_trial_logdens(const orc_model *m, const orc_data *d, const double *theta, double *out)
{
    check_grouped(d);
    int i = 0;
    while (i < d->n_trial) {
        int c = d->cell[i], j = i;
        while (j < d->n_trial && d->cell[j] == c) ++j;
        cell_density(m, theta, c, NULL, d->rt + i, j - i, out + i);
        for (int k = i; k < j; ++k) out[k] = log_density(m, out[k]);
        i = j;
    }
}

/* likelihood_class::sumloglike -> lba_likelihood, @hdr/likelihood.h:272-292 and :73-108:
 * cells in model order (empty cells skipped, :82), trials in data order, one FP64 accumulator.
 * Draw order per call: n_acc uniforms per non-empty cell (@hdr/lba.h:117). */
double orc_sumloglike(const orc_model *m, const orc_data *d, const double *theta, orc_rng *r, const orc_addr *base)
{
    double u[16], dens[4096], *buf = dens, out = 0.0;
    int na = m->n_acc, i = 0, is_lba = m->type != ORC_MODEL_DDM; /* the DDM code draws no uniforms */
    if (is_lba && r && r->mode == 0 && r->burn_static_ctor && !r->first_like_done) {
        orc_addr a = {0};
        orc_uniform(r, &a);
        orc_uniform(r, &a);
    }
    if (r) r->first_like_done = 1;
    while (i < d->n_trial) {
        int c = d->cell[i], j = i;
        while (j < d->n_trial && d->cell[j] == c) ++j;
        for (int k = 0; k < na; ++k) {
            if (r && is_lba) {
                orc_addr a = *base;
                a.purpose = ORC_U_ST0;
                a.slot = (unsigned)(c * na + k);
                u[k] = orc_uniform(r, &a);
            } else
                u[k] = 0.0;
        }
        if (j - i > 4096) buf = (double *)malloc(sizeof(double) * (size_t)(j - i));
        cell_density(m, theta, c, u, d->rt + i, j - i, buf);
        for (int k = 0; k < j - i; ++k) out += log_density(m, buf[k]); /* :284-288, :298-303 */
        if (buf != dens) { free(buf); buf = dens; }
        i = j;
    }
    return out;
}

/* R-side init path: R/phi.R:3-13 (.sumlog) floors densities <= 0 at .Machine$double.eps; the
 * fixture goldens `log_likelihoods[,1]` were produced by this path (R/phi.R:176-177). */
double orc_sumloglike_rinit(const orc_model *m, const orc_data *d, const double *theta)
{
    double out = 0.0;
    int i = 0;
    check_grouped(d);
    while (i < d->n_trial) {
        int c = d->cell[i], j = i;
        double s = 0.0;
        while (j < d->n_trial && d->cell[j] == c) ++j;
        double *buf = (double *)malloc(sizeof(double) * (size_t)(j - i));
        cell_density(m, theta, c, NULL, d->rt + i, j - i, buf);
        int any = 0;
        for (int k = 0; k < j - i; ++k) if (buf[k] <= 0) any = 1;
        for (int k = 0; k < j - i; ++k) {
            double x = buf[k];
            if (any) x = isnan(x) ? DBL_EPSILON : (x > DBL_EPSILON ? x : DBL_EPSILON); /* pmax(xi, eps, na.rm=TRUE) */
            s += log(x);
        }
        out += s; /* sum over cells of per-cell sums */
        free(buf);
        i = j;
    }
    return out;
}

double orc_time_sumloglike(const orc_model *m, const orc_data *d, const double *thetas, int nchain, int reps)
{
    double v = 0.0;
    check_grouped(d);
    for (int r = 0; r < reps; ++r)
        for (int k = 0; k < nchain; ++k) v += orc_sumloglike(m, d, thetas + (size_t)k * m->npar, NULL, NULL);
    return v;
}

/* ------------------------------------------------------------------------------------------- */
/* priors                                                                                       */
/* ------------------------------------------------------------------------------------------- */

/* tnorm_class::set_parameters @hdr/tnorm.h:59-67 then d(x) :118-127 */
double orc_tnorm_d(double x, double mean, double sd, double lower, double upper, int log_p)
{
    double denom = orc_pnorm5(upper, mean, sd, 1, 0) - orc_pnorm5(lower, mean, sd, 1, 0);
    double log_denom = log(denom);
    if (x < lower || x > upper) return log_p ? -INFINITY : FLOOR_;
    return log_p ? orc_dnorm4(x, mean, sd, 1) - log_denom : orc_dnorm4(x, mean, sd, 0) / denom;
}

/* prior_class::dcauchy_trunc, @hdr/prior.h:25-58 */
static double dcauchy_trunc(double x, double loc, double scale, double lower, double upper, int log_p)
{
    if (0.0 >= scale || !(lower < upper)) return NAN;
    double den = orc_dcauchy(x, loc, scale, 0);
    double Fu = orc_pcauchy(upper, loc, scale, 1, 0), Fl = orc_pcauchy(lower, loc, scale, 1, 0);
    double out = 0.0;
    if (x >= lower && upper >= x) out = den / (Fu - Fl);
    if (log_p) return out > 0.0 ? log(out) : -INFINITY;
    return out;
}

/* prior_class::dprior(arma), @hdr/prior.h:331-412.  p0/p1 are passed separately because the
 * hierarchical sampler overwrites m_p0/m_p1 from phi before each call (src/de.cpp:250-251,
 * 599-600, 646-649) while lower/upper/dist/log_p stay fixed. */
void orc_dprior(const orc_prior *p, const double *p0, const double *p1, const double *x, double *out)
{
    for (int i = 0; i < p->npar; ++i) {
        int lg = p->log_p[i];
        double v;
        switch (p->dist[i]) {
        case ORC_TNORM: /* :349-350 */
            v = orc_tnorm_d(x[i], p0[i], p1[i], p->lower[i], p->upper[i], lg);
            break;
        case ORC_BETA_LU: { /* :358-367; m_beta_range taken per parameter as upper - lower */
            double range = p->upper[i] - p->lower[i];
            double xs = (x[i] - p->lower[i]) / range, den = -INFINITY;
            if (p0[i] >= 0 && p1[i] >= 0) den = orc_dbeta(xs, p0[i], p1[i], lg);
            v = lg ? den - log(range) : den / range;
            break;
        }
        case ORC_GAMMA_L: /* :375-376, lambda :336-340: x - lower when lower is finite */
            v = orc_dgamma(isfinite(p->lower[i]) ? x[i] - p->lower[i] : x[i], p0[i], p1[i], lg);
            break;
        case ORC_LNORM_L: /* :377 */
            v = orc_dlnorm(isfinite(p->lower[i]) ? x[i] - p->lower[i] : x[i], p0[i], p1[i], lg);
            break;
        case ORC_CAUCHY: /* :383-385 */
            v = dcauchy_trunc(x[i], p0[i], p1[i], p->lower[i], p->upper[i], lg);
            break;
        case ORC_UNIF: /* :393-394 */
            v = orc_dunif(x[i], p0[i], p1[i], lg);
            if (isnan(v)) v = -1e10;
            break;
        case ORC_NORM: /* :399 */
            v = orc_dnorm4(x[i], p0[i], p1[i], lg);
            break;
        default: /* :405-406 */
            v = NAN;
        }
        out[i] = v;
    }
}

/* prior_class::sumlogprior, @hdr/prior.h:469-476: arma::accu of dprior.
 * arma::accu on a small vector is a two-accumulator pairwise loop (accu_proxy_linear: even
 * indices into val1, odd into val2, tail into val1, return val1 + val2). */
double orc_sumlogprior(const orc_prior *p, const double *p0, const double *p1, const double *x)
{
    double buf[256], v1 = 0.0, v2 = 0.0;
    int n = p->npar, i;
    if (n > 256) abort();
    orc_dprior(p, p0 ? p0 : p->p0, p1 ? p1 : p->p1, x, buf);
    for (i = 0; i + 1 < n; i += 2) { v1 += buf[i]; v2 += buf[i + 1]; }
    if (i < n) v1 += buf[i];
    return v1 + v2;
}

/* de_class::sumloghlike, src/de.cpp:245-270 */
double orc_sumloghlike(const orc_prior *p_prior, const double *phi, int chain, const double *const *thetas, int nsubject)
{
    int np = p_prior->npar;
    double out = 0.0;
    for (int s = 0; s < nsubject; ++s) out += orc_sumlogprior(p_prior, phi, phi + np, thetas[s] + (size_t)chain * np);
    return out;
}

/* ------------------------------------------------------------------------------------------- */
/* chain selection                                                                              */
/* ------------------------------------------------------------------------------------------- */

typedef struct { int key; unsigned idx; unsigned pos; } packet;
static int packet_cmp(const void *a, const void *b)
{
    const packet *x = (const packet *)a, *y = (const packet *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

/* arma::shuffle as compiled into de.o (op_shuffle_meat.hpp + RcppArmadillo Alt_R_RNG.h:75):
 * key_i = (int) Rf_runif(0, 2147483647) in element order; sort (key, i) ascending by key.
 * std::sort is not stable; equal keys (probability ~ n^2 / 2^32 per shuffle) are broken here by
 * position, which is what std::sort's insertion-sort path does for n <= 16. */
static void shuffle_(const unsigned *v, int n, orc_rng *r, const orc_addr *base, unsigned purpose, unsigned *out)
{
    packet *pk = (packet *)malloc(sizeof(packet) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        orc_addr a = *base;
        a.purpose = purpose;
        a.slot = (unsigned)i;
        pk[i].key = (int)runif_(r, &a, 0.0, 2147483647.0);
        pk[i].idx = v[i];
        pk[i].pos = (unsigned)i;
    }
    qsort(pk, (size_t)n, sizeof(packet), packet_cmp);
    for (int i = 0; i < n; ++i) out[i] = pk[i].idx;
    free(pk);
}

/* de_class::get_chains, src/de.cpp:54-60 */
void orc_get_chains(int nchain, int k, int nsub, orc_rng *r, const orc_addr *base, unsigned *out)
{
    unsigned *v = (unsigned *)malloc(sizeof(unsigned) * (size_t)nchain * 2), *sh = v + nchain;
    int n = 0;
    for (int i = 0; i < nchain; ++i)
        if (i != k) v[n++] = (unsigned)i;
    shuffle_(v, n, r, base, ORC_U_PARTNER, sh);
    for (int i = 0; i < nsub; ++i) out[i] = sh[i];
    free(v);
}

static int uint_cmp(const void *a, const void *b)
{
    unsigned x = *(const unsigned *)a, y = *(const unsigned *)b;
    return x < y ? -1 : (x > y);
}

/* de_class::get_subchains, src/de.cpp:62-78 */
int orc_get_subchains(int nchain, orc_rng *r, const orc_addr *base, unsigned *out)
{
    orc_addr a = *base;
    a.purpose = ORC_U_MIG_N;
    a.slot = 0;
    double proportion = runif_(r, &a, 0.0, 1.0);
    unsigned n = (unsigned)ceil(nchain * proportion);
    if (n < 2u) n = 2u;
    if (n > (unsigned)nchain) n = (unsigned)nchain;
    unsigned *v = (unsigned *)malloc(sizeof(unsigned) * (size_t)nchain * 2), *sh = v + nchain;
    for (int i = 0; i < nchain; ++i) v[i] = (unsigned)i;
    shuffle_(v, nchain, r, base, ORC_U_MIG_KEYS, sh);
    for (unsigned i = 0; i < n; ++i) out[i] = sh[i];
    qsort(out, n, sizeof(unsigned), uint_cmp);
    free(v);
    return (int)n;
}

/* ------------------------------------------------------------------------------------------- */
/* samplers                                                                                     */
/* ------------------------------------------------------------------------------------------- */

/* theta_phi::store, @hdr/theta.h:61-74 */
void orc_store(orc_pop *p, unsigned i)
{
    if (i % (unsigned)p->thin != 0) return;
    p->store_i++;
    if (p->store_i < p->nmc) {
        size_t s = (size_t)p->store_i;
        memcpy(p->out_lp + s * p->nchain, p->lp, sizeof(double) * (size_t)p->nchain);
        memcpy(p->out_ll + s * p->nchain, p->ll, sizeof(double) * (size_t)p->nchain);
        memcpy(p->out_theta + s * p->nchain * p->npar, p->theta, sizeof(double) * (size_t)p->nchain * p->npar);
    }
}

/* Generic likelihood / prior hooks so the three crossover/migration variants of de.cpp share
 * one body here while keeping each variant's exact evaluation order. */
typedef struct {
    /* 1-level LBA (run_subject) or subject-in-hierarchy */
    const orc_model *m;
    const orc_data *d;
    /* hyper-only (run_hyper): data_theta [nsubject][npar_sub] scored under p_prior(phi) */
    const double *data_theta;
    /* phi level of the hierarchy: subject states */
    const double *const *subj_thetas;
    int nsubject;
    const orc_prior *prior;   /* prior of the moved vector: p_prior (subject), h_prior (phi / hyper) */
    const orc_prior *p_prior; /* hyper-likelihood object for phi / hyper levels */
    const double *phi_theta;  /* subject-in-hierarchy: phi's current state [nchain][2 npar] */
    int kind;                 /* 0 subject (fixed prior), 1 hyper-only, 2 phi of hierarchy, 3 subject in hierarchy */
} ctx_t;

static double like_(const ctx_t *c, const double *x, int chain, orc_rng *r, const orc_addr *base)
{
    switch (c->kind) {
    case 0:
    case 3:
        return orc_sumloglike(c->m, c->d, x, r, base);
    case 1: { /* likelihood_class::sumloghlike, @hdr/likelihood.h:257-271 */
        int np = c->p_prior->npar;
        double out = 0.0;
        for (int s = 0; s < c->nsubject; ++s) out += orc_sumlogprior(c->p_prior, x, x + np, c->data_theta + (size_t)s * np);
        return out;
    }
    default:
        return orc_sumloghlike(c->p_prior, x, chain, c->subj_thetas, c->nsubject);
    }
}

static double prior_(const ctx_t *c, const double *x, int phi_chain)
{
    if (c->kind == 3) { /* src/de.cpp:599-603 / 646-652 */
        int np = c->prior->npar;
        const double *ph = c->phi_theta + (size_t)phi_chain * 2 * np;
        return orc_sumlogprior(c->prior, ph, ph + np, x);
    }
    return orc_sumlogprior(c->prior, NULL, NULL, x);
}

/* update_theta, src/de.cpp:81-108 (and the inlined copy :433-463) */
static void accept_(orc_pop *t, int target, const double *tmp, double tmp_lp, double tmp_ll, double mh, orc_rng *r,
                    const orc_addr *base)
{
    if (isnan(mh)) return; /* no draw consumed */
    orc_addr a = *base;
    a.purpose = ORC_U_ACCEPT;
    a.slot = 0;
    if (runif_(r, &a, 0.0, 1.0) < mh) {
        memcpy(t->theta + (size_t)target * t->npar, tmp, sizeof(double) * (size_t)t->npar);
        t->lp[target] = tmp_lp;
        t->ll[target] = tmp_ll;
    }
}

/* crossover: src/de.cpp:111-155 (kind 0/1), :385-465 (kind 2), :567-613 (kind 3).
 * nmove = number of leading parameters perturbed when para_idx < 0 (m_nparameter for kinds 0-2,
 * m_half_nparameter for kind 3, :592). */
static void crossover_(const orc_de *de, orc_pop *t, const ctx_t *c, orc_rng *r, unsigned pop, unsigned iter, int para_idx,
                       int nmove)
{
    int nc = t->nchain, np = t->npar;
    double gamma = de->gamma_precursor / sqrt(2.0 * de->nparameter); /* src/de.cpp:12,24 */
    double *tmp = (double *)malloc(sizeof(double) * (size_t)np);
    double *snap_theta = NULL, *snap_lp = NULL, *snap_ll = NULL;
    const double *src = t->theta, *slp = t->lp, *sll = t->ll;
    int sched = de->jacobi;
    int nhalf = (sched == 1) ? 2 : 1; /* 1: two half-sweeps (even chains, then odd chains) */
    unsigned *cand = (unsigned *)malloc(sizeof(unsigned) * (size_t)nc * 2), *sh = cand + nc;
    if (sched) {
        snap_theta = (double *)malloc(sizeof(double) * (size_t)nc * np);
        snap_lp = (double *)malloc(sizeof(double) * (size_t)nc * 2);
        snap_ll = snap_lp + nc;
        src = snap_theta; slp = snap_lp; sll = snap_ll;
    }
    for (int half = 0; half < nhalf; ++half) {
        if (sched) { /* proposals of this (half-)sweep are made from its start state */
            memcpy(snap_theta, t->theta, sizeof(double) * (size_t)nc * np);
            memcpy(snap_lp, t->lp, sizeof(double) * (size_t)nc);
            memcpy(snap_ll, t->ll, sizeof(double) * (size_t)nc);
        }
        for (int i = 0; i < nc; ++i) {
            if (sched == 1 && (i & 1) != half) continue;
            orc_addr base = {pop, iter, para_idx < 0 ? 0u : (unsigned)para_idx, (unsigned)i, 0, 0};
            double cur_ll = sll[i];
            if (c->kind == 2) { /* :397-398 refresh the hyper-likelihood of the current phi */
                cur_ll = like_(c, src + (size_t)i * np, i, r, &base);
                t->ll[i] = cur_ll;
            }
            double cur = cur_ll + slp[i];
            memcpy(tmp, src + (size_t)i * np, sizeof(double) * (size_t)np);
            unsigned sub[2];
            if (sched == 1) { /* partners only from the other half, which stands still during this half-sweep */
                int n = 0;
                for (int k = 0; k < nc; ++k)
                    if ((k & 1) != half) cand[n++] = (unsigned)k;
                shuffle_(cand, n, r, &base, ORC_U_PARTNER, sh);
                sub[0] = sh[0];
                sub[1] = sh[1];
            } else {
                orc_get_chains(nc, i, 2, r, &base, sub);
            }
            const double *th0 = src + (size_t)sub[0] * np, *th1 = src + (size_t)sub[1] * np;
            orc_addr a = base;
            a.purpose = ORC_U_NOISE;
            if (para_idx >= 0) {
                a.slot = (unsigned)para_idx;
                tmp[para_idx] += runif_(r, &a, -de->rp, de->rp) + gamma * (th0[para_idx] - th1[para_idx]);
            } else {
                for (int j = 0; j < nmove; ++j) {
                    a.slot = (unsigned)j;
                    tmp[j] += runif_(r, &a, -de->rp, de->rp) + gamma * (th0[j] - th1[j]);
                }
            }
            double tmp_lp = prior_(c, tmp, i);
            double tmp_ll = like_(c, tmp, i, r, &base);
            double mh = exp(tmp_lp + tmp_ll - cur);
            accept_(t, i, tmp, tmp_lp, tmp_ll, mh, r, &base);
        }
    }
    free(tmp);
    free(cand);
    free(snap_theta);
    free(snap_lp);
}

/* migration: src/de.cpp:157-199 (kind 0/1), :467-565 (kind 2), :615-665 (kind 3) */
static void migration_(const orc_de *de, orc_pop *t, const ctx_t *c, orc_rng *r, unsigned pop, unsigned iter, int para_idx,
                       int nmove)
{
    int nc = t->nchain, np = t->npar;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)np);
    unsigned *S = (unsigned *)malloc(sizeof(unsigned) * (size_t)nc);
    orc_addr b0 = {pop, iter, para_idx < 0 ? 0u : (unsigned)para_idx, 0, 0, 0};
    int n = orc_get_subchains(nc, r, &b0, S);
    double *snap_theta = NULL, *snap_lp = NULL, *snap_ll = NULL;
    const double *src = t->theta, *slp = t->lp, *sll = t->ll;
    if (de->jacobi) {
        snap_theta = (double *)malloc(sizeof(double) * (size_t)nc * np);
        snap_lp = (double *)malloc(sizeof(double) * (size_t)nc * 2);
        snap_ll = snap_lp + nc;
        memcpy(snap_theta, t->theta, sizeof(double) * (size_t)nc * np);
        memcpy(snap_lp, t->lp, sizeof(double) * (size_t)nc);
        memcpy(snap_ll, t->ll, sizeof(double) * (size_t)nc);
        src = snap_theta; slp = snap_lp; sll = snap_ll;
    }
    for (int i = 0; i < n; ++i) {
        int cur_chain = (int)S[i];
        int next = (i + 1 == n) ? (int)S[0] : (int)S[i + 1];
        orc_addr base = {pop, iter, para_idx < 0 ? 0u : (unsigned)para_idx, (unsigned)cur_chain, 0, 0};
        double next_ll = sll[next];
        if (c->kind == 2) { /* :494-500 */
            if (!de->jacobi) { /* in-place order: refresh the source chain too (it may have just been replaced) */
                double l_cur = like_(c, src + (size_t)cur_chain * np, cur_chain, r, &base);
                t->ll[cur_chain] = l_cur;
            }
            /* snapshot order: every selected chain is `next` exactly once, so this refreshes them all */
            next_ll = like_(c, src + (size_t)next * np, next, r, &base);
            t->ll[next] = next_ll;
        }
        memcpy(tmp, src + (size_t)cur_chain * np, sizeof(double) * (size_t)np);
        orc_addr a = base;
        a.purpose = ORC_U_NOISE;
        if (para_idx >= 0) {
            a.slot = (unsigned)para_idx;
            tmp[para_idx] += runif_(r, &a, -de->rp, de->rp);
        } else {
            for (int j = 0; j < nmove; ++j) {
                a.slot = (unsigned)j;
                tmp[j] += runif_(r, &a, -de->rp, de->rp);
            }
        }
        double tmp_lp = prior_(c, tmp, cur_chain);
        double tmp_ll = like_(c, tmp, cur_chain, r, &base);
        double cur = slp[next] + next_ll;
        double mh = exp(tmp_lp + tmp_ll - cur);
        accept_(t, next, tmp, tmp_lp, tmp_ll, mh, r, &base);
    }
    free(tmp);
    free(S);
    free(snap_theta);
    free(snap_lp);
}

void orc_crossover_subject(const orc_de *de, orc_pop *t, const orc_prior *prior, const orc_model *m, const orc_data *d,
                           orc_rng *r, unsigned pop, unsigned iter, int para_idx)
{
    ctx_t c = {0};
    c.m = m; c.d = d; c.prior = prior; c.kind = 0;
    check_grouped(d);
    crossover_(de, t, &c, r, pop, iter, para_idx, de->nparameter);
}

void orc_migration_subject(const orc_de *de, orc_pop *t, const orc_prior *prior, const orc_model *m, const orc_data *d,
                           orc_rng *r, unsigned pop, unsigned iter, int para_idx)
{
    ctx_t c = {0};
    c.m = m; c.d = d; c.prior = prior; c.kind = 0;
    check_grouped(d);
    migration_(de, t, &c, r, pop, iter, para_idx, de->nparameter);
}

/* run_chains, src/de.cpp:201-242 */
static void run_chains_(const orc_de *de, orc_pop *t, const ctx_t *c, orc_rng *r, unsigned pop, unsigned n_iter)
{
    for (unsigned i = 1; i <= n_iter; ++i) {
        orc_addr a = {pop, i, 0, 0, ORC_U_DECIDE, 0};
        double rv = runif_(r, &a, 0.0, 1.0);
        if (rv < de->sub_migration_prob) {
            migration_(de, t, c, r, pop, i, -1, de->nparameter);
        } else if (de->is_pblocked) {
            for (int p = 0; p < de->nparameter; ++p) crossover_(de, t, c, r, pop, i, p, de->nparameter);
        } else {
            crossover_(de, t, c, r, pop, i, -1, de->nparameter);
        }
        orc_store(t, i);
    }
}

void orc_run_subject(const orc_de *de, orc_pop *t, const orc_prior *prior, const orc_model *m, const orc_data *d, orc_rng *r,
                     unsigned pop, unsigned n_iter)
{
    ctx_t c = {0};
    c.m = m; c.d = d; c.prior = prior; c.kind = 0;
    check_grouped(d);
    run_chains_(de, t, &c, r, pop, n_iter);
}

void orc_run_hyper(const orc_de *de, orc_pop *phi, const orc_prior *p_prior, const orc_prior *h_prior, const double *data_theta,
                   int nsubject, orc_rng *r, unsigned n_iter)
{
    ctx_t c = {0};
    c.data_theta = data_theta; c.nsubject = nsubject; c.prior = h_prior; c.p_prior = p_prior; c.kind = 1;
    run_chains_(de, phi, &c, r, ORC_POP_PHI, n_iter);
}

/* run_hchains, src/de.cpp:272-383 */
void orc_run_hier(const orc_de *de, orc_pop *phi, orc_pop *subj, int nsubject, const orc_prior *p_prior, const orc_prior *h_prior,
                  const orc_model *m, const orc_data *d, orc_rng *r, unsigned n_iter, unsigned first_subject_id)
{
    const double **st = (const double **)malloc(sizeof(double *) * (size_t)nsubject);
    int half = de->nparameter / 2;
    for (int s = 0; s < nsubject; ++s) { st[s] = subj[s].theta; check_grouped(&d[s]); }
    ctx_t cp = {0};
    cp.subj_thetas = st; cp.nsubject = nsubject; cp.prior = h_prior; cp.p_prior = p_prior; cp.kind = 2;
    for (unsigned it = 1; it <= n_iter; ++it) {
        if (de->is_hblocked) { /* :283-307 */
            for (int p = 0; p < de->nparameter; ++p) {
                orc_addr a = {ORC_POP_PHI, it, (unsigned)p, 0, ORC_U_DECIDE, 0};
                if (runif_(r, &a, 0.0, 1.0) < de->pop_migration_prob)
                    migration_(de, phi, &cp, r, ORC_POP_PHI, it, p, de->nparameter);
                else
                    crossover_(de, phi, &cp, r, ORC_POP_PHI, it, p, de->nparameter);
            }
        } else { /* :310-319 */
            orc_addr a = {ORC_POP_PHI, it, 0, 0, ORC_U_DECIDE, 0};
            if (runif_(r, &a, 0.0, 1.0) < de->pop_migration_prob)
                migration_(de, phi, &cp, r, ORC_POP_PHI, it, -1, de->nparameter);
            else
                crossover_(de, phi, &cp, r, ORC_POP_PHI, it, -1, de->nparameter);
        }
        for (int s = 0; s < nsubject; ++s) { /* :323-374 */
            unsigned pop = first_subject_id + (unsigned)s;
            ctx_t cs = {0};
            cs.m = m; cs.d = &d[s]; cs.prior = p_prior; cs.phi_theta = phi->theta; cs.kind = 3;
            if (de->is_pblocked) {
                for (int p = 0; p < half; ++p) {
                    orc_addr a = {pop, it, (unsigned)p, 0, ORC_U_DECIDE, 0};
                    if (runif_(r, &a, 0.0, 1.0) < de->sub_migration_prob)
                        migration_(de, &subj[s], &cs, r, pop, it, p, half);
                    else
                        crossover_(de, &subj[s], &cs, r, pop, it, p, half);
                }
            } else {
                orc_addr a = {pop, it, 0, 0, ORC_U_DECIDE, 0};
                if (runif_(r, &a, 0.0, 1.0) < de->sub_migration_prob)
                    migration_(de, &subj[s], &cs, r, pop, it, -1, half);
                else
                    crossover_(de, &subj[s], &cs, r, pop, it, -1, half);
            }
            orc_store(&subj[s], it);
        }
        orc_store(phi, it);
    }
    free(st);
}
