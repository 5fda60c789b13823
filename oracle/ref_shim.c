/* TEST INFRASTRUCTURE ONLY.
 * R-API shim that lets the reference's own object code (/root/reference/src/de.o) run without R.
 * de.o needs from libR only: Rf_runif, Rf_pnorm5, Rf_dnorm4, Rf_dunif, Rf_dbeta, Rf_dgamma,
 * Rf_dlnorm, Rf_dcauchy, Rf_pcauchy, R_NaN, R_NaReal, R_NegInf, Rprintf, REprintf,
 * R_FlushConsole, Rf_warning (nm -C src/de.o | grep ' U ').  The distribution functions come from
 * rmath_port.c (restated nmath algorithms); Rf_runif reads a caller-supplied uniform stream so
 * tests can feed the reference and the restatement the SAME draws.
 */
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "rmath_port.h"

double R_NaN, R_NaReal, R_NegInf, R_PosInf;
__attribute__((constructor)) static void init_consts(void)
{
    R_NaN = NAN; R_NaReal = NAN; R_NegInf = -INFINITY; R_PosInf = INFINITY;
}

static const double *g_stream = NULL;
static long g_stream_len = 0, g_stream_pos = 0;
static unsigned long long g_lcg = 0x9E3779B97F4A7C15ULL;

void ref_set_uniform_stream(const double *u, long n) { g_stream = u; g_stream_len = n; g_stream_pos = 0; }
long ref_uniform_stream_pos(void) { return g_stream_pos; }

static double next_uniform(void)
{
    if (g_stream) {
        if (g_stream_pos >= g_stream_len) {
            fprintf(stderr, "ref_shim: uniform stream exhausted at %ld\n", g_stream_pos);
            abort();
        }
        return g_stream[g_stream_pos++];
    }
    /* fallback generator when no stream is installed (timing runs): splitmix64 -> (0,1) */
    unsigned long long z = (g_lcg += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return ((double)(z >> 32) + 0.5) * (1.0 / 4294967296.0);
}

double Rf_runif(double a, double b)
{
    /* nmath/runif.c: a + (b - a) * unif_rand() */
    if (a == b) return a;
    return a + (b - a) * next_uniform();
}
double Rf_pnorm5(double x, double mu, double s, int lower, int lg) { return orc_pnorm5(x, mu, s, lower, lg); }
double Rf_dnorm4(double x, double mu, double s, int lg) { return orc_dnorm4(x, mu, s, lg); }
double Rf_dunif(double x, double a, double b, int lg) { return orc_dunif(x, a, b, lg); }
double Rf_dbeta(double x, double a, double b, int lg) { return orc_dbeta(x, a, b, lg); }
double Rf_dgamma(double x, double shape, double scale, int lg) { return orc_dgamma(x, shape, scale, lg); }
double Rf_dlnorm(double x, double ml, double sl, int lg) { return orc_dlnorm(x, ml, sl, lg); }
double Rf_dcauchy(double x, double l, double s, int lg) { return orc_dcauchy(x, l, s, lg); }
double Rf_pcauchy(double x, double l, double s, int lower, int lg) { return orc_pcauchy(x, l, s, lower, lg); }
void Rprintf(const char *fmt, ...) { (void)fmt; }
void REprintf(const char *fmt, ...) { (void)fmt; }
void R_FlushConsole(void) {}
void Rf_warning(const char *fmt, ...) { (void)fmt; }
