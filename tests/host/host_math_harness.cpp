// TEST INFRASTRUCTURE ONLY: compiles the engine's device math headers (gg_math.cuh, gg_lba.cuh,
// gg_rng.cuh) as plain host C++ so the arithmetic can be checked against the oracle / mpmath
// without a GPU.  The product never runs this code path; it exists to catch formula errors
// before spending GPU time.
#include "../../ggdmc_b200/csrc/gg_lba.cuh"
#include "../../ggdmc_b200/csrc/gg_rng.cuh"

extern "C" {
// P rows: A, B (NOT yet b), mean_v, sd_v, st0, t0
int hm_lba_cell(const double *P, int n_acc, const unsigned char *posdrift, const double *rt, int n, double *out)
{
    gg::CellAcc e[16];
    uint8_t cls = gg::kCellRegular;
    for (int j = 0; j < n_acc; ++j) {
        gg::cellacc_build(e[j], P[0 * n_acc + j], P[1 * n_acc + j], P[2 * n_acc + j], P[3 * n_acc + j], P[4 * n_acc + j],
                          P[5 * n_acc + j], posdrift[j] != 0, 0.0);
        cls = gg::cell_class_update(cls, P[0 * n_acc + j], P[1 * n_acc + j], P[2 * n_acc + j], P[3 * n_acc + j], P[4 * n_acc + j],
                                    P[5 * n_acc + j]);
    }
    for (int i = 0; i < n; ++i) out[i] = gg::n1pdf_any<0>(cls, rt[i], e, n_acc);
    return cls;
}
// the same with the uniforms of `t0 + st0 U` (one per accumulator)
int hm_lba_cell_u(const double *P, int n_acc, const unsigned char *posdrift, const double *u_st0, const double *rt, int n, double *out)
{
    gg::CellAcc e[16];
    uint8_t cls = gg::kCellRegular;
    for (int j = 0; j < n_acc; ++j) {
        gg::cellacc_build(e[j], P[0 * n_acc + j], P[1 * n_acc + j], P[2 * n_acc + j], P[3 * n_acc + j], P[4 * n_acc + j],
                          P[5 * n_acc + j], posdrift[j] != 0, u_st0[j]);
        cls = gg::cell_class_update(cls, P[0 * n_acc + j], P[1 * n_acc + j], P[2 * n_acc + j], P[3 * n_acc + j], P[4 * n_acc + j],
                                    P[5 * n_acc + j]);
    }
    for (int i = 0; i < n; ++i) out[i] = gg::n1pdf_any<0>(cls, rt[i], e, n_acc);
    return cls;
}
// The trial functions of the sampler's hot loop on a regular 2-accumulator cell: consecutive trials in pairs through
// n1pdf_fast2<2> (a last odd trial and trials with rt <= t0 through n1pdf_any, like the cold loop); how[i] = 2 for fast2.
int hm_lba_cell_hot2(const double *P, const unsigned char *posdrift, const double *rt, int n, double *out, int *how)
{
    gg::CellAcc e[2];
    uint8_t cls = gg::kCellRegular;
    for (int j = 0; j < 2; ++j) {
        gg::cellacc_build(e[j], P[0 * 2 + j], P[1 * 2 + j], P[2 * 2 + j], P[3 * 2 + j], P[4 * 2 + j], P[5 * 2 + j], posdrift[j] != 0, 0.0);
        cls = gg::cell_class_update(cls, P[0 * 2 + j], P[1 * 2 + j], P[2 * 2 + j], P[3 * 2 + j], P[4 * 2 + j], P[5 * 2 + j]);
    }
    for (int i = 0; i < n; i += 2) {
        const bool pair = i + 1 < n && cls == gg::kCellRegular && gg::n1pdf_fast_ok<2>(rt[i], e, 2) && gg::n1pdf_fast_ok<2>(rt[i + 1], e, 2);
        if (pair) {
            gg::n1pdf_fast2<2>(rt[i], e, rt[i + 1], e, 2, out[i], out[i + 1]);
            how[i] = how[i + 1] = 2;
        } else {
            for (int h = 0; h < 2 && i + h < n; ++h) {
                out[i + h] = gg::n1pdf_any<2>(cls, rt[i + h], e, 2);
                how[i + h] = 1;
            }
        }
    }
    return cls;
}
// one-trial fast path of the 3- and 4-accumulator hot loops
int hm_lba_cell_hot1(const double *P, int n_acc, const unsigned char *posdrift, const double *rt, int n, double *out)
{
    gg::CellAcc e[16];
    uint8_t cls = gg::kCellRegular;
    for (int j = 0; j < n_acc; ++j) {
        gg::cellacc_build(e[j], P[0 * n_acc + j], P[1 * n_acc + j], P[2 * n_acc + j], P[3 * n_acc + j], P[4 * n_acc + j],
                          P[5 * n_acc + j], posdrift[j] != 0, 0.0);
        cls = gg::cell_class_update(cls, P[0 * n_acc + j], P[1 * n_acc + j], P[2 * n_acc + j], P[3 * n_acc + j], P[4 * n_acc + j],
                                    P[5 * n_acc + j]);
    }
    for (int i = 0; i < n; ++i)
        out[i] = (cls == gg::kCellRegular && gg::n1pdf_fast_ok<0>(rt[i], e, n_acc)) ? gg::n1pdf_fast<0>(rt[i], e, n_acc)
                                                                                     : gg::n1pdf_any<0>(cls, rt[i], e, n_acc);
    return cls;
}
double hm_pnorm_std(double z) { return gg::pnorm_std(z); }
double hm_dnorm_std(double z) { return gg::dnorm_std(z); }
double hm_pnorm5(double x, double mu, double s, int lower) { return gg::pnorm5(x, mu, s, lower != 0); }
double hm_dnorm4(double x, double mu, double s, int lg) { return gg::dnorm4(x, mu, s, lg != 0); }
double hm_dprior1(int dist, double x, double p0, double p1, double lo, double up, int lg)
{
    return gg::dprior1(dist, x, p0, p1, lo, up, lg != 0);
}
void hm_philox(const unsigned *ctr, const unsigned *key, unsigned *out)
{
    gg::U4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    gg::U4 r = gg::philox4x32_10(c, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
double hm_draw_uniform(unsigned long long seed, unsigned pop, unsigned iter, unsigned sweep, unsigned chain,
                       unsigned purpose, unsigned slot)
{
    gg::DrawAddr a = {seed, pop, iter, sweep, chain};
    return gg::draw_uniform(a, purpose, slot);
}
}

#include "../../ggdmc_b200/csrc/gg_fastmath.cuh"
extern "C" {
void hm_norm_pair(const double *z, int n, double *cdf, double *pdf)
{
    for (int i = 0; i < n; ++i) {
        gg::fm::Pair p = gg::fm::norm_pair(z[i]);
        cdf[i] = p.cdf;
        pdf[i] = p.pdf;
    }
}
// the trial loop's batched twins (table-driven exponential): four arguments in lock step / two at a time
void hm_norm_pairs_hot(const double *z, int n, double *cdf4, double *pdf4, double *cdf2, double *pdf2)
{
    for (int i = 0; i + 4 <= n; i += 4) {
        double zz[4] = {z[i], z[i + 1], z[i + 2], z[i + 3]}, c[4], p[4];
        gg::fm::norm_pairs_stepmajor<4>(zz, c, p);
        for (int k = 0; k < 4; ++k) { cdf4[i + k] = c[k]; pdf4[i + k] = p[k]; }
        for (int h = 0; h < 4; h += 2) {
            double z2[2] = {zz[h], zz[h + 1]}, c2[2], p2[2];
            gg::fm::norm_pairs_finite<2>(z2, c2, p2);
            cdf2[i + h] = c2[0]; cdf2[i + h + 1] = c2[1]; pdf2[i + h] = p2[0]; pdf2[i + h + 1] = p2[1];
        }
    }
}
double hm_rcp_pos(double d) { return gg::fm::rcp_pos(d); }
void hm_norm_cdf_lowlatency(const double *z, int n, double *cdf)
{
    for (int i = 0; i < n; ++i) cdf[i] = gg::fm::norm_cdf_lowlatency(z[i]);
}
}

#include "../../ggdmc_b200/csrc/gg_ddm.cuh"
extern "C" {
// P = column 0 of the ten DDM rows (a, d, precision, s, st0, sv, sz, t0, v, z); returns validate_parameters()
int hm_ddm_ceil_sqrt(double x) { return gg::ddm_ceil_sqrt(x); }
int hm_ddm_cell(const double *P, int is_upper, const double *rt, int n, double *out)
{
    gg::DdmCell q;
    gg::ddmcell_build(q, P, is_upper != 0);
    for (int i = 0; i < n; ++i) out[i] = gg::ddm_density(q, rt[i]);
    return q.valid;
}
}
