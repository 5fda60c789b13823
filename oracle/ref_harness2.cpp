// TEST INFRASTRUCTURE ONLY.
// Tier-2 reference harness: drives the REFERENCE'S OWN SAMPLER CODE -- de_class::crossover and
// de_class::migration of /root/reference/src/de.o (src/de.cpp:111-199), and through them
// prior_class::sumlogprior, likelihood_class::sumloglike -> lba_likelihood ->
// design_class::set_parameter_values -> lba_class (all compiled into de.o from ggdmcHeaders) -- on
// hand-built objects.  The constructors of those classes are not in de.o (they were emitted into the
// unshipped de2R.o), so the objects are laid out here byte for byte from de.o's DWARF
// (readelf --debug-dump=info; offsets quoted below): real libstdc++ containers placement-constructed
// at the recorded offsets, Armadillo matrices as their plain 176-byte headers.  No member function of
// any reference class is defined here; every sampler / density instruction executed is de.o's.
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

// ---- opaque storage with the reference's class names (so the symbols in de.o resolve) -------------
class theta_class { public: alignas(16) unsigned char raw[2592]; };                         // theta_phi: 2592 B
namespace prior { class prior_class { public: alignas(16) unsigned char raw[592]; }; }      // 592 B
namespace likelihood {
class likelihood_class { // 416 B
  public:
    alignas(16) unsigned char raw[416];
    void ddm_likelihood(const std::vector<double> &theta, bool debug); // @hdr/likelihood.h:129-161, body in de.o (weak symbol)
};
} // namespace likelihood
namespace design { class design_class { public: alignas(16) unsigned char raw[448]; }; }    // 448 B
namespace tnorm {
struct tnorm_class { double m_mean, m_sd, m_lower, m_upper; bool m_lower_tail, m_log_p; double m_denom, m_log_denom; };
static_assert(sizeof(tnorm_class) == 56, "tnorm layout");
}

struct DEInput {
    double pop_migration_prob, sub_migration_prob, gamma_precursor, rp;
    bool is_hblocked, is_pblocked;
    unsigned nparameter, nchain;
    bool pop_debug, sub_debug;
};

typedef std::shared_ptr<theta_class> ThetaPtr;
typedef std::shared_ptr<prior::prior_class> PriorPtr;
typedef std::shared_ptr<likelihood::likelihood_class> LPtr;

class de_class {
  public:
    alignas(16) char storage[704];
    de_class(const DEInput &);
    de_class(const DEInput &, unsigned int nsubject);
    ~de_class();
    void crossover(ThetaPtr t_ptr, PriorPtr p_ptr, LPtr l_ptr, bool debug, size_t para_idx);
    void migration(ThetaPtr t_ptr, PriorPtr p_ptr, LPtr l_ptr, bool debug, size_t para_idx);
    void run_chains(ThetaPtr t_ptr, PriorPtr p_ptr, LPtr l_ptr, bool debug);
    void run_hchains(ThetaPtr phi_ptr, std::vector<LPtr> l_ptrs, std::vector<ThetaPtr> subj_thetas, PriorPtr hyper_likelihood,
                     PriorPtr h_prior, bool pop_debug, bool sub_debug);
};

namespace {

// arma::Mat<double> / arma::Col<double> header (176 B): n_rows @0, n_cols @4, n_elem @8, n_alloc @12,
// vec_state @16 (u16), mem_state @18 (u16), mem @32, mem_local[16] @48
void set_mat(unsigned char *at, unsigned n_rows, unsigned n_cols, bool is_col, double *mem)
{
    std::memset(at, 0, 176);
    unsigned n = n_rows * n_cols;
    std::memcpy(at + 0, &n_rows, 4);
    std::memcpy(at + 4, &n_cols, 4);
    std::memcpy(at + 8, &n, 4);
    std::memcpy(at + 12, &n, 4); // n_alloc: heap-owned
    unsigned short vs = is_col ? 1 : 0, ms = 0;
    std::memcpy(at + 16, &vs, 2);
    std::memcpy(at + 18, &ms, 2);
    std::memcpy(at + 32, &mem, 8);
}

double *amalloc(size_t n)
{
    void *p = nullptr;
    if (posix_memalign(&p, 64, sizeof(double) * (n ? n : 1))) std::abort();
    return static_cast<double *>(p);
}

template <class T, class... A>
T *put(unsigned char *raw, size_t off, A &&...a)
{
    return new (raw + off) T(std::forward<A>(a)...);
}

struct Built {
    theta_class *theta;
    prior::prior_class *prior;
    likelihood::likelihood_class *like;
    design::design_class *design;
    double *used_theta, *used_lp, *used_ll;
};

typedef std::vector<std::string> VS;
typedef std::vector<double> VD;
typedef std::vector<unsigned> VU;

// 0 = "lba" (6 core rows, posdrift per accumulator), 1 = "fastdm" (10 core rows, posdrift per cell = upper-boundary flag);
// set with ref2_set_model_type before building objects
int g_model_type = 0;
int core_rows() { return g_model_type == 1 ? 10 : 6; }
const char *model_str() { return g_model_type == 1 ? "fastdm" : "lba"; }

design::design_class *build_design(int n_acc, int n_cell, const int *param_src, const double *const_val)
{
    const int rows = core_rows();
    auto *d = new design::design_class();
    std::memset(d->raw, 0, sizeof(d->raw));
    unsigned char *r = d->raw;
    *reinterpret_cast<size_t *>(r + 0) = (size_t)rows;                                   // m_n_core_parameter @0
    put<VS>(r, 8, g_model_type == 1 ? VS{"a", "d", "precision", "s", "st0", "sv", "sz", "t0", "v", "z"}
                                    : VS{"A", "B", "mean_v", "sd_v", "st0", "t0"});      // m_core_parameter_names @8
    put<VS>(r, 32, VS((size_t)n_acc, "acc"));                                            // m_accumulator_names @32
    put<VS>(r, 56, VS((size_t)n_cell, "cell"));                                          // m_cell_names @56
    put<VS>(r, 80);                                                                      // m_parameter_x_condition_names @80
    *reinterpret_cast<size_t *>(r + 104) = (size_t)n_acc;                                // m_n_accumulator @104
    *reinterpret_cast<size_t *>(r + 112) = (size_t)n_cell;                               // m_n_cell @112
    *reinterpret_cast<size_t *>(r + 120) = 0;                                            // m_n_parameter_x_condition @120
    put<std::map<std::string, double>>(r, 128);                                          // m_constants @128
    put<VS>(r, 176);                                                                     // m_constant_names @176
    put<VD>(r, 200);                                                                     // m_constant_values @200
    put<std::vector<std::vector<std::vector<bool>>>>(r, 224);                            // m_model_boolean @224
    put<std::vector<VU>>(r, 248);                                                        // m_node_1_index @248
    put<VS>(r, 272);                                                                     // m_free_parameter_names @272
    *reinterpret_cast<size_t *>(r + 296) = 0;                                            // m_n_free_parameter @296
    put<std::string>(r, 304, model_str());                                               // m_model_str @304
    typedef std::vector<std::vector<std::vector<VU>>> Map4;
    put<Map4>(r, 336);                                                                   // m_tmp_param_map @336
    Map4 pm((size_t)n_acc, std::vector<std::vector<VU>>((size_t)n_cell, std::vector<VU>((size_t)rows, VU(2, 0u))));
    std::vector<std::vector<VD>> mat((size_t)n_cell, std::vector<VD>((size_t)rows, VD((size_t)n_acc, 0.0)));
    for (int c = 0; c < n_cell; ++c)
        for (int row = 0; row < rows; ++row)
            for (int j = 0; j < n_acc; ++j) {
                int s = param_src[((size_t)c * rows + row) * n_acc + j];
                if (s >= 0) {
                    pm[j][c][row][0] = (unsigned)s; // [0] = index into theta, [1] = is_free (design_light.h:323, 329-330)
                    pm[j][c][row][1] = 1u;
                } else {
                    mat[c][row][j] = const_val[-1 - s]; // constants are pre-filled in m_parameter_matrix
                }
            }
    put<Map4>(r, 360, pm);                                                               // m_param_map @360
    put<std::vector<bool>>(r, 384);                                                      // m_is_free_parameter @384
    put<std::vector<std::vector<VD>>>(r, 424, mat);                                      // m_parameter_matrix @424
    return d;
}

likelihood::likelihood_class *build_like(design::design_class *d, int n_acc, int n_cell, const unsigned char *posdrift, const double *rt,
                                         const unsigned short *cell, int n_trial)
{
    auto *l = new likelihood::likelihood_class();
    std::memset(l->raw, 0, sizeof(l->raw));
    unsigned char *r = l->raw;
    put<std::shared_ptr<design::design_class>>(r, 0, d, [](design::design_class *) {});   // m_model @0
    std::vector<VD> rts((size_t)n_cell);
    for (int i = 0; i < n_trial; ++i) rts[cell[i]].push_back(rt[i]);
    const int n_pd = g_model_type == 1 ? n_cell : n_acc; // the DDM path indexes m_is_positive_drift by CELL (@hdr/likelihood.h:142)
    std::vector<bool> empty((size_t)n_cell), pd((size_t)n_pd);
    for (int c = 0; c < n_cell; ++c) empty[c] = rts[c].empty();
    for (int j = 0; j < n_pd; ++j) pd[j] = posdrift[j] != 0;
    put<std::vector<VD>>(r, 16, rts);                                                     // m_data_rt @16
    put<std::vector<VD>>(r, 40, rts);                                                     // m_rt @40
    put<std::vector<VD>>(r, 64, std::vector<VD>((size_t)n_cell));                         // m_density @64
    put<VS>(r, 88);                                                                       // m_data_cell_names @88
    put<std::string>(r, 112, model_str());                                                // m_model_str @112
    put<std::vector<bool>>(r, 144, empty);                                                // m_is_empty_cell @144
    put<std::vector<bool>>(r, 184, pd);                                                   // m_is_positive_drift @184
    set_mat(r + 224, 0, 0, false, amalloc(1));                                            // m_theta_data @224
    put<std::shared_ptr<prior::prior_class>>(r, 400);                                     // m_p_prior @400
    return l;
}

prior::prior_class *build_prior(int npar, const double *p0, const double *p1, const double *lower, const double *upper, const int *dist,
                                const unsigned char *log_p)
{
    auto *p = new prior::prior_class();
    std::memset(p->raw, 0, sizeof(p->raw));
    unsigned char *r = p->raw;
    put<VD>(r, 0, VD(p0, p0 + npar));                                                     // m_stdp0 @0
    put<VD>(r, 24, VD(p1, p1 + npar));                                                    // m_stdp1 @24
    put<VD>(r, 48, VD(lower, lower + npar));                                              // m_lower @48
    put<VD>(r, 72, VD(upper, upper + npar));                                              // m_upper @72
    VU dc((size_t)npar), tix;
    std::vector<bool> lg((size_t)npar);
    std::vector<tnorm::tnorm_class> tn;
    for (int i = 0; i < npar; ++i) {
        dc[i] = (unsigned)dist[i];
        lg[i] = log_p[i] != 0;
        if (dist[i] == 1) { // one tnorm object per TNORM parameter, in parameter order (prior.h:349: running index)
            tnorm::tnorm_class t;
            std::memset(&t, 0, sizeof(t));
            t.m_lower = lower[i]; t.m_upper = upper[i]; t.m_lower_tail = true; t.m_log_p = log_p[i] != 0;
            tix.push_back((unsigned)i);
            tn.push_back(t);
        }
    }
    put<VU>(r, 96, dc);                                                                   // m_dist_code @96
    put<std::vector<bool>>(r, 120, lg);                                                   // m_log_p @120
    *reinterpret_cast<unsigned *>(r + 160) = (unsigned)npar;                              // m_nparameter @160
    double *m0 = amalloc(npar), *m1 = amalloc(npar);
    std::memcpy(m0, p0, sizeof(double) * npar);
    std::memcpy(m1, p1, sizeof(double) * npar);
    set_mat(r + 176, (unsigned)npar, 1, true, m0);                                        // m_p0 @176
    set_mat(r + 352, (unsigned)npar, 1, true, m1);                                        // m_p1 @352
    put<VU>(r, 528, tix);                                                                 // m_tnorm_index @528
    put<std::vector<tnorm::tnorm_class>>(r, 552, tn);                                     // m_tnorm_objects @552
    *reinterpret_cast<double *>(r + 576) = 1.0;                                           // m_beta_range @576
    *reinterpret_cast<double *>(r + 584) = 0.0;                                           // m_log_beta_range @584
    return p;
}

// arma::Cube<double> header (640 B): n_rows @0, n_cols @4, n_elem_slice @8, n_slices @12, n_elem @16, n_alloc @20,
// mem_state @24, mem @32, mat_ptrs @40 (lazily created per-slice Mat views), mat_mutex @48, mat_ptrs_local @96, mem_local @128
void set_cube(unsigned char *at, unsigned n_rows, unsigned n_cols, unsigned n_slices, double *mem)
{
    std::memset(at, 0, 640);
    unsigned nes = n_rows * n_cols, n = nes * n_slices, z = 0;
    std::memcpy(at + 0, &n_rows, 4);
    std::memcpy(at + 4, &n_cols, 4);
    std::memcpy(at + 8, &nes, 4);
    std::memcpy(at + 12, &n_slices, 4);
    std::memcpy(at + 16, &n, 4);
    std::memcpy(at + 20, &n, 4);
    std::memcpy(at + 24, &z, 4);
    std::memcpy(at + 32, &mem, 8);
    void **ptrs = static_cast<void **>(std::calloc(n_slices ? n_slices : 1, sizeof(void *)));
    std::memcpy(at + 40, &ptrs, 8);
}

struct ThetaBuf { // heap storage behind one hand-built theta_class
    theta_class *obj;
    double *used_theta, *used_lp, *used_ll, *theta, *lp, *ll;
    int npar, nchain, nmc;
};

// theta_phi layout (DWARF): m_nmc @0, m_nchain @4, m_thin @8, m_nparameter @12, m_report_length @16,
// m_max_init_attempts @20, m_is_print @24, m_start @28, m_store_i @32, m_nsample @36, m_previous_nmc @40,
// m_theta @48 (cube), m_previous_theta @688, m_lp @1328, m_ll @1504, m_used_theta @1680, m_previous_lp @1856,
// m_previous_ll @2032, m_used_lp @2208, m_used_ll @2384, m_pnames @2560
ThetaBuf build_theta_full(int npar, int nchain, int nmc, int thin, const double *theta0, const double *lp0, const double *ll0)
{
    ThetaBuf b;
    b.npar = npar; b.nchain = nchain; b.nmc = nmc;
    b.obj = new theta_class();
    std::memset(b.obj->raw, 0, sizeof(b.obj->raw));
    unsigned char *r = b.obj->raw;
    auto setu = [&](size_t off, unsigned v) { std::memcpy(r + off, &v, 4); };
    setu(0, (unsigned)nmc); setu(4, (unsigned)nchain); setu(8, (unsigned)thin); setu(12, (unsigned)npar);
    setu(16, 1000000u); setu(20, 1000u);
    r[24] = 0; // m_is_print
    setu(28, 1u); setu(32, 0u); setu(36, (unsigned)(1 + (nmc - 1) * thin)); setu(40, 0u);
    const size_t blk = (size_t)npar * nchain;
    b.theta = amalloc(blk * nmc); b.lp = amalloc((size_t)nchain * nmc); b.ll = amalloc((size_t)nchain * nmc);
    b.used_theta = amalloc(blk); b.used_lp = amalloc(nchain); b.used_ll = amalloc(nchain);
    for (size_t i = 0; i < blk * nmc; ++i) b.theta[i] = 0.0;
    for (size_t i = 0; i < (size_t)nchain * nmc; ++i) b.lp[i] = b.ll[i] = 0.0;
    std::memcpy(b.used_theta, theta0, sizeof(double) * blk);
    std::memcpy(b.used_lp, lp0, sizeof(double) * nchain);
    std::memcpy(b.used_ll, ll0, sizeof(double) * nchain);
    std::memcpy(b.theta, theta0, sizeof(double) * blk);
    std::memcpy(b.lp, lp0, sizeof(double) * nchain);
    std::memcpy(b.ll, ll0, sizeof(double) * nchain);
    set_cube(r + 48, (unsigned)npar, (unsigned)nchain, (unsigned)nmc, b.theta);
    set_cube(r + 688, 0, 0, 0, amalloc(1));
    set_mat(r + 1328, (unsigned)nchain, (unsigned)nmc, false, b.lp);
    set_mat(r + 1504, (unsigned)nchain, (unsigned)nmc, false, b.ll);
    set_mat(r + 1680, (unsigned)npar, (unsigned)nchain, false, b.used_theta);
    set_mat(r + 1856, 0, 0, false, amalloc(1));
    set_mat(r + 2032, 0, 0, false, amalloc(1));
    set_mat(r + 2208, (unsigned)nchain, 1, true, b.used_lp);
    set_mat(r + 2384, (unsigned)nchain, 1, true, b.used_ll);
    put<VS>(r, 2560, VS((size_t)npar, "p"));
    return b;
}

theta_class *build_theta(int npar, int nchain, double *used_theta, double *used_lp, double *used_ll)
{
    ThetaBuf b = build_theta_full(npar, nchain, 2, 1, used_theta, used_lp, used_ll);
    // point the "used" matrices at the caller's buffers so results can be read back directly
    set_mat(b.obj->raw + 1680, (unsigned)npar, (unsigned)nchain, false, used_theta);
    set_mat(b.obj->raw + 2208, (unsigned)nchain, 1, true, used_lp);
    set_mat(b.obj->raw + 2384, (unsigned)nchain, 1, true, used_ll);
    return b.obj;
}

void read_back(const ThetaBuf &b, double *theta, double *lp, double *ll, double *out_theta, double *out_lp, double *out_ll)
{
    const size_t blk = (size_t)b.npar * b.nchain;
    if (theta) std::memcpy(theta, b.used_theta, sizeof(double) * blk);
    if (lp) std::memcpy(lp, b.used_lp, sizeof(double) * b.nchain);
    if (ll) std::memcpy(ll, b.used_ll, sizeof(double) * b.nchain);
    if (out_theta) std::memcpy(out_theta, b.theta, sizeof(double) * blk * b.nmc); // npar x nchain x nmc col-major == [nmc][nchain][npar]
    if (out_lp) std::memcpy(out_lp, b.lp, sizeof(double) * (size_t)b.nchain * b.nmc); // nchain x nmc col-major == [nmc][nchain]
    if (out_ll) std::memcpy(out_ll, b.ll, sizeof(double) * (size_t)b.nchain * b.nmc);
}

} // namespace

extern "C" {

void ref2_set_model_type(int type) { g_model_type = type == 1 ? 1 : 0; }

// likelihood_class::ddm_likelihood of de.o (@hdr/likelihood.h:129-161) on a hand-built "fastdm" likelihood object:
// design_class::set_parameter_values, ddm_class::set_parameters / validate_parameters / dddm all run from de.o.
// Trials must be grouped by ascending cell; out[i] = m_density[cell][k] of trial i (1e-10 for invalid cells).
void ref2_ddm_density(int n_acc, int n_cell, const int *param_src, const double *const_val, const unsigned char *is_upper,
                      const double *rt, const unsigned short *cell, int n_trial, const double *theta, int npar, double *out)
{
    const int saved = g_model_type;
    g_model_type = 1;
    design::design_class *d = build_design(n_acc, n_cell, param_src, const_val);
    likelihood::likelihood_class *l = build_like(d, n_acc, n_cell, is_upper, rt, cell, n_trial);
    g_model_type = saved;
    l->ddm_likelihood(std::vector<double>(theta, theta + npar), false);
    const std::vector<VD> &dens = *reinterpret_cast<std::vector<VD> *>(l->raw + 64); // m_density @64
    int i = 0;
    for (int c = 0; c < n_cell; ++c)
        for (double x : dens[c]) out[i++] = x;
}

// One sweep of the reference's 1-level sampler on chains theta[nchain][npar] (updated in place):
// kind 0 = de_class::crossover (src/de.cpp:111-155), kind 1 = de_class::migration (:157-199);
// para_idx < 0 = all parameters.  Uniforms come from the stream installed with ref_set_uniform_stream.
// The very first likelihood call of the process additionally consumes 2 uniforms (static lba_obj).
void ref2_sweep_subject(int kind, int para_idx, int nchain, int nparameter, double gamma_precursor, double rp, int n_acc, int n_cell,
                        const int *param_src, const double *const_val, const unsigned char *posdrift, const double *rt,
                        const unsigned short *cell, int n_trial, int npar, const double *p0, const double *p1, const double *lower,
                        const double *upper, const int *dist, const unsigned char *log_p, double *theta, double *lp, double *ll)
{
    design::design_class *d = build_design(n_acc, n_cell, param_src, const_val);
    likelihood::likelihood_class *l = build_like(d, n_acc, n_cell, posdrift, rt, cell, n_trial);
    prior::prior_class *p = build_prior(npar, p0, p1, lower, upper, dist, log_p);
    double *ut = amalloc((size_t)npar * nchain), *ulp = amalloc(nchain), *ull = amalloc(nchain);
    std::memcpy(ut, theta, sizeof(double) * (size_t)npar * nchain);
    std::memcpy(ulp, lp, sizeof(double) * nchain);
    std::memcpy(ull, ll, sizeof(double) * nchain);
    theta_class *t = build_theta(npar, nchain, ut, ulp, ull);
    DEInput in;
    std::memset(&in, 0, sizeof(in));
    in.gamma_precursor = gamma_precursor;
    in.rp = rp;
    in.nparameter = (unsigned)nparameter;
    in.nchain = (unsigned)nchain;
    {
        de_class de(in);
        ThetaPtr tp(t, [](theta_class *) {});
        PriorPtr pp(p, [](prior::prior_class *) {});
        LPtr lptr(l, [](likelihood::likelihood_class *) {});
        const size_t pi = para_idx < 0 ? (size_t)-1 : (size_t)para_idx; // SIZE_T_MAX = all parameters (src/de.h:6)
        if (kind == 0) de.crossover(tp, pp, lptr, false, pi);
        else de.migration(tp, pp, lptr, false, pi);
    }
    std::memcpy(theta, ut, sizeof(double) * (size_t)npar * nchain);
    std::memcpy(lp, ulp, sizeof(double) * nchain);
    std::memcpy(ll, ull, sizeof(double) * nchain);
    // the hand-built objects are leaked on purpose (tests only; their destructors are not in de.o)
}

// run_subject's sampler: de_class::run_chains (src/de.cpp:201-242) on one subject; out_* are
// [nmc][nchain][npar] / [nmc][nchain] like posterior@theta etc.
void ref2_run_chains(int nchain, int nparameter, double sub_migration_prob, double gamma_precursor, double rp, int is_pblocked, int nmc,
                     int thin, int n_acc, int n_cell, const int *param_src, const double *const_val, const unsigned char *posdrift,
                     const double *rt, const unsigned short *cell, int n_trial, int npar, const double *p0, const double *p1,
                     const double *lower, const double *upper, const int *dist, const unsigned char *log_p, const double *theta0,
                     const double *lp0, const double *ll0, double *out_theta, double *out_lp, double *out_ll)
{
    design::design_class *d = build_design(n_acc, n_cell, param_src, const_val);
    likelihood::likelihood_class *l = build_like(d, n_acc, n_cell, posdrift, rt, cell, n_trial);
    prior::prior_class *p = build_prior(npar, p0, p1, lower, upper, dist, log_p);
    ThetaBuf tb = build_theta_full(npar, nchain, nmc, thin, theta0, lp0, ll0);
    DEInput in;
    std::memset(&in, 0, sizeof(in));
    in.sub_migration_prob = sub_migration_prob; in.gamma_precursor = gamma_precursor; in.rp = rp;
    in.is_pblocked = is_pblocked != 0; in.nparameter = (unsigned)nparameter; in.nchain = (unsigned)nchain;
    {
        de_class de(in);
        de.run_chains(ThetaPtr(tb.obj, [](theta_class *) {}), PriorPtr(p, [](prior::prior_class *) {}),
                      LPtr(l, [](likelihood::likelihood_class *) {}), false);
    }
    read_back(tb, nullptr, nullptr, nullptr, out_theta, out_lp, out_ll);
}

// run()'s sampler: de_class::run_hchains (src/de.cpp:272-383).  Subject s: trials [off[s], off[s+1]) of rt / cell,
// start state subj_theta0 + s * nchain * npar etc.; phi vectors have 2 * npar entries.
void ref2_run_hchains(int nchain, int nparameter, double pop_migration_prob, double sub_migration_prob, double gamma_precursor, double rp,
                      int is_hblocked, int is_pblocked, int nmc, int thin, int n_subject, int n_acc, int n_cell, const int *param_src,
                      const double *const_val, const unsigned char *posdrift, const long long *off, const double *rt,
                      const unsigned short *cell, int npar, const double *pp0, const double *pp1, const double *plower,
                      const double *pupper, const int *pdist, const unsigned char *plog_p, const double *hp0, const double *hp1,
                      const double *hlower, const double *hupper, const int *hdist, const unsigned char *hlog_p, const double *phi_theta0,
                      const double *phi_lp0, const double *phi_ll0, const double *subj_theta0, const double *subj_lp0,
                      const double *subj_ll0, double *phi_out_theta, double *phi_out_lp, double *phi_out_ll, double *subj_out_theta,
                      double *subj_out_lp, double *subj_out_ll)
{
    design::design_class *d = build_design(n_acc, n_cell, param_src, const_val);
    prior::prior_class *hyper_like = build_prior(npar, pp0, pp1, plower, pupper, pdist, plog_p);
    prior::prior_class *h_prior = build_prior(2 * npar, hp0, hp1, hlower, hupper, hdist, hlog_p);
    std::vector<LPtr> l_ptrs;
    std::vector<ThetaPtr> t_ptrs;
    std::vector<ThetaBuf> bufs;
    for (int s = 0; s < n_subject; ++s) {
        likelihood::likelihood_class *l = build_like(d, n_acc, n_cell, posdrift, rt + off[s], cell + off[s], (int)(off[s + 1] - off[s]));
        l_ptrs.emplace_back(l, [](likelihood::likelihood_class *) {});
        ThetaBuf tb = build_theta_full(npar, nchain, nmc, thin, subj_theta0 + (size_t)s * nchain * npar, subj_lp0 + (size_t)s * nchain,
                                       subj_ll0 + (size_t)s * nchain);
        bufs.push_back(tb);
        t_ptrs.emplace_back(tb.obj, [](theta_class *) {});
    }
    ThetaBuf pb = build_theta_full(2 * npar, nchain, nmc, thin, phi_theta0, phi_lp0, phi_ll0);
    DEInput in;
    std::memset(&in, 0, sizeof(in));
    in.pop_migration_prob = pop_migration_prob; in.sub_migration_prob = sub_migration_prob; in.gamma_precursor = gamma_precursor;
    in.rp = rp; in.is_hblocked = is_hblocked != 0; in.is_pblocked = is_pblocked != 0;
    in.nparameter = (unsigned)nparameter; in.nchain = (unsigned)nchain;
    {
        de_class de(in, (unsigned)n_subject);
        de.run_hchains(ThetaPtr(pb.obj, [](theta_class *) {}), l_ptrs, t_ptrs, PriorPtr(hyper_like, [](prior::prior_class *) {}),
                       PriorPtr(h_prior, [](prior::prior_class *) {}), false, false);
    }
    read_back(pb, nullptr, nullptr, nullptr, phi_out_theta, phi_out_lp, phi_out_ll);
    const size_t blk = (size_t)nmc * nchain * npar, blk1 = (size_t)nmc * nchain;
    for (int s = 0; s < n_subject; ++s)
        read_back(bufs[s], nullptr, nullptr, nullptr, subj_out_theta + s * blk, subj_out_lp + s * blk1, subj_out_ll + s * blk1);
}

// The first likelihood call of a process constructs `static lba_class lba_obj` (@hdr/likelihood.h:77) and
// consumes 2 uniforms; tests call this once so that later calls have the steady-state draw order.
void ref_set_uniform_stream(const double *u, long n);
void ref2_prime(void)
{
    static bool done = false;
    if (done) return;
    done = true;
    static double u[64];
    for (double &x : u) x = 0.5;
    ref_set_uniform_stream(u, 64);
    const int ps[12] = {-1, -1, -2, -2, -3, -3, -4, -4, -5, -5, -6, -6};
    const double cv[6] = {0.5, 0.5, 1.0, 1.0, 0.0, 0.2};
    const unsigned char pd[2] = {1, 1}, lg[1] = {1};
    const double rt[1] = {0.6}, p0[1] = {0.0}, p1[1] = {1.0}, lo[1] = {0.0}, up[1] = {1.0};
    const unsigned short cl[1] = {0};
    const int di[1] = {6};
    double th[3] = {0.5, 0.4, 0.3}, lp[3] = {0, 0, 0}, ll[3] = {0, 0, 0};
    ref2_sweep_subject(0, -1, 3, 1, 2.38, 0.001, 2, 1, ps, cv, pd, rt, cl, 1, 1, p0, p1, lo, up, di, lg, th, lp, ll);
    ref_set_uniform_stream(nullptr, 0);
}
}
