// TEST INFRASTRUCTURE ONLY.
// C wrappers around the REFERENCE'S OWN OBJECT CODE (/root/reference/src/de.o, built by the
// package author with g++ 13.3.0).  The class declarations below only have to mangle to the
// symbols de.o defines and to reserve enough storage (sizes from de.o's DWARF);
// every function body executed comes from de.o.  Built into oracle/_ref/libggdmc_ref.so by
// oracle/Makefile; no reference SOURCE is copied.
#include <cstring>
#include <string>
#include <vector>

namespace lba {
class lba_class {
  public:
    alignas(8) char storage[312];
    lba_class(const std::vector<std::vector<double>> &P, const std::vector<bool> &is_positive_drift,
              const std::vector<double> &time_par);
    ~lba_class();
    void set_parameters(const std::vector<std::vector<double>> &P, const std::vector<bool> &is_positive_drift);
    bool validate_parameters(bool debug);
    std::vector<double> dlba(const std::vector<double> &rt);
};
} // namespace lba

struct DEInput {
    double pop_migration_prob, sub_migration_prob, gamma_precursor, rp;
    bool is_hblocked, is_pblocked;
    unsigned nparameter, nchain;
    bool pop_debug, sub_debug;
};
static_assert(sizeof(DEInput) == 48, "DEInput layout");

struct UVec { // arma::Col<unsigned int> look-alike (112 B), returned via hidden pointer
    unsigned n_rows, n_cols, n_elem, n_alloc;
    unsigned short vec_state, mem_state;
    char pad[12];
    const unsigned *mem;
    char pad2[8];
    unsigned mem_local[16];
    ~UVec();
    UVec(const UVec &) = delete;
};
static_assert(sizeof(UVec) == 112, "UVec layout");
UVec::~UVec()
{
    if (n_alloc > 0 && mem) free((void *)mem);
}

class de_class {
  public:
    alignas(16) char storage[704];
    de_class(const DEInput &);
    ~de_class();
    UVec get_chains(unsigned k, unsigned nsubchain);
    UVec get_subchains();
};

namespace tnorm {
class tnorm_class {
  public:
    double m_mean, m_sd, m_lower, m_upper;
    bool m_lower_tail, m_log_p;
    double m_denom, m_log_denom;
    void set_parameters(double mean, double sd);
    double d(double x) const;
};
static_assert(sizeof(tnorm_class) == 56, "tnorm layout");
} // namespace tnorm

extern "C" {

// n1PDF for one cell: P is 6 rows (A, b, mean_v, sd_v, st0, t0) x n_acc, row-major.
// Returns validate_parameters(); out[i] = dlba(rt)[i] if valid else 1e-10 (likelihood.h:105 rule).
int ref_lba_cell(const double *P, int n_acc, const unsigned char *posdrift, const double *rt, int n, double *out)
{
    std::vector<std::vector<double>> Pm(6, std::vector<double>(n_acc));
    for (int r = 0; r < 6; ++r)
        for (int j = 0; j < n_acc; ++j) Pm[r][j] = P[r * n_acc + j];
    std::vector<bool> pd(n_acc);
    for (int j = 0; j < n_acc; ++j) pd[j] = posdrift[j] != 0;
    // dummy construction parameters, then set_parameters like lba_likelihood does per cell
    std::vector<std::vector<double>> P0 = {{0.5, 0.5}, {1.0, 1.0}, {1.0, 1.0}, {1.0, 1.0}, {0.0, 0.0}, {0.1, 0.1}};
    std::vector<bool> pd0 = {true, true};
    std::vector<double> tp = {0.0, 10.0, 0.01};
    lba::lba_class obj(P0, pd0, tp);
    obj.set_parameters(Pm, pd);
    bool valid = obj.validate_parameters(false);
    std::vector<double> rtv(rt, rt + n);
    if (valid) {
        std::vector<double> d = obj.dlba(rtv);
        for (int i = 0; i < n; ++i) out[i] = d[i];
    } else {
        for (int i = 0; i < n; ++i) out[i] = 1e-10;
    }
    return valid ? 1 : 0;
}

static DEInput mk(unsigned nchain, unsigned npar)
{
    DEInput in;
    std::memset(&in, 0, sizeof(in));
    in.pop_migration_prob = 0.0;
    in.sub_migration_prob = 0.0;
    in.gamma_precursor = 2.38;
    in.rp = 0.001;
    in.nparameter = npar;
    in.nchain = nchain;
    return in;
}

// de_class::get_chains(k, nsub) -> out[nsub]; uniforms come from the installed stream.
void ref_get_chains(unsigned nchain, unsigned k, unsigned nsub, unsigned *out)
{
    DEInput in = mk(nchain, 4);
    de_class de(in);
    UVec v = de.get_chains(k, nsub);
    for (unsigned i = 0; i < v.n_elem; ++i) out[i] = v.mem[i];
}

// de_class::get_subchains() -> returns n, out[n] sorted ascending.
unsigned ref_get_subchains(unsigned nchain, unsigned *out)
{
    DEInput in = mk(nchain, 4);
    de_class de(in);
    UVec v = de.get_subchains();
    for (unsigned i = 0; i < v.n_elem; ++i) out[i] = v.mem[i];
    return v.n_elem;
}

// tnorm_class: set_parameters(mean, sd) then d(x)
double ref_tnorm_d(double x, double mean, double sd, double lower, double upper, int log_p)
{
    tnorm::tnorm_class t;
    std::memset(&t, 0, sizeof(t));
    t.m_lower = lower;
    t.m_upper = upper;
    t.m_lower_tail = true;
    t.m_log_p = log_p != 0;
    t.set_parameters(mean, sd);
    return t.d(x);
}
}
