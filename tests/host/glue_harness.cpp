// Test infrastructure: compiles ggdmc_b200/r/ggdmc_b200_glue.cpp against the Rcpp stand-in (mock_rcpp/Rcpp.h) and exposes
// (1) a tiny C API to assemble R-like objects from Python and (2) the glue's three entry points and its flattening rules.
#include "../../ggdmc_b200/r/ggdmc_b200_glue.cpp"

using Rcpp::Node;
using Rcpp::NodeP;
using Rcpp::RObject;

namespace {
RObject *box(const RObject &o) { return new RObject(o); } // handles live for the length of a test process
thread_local std::string g_err;
} // namespace

extern "C" {
void *gh_real(const double *v, long n) { Rcpp::NumericVector x(v, v + n); return box(x); }
void *gh_int(const int *v, long n) { Rcpp::IntegerVector x(v, v + n); return box(x); }
void *gh_lgl(const int *v, long n) { Rcpp::LogicalVector x(v, v + n); return box(x); }
void *gh_str(const char **v, long n)
{
    std::vector<std::string> s;
    for (long i = 0; i < n; ++i) s.emplace_back(v[i]);
    return box(RObject(s));
}
void *gh_list(long n) { return box(Rcpp::List((R_xlen_t)n)); }
void *gh_s4(const char *klass) { return box(Rcpp::S4(std::string(klass))); }
void gh_set_attr(void *o, const char *name, void *v) { static_cast<RObject *>(o)->p->attrs[name] = static_cast<RObject *>(v)->p; }
void gh_list_set(void *o, long i, void *v) { static_cast<RObject *>(o)->p->list.at((size_t)i) = static_cast<RObject *>(v)->p; }
void *gh_get_attr(void *o, const char *name)
{
    auto &a = static_cast<RObject *>(o)->p->attrs;
    auto it = a.find(name);
    return it == a.end() ? nullptr : box(RObject(it->second));
}
void *gh_list_get(void *o, long i) { return box(RObject(static_cast<RObject *>(o)->p->list.at((size_t)i))); }
long gh_length(void *o) { return static_cast<RObject *>(o)->p->length(); }
int gh_kind(void *o) { return (int)static_cast<RObject *>(o)->p->kind; }
const double *gh_real_ptr(void *o) { return static_cast<RObject *>(o)->p->real.data(); }
const int *gh_int_ptr(void *o) { return static_cast<RObject *>(o)->p->ints.data(); }
const char *gh_str_at(void *o, long i) { return static_cast<RObject *>(o)->p->str.at((size_t)i).c_str(); }
const char *gh_last_error() { return g_err.c_str(); }
void gh_set_option(const char *name, const char *value) // options(name = value); value == NULL removes it
{
    if (value) Rcpp::mock_options()[name] = value;
    else Rcpp::mock_options().erase(name);
}
int gh_schedule_option() // what FlatConfig would put into ggdmc_config_t::schedule; -1 + gh_last_error() on error
{
    try { return schedule_option(); } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// the glue's entry points (R: .Call('_ggdmc_run_subject' | '_ggdmc_run_hyper' | '_ggdmc_run', ...)); null + gh_last_error() on error
void *gh_run_subject(void *config, void *dmi, void *samples)
{
    try {
        return box(run_subject(Rcpp::S4(*static_cast<RObject *>(config)), Rcpp::S4(*static_cast<RObject *>(dmi)),
                               Rcpp::S4(*static_cast<RObject *>(samples))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void *gh_run_hyper(void *config, void *dmi, void *samples)
{
    try {
        return box(run_hyper(Rcpp::S4(*static_cast<RObject *>(config)), Rcpp::S4(*static_cast<RObject *>(dmi)),
                             Rcpp::S4(*static_cast<RObject *>(samples))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void *gh_run(void *config, void *dmis, void *samples)
{
    try {
        return box(run(Rcpp::S4(*static_cast<RObject *>(config)), Rcpp::List(*static_cast<RObject *>(dmis)),
                       Rcpp::List(*static_cast<RObject *>(samples))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void *gh_run_subject_batch(void *configs, void *dmi, void *samples)
{
    try {
        return box(run_subject_batch(Rcpp::List(*static_cast<RObject *>(configs)), Rcpp::S4(*static_cast<RObject *>(dmi)),
                                     Rcpp::List(*static_cast<RObject *>(samples))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void *gh_run_batch(void *configs, void *dmis, void *samples)
{
    try {
        return box(run_batch(Rcpp::List(*static_cast<RObject *>(configs)), Rcpp::List(*static_cast<RObject *>(dmis)),
                             Rcpp::List(*static_cast<RObject *>(samples))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void *gh_sumloglike_init_batch(void *dmis, void *theta)
{
    try {
        return box(sumloglike_init_batch(Rcpp::List(*static_cast<RObject *>(dmis)), Rcpp::NumericVector(*static_cast<RObject *>(theta))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void *gh_sumlogprior_batch(void *prior, void *x, void *p0, void *p1)
{
    try {
        return box(sumlogprior_batch(Rcpp::List(*static_cast<RObject *>(prior)), Rcpp::NumericMatrix(*static_cast<RObject *>(x)),
                                     Rcpp::NumericVector(*static_cast<RObject *>(p0)), Rcpp::NumericVector(*static_cast<RObject *>(p1))));
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

// the glue's flattening rules on their own (no GPU): param_src [n_cell][6][n_acc], trials of one dmi, the start slice
int gh_flatten_model(void *dmi, int *param_src, long cap, int *dims /* n_acc, n_cell, npar, n_const */)
{
    try {
        FlatModel m = flatten_model(Rcpp::S4(*static_cast<RObject *>(dmi)));
        if ((long)m.param_src.size() > cap) return -2;
        for (size_t i = 0; i < m.param_src.size(); ++i) param_src[i] = m.param_src[i];
        dims[0] = m.c.n_acc; dims[1] = m.c.n_cell; dims[2] = m.c.npar; dims[3] = m.c.n_const + 1000 * m.c.type; // model type in the thousands
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
long gh_flatten_trials(void *dmi, double *rt, unsigned short *cell, long cap)
{
    try {
        FlatModel m = flatten_model(Rcpp::S4(*static_cast<RObject *>(dmi)));
        FlatTrials t;
        t.add(Rcpp::S4(*static_cast<RObject *>(dmi)), m.cell_names);
        t.finish();
        if ((long)t.rt.size() > cap) return -2;
        for (size_t i = 0; i < t.rt.size(); ++i) { rt[i] = t.rt[i]; cell[i] = t.cell[i]; }
        return (long)t.rt.size();
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
long gh_start_slice(void *samples, double *theta, double *lp, double *ll, long cap)
{
    try {
        StartState s(Rcpp::S4(*static_cast<RObject *>(samples)));
        if ((long)s.theta.size() > cap) return -2;
        for (size_t i = 0; i < s.theta.size(); ++i) theta[i] = s.theta[i];
        for (size_t i = 0; i < s.lp.size(); ++i) { lp[i] = s.lp[i]; ll[i] = s.ll[i]; }
        return (long)s.lp.size();
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
}
