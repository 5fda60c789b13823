// ggdmc_b200 -- R glue: the three .Call routines of ggdmc, re-implemented on top of the C ABI.
//
// Drop this file into ggdmc's src/ in place of de.cpp / de2R.cpp (keep RcppExports.cpp and
// type_casting.h's new_posterior), add `PKG_LIBS += -L<repo>/ggdmc_b200 -lggdmc_b200` to
// src/Makevars and -I<repo>/include to PKG_CPPFLAGS.  It exports the same three functions
//     run_subject(config_r, dmi, samples), run_hyper(config_r, dmi, samples), run(config_r, dmis, samples)
// (src/de2R.cpp:8-171), so R/RcppExports.R, R/sampling.R and every user script stay unchanged.
//
// The build image has no R, Rcpp or Armadillo, so this file is compiled there against a small stand-in for the Rcpp
// calls it makes (tests/host/mock_rcpp/Rcpp.h) and driven with R-like objects by tests/glue_mock.py: its flattening
// rules are checked against the reference's fixtures (tests/test_glue_cpu.py) and its three entry points return the
// same posterior objects, bit for bit, as the Python mirror of the interface (tests/test_gpu_api.py).  Against real
// Rcpp it needs nothing else: every call used here has Rcpp's spelling and semantics.
//
// [[Rcpp::depends(Rcpp)]]
#include <Rcpp.h>
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "ggdmc_b200.h"

namespace {

// core-parameter rows in the order design_class::set_parameter_values fills them (@hdr/design_light.h:314-344):
// alphabetical; ddm_class::set_parameters reads the DDM's by index (@hdr/ddm.h:194-214)
const char *kCoreLba[GGDMC_LBA_ROWS] = {"A", "B", "mean_v", "sd_v", "st0", "t0"};
const char *kCoreDdm[GGDMC_DDM_ROWS] = {"a", "d", "precision", "s", "st0", "sv", "sz", "t0", "v", "z"};

struct FlatModel {
    std::vector<int32_t> param_src;
    std::vector<double> const_val;
    std::vector<uint8_t> posdrift;
    std::vector<std::string> cell_names, pnames;
    ggdmc_model_t c{};
};

// model_boolean + node_1_index + constants + pnames -> param_src
FlatModel flatten_model(const Rcpp::S4 &dmi)
{
    Rcpp::S4 model = dmi.slot("model");
    const std::string type = Rcpp::as<std::string>(model.slot("type"));
    if (type != "lba" && type != "fastdm") Rcpp::stop("Undefined model type"); // @hdr/likelihood.h:312
    const bool ddm = type == "fastdm";
    const int rows = ddm ? GGDMC_DDM_ROWS : GGDMC_LBA_ROWS;
    const char *const *kCore = ddm ? kCoreDdm : kCoreLba;
    FlatModel m;
    std::vector<std::string> pxc = Rcpp::as<std::vector<std::string>>(model.slot("parameter_x_condition_names"));
    m.pnames = Rcpp::as<std::vector<std::string>>(model.slot("pnames"));
    m.cell_names = Rcpp::as<std::vector<std::string>>(model.slot("cell_names"));
    Rcpp::NumericVector constants = model.slot("constants");
    std::vector<std::string> cnames = Rcpp::as<std::vector<std::string>>(constants.names());
    m.const_val.assign(constants.begin(), constants.end());
    Rcpp::LogicalVector mb = model.slot("model_boolean");
    Rcpp::IntegerVector dim = mb.attr("dim");
    const int n_cell = dim[0], n_pxc = dim[1], n_acc = dim[2];
    Rcpp::IntegerMatrix n1 = dmi.slot("node_1_index");
    Rcpp::LogicalVector pd = dmi.slot("is_positive_drift");
    m.param_src.assign((size_t)n_cell * rows * n_acc, 0);
    for (int c = 0; c < n_cell; ++c)
        for (int j = 0; j < n_acc; ++j) {
            const int acc = n1(c, j);
            for (int r = 0; r < rows; ++r) {
                int found = -1, n_found = 0;
                for (int k = 0; k < n_pxc; ++k) {
                    const std::string core = pxc[k].substr(0, pxc[k].find('.'));
                    if (core == kCore[r] && mb[c + n_cell * (k + (size_t)n_pxc * acc)]) { found = k; ++n_found; }
                }
                // exactly one parameter_x_condition column feeds a (cell, accumulator, core parameter)
                if (n_found != 1)
                    Rcpp::stop(n_found ? "model_boolean has more than one source for a core parameter" : "model_boolean has no source for a core parameter");
                const int kNotFound = INT32_MIN;
                int src = kNotFound; // >= 0: free parameter, < 0: constant -1 - index
                for (size_t q = 0; q < m.pnames.size(); ++q)
                    if (m.pnames[q] == pxc[found]) src = (int)q;
                if (src == kNotFound)
                    for (size_t q = 0; q < cnames.size(); ++q)
                        if (cnames[q] == pxc[found]) src = -1 - (int)q;
                if (src == kNotFound) Rcpp::stop("parameter '" + pxc[found] + "' is neither a free parameter (pnames) nor a constant");
                m.param_src[((size_t)c * rows + r) * n_acc + j] = src;
            }
        }
    // "lba": one flag per accumulator; "fastdm": one per cell, TRUE = upper-boundary response (@hdr/likelihood.h:142)
    const int n_pd = ddm ? n_cell : n_acc;
    if (pd.size() != n_pd) Rcpp::stop("is_positive_drift has the wrong length for this model type");
    for (int j = 0; j < n_pd; ++j) m.posdrift.push_back(pd[j] ? 1 : 0);
    m.c.type = ddm ? GGDMC_MODEL_DDM : GGDMC_MODEL_LBA;
    m.c.n_acc = n_acc; m.c.n_cell = n_cell; m.c.npar = (int)m.pnames.size(); m.c.n_const = (int)m.const_val.size();
    m.c.param_src = m.param_src.data(); m.c.const_val = m.const_val.data(); m.c.posdrift = m.posdrift.data();
    return m;
}

// The ABI takes ONE model for all subjects of a call (true of every script of the reference, which builds every dmi from
// the same `model` object; src/de2R.cpp:129 would allow one likelihood object per dmi): refuse anything else loudly.
void require_same_model(const FlatModel &first, const Rcpp::S4 &dmi, int subject)
{
    const FlatModel other = flatten_model(dmi);
    if (other.c.type != first.c.type || other.param_src != first.param_src || other.const_val != first.const_val ||
        other.posdrift != first.posdrift || other.pnames != first.pnames || other.cell_names != first.cell_names)
        Rcpp::stop("all subjects of one call must share one model (subject " + std::to_string(subject + 1) + " differs from subject 1)");
}

struct FlatTrials {
    std::vector<int64_t> offset{0};
    std::vector<double> rt;
    std::vector<uint16_t> cell;
    ggdmc_trials_t c{};
    void add(const Rcpp::S4 &dmi, const std::vector<std::string> &cell_names)
    {
        Rcpp::List data = dmi.slot("data"); // named list cell_name -> RT vector, empty cells omitted
        std::vector<std::string> names = Rcpp::as<std::vector<std::string>>(data.names());
        for (size_t i = 0; i < names.size(); ++i) {
            size_t c = 0;
            while (c < cell_names.size() && cell_names[c] != names[i]) ++c;
            Rcpp::NumericVector v = data[i];
            for (double x : v) { rt.push_back(x); cell.push_back((uint16_t)c); }
        }
        offset.push_back((int64_t)rt.size());
    }
    void finish()
    {
        c.n_subject = (int32_t)offset.size() - 1;
        c.subject_offset = offset.data(); c.rt = rt.data(); c.cell = cell.data();
    }
};

struct FlatPrior {
    std::vector<double> p0, p1, lower, upper;
    std::vector<int32_t> dist;
    std::vector<uint8_t> log_p;
    ggdmc_prior_t c{};
    explicit FlatPrior(const Rcpp::List &pl)
    {
        for (R_xlen_t i = 0; i < pl.size(); ++i) {
            Rcpp::List e = pl[i];
            p0.push_back(Rcpp::as<double>(e["p0"])); p1.push_back(Rcpp::as<double>(e["p1"]));
            lower.push_back(Rcpp::as<double>(e["lower"])); upper.push_back(Rcpp::as<double>(e["upper"]));
            dist.push_back((int32_t)Rcpp::as<double>(e["dist_id"])); log_p.push_back(Rcpp::as<bool>(e["log_p"]) ? 1 : 0);
        }
        c.npar = (int32_t)p0.size(); c.p0 = p0.data(); c.p1 = p1.data(); c.lower = lower.data(); c.upper = upper.data();
        c.dist = dist.data(); c.log_p = log_p.data();
    }
};

// options(ggdmc.schedule = "reference" | "parallel" | "simultaneous"); unset = the two-half parallel schedule
int schedule_option()
{
    Rcpp::Environment base = Rcpp::Environment::base_env();
    Rcpp::Function get_option = base["getOption"];
    const std::string v = Rcpp::as<std::string>(get_option("ggdmc.schedule", "parallel"));
    if (v == "reference") return GGDMC_SCHEDULE_REFERENCE;
    if (v == "simultaneous") return GGDMC_SCHEDULE_SIMULTANEOUS;
    if (v != "parallel") Rcpp::stop("options(ggdmc.schedule) must be \"parallel\", \"reference\" or \"simultaneous\"");
    return GGDMC_SCHEDULE_PARALLEL;
}

struct FlatConfig {
    std::vector<uint64_t> seeds; // one Philox key per replicate
    ggdmc_config_t c{};
    // the replicates of one StartSampling call share everything but the seed (R/sampling.R:13-29 builds them that way)
    explicit FlatConfig(const Rcpp::List &configs) : FlatConfig(Rcpp::as<Rcpp::S4>(configs[0]))
    {
        seeds.clear();
        for (R_xlen_t r = 0; r < configs.size(); ++r) {
            Rcpp::S4 cr = Rcpp::as<Rcpp::S4>(configs[r]);
            seeds.push_back((uint64_t)Rcpp::as<double>(cr.slot("seed")));
        }
        c.n_replicate = (int32_t)seeds.size();
        c.seed = seeds.data();
    }
    explicit FlatConfig(const Rcpp::S4 &config_r)
    {
        Rcpp::S4 ti = config_r.slot("theta_input"), de = config_r.slot("de_input");
        c.nmc = ti.slot("nmc"); c.nchain = ti.slot("nchain"); c.thin = ti.slot("thin");
        c.report_length = Rcpp::as<bool>(ti.slot("is_print")) ? Rcpp::as<int>(ti.slot("report_length")) : 0;
        c.pop_migration_prob = de.slot("pop_migration_prob"); c.sub_migration_prob = de.slot("sub_migration_prob");
        c.gamma_precursor = de.slot("gamma_precursor"); c.rp = de.slot("rp");
        c.is_hblocked = Rcpp::as<bool>(de.slot("is_hblocked")); c.is_pblocked = Rcpp::as<bool>(de.slot("is_pblocked"));
        c.nparameter = de.slot("nparameter");
        c.schedule = schedule_option();
        c.n_replicate = 1; c.device = -1;
        seeds.push_back((uint64_t)Rcpp::as<double>(config_r.slot("seed"))); // config@seed -> Philox key (R/model-class.R:1514-1515)
        c.seed = seeds.data();
    }
};

// start state = last fully finite slice of the incoming `posterior` (fresh init: slice 1 only)
struct StartState {
    std::vector<double> theta, lp, ll;
    ggdmc_start_t c{};
    StartState() = default;
    StartState(const Rcpp::S4 &samples) { append(samples); }
    // one more replicate: the ABI takes [n_replicate][nchain][npar] per population
    void append(const Rcpp::S4 &samples)
    {
        Rcpp::NumericVector th = samples.slot("theta");
        Rcpp::IntegerVector d = th.attr("dim");
        const size_t npar = d[0], nchain = d[1], nmc = d[2], blk = npar * nchain;
        size_t s = nmc;
        bool found = false;
        while (!found && s-- > 0) {
            found = true;
            for (size_t i = 0; i < blk && found; ++i) found = R_finite(th[s * blk + i]);
        }
        if (!found) Rcpp::stop("samples has no slice with finite thetas");
        Rcpp::NumericMatrix lpm = samples.slot("summed_log_prior"), llm = samples.slot("log_likelihoods");
        theta.insert(theta.end(), th.begin() + s * blk, th.begin() + (s + 1) * blk); // npar x nchain col-major == [nchain][npar]
        for (size_t k = 0; k < nchain; ++k) { lp.push_back(lpm(k, s)); ll.push_back(llm(k, s)); }
        c.theta = theta.data(); c.lp = lp.data(); c.ll = ll.data();
    }
};

void progress_cb(int32_t i, void *) { Rcpp::Rcout << i << " "; } // theta_phi::print_progress, @hdr/theta.h:76-85

// replicate r of the sample arrays of one population -> one `posterior`
Rcpp::S4 make_posterior(const ggdmc_samples_t &s, const std::vector<std::string> &pnames, int thin, int r = 0)
{
    Rcpp::S4 out("posterior"); // src/type_casting.h:10-26
    const size_t blk = (size_t)s.npar * s.nchain * s.nmc, blk1 = (size_t)s.nchain * s.nmc;
    Rcpp::NumericVector th(s.theta + r * blk, s.theta + (r + 1) * blk);
    th.attr("dim") = Rcpp::IntegerVector::create(s.npar, s.nchain, s.nmc);
    Rcpp::NumericMatrix lp(s.nchain, s.nmc, s.lp + r * blk1), ll(s.nchain, s.nmc, s.ll + r * blk1);
    out.slot("theta") = th; out.slot("summed_log_prior") = lp; out.slot("log_likelihoods") = ll;
    out.slot("start") = 1; out.slot("npar") = s.npar; out.slot("pnames") = pnames;
    out.slot("nmc") = s.nmc; out.slot("thin") = thin; out.slot("nchain") = s.nchain;
    return out;
}

struct OutBuf {
    std::vector<double> theta, lp, ll;
    ggdmc_samples_t c{};
    OutBuf(int npar, int nchain, int nmc, int n_rep = 1)
        : theta((size_t)n_rep * npar * nchain * nmc), lp((size_t)n_rep * nchain * nmc), ll((size_t)n_rep * nchain * nmc)
    {
        c.npar = npar; c.nchain = nchain; c.nmc = nmc; c.theta = theta.data(); c.lp = lp.data(); c.ll = ll.data();
    }
};

} // namespace

// [[Rcpp::export]]
Rcpp::S4 run_subject(const Rcpp::S4 &config_r, const Rcpp::S4 &dmi, const Rcpp::S4 &samples)
{
    Rcpp::S4 priors = config_r.slot("prior");
    FlatModel m = flatten_model(dmi);
    FlatTrials t; t.add(dmi, m.cell_names); t.finish();
    FlatPrior pp(priors.slot("p_prior"));
    FlatConfig cfg(config_r);
    StartState st(samples);
    OutBuf out(m.c.npar, cfg.c.nchain, cfg.c.nmc);
    char err[256] = {0};
    if (ggdmc_b200_run_subject(&m.c, &t.c, &pp.c, &cfg.c, &st.c, &out.c, progress_cb, nullptr, err)) Rcpp::stop(err);
    Rcpp::Rcout << std::endl;
    return make_posterior(out.c, m.pnames, cfg.c.thin);
}

// [[Rcpp::export]]
Rcpp::S4 run_hyper(const Rcpp::S4 &config_r, const Rcpp::S4 &dmi, const Rcpp::S4 &samples)
{
    Rcpp::S4 priors = config_r.slot("prior");
    FlatPrior pp(priors.slot("p_prior")), hp(priors.slot("h_prior"));
    Rcpp::NumericMatrix data = dmi.slot("data"); // nsubject x npar
    std::vector<double> x((size_t)data.nrow() * data.ncol());
    for (int s = 0; s < data.nrow(); ++s)
        for (int p = 0; p < data.ncol(); ++p) x[(size_t)s * data.ncol() + p] = data(s, p);
    FlatConfig cfg(config_r);
    StartState st(samples);
    OutBuf out(hp.c.npar, cfg.c.nchain, cfg.c.nmc);
    char err[256] = {0};
    if (ggdmc_b200_run_hyper(&pp.c, &hp.c, x.data(), data.nrow(), &cfg.c, &st.c, &out.c, progress_cb, nullptr, err)) Rcpp::stop(err);
    Rcpp::Rcout << std::endl;
    Rcpp::S4 ti = config_r.slot("theta_input");
    return make_posterior(out.c, Rcpp::as<std::vector<std::string>>(ti.slot("pnames")), cfg.c.thin);
}

// [[Rcpp::export]]
Rcpp::List run(const Rcpp::S4 &config_r, const Rcpp::List &dmis, const Rcpp::List &samples)
{
    Rcpp::S4 priors = config_r.slot("prior");
    FlatPrior pp(priors.slot("p_prior")), hp(priors.slot("h_prior"));
    const int S = dmis.size();
    FlatModel m = flatten_model(Rcpp::as<Rcpp::S4>(dmis[0]));
    FlatTrials t;
    for (int s = 0; s < S; ++s) {
        if (s > 0) require_same_model(m, Rcpp::as<Rcpp::S4>(dmis[s]), s);
        t.add(Rcpp::as<Rcpp::S4>(dmis[s]), m.cell_names);
    }
    t.finish();
    FlatConfig cfg(config_r);
    Rcpp::List subj_r = samples["subject_theta"];
    std::vector<StartState> starts;
    std::vector<ggdmc_start_t> starts_c;
    for (int s = 0; s < S; ++s) starts.emplace_back(Rcpp::as<Rcpp::S4>(subj_r[s]));
    for (auto &s : starts) starts_c.push_back(s.c);
    StartState phi_start(Rcpp::as<Rcpp::S4>(samples["phi"]));
    // one allocation for all subjects: the engine then returns them in a single device->host copy
    const size_t blk = (size_t)m.c.npar * cfg.c.nchain * cfg.c.nmc, blk1 = (size_t)cfg.c.nchain * cfg.c.nmc;
    std::vector<double> big_t(blk * S), big_lp(blk1 * S), big_ll(blk1 * S);
    std::vector<ggdmc_samples_t> outs(S);
    for (int s = 0; s < S; ++s) {
        outs[s].npar = m.c.npar; outs[s].nchain = cfg.c.nchain; outs[s].nmc = cfg.c.nmc;
        outs[s].theta = big_t.data() + s * blk; outs[s].lp = big_lp.data() + s * blk1; outs[s].ll = big_ll.data() + s * blk1;
    }
    OutBuf phi_out(hp.c.npar, cfg.c.nchain, cfg.c.nmc);
    char err[256] = {0};
    if (ggdmc_b200_run(&m.c, &t.c, &pp.c, &hp.c, &cfg.c, &phi_start.c, starts_c.data(), &phi_out.c, outs.data(), progress_cb, nullptr, err))
        Rcpp::stop(err);
    Rcpp::Rcout << std::endl;
    Rcpp::List theta_out(S);
    for (int s = 0; s < S; ++s) theta_out[s] = make_posterior(outs[s], m.pnames, cfg.c.thin);
    Rcpp::S4 ti = config_r.slot("theta_input");
    return Rcpp::List::create(Rcpp::Named("phi") = make_posterior(phi_out.c, Rcpp::as<std::vector<std::string>>(ti.slot("pnames")), cfg.c.thin),
                              Rcpp::Named("subject_theta") = theta_out);
}

// ---- the ncore replicates of a StartSampling call as ONE call ---------------------------------------------------
// The reference forks ncore R processes (parallel_lapply, R/sampling.R:13-121), one replicate each.  Forking after
// CUDA is initialised is not supported and ncore processes would time-share one GPU anyway, so the replicates become
// a batch dimension of the engine: parallel_lapply calls these once with the list of configs.

// [[Rcpp::export]]
Rcpp::List run_subject_batch(const Rcpp::List &configs, const Rcpp::S4 &dmi, const Rcpp::List &samples)
{
    const int R = configs.size();
    if (R < 1 || samples.size() != R) Rcpp::stop("need one start object per config");
    Rcpp::S4 config0 = Rcpp::as<Rcpp::S4>(configs[0]);
    Rcpp::S4 priors = config0.slot("prior");
    FlatModel m = flatten_model(dmi);
    FlatTrials t; t.add(dmi, m.cell_names); t.finish();
    FlatPrior pp(priors.slot("p_prior"));
    FlatConfig cfg(configs);
    StartState st;
    for (int r = 0; r < R; ++r) st.append(Rcpp::as<Rcpp::S4>(samples[r]));
    OutBuf out(m.c.npar, cfg.c.nchain, cfg.c.nmc, R);
    char err[256] = {0};
    if (ggdmc_b200_run_subject(&m.c, &t.c, &pp.c, &cfg.c, &st.c, &out.c, progress_cb, nullptr, err)) Rcpp::stop(err);
    Rcpp::Rcout << std::endl;
    Rcpp::List fits(R);
    for (int r = 0; r < R; ++r) fits[r] = make_posterior(out.c, m.pnames, cfg.c.thin, r);
    return fits;
}

// [[Rcpp::export]]
Rcpp::List run_batch(const Rcpp::List &configs, const Rcpp::List &dmis, const Rcpp::List &samples)
{
    const int R = configs.size(), S = dmis.size();
    if (R < 1 || samples.size() != R) Rcpp::stop("need one start object per config");
    Rcpp::S4 config0 = Rcpp::as<Rcpp::S4>(configs[0]);
    Rcpp::S4 priors = config0.slot("prior");
    FlatPrior pp(priors.slot("p_prior")), hp(priors.slot("h_prior"));
    FlatModel m = flatten_model(Rcpp::as<Rcpp::S4>(dmis[0]));
    FlatTrials t;
    for (int s = 0; s < S; ++s) {
        if (s > 0) require_same_model(m, Rcpp::as<Rcpp::S4>(dmis[s]), s);
        t.add(Rcpp::as<Rcpp::S4>(dmis[s]), m.cell_names);
    }
    t.finish();
    FlatConfig cfg(configs);
    std::vector<StartState> starts(S);
    StartState phi_start;
    for (int r = 0; r < R; ++r) {
        Rcpp::List sr = samples[r]; // list(phi = <posterior>, subject_theta = list(<posterior> ...)) of replicate r
        phi_start.append(Rcpp::as<Rcpp::S4>(sr["phi"]));
        Rcpp::List subj_r = sr["subject_theta"];
        for (int s = 0; s < S; ++s) starts[s].append(Rcpp::as<Rcpp::S4>(subj_r[s]));
    }
    std::vector<ggdmc_start_t> starts_c;
    for (auto &s : starts) starts_c.push_back(s.c);
    const size_t blk = (size_t)R * m.c.npar * cfg.c.nchain * cfg.c.nmc, blk1 = (size_t)R * cfg.c.nchain * cfg.c.nmc;
    std::vector<double> big_t(blk * S), big_lp(blk1 * S), big_ll(blk1 * S);
    std::vector<ggdmc_samples_t> outs(S);
    for (int s = 0; s < S; ++s) {
        outs[s].npar = m.c.npar; outs[s].nchain = cfg.c.nchain; outs[s].nmc = cfg.c.nmc;
        outs[s].theta = big_t.data() + s * blk; outs[s].lp = big_lp.data() + s * blk1; outs[s].ll = big_ll.data() + s * blk1;
    }
    OutBuf phi_out(hp.c.npar, cfg.c.nchain, cfg.c.nmc, R);
    char err[256] = {0};
    if (ggdmc_b200_run(&m.c, &t.c, &pp.c, &hp.c, &cfg.c, &phi_start.c, starts_c.data(), &phi_out.c, outs.data(), progress_cb, nullptr, err))
        Rcpp::stop(err);
    Rcpp::Rcout << std::endl;
    Rcpp::S4 ti = config0.slot("theta_input");
    const std::vector<std::string> phi_names = Rcpp::as<std::vector<std::string>>(ti.slot("pnames"));
    Rcpp::List fits(R);
    for (int r = 0; r < R; ++r) {
        Rcpp::List theta_out(S);
        for (int s = 0; s < S; ++s) theta_out[s] = make_posterior(outs[s], m.pnames, cfg.c.thin, r);
        fits[r] = Rcpp::List::create(Rcpp::Named("phi") = make_posterior(phi_out.c, phi_names, cfg.c.thin, r),
                                     Rcpp::Named("subject_theta") = theta_out);
    }
    return fits;
}

// ---- scoring candidate start values in bulk (initialise_theta / initialise_phi, R/phi.R:141-332) -----------------
// The reference scores ONE candidate per R -> C++ round trip (ggdmcLikelihood::compute_subject_likelihood + .sumlog,
// ggdmcPrior::dprior) inside `repeat` loops.  These two routines score all candidates of all chains of all subjects in
// one call each, so initialise_* reduces to: draw a batch with rprior(), score it, keep the first valid per chain.

// theta: numeric array npar x n_candidate x n_subject (column-major, like R); returns n_candidate x n_subject matrix of
// sum log-likelihoods under `.sumlog`'s rule (densities <= 0 count as .Machine$double.eps, R/phi.R:3-13)
// [[Rcpp::export]]
Rcpp::NumericMatrix sumloglike_init_batch(const Rcpp::List &dmis, const Rcpp::NumericVector &theta)
{
    const int S = dmis.size();
    FlatModel m = flatten_model(Rcpp::as<Rcpp::S4>(dmis[0]));
    FlatTrials t;
    for (int s = 0; s < S; ++s) {
        if (s > 0) require_same_model(m, Rcpp::as<Rcpp::S4>(dmis[s]), s);
        t.add(Rcpp::as<Rcpp::S4>(dmis[s]), m.cell_names);
    }
    t.finish();
    Rcpp::IntegerVector d = theta.attr("dim");
    if (d.size() != 3 || d[0] != m.c.npar || d[2] != S) Rcpp::stop("theta must be npar x n_candidate x n_subject");
    const int n_cand = d[1];
    std::vector<double> out((size_t)S * n_cand);
    char err[256] = {0};
    // npar x n_candidate x n_subject column-major IS [n_subject][n_candidate][npar] row-major: no copy needed
    if (ggdmc_b200_sumloglike_init(&m.c, &t.c, &theta[0], n_cand, out.data(), err)) Rcpp::stop(err);
    return Rcpp::NumericMatrix(n_cand, S, out.data());
}

// x: npar x n matrix of parameter vectors; p0 / p1: NULL-length (use the prior's own location / scale) or npar x n
// (the phi-driven case: the hyper-likelihood of theta under candidate phi); returns n sum log-priors
// [[Rcpp::export]]
Rcpp::NumericVector sumlogprior_batch(const Rcpp::List &prior, const Rcpp::NumericMatrix &x, const Rcpp::NumericVector &p0,
                                      const Rcpp::NumericVector &p1)
{
    FlatPrior pr(prior);
    if (x.nrow() != pr.c.npar) Rcpp::stop("x must be npar x n");
    const int n = x.ncol();
    const bool ovr = p0.size() > 0;
    if (ovr && (p0.size() != (R_xlen_t)pr.c.npar * n || p1.size() != p0.size())) Rcpp::stop("p0 / p1 must be npar x n");
    Rcpp::NumericVector out((R_xlen_t)n);
    char err[256] = {0};
    if (ggdmc_b200_sumlogprior(&pr.c, &x[0], ovr ? &p0[0] : nullptr, ovr ? &p1[0] : nullptr, n, &out[0], err)) Rcpp::stop(err);
    return out;
}

