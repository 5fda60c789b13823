/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the ggdmc DE-MCMC / LBA hot path.
 *
 * This is a plain-C restatement of the reference's algorithm (src/de.cpp and the ggdmcHeaders
 * code compiled into src/de.o; decoded arithmetic in SURVEY.md Appendix A/B).  It exists to CHECK
 * the CUDA engine (tests/, __graft_entry__.smoke()) and to be TIMED as the CPU baseline
 * (bench.py cpu_baseline / --impl reference).  Nothing in ggdmc_b200/ links, imports or calls it.
 *
 * Parity status: PINNED.  (1) against the reference's known-answer fixtures
 * tests/testthat/Group1/data/lba_data[2-6].rda (the .npz files under tests/golden/, made by
 * tests/golden/make_golden.py), (2) against the reference's own object code src/de.o linked into
 * oracle/_ref/libggdmc_ref.so (lba_class::dlba, tnorm_class, de_class::get_chains/get_subchains,
 * de_class::crossover/migration under an injected uniform stream; for the DDM ("fastdm") family:
 * likelihood_class::ddm_likelihood and de_class::run_chains on a "fastdm" likelihood object -- bit-identical).
 */
#ifndef GGDMC_ORACLE_H
#define GGDMC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* prior::DistributionType, @hdr/prior.h:186 */
enum { ORC_TNORM = 1, ORC_BETA_LU = 2, ORC_GAMMA_L = 3, ORC_LNORM_L = 4, ORC_CAUCHY = 5, ORC_UNIF = 6, ORC_NORM = 7 };

/* draw purposes for the counter-addressed uniform source */
enum { ORC_U_DECIDE = 0, ORC_U_PARTNER = 1, ORC_U_NOISE = 2, ORC_U_ST0 = 3, ORC_U_ACCEPT = 4, ORC_U_MIG_N = 5, ORC_U_MIG_KEYS = 6 };

#define ORC_POP_PHI 0xFFFFFFFFu

typedef struct {
    unsigned pop;     /* global subject index, or ORC_POP_PHI */
    unsigned iter;    /* DE-MCMC iteration (1-based like de.cpp's loop variable) */
    unsigned sweep;   /* parameter index in blocked sweeps, else 0 */
    unsigned chain;   /* chain the draw belongs to (source chain for migration steps) */
    unsigned purpose; /* ORC_U_* */
    unsigned slot;    /* running index within the purpose */
} orc_addr;

/* Uniform source.  mode 0: sequential stream in the reference's draw order;
 * mode 1: counter-addressed Philox4x32-10, u = (word + 0.5) * 2^-32. */
typedef struct {
    int mode;
    const double *u;
    long n, pos;
    unsigned long long seed;
    int burn_static_ctor; /* mode 0: consume 2 extra draws before the first likelihood call
                             (one-off `static lba_class lba_obj`, @hdr/likelihood.h:77) */
    int first_like_done;
    double *rec;          /* optional: every uniform handed out is also appended here (replay a counter-addressed run
                             as the sequential stream the reference's Rf_runif would have consumed) */
    long rec_n, rec_cap;
} orc_rng;

/* likelihood_class::resolve_string codes (@hdr/likelihood.h:279): "lba" and "fastdm" */
enum { ORC_MODEL_LBA = 0, ORC_MODEL_DDM = 1 };
#define ORC_LBA_ROWS 6  /* A, B, mean_v, sd_v, st0, t0 */
#define ORC_DDM_ROWS 10 /* a, d, precision, s, st0, sv, sz, t0, v, z (alphabetical, like every core-parameter table) */

typedef struct {
    int n_acc, n_cell, npar;
    const int *param_src;          /* [n_cell][rows][n_acc]; >=0: theta index, <0: const_val[-1-k]; rows = 6 (LBA) or 10 (DDM) */
    const double *const_val;
    const unsigned char *posdrift; /* LBA: [n_acc] is_positive_drift.  DDM: [n_cell], non-zero = the cell's response is
                                      the UPPER boundary (the same dmi@is_positive_drift slot, @hdr/likelihood.h:142) */
    int type;                      /* ORC_MODEL_* */
} orc_model;

typedef struct {
    int n_trial;
    const double *rt;             /* [n_trial], any order */
    const unsigned short *cell;   /* [n_trial] model cell index of every trial */
} orc_data;

typedef struct {
    int npar;
    const double *p0, *p1, *lower, *upper;
    const int *dist;
    const unsigned char *log_p;
} orc_prior;

typedef struct {
    double pop_migration_prob, sub_migration_prob, gamma_precursor, rp;
    int is_hblocked, is_pblocked;
    int nparameter; /* de_input@nparameter: npar (1 level) or 2*npar (hierarchical) */
    int nchain;
    int jacobi;     /* schedule.  0: reference order (chains swept in place, one after another);
                       1: two half-sweeps -- even chains move together with partners drawn from the odd
                          chains, then the odd chains with partners from the even ones (migration: all
                          selected chains at once);
                       2: all chains of a sweep proposed at once from the sweep-start state */
} orc_de;

typedef struct {
    int npar, nchain, nmc, thin;
    double *theta;              /* current state, [nchain][npar] (= npar x nchain column-major) */
    double *lp, *ll;            /* [nchain] */
    double *out_theta;          /* [nmc][nchain][npar] (= npar x nchain x nmc column-major) */
    double *out_lp, *out_ll;    /* [nmc][nchain] */
    int store_i;
} orc_pop;

/* --- uniform source ------------------------------------------------------------------------ */
void orc_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4]);
double orc_uniform(orc_rng *r, const orc_addr *a);

/* --- LBA node-1 density (@hdr/lba.h, @hdr/design_light.h, @hdr/likelihood.h) ----------------- */
void orc_cell_params(const orc_model *m, const double *theta, int cell, double *P /* [6][n_acc] */);
int orc_lba_cell(const double *P, int n_acc, const unsigned char *posdrift, const double *u_st0,
                 const double *rt, int n, double *out);
/* per-trial log density in the caller's trial order; u_st0 = NULL means all-zero st0 draws */
void orc_trial_logdens(const orc_model *m, const orc_data *d, const double *theta, double *out);
double orc_sumloglike(const orc_model *m, const orc_data *d, const double *theta, orc_rng *r, const orc_addr *base);
/* R init path (R/phi.R:3-13): densities <= 0 floored at DBL_EPSILON before the log */
double orc_sumloglike_rinit(const orc_model *m, const orc_data *d, const double *theta);

/* --- DDM ("fastdm") density (@hdr/ddm.h as compiled into src/de.o; @hdr/likelihood.h:129-161, 295-305) ----------- */
/* One cell: P = column 0 of the 10 rows (a, d, precision, s, st0, sv, sz, t0, v, z); returns
 * validate_parameters(); out[i] = g(rt[i]) if valid, else 1e-10 (likelihood.h:158). */
int orc_ddm_cell(const double *P, int is_upper, const double *rt, int n, double *out);
/* measurement aid: {series evaluations, small-time terms, large-time terms} since the last reset */
void orc_ddm_counters(long long out[3], int reset);

/* --- priors (@hdr/prior.h, @hdr/tnorm.h) ------------------------------------------------------ */
double orc_tnorm_d(double x, double mean, double sd, double lower, double upper, int log_p);
void orc_dprior(const orc_prior *p, const double *p0, const double *p1, const double *x, double *out);
double orc_sumlogprior(const orc_prior *p, const double *p0, const double *p1, const double *x);
/* de_class::sumloghlike, src/de.cpp:245-270: thetas[s] points at subject s' [nchain][npar] state */
double orc_sumloghlike(const orc_prior *p_prior, const double *phi, int chain, const double *const *thetas, int nsubject);

/* --- chain selection (src/de.cpp:54-78, arma::shuffle) --------------------------------------- */
void orc_get_chains(int nchain, int k, int nsub, orc_rng *r, const orc_addr *base, unsigned *out);
int orc_get_subchains(int nchain, orc_rng *r, const orc_addr *base, unsigned *out);

/* --- samplers (src/de.cpp) --------------------------------------------------------------------- */
void orc_store(orc_pop *p, unsigned i);
/* one sweep; para_idx < 0 means all parameters */
void orc_crossover_subject(const orc_de *de, orc_pop *t, const orc_prior *prior, const orc_model *m, const orc_data *d,
                           orc_rng *r, unsigned pop, unsigned iter, int para_idx);
void orc_migration_subject(const orc_de *de, orc_pop *t, const orc_prior *prior, const orc_model *m, const orc_data *d,
                           orc_rng *r, unsigned pop, unsigned iter, int para_idx);
/* run_chains for run_subject (src/de.cpp:201-242, src/de2R.cpp:8-23): n_iter = (nmc-1)*thin */
void orc_run_subject(const orc_de *de, orc_pop *t, const orc_prior *prior, const orc_model *m, const orc_data *d,
                     orc_rng *r, unsigned pop, unsigned n_iter);
/* run_chains for run_hyper (src/de2R.cpp:30-47): data_theta is [nsubject][npar] */
void orc_run_hyper(const orc_de *de, orc_pop *phi, const orc_prior *p_prior, const orc_prior *h_prior,
                   const double *data_theta, int nsubject, orc_rng *r, unsigned n_iter);
/* run_hchains (src/de.cpp:272-383) */
void orc_run_hier(const orc_de *de, orc_pop *phi, orc_pop *subj, int nsubject, const orc_prior *p_prior,
                  const orc_prior *h_prior, const orc_model *m, const orc_data *d, orc_rng *r, unsigned n_iter,
                  unsigned first_subject_id);

/* bounded timing helper for bench.py: evaluates sumloglike `reps` times, returns the last value */
double orc_time_sumloglike(const orc_model *m, const orc_data *d, const double *thetas, int nchain, int reps);

#ifdef __cplusplus
}
#endif
#endif
