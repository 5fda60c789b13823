/* ggdmc_b200 -- C ABI of the B200-native DE-MCMC sampling engine (LBA and DDM likelihoods).
 *
 * Drop-in boundary: the three routines ggdmc registers for .Call
 *     _ggdmc_run_subject, _ggdmc_run_hyper, _ggdmc_run      (src/RcppExports.cpp:16-64)
 * whose C++ bodies are run_subject / run_hyper / run         (src/de2R.cpp:8-23, 30-47, 123-171).
 * The reference marshals S4 objects into C++ classes there; this ABI takes the same information
 * as plain arrays (the R glue in ggdmc_b200/r/ggdmc_b200_glue.cpp and the Python mirror in
 * ggdmc_b200/api.py do the flattening), runs the whole fit on the GPU and fills caller-owned host
 * arrays laid out exactly like the reference's `posterior` slots.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is HOST memory owned by the caller
 *   - arrays are row-major in the index order written in brackets; a [nmc][nchain][npar] block is
 *     bit-for-bit R's column-major  npar x nchain x nmc  array (posterior@theta)
 *   - every entry point returns 0 on success, non-zero on error with a message in err[256]
 *     (the reference throws std::runtime_error, turned into an R error by END_RCPP,
 *     src/RcppExports.cpp:17-26); numerical trouble never errors (NaN MH ratio => reject)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     GGDMC_ERR_CUDA.
 */
#ifndef GGDMC_B200_H
#define GGDMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGDMC_B200_ABI_VERSION 2

enum ggdmc_status {
    GGDMC_OK = 0,
    GGDMC_ERR_ARG = 1,   /* bad argument / unsupported configuration */
    GGDMC_ERR_CUDA = 2,  /* CUDA runtime failure or no device */
    GGDMC_ERR_COMM = 3,  /* NCCL failure */
    GGDMC_ERR_CHAINS = 4 /* "Require three or more chains." (src/de.cpp:7-10) */
};

/* prior::DistributionType (@hdr/prior.h:186), the values R's `dist_id` carries */
enum ggdmc_dist {
    GGDMC_TNORM = 1, GGDMC_BETA_LU = 2, GGDMC_GAMMA_L = 3, GGDMC_LNORM_L = 4,
    GGDMC_CAUCHY = 5, GGDMC_UNIF = 6, GGDMC_NORM = 7
};

/* How the chains of one population are swept (DESIGN.md "Schedules").  Populations always run in
 * parallel.
 *   REFERENCE   : chains updated in place one after another, exactly like src/de.cpp:119-150 /
 *                 :167-194 (later chains see earlier chains' new states).
 *   PARALLEL    : (default) each crossover sweep is two half-sweeps: the even chains move together
 *                 with difference partners drawn from the odd chains, then the odd chains with
 *                 partners from the even ones.  The partners stand still while a half moves, so every
 *                 half-sweep is an exact Metropolis update -- same target distribution as REFERENCE,
 *                 different trajectory.  Needs nchain >= 4 (else REFERENCE is used).
 *   SIMULTANEOUS: every chain proposed from the sweep-start state and accepted together: one launch
 *                 per sweep, but only approximately invariant for small nchain.
 * Migration sweeps move all selected chains at once in PARALLEL and SIMULTANEOUS. */
enum ggdmc_schedule { GGDMC_SCHEDULE_REFERENCE = 0, GGDMC_SCHEDULE_PARALLEL = 1, GGDMC_SCHEDULE_SIMULTANEOUS = 2 };

/* model@type as likelihood_class::resolve_string reads it (@hdr/likelihood.h:279): "lba" -> lba_likelihood
 * (:73-108), "fastdm" -> ddm_likelihood (:129-161).  ("hyper" is ggdmc_b200_run_hyper; anything else is the
 * reference's "Undefined model type", :312.) */
enum ggdmc_model_type { GGDMC_MODEL_LBA = 0, GGDMC_MODEL_DDM = 1 };
#define GGDMC_LBA_ROWS 6  /* A, B, mean_v, sd_v, st0, t0 */
#define GGDMC_DDM_ROWS 10 /* a, d, precision, s, st0, sv, sz, t0, v, z */

/* dmi@model + dmi@node_1_index + dmi@is_positive_drift, flattened (SURVEY.md A.1;
 * replaces design_class, @hdr/design_light.h:77-344).  One model is shared by all subjects. */
typedef struct ggdmc_model {
    int32_t n_acc;            /* accumulators */
    int32_t n_cell;           /* design cells */
    int32_t npar;             /* free parameters of one subject (length of theta) */
    int32_t n_const;
    const int32_t *param_src; /* [n_cell][rows][n_acc]; rows = the model family's core parameters in alphabetical
                                 order (6 for the LBA, 10 for the DDM, above); column 0 = the responding
                                 accumulator (the only column the DDM reads, @hdr/ddm.h:194-214);
                                 >= 0: index into theta, < 0: const_val[-1-k] */
    const double *const_val;  /* [n_const] */
    const uint8_t *posdrift;  /* dmi@is_positive_drift.  LBA: [n_acc].  DDM: [n_cell], non-zero = the cell's
                                 response is the upper boundary (@hdr/likelihood.h:142 indexes it by cell) */
    int32_t type;             /* enum ggdmc_model_type */
} ggdmc_model_t;

/* dmi@data of every subject (replaces likelihood_class::m_rt, @hdr/likelihood.h:10).  Trials may
 * come in any order; the engine groups them by cell on upload. */
typedef struct ggdmc_trials {
    int32_t n_subject;
    const int64_t *subject_offset; /* [n_subject + 1] into rt / cell */
    const double *rt;              /* response times */
    const uint16_t *cell;          /* model cell index of every trial */
} ggdmc_trials_t;

/* prior@p_prior or prior@h_prior (replaces prior_class, @hdr/prior.h:10) */
typedef struct ggdmc_prior {
    int32_t npar;
    const double *p0, *p1, *lower, *upper; /* [npar] */
    const int32_t *dist;                    /* [npar] enum ggdmc_dist */
    const uint8_t *log_p;                   /* [npar] */
} ggdmc_prior_t;

/* config@theta_input + config@de_input + config@seed (R/model-class.R:36-57, 1288-1313,
 * 1467-1485; replaces ThetaInput / DEInput, @hdr/theta_helpers.h:6,54) */
typedef struct ggdmc_config {
    int32_t nmc;                /* stored samples per chain, slot 0 = start state */
    int32_t nchain;
    int32_t thin;
    int32_t report_length;      /* progress callback period in stored samples (0 = silent) */
    double pop_migration_prob;
    double sub_migration_prob;
    double gamma_precursor;     /* 2.38 */
    double rp;                  /* 0.001 */
    int32_t is_hblocked;
    int32_t is_pblocked;
    int32_t nparameter;         /* de_input@nparameter: sets gamma = precursor / sqrt(2 nparameter)
                                   (src/de.cpp:12); 2*npar in hierarchical fits */
    int32_t schedule;           /* enum ggdmc_schedule */
    int32_t n_replicate;        /* independent replicate runs batched in one call (the reference
                                   forks `ncore` processes instead, R/sampling.R:26-55) */
    int32_t device;             /* CUDA device ordinal, -1 = current */
    const uint64_t *seed;       /* [n_replicate] config@seed of each replicate -> Philox key */
    /* subject sharding (one process per GPU): this process owns global subjects
       [subject_begin, subject_begin + trials.n_subject); 0 / total when not sharded */
    int32_t subject_begin;
    int32_t n_subject_total;
} ggdmc_config_t;

/* A `posterior` object (R/model-class.R:238-254) for n_replicate replicates of one population.
 * On input: the state to continue from is slot `start_slot` of every replicate.
 * On output: slots 0..nmc-1 filled like theta_phi::store does (@hdr/theta.h:61-74). */
typedef struct ggdmc_samples {
    int32_t npar, nchain, nmc;
    double *theta; /* [n_replicate][nmc][nchain][npar] */
    double *lp;    /* [n_replicate][nmc][nchain]  summed_log_prior */
    double *ll;    /* [n_replicate][nmc][nchain]  log_likelihoods */
} ggdmc_samples_t;

/* start state of one population (all replicates) */
typedef struct ggdmc_start {
    const double *theta; /* [n_replicate][nchain][npar] */
    const double *lp;    /* [n_replicate][nchain] */
    const double *ll;    /* [n_replicate][nchain] */
} ggdmc_start_t;

typedef void (*ggdmc_progress_fn)(int32_t stored_sample, void *user);

/* ---- the three reference entry points ------------------------------------------------------ */

/* run_subject (src/de2R.cpp:8-23 -> de_class::run_chains, src/de.cpp:201-242).
 * trials->n_subject must be 1. */
int ggdmc_b200_run_subject(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const ggdmc_prior_t *p_prior,
                           const ggdmc_config_t *cfg, const ggdmc_start_t *start, ggdmc_samples_t *out,
                           ggdmc_progress_fn progress, void *user, char err[256]);

/* run_hyper (src/de2R.cpp:30-47): data_theta is hyper_dmi@data, [n_subject][p_prior->npar]. */
int ggdmc_b200_run_hyper(const ggdmc_prior_t *p_prior, const ggdmc_prior_t *h_prior, const double *data_theta,
                         int32_t n_subject, const ggdmc_config_t *cfg, const ggdmc_start_t *start,
                         ggdmc_samples_t *out, ggdmc_progress_fn progress, void *user, char err[256]);

/* run (src/de2R.cpp:123-171 -> de_class::run_hchains, src/de.cpp:272-383).
 * subj_start / subj_out are arrays of trials->n_subject entries (the local subjects). */
int ggdmc_b200_run(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const ggdmc_prior_t *p_prior,
                   const ggdmc_prior_t *h_prior, const ggdmc_config_t *cfg, const ggdmc_start_t *phi_start,
                   const ggdmc_start_t *subj_start, ggdmc_samples_t *phi_out, ggdmc_samples_t *subj_out,
                   ggdmc_progress_fn progress, void *user, char err[256]);

/* ---- density entry points (what ggdmcLikelihood / ggdmcPrior expose to R; used by
 *      initialise_theta, R/phi.R:166-201, and by the parity tests) ---------------------------- */

/* likelihood_class::lba_likelihood / ddm_likelihood (@hdr/likelihood.h:73-108, 129-161): log density of
 * every trial of ONE subject for n_theta parameter vectors as sumloglike takes it (:288 LBA: log(n1PDF);
 * :303 DDM: log(max(density, DBL_MIN))); out[n_theta][n_trial] in the caller's trial order. */
int ggdmc_b200_trial_logdens(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta,
                             int32_t n_theta, double *out, char err[256]);

/* The same per-trial log densities (LBA only), but computed by the trial loops of the SAMPLER's likelihood code
 * (two trials of a thread advanced in lock step for 2-accumulator models, the hot / cold split, the running product) --
 * a parity probe of the production path, not something the reference exposes.  The draws of `t0 + st0 U`
 * (@hdr/lba.h:117) are the addressed draws of (seed, population pop, iteration iter, sweep 0, chain k) for row k of
 * theta, as in a fit.  sums (may be NULL) receives the n_theta summed log-likelihoods as the sampler would see them. */
int ggdmc_b200_trial_logdens_hot(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta,
                                 int32_t n_theta, uint64_t seed, uint32_t pop, uint32_t iter, double *out, double *sums,
                                 char err[256]);

/* likelihood_class::sumloglike (@hdr/likelihood.h:272-317) for every subject x n_theta vectors:
 * theta [n_subject][n_theta][npar] -> out [n_subject][n_theta]. */
int ggdmc_b200_sumloglike(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta,
                          int32_t n_theta, double *out, char err[256]);

/* The same sum under the R-side initialisation rule (R/phi.R:3-13, `.sumlog`, used by initialise_theta
 * R/phi.R:166-201): a density <= 0 is replaced by .Machine$double.eps before the log, so a start value
 * whose likelihood the sampler would score -Inf still gets a finite (very low) score.  The replacement is made trial by
 * trial; R applies pmax(xi, eps) to a whole cell once one of its densities is <= 0, which differs only for DDM cells
 * that also hold densities in (0, eps). */
int ggdmc_b200_sumloglike_init(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta,
                               int32_t n_theta, double *out, char err[256]);

/* prior_class::sumlogprior (@hdr/prior.h:469-476): x [n][npar] -> out [n].  p0/p1 may be NULL
 * (use the prior's own) or [n][npar] per-vector overrides (the phi-driven case, src/de.cpp:599-600). */
int ggdmc_b200_sumlogprior(const ggdmc_prior_t *prior, const double *x, const double *p0, const double *p1, int32_t n,
                           double *out, char err[256]);

/* de_class::get_chains / get_subchains (src/de.cpp:54-78) evaluated on the device from explicit
 * uniforms, one problem per row: u_partner [n][nchain-1], out_partner [n][2] for chain k[n];
 * u_mig [n][nchain+1] (proportion first), out_mig [n][nchain] padded with -1, out_nmig [n]. */
int ggdmc_b200_select_chains(int32_t nchain, int32_t n, const int32_t *k, const double *u_partner,
                             int32_t *out_partner, const double *u_mig, int32_t *out_mig, int32_t *out_nmig,
                             char err[256]);

/* ---- resident engine (bench / long-lived callers): inputs stay in HBM between steps -------- */

typedef struct ggdmc_engine ggdmc_engine_t;

/* Builds the same device state `ggdmc_b200_run` builds (h_prior == NULL: independent subjects,
 * i.e. run_subject on every subject of `trials`).  out storage lives on the device. */
int ggdmc_b200_engine_create(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const ggdmc_prior_t *p_prior,
                             const ggdmc_prior_t *h_prior, const ggdmc_config_t *cfg, const ggdmc_start_t *phi_start,
                             const ggdmc_start_t *subj_start, ggdmc_engine_t **engine, char err[256]);
/* Advance n_iter DE-MCMC iterations; *elapsed_ms (may be NULL) = CUDA-event time on the engine's
 * stream around exactly those iterations. */
int ggdmc_b200_engine_iterate(ggdmc_engine_t *engine, int32_t n_iter, float *elapsed_ms, char err[256]);
/* Same, but a buffer of flush_bytes (> L2) is overwritten before every iteration, outside the
 * timed brackets; *elapsed_ms = sum of the per-iteration CUDA-event times (benchmark hygiene). */
int ggdmc_b200_engine_iterate_flushed(ggdmc_engine_t *engine, int32_t n_iter, int64_t flush_bytes, float *elapsed_ms,
                                      char err[256]);
/* Evaluate the likelihood kernel alone `reps` times on the engine's current proposals;
 * *elapsed_ms = mean CUDA-event time of one launch; *n_trial_lik = trial-likelihoods per launch. */
int ggdmc_b200_engine_time_likelihood(ggdmc_engine_t *engine, int32_t reps, float *elapsed_ms, int64_t *n_trial_lik,
                                      char err[256]);
/* Copy the current state of the local populations back: subject theta [n_subject][n_replicate][nchain][npar],
 * phi theta [n_replicate][nchain][2 npar] etc.
 * Any pointer may be NULL. */
int ggdmc_b200_engine_state(ggdmc_engine_t *engine, double *phi_theta, double *phi_lp, double *phi_ll,
                            double *subj_theta, double *subj_lp, double *subj_ll, char err[256]);
/* enable CUDA-event bracketing of every likelihood-kernel launch (off by default) */
int ggdmc_b200_engine_profile(ggdmc_engine_t *engine, int32_t enable, char err[256]);
/* Counters since the last call (then reset): trial-likelihoods actually evaluated by the likelihood
 * kernel (device counter), and -- when profiling is on -- the summed CUDA-event time and number of
 * its launches.  Any pointer may be NULL. */
int ggdmc_b200_engine_counters(ggdmc_engine_t *engine, int64_t *trial_lik, double *like_ms, int64_t *like_launches,
                               char err[256]);
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t ggdmc_b200_engine_launch_count(const ggdmc_engine_t *engine);
/* 1: the engine's iterations run inside the persistent sampler kernel (the PARALLEL schedule of an LBA fit without
 * per-parameter sweeps; whole iterations per launch), 0: as a sequence of launches per iteration.  With profiling on,
 * ggdmc_b200_engine_counters then reports the launches of that kernel instead of the likelihood kernel's. */
int32_t ggdmc_b200_engine_is_persistent(const ggdmc_engine_t *engine);
void ggdmc_b200_engine_destroy(ggdmc_engine_t *engine);

/* ---- multi-GPU (one process per GPU; subjects sharded; phi replicated) ---------------------- */

/* NCCL bootstrap: rank 0 calls get_unique_id, the bytes travel by any out-of-band channel
 * (bench.py uses torch.distributed), every rank calls comm_init before creating engines.
 *
 * Lock step.  A sharded fit exchanges the phi-level sums once per phi half-sweep through a peer-memory window (or
 * ncclAllReduce with GGDMC_B200_NO_P2P=1), so every rank must create its engine and call run / iterate with the same
 * configuration and the same number of iterations.  The ranks do NOT have to arrive at the same time: the end of
 * engine creation and the start of every run / iterate call are peer barriers, and inside an exchange a rank waits up
 * to GGDMC_B200_PEER_TIMEOUT_S seconds (default 120) for its peers.  If a peer never arrives, the waiting rank takes
 * no MH decision on the incomplete sums, its call returns GGDMC_ERR_COMM ("peer exchange timed out"), and the
 * communicator stays in that error state -- every later engine creation fails with GGDMC_ERR_COMM -- until
 * ggdmc_b200_comm_finalize() and a new ggdmc_b200_comm_init().  One communicator per process: sharded fits of one
 * process run one after the other. */
int ggdmc_b200_comm_unique_id(uint8_t id[128], char err[256]);
int ggdmc_b200_comm_init(int32_t n_rank, int32_t rank, const uint8_t id[128], int32_t device, char err[256]);
void ggdmc_b200_comm_finalize(void);

/* ---- utilities ----------------------------------------------------------------------------- */
int ggdmc_b200_abi_version(void);
int ggdmc_b200_device_count(void);
/* FP64 FMA peak of the current device measured with a dependent-chain-free DFMA microbenchmark
 * (roofline denominator; MEASURED_PEAKS.json has no FP64 entry).  Returns TFLOP/s, < 0 on error. */
double ggdmc_b200_measure_fp64_tflops(int32_t device, char err[256]);
/* Philox4x32-10 block (test hook for the counter-based uniform source) */
void ggdmc_b200_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* GGDMC_B200_H */
