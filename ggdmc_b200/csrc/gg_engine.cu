// ggdmc_b200 -- host side of the engine: device state, launch sequences, C ABI.
//
// Replaces the C++ side of the reference's .Call boundary: run_subject / run_hyper / run
// (src/de2R.cpp:8-171) and the drivers de_class::run_chains / run_hchains (src/de.cpp:201-242,
// 272-383).  No PyTorch, no CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/ggdmc_b200.h"
#include "gg_kernels.cuh"
#include "gg_sampler.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "gg_host.cuh"
#include "gg_engine.cuh"

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
namespace {
int fail(char err[256], const std::exception &e, int code)
{
    if (err) {
        std::snprintf(err, 256, "%s", e.what());
    }
    return code;
}
#define GG_TRY try {
#define GG_CATCH                                                    \
    }                                                               \
    catch (const Error &e) { return fail(err, e, e.code); }         \
    catch (const std::exception &e) { return fail(err, e, GGDMC_ERR_ARG); } \
    return GGDMC_OK;
} // namespace

extern "C" {

int ggdmc_b200_abi_version(void) { return GGDMC_B200_ABI_VERSION; }

int ggdmc_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int ggdmc_b200_engine_create(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const ggdmc_prior_t *p_prior,
                             const ggdmc_prior_t *h_prior, const ggdmc_config_t *cfg, const ggdmc_start_t *phi_start,
                             const ggdmc_start_t *subj_start, ggdmc_engine_t **engine, char err[256])
{
    GG_TRY
    require(engine != nullptr, "null engine pointer");
    *engine = nullptr;
    require(model && trials && p_prior && cfg && subj_start, "null argument");
    require(h_prior == nullptr || phi_start != nullptr, "hierarchical fit needs a phi start state");
    auto *e = new ggdmc_engine();
    try {
        e->create_lba(model, trials, p_prior, h_prior, cfg, phi_start, subj_start);
    } catch (...) {
        delete e;
        throw;
    }
    *engine = e;
    GG_CATCH
}

int ggdmc_b200_engine_iterate(ggdmc_engine_t *engine, int32_t n_iter, float *elapsed_ms, char err[256])
{
    GG_TRY
    require(engine && n_iter >= 0, "bad arguments");
    engine->iterate(n_iter, elapsed_ms, nullptr, nullptr, 0);
    GG_CATCH
}

int ggdmc_b200_engine_iterate_flushed(ggdmc_engine_t *engine, int32_t n_iter, int64_t flush_bytes, float *elapsed_ms, char err[256])
{
    GG_TRY
    require(engine && n_iter >= 0 && flush_bytes > 0, "bad arguments");
    engine->iterate_flushed(n_iter, (size_t)flush_bytes, elapsed_ms);
    GG_CATCH
}

int ggdmc_b200_engine_time_likelihood(ggdmc_engine_t *engine, int32_t reps, float *elapsed_ms, int64_t *n_trial_lik, char err[256])
{
    GG_TRY
    require(engine && reps >= 1 && engine->kind != 1, "bad arguments");
    ggdmc_engine &e = *engine;
    CUDA_CHECK(cudaSetDevice(e.device));
    // proposals for every chain (crossover sweep), then the likelihood kernel alone, `reps` times
    Level &L = e.subj.L;
    const double saved = L.mig_prob;
    L.mig_prob = 0.0;
    k_sweep_begin<<<L.npop, 128, (size_t)2 * e.C * sizeof(int), e.stream>>>(L, e.d_iter.p, 0, 1, -1);
    k_propose<kProposeWarps><<<(L.npop * e.C + kProposeWarps - 1) / kProposeWarps, kProposeWarps * 32, (size_t)kProposeWarps * e.D * 8, e.stream>>>(L, e.d_iter.p, 0, -1, -1);
    L.mig_prob = saved;
    launch_like(L, e.model, e.trials.d, e.d_iter.p, 0, -1, -1, e.ll_part.p, e.stream); // warm-up
    CUDA_CHECK(cudaEventRecord(e.ev0, e.stream));
    for (int i = 0; i < reps; ++i) launch_like(L, e.model, e.trials.d, e.d_iter.p, 0, -1, -1, e.ll_part.p, e.stream);
    CUDA_CHECK(cudaEventRecord(e.ev1, e.stream));
    CUDA_CHECK(cudaStreamSynchronize(e.stream));
    e.launches += 3 + reps;
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e.ev0, e.ev1));
    if (const char *path = std::getenv("GGDMC_B200_BLOCKTRACE")) { // diagnostics: one more launch with per-block time stamps
        const size_t nblk = (size_t)L.npop * e.C * e.trials.d.nsplit;
        DBuf<unsigned long long> bt;
        bt.alloc(5 * nblk + 2);
        bt.zero();
        CUDA_CHECK(cudaStreamSynchronize(0));
        TrialData T = e.trials.d;
        T.btrace = bt.p + 2;
        k_stamp<<<1, 1, 0, e.stream>>>(bt.p);
        launch_like(L, e.model, T, e.d_iter.p, 0, -1, -1, e.ll_part.p, e.stream);
        k_stamp<<<1, 1, 0, e.stream>>>(bt.p + 1);
        CUDA_CHECK(cudaStreamSynchronize(e.stream));
        std::vector<unsigned long long> h(5 * nblk + 2);
        CUDA_CHECK(cudaMemcpy(h.data(), bt.p, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE *f = std::fopen(path, "wb")) {
            std::fwrite(h.data(), 8, h.size(), f);
            std::fclose(f);
        }
    }
    if (elapsed_ms) *elapsed_ms = ms / reps;
    if (n_trial_lik) *n_trial_lik = (int64_t)e.R * e.C * e.trials.total;
    GG_CATCH
}

int ggdmc_b200_engine_state(ggdmc_engine_t *engine, double *phi_theta, double *phi_lp, double *phi_ll, double *subj_theta,
                            double *subj_lp, double *subj_ll, char err[256])
{
    GG_TRY
    require(engine != nullptr, "null engine");
    ggdmc_engine &e = *engine;
    CUDA_CHECK(cudaSetDevice(e.device));
    CUDA_CHECK(cudaStreamSynchronize(e.stream));
    auto get = [&](double *dst, DBuf<double> &src) {
        if (dst && src.n) CUDA_CHECK(cudaMemcpy(dst, src.p, src.n * 8, cudaMemcpyDeviceToHost));
    };
    get(phi_theta, e.phi.theta); get(phi_lp, e.phi.lp); get(phi_ll, e.phi.ll);
    get(subj_theta, e.subj.theta); get(subj_lp, e.subj.lp); get(subj_ll, e.subj.ll);
    GG_CATCH
}

int64_t ggdmc_b200_engine_launch_count(const ggdmc_engine_t *engine) { return engine ? engine->launches : 0; }
int32_t ggdmc_b200_engine_is_persistent(const ggdmc_engine_t *engine) { return engine && engine->persist ? 1 : 0; }

int ggdmc_b200_engine_profile(ggdmc_engine_t *engine, int32_t enable, char err[256])
{
    GG_TRY
    require(engine != nullptr, "null engine");
    engine->profile = enable != 0;
    GG_CATCH
}

int ggdmc_b200_engine_counters(ggdmc_engine_t *engine, int64_t *trial_lik, double *like_ms, int64_t *like_launches, char err[256])
{
    GG_TRY
    require(engine != nullptr, "null engine");
    ggdmc_engine &e = *engine;
    CUDA_CHECK(cudaSetDevice(e.device));
    CUDA_CHECK(cudaStreamSynchronize(e.stream));
    unsigned long long n = 0;
    if (e.trials.counter.p) {
        CUDA_CHECK(cudaMemcpy(&n, e.trials.counter.p, sizeof(n), cudaMemcpyDeviceToHost));
        e.trials.counter.zero();
    }
    if (trial_lik) *trial_lik = (int64_t)n;
    if (like_ms) *like_ms = e.like_ms;
    if (like_launches) *like_launches = e.like_launches;
    e.like_ms = 0.0;
    e.like_launches = 0;
    GG_CATCH
}

void ggdmc_b200_engine_destroy(ggdmc_engine_t *engine)
{
    if (!engine) return;
    cudaSetDevice(engine->device);
    delete engine;
}

int ggdmc_b200_run_subject(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const ggdmc_prior_t *p_prior,
                           const ggdmc_config_t *cfg, const ggdmc_start_t *start, ggdmc_samples_t *out,
                           ggdmc_progress_fn progress, void *user, char err[256])
{
    GG_TRY
    require(model && trials && p_prior && cfg && start && out, "null argument");
    require(trials->n_subject == 1, "run_subject takes exactly one subject");
    ggdmc_engine e;
    e.create_lba(model, trials, p_prior, nullptr, cfg, nullptr, start);
    e.stream_results_to(e.subj, 1, out);
    e.short_call = (cfg->nmc - 1) * cfg->thin < 48;
    e.iterate((cfg->nmc - 1) * cfg->thin, nullptr, progress, user, cfg->report_length); // m_nsample - 1 iterations
    GG_CATCH
}

int ggdmc_b200_run_hyper(const ggdmc_prior_t *p_prior, const ggdmc_prior_t *h_prior, const double *data_theta, int32_t n_subject,
                         const ggdmc_config_t *cfg, const ggdmc_start_t *start, ggdmc_samples_t *out,
                         ggdmc_progress_fn progress, void *user, char err[256])
{
    GG_TRY
    require(cfg && start && out, "null argument");
    ggdmc_engine e;
    e.create_hyper(p_prior, h_prior, data_theta, n_subject, cfg, start);
    e.stream_results_to(e.phi, 1, out);
    e.short_call = (cfg->nmc - 1) * cfg->thin < 48;
    e.iterate((cfg->nmc - 1) * cfg->thin, nullptr, progress, user, cfg->report_length);
    GG_CATCH
}

int ggdmc_b200_run(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const ggdmc_prior_t *p_prior,
                   const ggdmc_prior_t *h_prior, const ggdmc_config_t *cfg, const ggdmc_start_t *phi_start,
                   const ggdmc_start_t *subj_start, ggdmc_samples_t *phi_out, ggdmc_samples_t *subj_out,
                   ggdmc_progress_fn progress, void *user, char err[256])
{
    GG_TRY
    require(model && trials && p_prior && h_prior && cfg && phi_start && subj_start && phi_out && subj_out, "null argument");
    PhaseTimer pt;
    {
        ggdmc_engine e;
        e.create_lba(model, trials, p_prior, h_prior, cfg, phi_start, subj_start);
        pt.lap("create");
        e.stream_results_to(e.subj, e.S, subj_out);
        e.stream_results_to(e.phi, 1, phi_out);
        e.short_call = (cfg->nmc - 1) * cfg->thin < 48;
        e.iterate((cfg->nmc - 1) * cfg->thin, nullptr, progress, user, cfg->report_length);
        pt.lap("iterate+download");
    }
    pt.lap("destroy");
    GG_CATCH
}

int ggdmc_b200_trial_logdens(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta, int32_t n_theta,
                             double *out, char err[256])
{
    GG_TRY
    require(model && trials && theta && out && n_theta >= 1, "bad arguments");
    require(trials->n_subject == 1, "trial_logdens takes exactly one subject");
    pick_device(-1);
    ModelDev M;
    M.upload(model);
    TrialsDev T;
    T.upload(trials, model->n_cell, true, model->type == GGDMC_MODEL_DDM);
    const int ntr = T.h_count[0];
    DBuf<double> d_theta, d_out;
    d_theta.upload(theta, (size_t)n_theta * model->npar);
    d_out.alloc((size_t)n_theta * std::max(ntr, 1));
    const bool ddm = M.type == GGDMC_MODEL_DDM;
    const size_t sm = ddm ? (size_t)M.d.n_cell * sizeof(DdmCell) : like_smem(M.d, 128);
    require(sm <= 220 * 1024, "cell table does not fit in shared memory");
    if (ddm) allow_smem(k_trial_logdens_ddm<128>, sm);
    else allow_smem(k_trial_logdens<128>, sm);
    if (ntr > 0) {
        dim3 grid(n_theta, std::min(64, (ntr + 127) / 128));
        if (ddm) k_trial_logdens_ddm<128><<<grid, 128, sm>>>(M.d, T.rt.p, T.cell.p, ntr, d_theta.p, d_out.p);
        else k_trial_logdens<128><<<grid, 128, sm>>>(M.d, T.rt.p, T.cell.p, ntr, d_theta.p, d_out.p);
        CUDA_CHECK(cudaGetLastError());
        std::vector<double> h((size_t)n_theta * ntr);
        CUDA_CHECK(cudaMemcpy(h.data(), d_out.p, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int k = 0; k < n_theta; ++k)
            for (int i = 0; i < ntr; ++i) out[(size_t)k * ntr + T.order[0][i]] = h[(size_t)k * ntr + i];
    }
    GG_CATCH
}

int ggdmc_b200_trial_logdens_hot(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta, int32_t n_theta,
                                 uint64_t seed, uint32_t pop, uint32_t iter, double *out, double *sums, char err[256])
{
    GG_TRY
    require(model && trials && theta && out && n_theta >= 1, "bad arguments");
    require(trials->n_subject == 1, "trial_logdens takes exactly one subject");
    require(model->type == GGDMC_MODEL_LBA, "trial_logdens_hot is an LBA probe");
    pick_device(-1);
    ModelDev M;
    M.upload(model);
    TrialsDev T;
    T.upload(trials, model->n_cell, true);
    T.set_chunking(n_theta);
    T.d.counter = nullptr;
    const int ntr = T.h_count[0];
    DBuf<double> d_theta, d_out, d_sums;
    d_theta.upload(theta, (size_t)n_theta * model->npar);
    d_out.alloc((size_t)n_theta * std::max(ntr, 1));
    d_sums.alloc((size_t)n_theta * T.d.nsplit);
    if (ntr > 0) {
        switch (M.d.n_acc) {
        case 2: launch_trial_logdens_hot<2>(M.d, T.d, d_theta.p, n_theta, ntr, seed, pop, iter, d_out.p, d_sums.p); break;
        case 3: launch_trial_logdens_hot<3>(M.d, T.d, d_theta.p, n_theta, ntr, seed, pop, iter, d_out.p, d_sums.p); break;
        case 4: launch_trial_logdens_hot<4>(M.d, T.d, d_theta.p, n_theta, ntr, seed, pop, iter, d_out.p, d_sums.p); break;
        default: launch_trial_logdens_hot<0>(M.d, T.d, d_theta.p, n_theta, ntr, seed, pop, iter, d_out.p, d_sums.p);
        }
        std::vector<double> h((size_t)n_theta * ntr), hs((size_t)n_theta * T.d.nsplit);
        CUDA_CHECK(cudaMemcpy(h.data(), d_out.p, h.size() * 8, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(hs.data(), d_sums.p, hs.size() * 8, cudaMemcpyDeviceToHost));
        for (int k = 0; k < n_theta; ++k) {
            for (int i = 0; i < ntr; ++i) out[(size_t)k * ntr + T.order[0][i]] = h[(size_t)k * ntr + i];
            if (sums) {
                double v = 0.0;
                for (int q = 0; q < T.d.nsplit; ++q) v += hs[(size_t)k * T.d.nsplit + q];
                sums[k] = v;
            }
        }
    } else if (sums) {
        for (int k = 0; k < n_theta; ++k) sums[k] = 0.0;
    }
    GG_CATCH
}

namespace {
void sumloglike_impl(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta, int32_t n_theta, double zero_floor,
                     double *out)
{
    require(model && trials && theta && out && n_theta >= 1, "bad arguments");
    pick_device(-1);
    ModelDev M;
    M.upload(model);
    TrialsDev T;
    T.upload(trials, model->n_cell, false, model->type == GGDMC_MODEL_DDM);
    T.d.zero_floor = zero_floor;
    const int S = T.S, D = model->npar;
    T.set_chunking((int64_t)S * n_theta);
    DBuf<double> d_theta, d_part;
    DBuf<int> d_target, d_mode;
    DBuf<uint64_t> d_seed;
    DBuf<uint32_t> d_iter;
    const size_t n = (size_t)S * n_theta;
    d_theta.upload(theta, n * D);
    d_part.alloc(n * T.d.nsplit);
    d_target.alloc(n); d_target.zero();
    d_mode.alloc(S); d_mode.zero();
    uint64_t z64 = 0; uint32_t z32 = 0;
    d_seed.upload(&z64, 1); d_iter.upload(&z32, 1);
    Level L{};
    L.npop = S; L.nchain = n_theta; L.npar = D; L.n_rep = 1; L.prop = d_theta.p; L.target = d_target.p;
    L.mode = d_mode.p; L.seed = d_seed.p;
    launch_like(L, M, T.d, d_iter.p, 0, -1, -1, d_part.p, 0);
    std::vector<double> h(n * T.d.nsplit);
    CUDA_CHECK(cudaMemcpy(h.data(), d_part.p, h.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) {
        double v = 0.0;
        for (int k = 0; k < T.d.nsplit; ++k) v += h[i * T.d.nsplit + k];
        out[i] = v;
    }
}
} // namespace

int ggdmc_b200_sumloglike(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta, int32_t n_theta,
                          double *out, char err[256])
{
    GG_TRY
    sumloglike_impl(model, trials, theta, n_theta, 0.0, out);
    GG_CATCH
}

int ggdmc_b200_sumloglike_init(const ggdmc_model_t *model, const ggdmc_trials_t *trials, const double *theta, int32_t n_theta,
                               double *out, char err[256])
{
    GG_TRY
    sumloglike_impl(model, trials, theta, n_theta, 2.220446049250313e-16 /* .Machine$double.eps */, out);
    GG_CATCH
}

int ggdmc_b200_sumlogprior(const ggdmc_prior_t *prior, const double *x, const double *p0, const double *p1, int32_t n, double *out,
                           char err[256])
{
    GG_TRY
    require(prior && x && out && n >= 1, "bad arguments");
    pick_device(-1);
    PriorDev P;
    P.upload(prior);
    DBuf<double> dx, d0, d1, dout;
    const size_t m = (size_t)n * prior->npar;
    dx.upload(x, m);
    if (p0) d0.upload(p0, m);
    if (p1) d1.upload(p1, m);
    dout.alloc(n);
    k_sumlogprior<<<(n + 127) / 128, 128>>>(P.d, dx.p, d0.p, d1.p, n, dout.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpy(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    GG_CATCH
}

int ggdmc_b200_select_chains(int32_t nchain, int32_t n, const int32_t *k, const double *u_partner, int32_t *out_partner,
                             const double *u_mig, int32_t *out_mig, int32_t *out_nmig, char err[256])
{
    GG_TRY
    require(nchain >= 3 && n >= 1, "bad arguments");
    require(!u_partner || (k && out_partner), "partner selection needs k and out_partner");
    require(!u_mig || (out_mig && out_nmig), "migration selection needs out_mig and out_nmig");
    pick_device(-1);
    DBuf<int> dk, dop, dom, don;
    DBuf<double> dup, dum;
    if (u_partner) { dk.upload(k, n); dup.upload(u_partner, (size_t)n * (nchain - 1)); dop.alloc((size_t)2 * n); }
    if (u_mig) { dum.upload(u_mig, (size_t)n * (nchain + 1)); dom.alloc((size_t)n * nchain); don.alloc(n); }
    k_select_chains<<<(n + 63) / 64, 64>>>(nchain, n, dk.p, dup.p, dop.p, dum.p, dom.p, don.p);
    CUDA_CHECK(cudaGetLastError());
    if (u_partner) CUDA_CHECK(cudaMemcpy(out_partner, dop.p, (size_t)2 * n * 4, cudaMemcpyDeviceToHost));
    if (u_mig) {
        CUDA_CHECK(cudaMemcpy(out_mig, dom.p, (size_t)n * nchain * 4, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(out_nmig, don.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    GG_CATCH
}

void ggdmc_b200_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    out[0] = out[1] = out[2] = out[3] = 0;
    uint32_t *d = nullptr;
    if (cudaMalloc(&d, 10 * sizeof(uint32_t)) != cudaSuccess) return;
    cudaMemcpy(d, ctr, 16, cudaMemcpyHostToDevice);
    cudaMemcpy(d + 4, key, 8, cudaMemcpyHostToDevice);
    k_philox<<<1, 1>>>(d, d + 4, d + 6);
    cudaMemcpy(out, d + 6, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
}

double ggdmc_b200_measure_fp64_tflops(int32_t device, char err[256])
{
    try {
        pick_device(device);
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device < 0 ? 0 : device));
        const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
        DBuf<double> out;
        out.alloc((size_t)blocks * threads);
        cudaEvent_t a, b;
        CUDA_CHECK(cudaEventCreate(&a));
        CUDA_CHECK(cudaEventCreate(&b));
        k_dfma_peak<<<blocks, threads>>>(out.p, 1000, 0.999999, 1e-9);
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            CUDA_CHECK(cudaEventRecord(a));
            k_dfma_peak<<<blocks, threads>>>(out.p, iters, 0.999999, 1e-9);
            CUDA_CHECK(cudaEventRecord(b));
            CUDA_CHECK(cudaEventSynchronize(b));
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
            double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        return best;
    } catch (const std::exception &e) {
        fail(err, e, 0);
        return -1.0;
    }
}

int ggdmc_b200_comm_unique_id(uint8_t id[128], char err[256])
{
    GG_TRY
    g_nccl.load();
    Nccl::UniqueId u;
    g_nccl.check(g_nccl.GetUniqueId(&u), "ncclGetUniqueId");
    std::memcpy(id, u.internal, 128);
    GG_CATCH
}

int ggdmc_b200_comm_init(int32_t n_rank, int32_t rank, const uint8_t id[128], int32_t device, char err[256])
{
    GG_TRY
    require(n_rank >= 1 && rank >= 0 && rank < n_rank, "bad rank");
    pick_device(device);
    if (n_rank == 1) { g_nccl.n_rank = 1; g_nccl.rank = 0; return GGDMC_OK; }
    g_nccl.load();
    Nccl::UniqueId u;
    std::memcpy(u.internal, id, 128);
    g_nccl.check(g_nccl.CommInitRank(&g_nccl.comm, n_rank, u, rank), "ncclCommInitRank");
    g_nccl.n_rank = n_rank;
    g_nccl.rank = rank;
    g_p2p.setup(g_nccl);
    GG_CATCH
}

void ggdmc_b200_comm_finalize(void)
{
    if (g_p2p.base) g_p2p.teardown(g_nccl.rank, g_nccl.n_rank);
    if (g_nccl.comm) {
        g_nccl.CommDestroy(g_nccl.comm);
        g_nccl.comm = nullptr;
    }
    g_nccl.n_rank = 1;
    g_nccl.rank = 0;
}

} // extern "C"
