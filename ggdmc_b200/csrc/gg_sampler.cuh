// ggdmc_b200 -- the persistent sampler kernel: whole DE-MCMC iterations of the PARALLEL schedule in ONE launch.
//
// Replaces, for the default schedule, the per-half-sweep launch chain  k_sweep_begin -> k_propose -> k_like -> k_accept
// (x 2 half-sweeps x subject groups) + k_phi_half x 2 + k_phi_consts + k_store_advance  = 19 launches per iteration
// (run_hchains body, src/de.cpp:281-381; run_chains body :208-240) by a queue of work items that a fixed set of
// resident CTAs (one per SM x occupancy) drains:
//
//   PHI item  (iteration t, half h, replicate r, phi chain c, subject split)
//       proposal of phi chain c (if it moves in this half) + its share of the two hyper-likelihood sums
//       (de_class::sumloghlike, src/de.cpp:245-270); the CTA that finishes the half last sums the splits, exchanges the
//       sums with the peer GPUs through the peer-memory window, takes every MH decision of the half (:397-463, :494-549),
//       and after half 1 refreshes the constants of the phi-driven subject prior, stores the thinned sample and draws the
//       next iteration's migration decision.
//   SUBJECT item (iteration t, half h, population p, slot, trial chunk)
//       proposal of chain 2 slot + h (crossover; migration: sweep positions slot and slot + nslot in half 0) made by warp 0
//       into shared memory, cell table built from it, trial loop, block reduction (de.cpp:567-613, 615-665 and
//       @hdr/likelihood.h:73-108, 272-292); the CTA that finishes the population's half last takes the half's MH decisions
//       (update_theta, :81-108) -- after half 1 also theta_phi::store (@hdr/theta.h:61-74) and the next iteration's
//       migration decision (get_subchains, :62-78).
//
// Dependencies are device-side flags instead of kernel boundaries, so nothing waits for a whole grid to drain:
//   subject item (p, t, h)        needs  pop_done[p] >= 2 t + h      (the population's previous half is accepted)
//   subject MH test (p, t, h)     needs  phi_done    >= 2 t + 2      (this iteration's phi: the prior of theta)
//   PHI item (t, 0)               needs  all_done    >= npop (t - 1) (every local population finished iteration t - 1)
//   PHI item (t, 1)               needs  phi_done    >= 2 t + 1
// Items are handed out in queue order  [PHI half 0][PHI half 1][SUBJECT half 0][SUBJECT half 1]  per iteration, and
// every dependency points to an item EARLIER in the queue, i.e. to an item some running CTA has already taken: the
// lowest unfinished item can always proceed, whatever the number of resident CTAs (no co-residency assumption).
// A wait that exceeds spin_ns raises `abort`; every CTA then leaves and the host reports the error.
//
// All state other CTAs rewrite while the kernel runs is read with ld.global.cg (ldm()); flags are published with a
// device-scope fence + release store after a block barrier and read with acquire loads.
#pragma once
#include "gg_kernels.cuh"

namespace gg {

struct SamplerSync {
    unsigned long long *queue;     // next work item of the launch
    unsigned int *exit_ctr;        // CTAs that have left the launch (the last one resets the queue)
    unsigned int *pop_arrive;      // [npop] finished items of the population's current half
    unsigned int *pop_done;        // [npop] 2 t + h + 1 once half h of iteration t is accepted (and stored)
    unsigned long long *all_done;  // (population, iteration) pairs completed since iteration 1
    unsigned int *phi_arrive;      // finished PHI items of the current half
    unsigned int *phi_done;        // 2 t + h + 1
    int *abort;                    // != 0: a wait timed out (1 peer exchange, 2 local flag); every CTA leaves
    unsigned long long spin_ns;    // bound of every local wait
};

struct SamplerArgs {
    Level S, P;         // subject level (all local populations), phi level (hier only)
    DevModel M;
    TrialData T;
    HyperArgs H;
    P2PWindow w;
    SamplerSync y;
    double *ll_part;    // [npop][C][nsplit]
    double *hpart, *hsum, *phi_consts;
    uint32_t *d_iter;   // set to t_end by the last CTA to leave (the multi-launch kernels read it)
    uint32_t t_begin, t_end; // iterations [t_begin, t_end)
    int hier;           // 1: phi level present (run), 0: independent subjects (run_subject)
    int use_p2p;        // phi sums are exchanged with peer GPUs
    int decide_once;    // run_chains draws the migration decision once per iteration (src/de.cpp:210)
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// thread 0 of the block waits until *flag >= need; false: aborted / timed out
__device__ __forceinline__ bool spin_until_u32(const unsigned int *flag, unsigned int need, const SamplerSync &y)
{
    if (ld_acquire_u32(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        __nanosleep(40);
        if (ld_acquire_u32(flag) >= need) return true;
        if (*(volatile int *)y.abort) return false;
        if (globaltimer_ns() - t0 > y.spin_ns) {
            atomicCAS(y.abort, 0, 2);
            return false;
        }
    }
}
__device__ __forceinline__ bool spin_until_u64(const unsigned long long *flag, unsigned long long need, const SamplerSync &y)
{
    if (ld_acquire_u64(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        __nanosleep(40);
        if (ld_acquire_u64(flag) >= need) return true;
        if (*(volatile int *)y.abort) return false;
        if (globaltimer_ns() - t0 > y.spin_ns) {
            atomicCAS(y.abort, 0, 2);
            return false;
        }
    }
}

// theta_phi::store of ONE population (block-wide): iteration `iter` goes to slot iter / thin when thin divides it
__device__ __forceinline__ void store_pop(const Level &L, int p, uint32_t iter)
{
    if (iter % (uint32_t)L.thin != 0) return;
    const uint32_t slot = iter / (uint32_t)L.thin;
    if (slot >= (uint32_t)L.nmc) return;
    const int C = L.nchain, CD = L.nchain * L.npar;
    const double *th = L.theta + (size_t)p * CD;
    double *o = L.out_theta + ((size_t)p * L.nmc + slot) * CD;
    for (int i = threadIdx.x; i < CD; i += blockDim.x) o[i] = ldm(th + i);
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        L.out_lp[((size_t)p * L.nmc + slot) * C + i] = ldm(L.lp + (size_t)p * C + i);
        L.out_ll[((size_t)p * L.nmc + slot) * C + i] = ldm(L.ll + (size_t)p * C + i);
    }
}

// constants of the phi-driven truncated-normal prior of the subject level for phi chain (r, c) (what k_phi_consts
// computes for all chains); threads [0, D) of the block
__device__ __forceinline__ void phi_consts_chain(const Level &P, const DevPrior &like, int D, int r, int c, double *consts, int tid, int nthr)
{
    const size_t rc = (size_t)r * P.nchain + c;
    const double *phi = P.theta + rc * 2 * D;
    for (int d = tid; d < D; d += nthr) {
        const double m = ldm(phi + d), sd = ldm(phi + D + d);
        double inv = 0.0, K = NAN;
        if (like.dist[d] == 1 && like.log_p[d] != 0 && sd > 0.0 && isfinite(sd) && isfinite(m)) {
            const double den = pnorm5(like.upper[d], m, sd, true) - pnorm5(like.lower[d], m, sd, true);
            inv = 1.0 / sd;
            K = kLnSqrt2Pi + log(sd) + log(den);
        }
        consts[2 * (rc * D + d)] = inv;
        consts[2 * (rc * D + d) + 1] = K;
    }
}

// Shared memory of a sampler CTA (bytes): the larger of the two item layouts + a common tail.
//   SUBJECT: like_smem (cell table, reduction scratch, classes) | theta' [D] | prior scratch [D]
//   PHI    : hyper_block 6 D + 2 BLOCK/32 | proposal [2 D] | prior scratch [2 D]           (doubles)
//   tail   : migration keys / ranks [2 C] ints | control ints [8]
__host__ __device__ inline size_t sampler_like_bytes(int n_cell, int n_acc, int block)
{
    size_t b = (size_t)n_cell * n_acc * sizeof(CellAcc) + (size_t)(block / 32) * 8 + (size_t)n_cell * (1 + n_acc);
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t sampler_item_bytes(int n_cell, int n_acc, int D, int block, int hier)
{
    const size_t subj = sampler_like_bytes(n_cell, n_acc, block) + (size_t)2 * D * 8;
    const size_t phi = hier ? (size_t)(6 * D + 2 * (block / 32) + 4 * D) * 8 : 0;
    return ((subj > phi ? subj : phi) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t sampler_smem_bytes(int n_cell, int n_acc, int D, int C, int block, int hier)
{
    return sampler_item_bytes(n_cell, n_acc, D, block, hier) + (size_t)(2 * C + 8) * sizeof(int);
}

template <int NACC, int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_sampler(SamplerArgs A)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const Level &S = A.S;
    const Level &P = A.P;
    const SamplerSync &y = A.y;
    const int C = S.nchain, D = S.npar, tid = threadIdx.x;
    const int nslot = (C + 1) / 2, nsplit = A.T.nsplit;
    const size_t item_bytes = sampler_item_bytes(A.M.n_cell, A.M.n_acc, D, BLOCK, A.hier);
    int *sm_keys = reinterpret_cast<int *>(sm_raw + item_bytes); // [2 C]
    int *ctl = sm_keys + 2 * C;                                  // [8]: 0 ok flag, 1 last flag, 2-3 sweep_begin scratch, 4 phi k, 5-6 item (u64)
    unsigned long long *sm_item = reinterpret_cast<unsigned long long *>(ctl + 6);

    const unsigned long long n_phi = A.hier ? (unsigned long long)P.npop * C * A.H.nsplit : 0ull; // items of one phi half
    const unsigned long long n_sub = (unsigned long long)S.npop * nslot * nsplit;                 // items of one subject half
    const unsigned long long per_iter = 2 * n_phi + 2 * n_sub;
    const unsigned long long total = per_iter * (unsigned long long)(A.t_end - A.t_begin);

    for (;;) {
        __syncthreads(); // shared memory of the previous item is free
        if (tid == 0) {
            unsigned long long it = atomicAdd(y.queue, 1ull);
            if (*(volatile int *)y.abort) it = ~0ull;
            *sm_item = it;
        }
        __syncthreads();
        const unsigned long long item = *sm_item;
        if (item >= total) break;
        const uint32_t t = A.t_begin + (uint32_t)(item / per_iter);
        unsigned long long j = item % per_iter;

        if (j < 2 * n_phi) {
            // ------------------------------------------------------------------ PHI item
            const int h = j >= n_phi ? 1 : 0;
            if (h) j -= n_phi;
            const int Hs = A.H.nsplit;
            const int split = (int)(j % Hs);
            const int rc = (int)(j / Hs);
            const int r = rc / C, c = rc - r * C;
            if (tid == 0) {
                bool ok = h == 0 ? spin_until_u64(y.all_done, (unsigned long long)S.npop * (t - 1), y)
                                 : spin_until_u32(y.phi_done, 2 * t + 1, y);
                ctl[0] = ok;
            }
            __syncthreads();
            if (!ctl[0]) break;
            double *sm_h = reinterpret_cast<double *>(sm_raw);
            const int k = phi_half_part<BLOCK>(P, A.H, t, 0, h, r, c, split, A.hpart, sm_h, &ctl[4]);
            // a chain that does not move in half 1 has its final value of this iteration: its prior constants can be made now
            if (h == 1 && k < 0 && split == 0) phi_consts_chain(P, A.H.like, D, r, c, A.phi_consts, tid, BLOCK);
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                ctl[1] = atomicAdd(y.phi_arrive, 1u) == (unsigned int)(n_phi - 1);
            }
            __syncthreads();
            if (!ctl[1]) continue;
            __threadfence();
            if (!phi_half_finish<BLOCK>(P, A.H, t, 0, A.hpart, A.hsum, A.w, A.use_p2p)) {
                if (tid == 0) atomicCAS(y.abort, 0, 1);
                break;
            }
            __syncthreads();
            if (h == 1) {
                // chains that moved in this half (crossover: the odd ones; migration moves everything in half 0)
                for (int r2 = 0; r2 < P.npop; ++r2)
                    if (ldm(P.mode + r2) == 0)
                        for (int c2 = 1; c2 < C; c2 += 2) phi_consts_chain(P, A.H.like, D, r2, c2, A.phi_consts, tid, BLOCK);
                for (int r2 = 0; r2 < P.npop; ++r2) store_pop(P, r2, t);
                __syncthreads();
                for (int r2 = 0; r2 < P.npop; ++r2) {
                    sweep_begin_pop(P, r2, t + 1, 0, 0, -1, sm_keys, ctl + 2);
                    __syncthreads();
                }
            }
            __syncthreads();
            if (tid == 0) {
                *y.phi_arrive = 0;
                __threadfence();
                st_release_u32(y.phi_done, 2 * t + h + 1);
            }
            continue;
        }

        // ---------------------------------------------------------------------- SUBJECT item
        j -= 2 * n_phi;
        const int h = j >= n_sub ? 1 : 0;
        if (h) j -= n_sub;
        const int split = (int)(j % nsplit);
        const int ps = (int)(j / nsplit);
        const int p = ps / nslot, slot = ps - p * nslot;
        if (tid == 0) ctl[0] = spin_until_u32(y.pop_done + p, 2 * t + h, y);
        __syncthreads();
        if (!ctl[0]) break;
        const int mode = ldm(S.mode + p);
        const int nsteps = mode ? ldm(S.mig_n + p) : C;
        double *sm_theta = reinterpret_cast<double *>(sm_raw + sampler_like_bytes(A.M.n_cell, A.M.n_acc, BLOCK));
        double *sm_scratch = sm_theta + D;
        // sweep positions of this item: crossover -> chain 2 slot + h; migration (all of it in half 0) -> slot, slot + nslot
        int k = -1, k2 = -1;
        if (mode == 0) {
            k = 2 * slot + h;
        } else if (mode == 1 && h == 0) {
            k = slot;
            k2 = slot + nslot;
        }
        for (; k >= 0; k = k2, k2 = -1) {
            if (k >= nsteps) continue;
            if (tid < 32) {
                int src, tgt;
                double lp;
                propose_position(S, p, k, mode, nsteps, -1, t, 0, h, tid, sm_scratch, sm_theta, src, tgt, lp);
                if (split == 0) {
                    double *pr = S.prop + ((size_t)p * C + src) * D;
                    for (int d = tid; d < D; d += 32) pr[d] = sm_theta[d];
                    if (tid == 0) {
                        if (!S.prior_ovr) S.prop_lp[p * C + src] = lp;
                        S.target[p * C + src] = tgt;
                    }
                }
                if (tid == 0) ctl[4] = src;
            }
            __syncthreads();
            const int src = ctl[4];
            const double v = like_eval<NACC, BLOCK>(A.M, A.T, sm_theta, make_addr(S, p, t, 0, src), p / S.n_rep, split, sm_raw);
            if (tid == 0) A.ll_part[((size_t)p * C + src) * nsplit + split] = v;
            __syncthreads(); // table, theta' and scratch are reused by a second sweep position
        }
        if (tid == 0) {
            __threadfence();
            ctl[1] = atomicAdd(y.pop_arrive + p, 1u) == (unsigned int)(nslot * nsplit - 1);
        }
        __syncthreads();
        if (!ctl[1]) continue;
        // last item of (p, t, h): the half's MH tests
        __threadfence();
        if (A.hier) {
            if (tid == 0) ctl[0] = spin_until_u32(y.phi_done, 2 * t + 2, y);
            __syncthreads();
            if (!ctl[0]) break;
        }
        for (int src = tid; src < C; src += BLOCK) accept_one(S, p, src, t, 0, A.ll_part, nsplit);
        __syncthreads();
        if (h == 1) {
            store_pop(S, p, t);
            __syncthreads();
            sweep_begin_pop(S, p, t + 1, 0, A.decide_once, -1, sm_keys, ctl + 2);
        }
        __syncthreads();
        if (tid == 0) {
            y.pop_arrive[p] = 0;
            __threadfence();
            st_release_u32(y.pop_done + p, 2 * t + h + 1);
            if (h == 1) atomicAdd(y.all_done, 1ull);
        }
    }

    // leave: the last CTA out re-arms the queue for the next launch and publishes the iteration counter
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(y.exit_ctr, 1u) == gridDim.x - 1) {
            *y.exit_ctr = 0;
            *y.queue = 0ull;
            *A.d_iter = A.t_end;
            __threadfence();
        }
    }
}

} // namespace gg
