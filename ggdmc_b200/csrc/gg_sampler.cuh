// ggdmc_b200 -- the persistent sampler kernel: whole DE-MCMC iterations of the PARALLEL schedule in ONE launch.
//
// Replaces, for the default schedule, the per-half-sweep launch chain  k_sweep_begin -> k_propose -> k_like -> k_accept
// (x 2 half-sweeps x subject groups) + k_phi_half x 2 + k_phi_consts + k_store_advance  = 19 launches per iteration
// (run_hchains body, src/de.cpp:281-381; run_chains body :208-240).  Every WARP of the launch is an independent worker
// that drains one device-wide queue of work items; nothing inside the loop is block-wide (no __syncthreads, no idle
// partner warp while one warp proposes), so an SM's 24 resident warps sit at 24 different points of their items and
// the FP64 pipe always finds trial loops to run.
//
//   SUBJECT item (iteration t, half h, population p, slot, trial chunk)
//       proposal of chain 2 slot + h (crossover; migration: sweep positions slot and slot + nslot in half 0) into shared
//       memory, row table built from it, trial loop, warp reduction (de.cpp:567-613, 615-665 and @hdr/likelihood.h:73-108,
//       272-292).  The worker that finishes a crossover proposal's last trial chunk takes its MH decision on the spot
//       (update_theta, :81-108): within a half-sweep nobody else reads or writes that chain.  Only if the decision needs a
//       phi that is not there yet (half 0 of a hierarchy) the proposal is parked in global memory for the ACCEPT items.
//       A migration sweep rewrites chains other proposals start from, so its decisions are taken together by the worker
//       that finishes the population's half last.  That worker also closes the half: after half 1 theta_phi::store
//       (@hdr/theta.h:61-74) and the next iteration's migration decision (get_subchains, :62-78).
//   PHI item  (iteration t, half h, replicate r, phi chain c, subject split)
//       proposal of phi chain c (if it moves in this half), the constants of the subject prior it would drive, and its
//       share of the two hyper-likelihood sums (de_class::sumloghlike, src/de.cpp:245-270); the worker that finishes the
//       half last sums the splits, exchanges the sums with the peer GPUs through the peer-memory window, takes every MH
//       decision of the half (:397-463, :494-549), and after half 1 stores the thinned sample and draws the next
//       iteration's migration decision.
//   ACCEPT item (iteration t, population p, group of 32 chains; hierarchy only)
//       the parked MH decisions of the population's half 0 (they need this iteration's phi: the prior of theta), then
//       closes the half.
//
// The queue order of an iteration is a list of segments made by the host (sampler_segments in gg_engine.cu):
//   [SUBJECT half 0][PHI half 0][PHI half 1][ACCEPT][SUBJECT half 1]                       small and medium fits
//   [SUBJECT half 0, first part][PHI half 0][.. second part][PHI half 1][.. rest][ACCEPT][SUBJECT half 1]   large fits
// Dependencies are device-side flags instead of kernel boundaries, so nothing waits for a whole grid to drain:
//   SUBJECT (p, t, h)   needs  pop_done[p] >= 2 t + h        (the population's previous half is closed)
//   PHI (t, 0)          needs  all_done    >= npop (t - 1)   (every local population finished iteration t - 1)
//   PHI (t, 1)          needs  phi_done    >= 2 t + 1
//   ACCEPT (p, t)       needs  pop_arrive[p] complete and phi_done >= 2 t + 2
// Every dependency points to an item EARLIER in the queue, i.e. to an item some running worker has already taken: the
// lowest unfinished item can always proceed, whatever the number of resident warps (no co-residency assumption).  In a
// large fit a worker reaches the phi items when the previous iteration is long finished and the ACCEPT items when phi is
// (no worker ever spins), the phi step runs underneath half 0, and most of half 0 finds phi ready and decides on the spot.
// A wait that exceeds spin_ns raises `abort`; every worker then leaves and the host reports the error.
//
// All state other workers rewrite while the kernel runs is read with ld.global.cg (ldm()); flags are published with a
// device-scope fence + release store after a warp barrier and read with acquire loads.
#pragma once
#include "gg_kernels.cuh"

namespace gg {

constexpr int kPopFlagStride = 32; // ints: every population's flags live in a 128-byte line of their own

struct SamplerSync {
    unsigned long long *queue;     // next work item of the launch
    unsigned int *exit_ctr;        // workers that have left the launch (the last one resets the queue)
    unsigned int *pop_flags;       // [npop][kPopFlagStride]: [0] done = 2 t + h + 1 once half h of iteration t is closed,
                                   //                          [1] arrive = finished items of the population's current half
    unsigned int *chain_arrive;    // [npop][C] finished trial chunks of the chain's current proposal (nsplit > 1)
    unsigned long long *all_done;  // (population, iteration) pairs completed since iteration 1
    unsigned int *phi_arrive;      // finished PHI items of the current half
    unsigned int *phi_done;        // 2 t + h + 1
    int *abort;                    // != 0: a wait timed out (1 peer exchange, 2 local flag); every worker leaves
    unsigned long long spin_ns;    // bound of every local wait
};

enum : int { kItemSubject = 0, kItemPhi = 1, kItemAccept = 2 };
constexpr int kMaxSeg = 8;

struct SamplerArgs {
    Level S, P;         // subject level (all local populations), phi level (hier only)
    DevModel M;
    TrialData T;
    HyperArgs H;
    P2PWindow w;
    SamplerSync y;
    double *ll_part;    // [npop][C][nsplit]
    double *hpart, *hsum;
    uint32_t *d_iter;   // set to t_end by the last worker to leave (the multi-launch kernels read it)
    uint32_t t_begin, t_end; // iterations [t_begin, t_end)
    int hier;           // 1: phi level present (run), 0: independent subjects (run_subject)
    int use_p2p;        // phi sums are exchanged with peer GPUs
    int decide_once;    // run_chains draws the migration decision once per iteration (src/de.cpp:210)
    int stage_bytes, warp_bytes; // shared memory: model tables staged once per CTA, then one region per warp
    // queue order of one iteration: segment i holds items [seg_first[i], seg_first[i] + seg_count[i]) of (kind, half)
    int n_seg, seg_kind[kMaxSeg], seg_half[kMaxSeg];
    unsigned long long seg_first[kMaxSeg], seg_count[kMaxSeg], per_iter;
    // diagnostics (GGDMC_B200_ITEMTRACE): 8 stamps per item of the launch's first trace_cap items --
    // taken, dependency met, proposal made, table built, trial loop done, finished (ns); SM id; item kind
    unsigned long long *trace;
    unsigned long long trace_cap;
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

// ONE thread waits until *flag >= need (polls back off to 256 ns); false: aborted / timed out
__device__ __forceinline__ bool spin_until_u32(const unsigned int *flag, unsigned int need, const SamplerSync &y)
{
    if (ld_acquire_u32(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int ns = 32;
    for (;;) {
        __nanosleep(ns);
        if (ld_acquire_u32(flag) >= need) return true;
        if (ns < 256) ns += ns;
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
    unsigned int ns = 32;
    for (;;) {
        __nanosleep(ns);
        if (ld_acquire_u64(flag) >= need) return true;
        if (ns < 256) ns += ns;
        if (*(volatile int *)y.abort) return false;
        if (globaltimer_ns() - t0 > y.spin_ns) {
            atomicCAS(y.abort, 0, 2);
            return false;
        }
    }
}

// theta_phi::store of ONE population by one warp: iteration `iter` goes to slot iter / thin when thin divides it
__device__ __forceinline__ void store_pop(const Level &L, int p, uint32_t iter, int lane)
{
    if (iter % (uint32_t)L.thin != 0) return;
    const uint32_t slot = iter / (uint32_t)L.thin;
    if (slot >= (uint32_t)L.nmc) return;
    const int C = L.nchain, CD = L.nchain * L.npar;
    const double *th = L.theta + (size_t)p * CD;
    double *o = L.out_theta + ((size_t)p * L.nmc + slot) * CD;
    for (int i = lane; i < CD; i += 32) o[i] = ldm(th + i);
    for (int i = lane; i < C; i += 32) {
        L.out_lp[((size_t)p * L.nmc + slot) * C + i] = ldm(L.lp + (size_t)p * C + i);
        L.out_ll[((size_t)p * L.nmc + slot) * C + i] = ldm(L.ll + (size_t)p * C + i);
    }
}

// Shared memory of a sampler CTA (bytes):  [model stage][region of warp 0][region of warp 1] ...
//   stage : row_of [n_cell n_acc] u16 | row_src [8 n_row] int | const_val [max(1, n_const)] double
//   region: the larger of   SUBJECT  like_smem_bytes (row table, classes) | theta' [D] | prior scratch [D]
//                           PHI      hyper_block 8 D + 2 | proposal [2 D] | prior scratch [2 D]     (doubles)
//                           migration keys / ranks [2 C] ints (sweep_begin_pop; the item's table is dead by then)
//           + 4 control ints
__host__ __device__ inline size_t sampler_stage_bytes(int n_cell, int n_acc, int n_row, int n_const)
{
    size_t b = (((size_t)n_cell * n_acc * 2 + 7) & ~(size_t)7) + (size_t)n_row * 8 * sizeof(int);
    b = (b + 7) & ~(size_t)7;
    b += (size_t)(n_const > 0 ? n_const : 1) * 8;
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t sampler_warp_bytes(int n_cell, int n_row, int D, int C, int hier)
{
    size_t b = like_smem_bytes(n_row, n_cell, 32) + (size_t)2 * D * 8;
    const size_t phi = hier ? (size_t)(12 * D + 2) * 8 : 0;
    const size_t keys = (size_t)2 * C * sizeof(int);
    b = b > phi ? b : phi;
    b = b > keys ? b : keys;
    b = (b + 7) & ~(size_t)7;
    return (b + 4 * sizeof(int) + 15) & ~(size_t)15;
}

constexpr int kSamplerMaxThreads = 256; // a CTA is 1 .. 8 independent warps; 24 warps per SM at 80 registers

template <int NACC>
__global__ void __launch_bounds__(kSamplerMaxThreads, 3) k_sampler(SamplerArgs A)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const Level &S = A.S;
    const Level &P = A.P;
    const SamplerSync &y = A.y;
    const int C = S.nchain, D = S.npar, lane = threadIdx.x & 31;
    const int nslot = (C + 1) / 2, nsplit = A.T.nsplit;
    constexpr unsigned FULL = 0xffffffffu;

    // ---- the model's tables, once per CTA ------------------------------------------------------------------------
    DevModel M = A.M;
    {
        uint16_t *s_row_of = reinterpret_cast<uint16_t *>(sm_raw);
        const int n_ent = M.n_cell * M.n_acc;
        int *s_row_src = reinterpret_cast<int *>(sm_raw + (((size_t)n_ent * 2 + 7) & ~(size_t)7));
        double *s_const = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(s_row_src) + ((((size_t)M.n_row * 8 * sizeof(int)) + 7) & ~(size_t)7));
        for (int i = threadIdx.x; i < n_ent; i += blockDim.x) s_row_of[i] = M.row_of[i];
        for (int i = threadIdx.x; i < M.n_row * 8; i += blockDim.x) s_row_src[i] = M.row_src[i];
        for (int i = threadIdx.x; i < M.n_const; i += blockDim.x) s_const[i] = M.const_val[i];
        __syncthreads(); // the only block-wide barrier of the kernel
        M.row_of = s_row_of;
        M.row_src = s_row_src;
        M.const_val = s_const;
    }
    unsigned char *wsm = sm_raw + A.stage_bytes + (size_t)(threadIdx.x >> 5) * A.warp_bytes;
    int *ctl = reinterpret_cast<int *>(wsm + A.warp_bytes - 4 * sizeof(int)); // [4] scratch of sweep_begin_pop / phi_half_part
    int *sm_keys = reinterpret_cast<int *>(wsm);                               // [2 C], aliases the item's table

    const unsigned long long n_phi = A.hier ? (unsigned long long)P.npop * C * A.H.nsplit : 0ull; // items of one phi half
    const int n_grp = (C + 31) / 32;                                                              // ACCEPT items per population
    const unsigned long long total = A.per_iter * (unsigned long long)(A.t_end - A.t_begin);
    const unsigned int per_pop_half = (unsigned int)(nslot * nsplit);

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) {
            item = atomicAdd(y.queue, 1ull);
            if (*(volatile int *)y.abort) item = ~0ull;
        }
        item = __shfl_sync(FULL, item, 0);
        if (item >= total) break;
        unsigned long long *tr = (A.trace && item < A.trace_cap && lane == 0) ? A.trace + 8 * item : nullptr;
        if (tr) {
            unsigned int sm, wslot;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            asm volatile("mov.u32 %0, %warpid;" : "=r"(wslot));
            tr[0] = globaltimer_ns();
            tr[2] = tr[3] = tr[4] = 0;
            tr[6] = sm | ((unsigned long long)wslot << 16) | ((unsigned long long)(threadIdx.x >> 5) << 32);
        }
        const uint32_t t = A.t_begin + (uint32_t)(item / A.per_iter);
        unsigned long long j = item % A.per_iter;
        int kind = 0, h = 0;
        for (int i = 0; i < A.n_seg; ++i) {
            if (j < A.seg_count[i]) {
                kind = A.seg_kind[i];
                h = A.seg_half[i];
                j += A.seg_first[i];
                break;
            }
            j -= A.seg_count[i];
        }
        if (tr) tr[7] = (unsigned long long)(kind == kItemSubject ? h : kind == kItemPhi ? 2 + h : 4);

        if (kind == kItemPhi) {
            // ------------------------------------------------------------------ PHI item
            const int Hs = A.H.nsplit;
            const int split = (int)(j % Hs);
            const int rc = (int)(j / Hs);
            const int r = rc / C, c = rc - r * C;
            int ok = 1;
            if (lane == 0)
                ok = h == 0 ? spin_until_u64(y.all_done, (unsigned long long)S.npop * (t - 1), y) : spin_until_u32(y.phi_done, 2 * t + 1, y);
            if (!__shfl_sync(FULL, ok, 0)) break;
            if (tr) tr[1] = globaltimer_ns();
            phi_half_part<32>(P, A.H, t, 0, h, r, c, split, A.hpart, reinterpret_cast<double *>(wsm), ctl);
            __syncwarp();
            if (tr) tr[4] = globaltimer_ns();
            int last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicAdd(y.phi_arrive, 1u) == (unsigned int)(n_phi - 1);
            }
            if (__shfl_sync(FULL, last, 0)) {
                __threadfence();
                if (!phi_half_finish<32>(P, A.H, t, 0, A.hpart, A.hsum, A.w, A.use_p2p)) {
                    if (lane == 0) atomicCAS(y.abort, 0, 1);
                    break;
                }
                __syncwarp();
                if (h == 1) {
                    __threadfence();
                    for (int r2 = 0; r2 < P.npop; ++r2) store_pop(P, r2, t, lane);
                    for (int r2 = 0; r2 < P.npop; ++r2) {
                        __syncwarp();
                        sweep_begin_pop<32>(P, r2, t + 1, 0, 0, -1, sm_keys, ctl);
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    *y.phi_arrive = 0;
                    __threadfence();
                    st_release_u32(y.phi_done, 2 * t + h + 1);
                }
            }
            if (tr) tr[5] = globaltimer_ns();
            continue;
        }

        int p, closer = 0;
        if (kind == kItemAccept) {
            // ------------------------------------------- ACCEPT item: the parked MH decisions of (p, t, half 0), 32 chains
            p = (int)(j / n_grp);
            const int g = (int)(j - (unsigned long long)p * n_grp);
            unsigned int *pf = y.pop_flags + (size_t)p * kPopFlagStride;
            int ok = 1;
            if (lane == 0) ok = spin_until_u32(pf + 1, per_pop_half, y) && spin_until_u32(y.phi_done, 2 * t + 2, y);
            if (!__shfl_sync(FULL, ok, 0)) break;
            if (tr) tr[1] = globaltimer_ns();
            accept_warp(S, p, 32 * g, min(C, 32 * g + 32), t, 0, A.ll_part, nsplit, lane);
            __syncwarp();
            int last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicAdd(pf + 1, 1u) == per_pop_half + (unsigned int)n_grp - 1;
            }
            closer = __shfl_sync(FULL, last, 0);
        } else {
            // ---------------------------------------------------------------------- SUBJECT item
            const int split = (int)(j % nsplit);
            const int ps = (int)(j / nsplit);
            p = ps / nslot;
            const int slot = ps - p * nslot;
            unsigned int *pf = y.pop_flags + (size_t)p * kPopFlagStride;
            int ok = 1;
            if (lane == 0) ok = spin_until_u32(pf, 2 * t + h, y);
            if (!__shfl_sync(FULL, ok, 0)) break;
            if (tr) tr[1] = globaltimer_ns();
            const int mode = ldm(S.mode + p);
            const int nsteps = mode ? ldm(S.mig_n + p) : C;
            double *sm_theta = reinterpret_cast<double *>(wsm + like_smem_bytes(M.n_row, M.n_cell, 32));
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
                int src, tgt;
                double lp;
                propose_position(S, p, k, mode, nsteps, -1, t, 0, h, lane, sm_scratch, sm_theta, src, tgt, lp);
                if (mode != 0 && split == 0) { // a migration sweep is decided later, all of it together: park the proposal
                    double *pr = S.prop + ((size_t)p * C + src) * D;
                    for (int d = lane; d < D; d += 32) pr[d] = sm_theta[d];
                    if (lane == 0) {
                        if (!S.prior_ovr) S.prop_lp[p * C + src] = lp;
                        S.target[p * C + src] = tgt;
                    }
                }
                __syncwarp();
                if (tr) tr[2] = globaltimer_ns();
                const double v = like_eval<NACC, 32>(M, A.T, sm_theta, make_addr(S, p, t, 0, src), p / S.n_rep, split, wsm, nullptr, tr ? tr + 3 : nullptr);
                int decide = 0;
                if (lane == 0) {
                    A.ll_part[((size_t)p * C + src) * nsplit + split] = v;
                    decide = mode == 0;
                    if (decide && nsplit > 1) { // the worker that finishes the proposal's last chunk decides
                        __threadfence();
                        unsigned int *ca = y.chain_arrive + (size_t)p * C + src;
                        decide = atomicAdd(ca, 1u) == (unsigned int)(nsplit - 1);
                        if (decide) {
                            *ca = 0;
                            __threadfence();
                        }
                    }
                    // half 0 of a hierarchy: the prior of theta is this iteration's phi -- decide now if it is there already
                    if (decide && A.hier && h == 0 && ld_acquire_u32(y.phi_done) < 2 * t + 2) decide = 2;
                }
                decide = __shfl_sync(FULL, decide, 0);
                __syncwarp();
                if (decide == 1) {
                    accept_self(S, p, src, t, 0, sm_theta, lp, A.ll_part, nsplit, v, sm_scratch, lane);
                } else if (decide == 2) { // park it for the ACCEPT item
                    double *pr = S.prop + ((size_t)p * C + src) * D;
                    for (int d = lane; d < D; d += 32) pr[d] = sm_theta[d];
                    if (lane == 0) S.target[p * C + src] = tgt;
                }
                __syncwarp(); // table, theta' and scratch are reused by a second sweep position
            }
            int last = 0;
            if (lane == 0) {
                __threadfence();
                last = atomicAdd(pf + 1, 1u) == per_pop_half - 1;
            }
            // the last item of (p, t, h) closes the half -- unless the ACCEPT items do (half 0 of a hierarchy)
            closer = __shfl_sync(FULL, last, 0) && (h == 1 || !A.hier);
            if (closer && mode != 0) { // migration: every decision of the sweep, now that every proposal is made
                __threadfence();
                accept_warp(S, p, 0, C, t, 0, A.ll_part, nsplit, lane);
                __syncwarp();
            }
        }
        if (closer) {
            if (h == 1) {
                __threadfence();
                store_pop(S, p, t, lane);
                __syncwarp();
                sweep_begin_pop<32>(S, p, t + 1, 0, A.decide_once, -1, sm_keys, ctl);
            }
            __syncwarp();
            if (lane == 0) {
                unsigned int *pf = y.pop_flags + (size_t)p * kPopFlagStride;
                pf[1] = 0;
                __threadfence();
                st_release_u32(pf, 2 * t + h + 1);
                if (h == 1) atomicAdd(y.all_done, 1ull);
            }
        }
        if (tr) tr[5] = globaltimer_ns();
    }

    // leave: the last worker out re-arms the queue for the next launch and publishes the iteration counter
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(y.exit_ctr, 1u) == gridDim.x * (blockDim.x >> 5) - 1) {
            *y.exit_ctr = 0;
            *y.queue = 0ull;
            *A.d_iter = A.t_end;
            __threadfence();
        }
    }
}

} // namespace gg
