// ggdmc_b200 -- the persistent sampler kernel: whole DE-MCMC iterations of the PARALLEL schedule in ONE launch.
//
// Replaces, for the default schedule, the per-half-sweep launch chain  k_sweep_begin -> k_propose -> k_like -> k_accept
// (x 2 half-sweeps x subject groups) + k_phi_half x 2 + k_phi_consts + k_store_advance  = 19 launches per iteration
// (run_hchains body, src/de.cpp:281-381; run_chains body :208-240).  Every WARP of the launch is an independent worker;
// nothing inside the loop is block-wide (no __syncthreads, no idle partner warp while one warp proposes), so an SM's 24
// resident warps sit at 24 different points of their items and the FP64 pipe always finds trial loops to run.
//
// Two sources of work:
//
// (1) the MAIN QUEUE, a plain counter over the SUBJECT items of the launch in the order
//         iteration t: [half 0: population 0 .. npop-1][half 1: population 0 .. npop-1], iteration t + 1: ...
//   SUBJECT item (iteration t, half h, population p, slot, trial chunk)
//       proposal of chain 2 slot + h (crossover; migration: sweep positions slot and slot + nslot in half 0) into shared
//       memory, row table built from it, trial loop, warp reduction (de.cpp:567-613, 615-665 and @hdr/likelihood.h:73-108,
//       272-292).  The worker that finishes a crossover proposal's last trial chunk takes its MH decision on the spot
//       (update_theta, :81-108): within a half-sweep nobody else reads or writes that chain.  Only if the decision needs a
//       phi that is not there yet (half 0 of a hierarchy) the proposal is parked in global memory.  A migration sweep
//       rewrites chains other proposals start from, so its decisions are taken together at the end of the half.
//       The worker that finishes the population's half last closes it: parked and migration decisions if phi is there,
//       after half 1 theta_phi::store (@hdr/theta.h:61-74) and the next iteration's migration decision (get_subchains,
//       :62-78), then the population's `done` flag.  It needs  pop_done[p] >= 2 t + h  (the previous half is closed).
//
// (2) URGENT work, published when it becomes runnable and taken by whichever worker looks first -- every worker looks
//     when it finishes an item and, every couple of microseconds, while it waits for a flag:
//   PHI item  (iteration t, half h, replicate r, phi chain c, subject split)
//       proposal of phi chain c, the constants of the subject prior it would drive, and its share of the two
//       hyper-likelihood sums (de_class::sumloghlike, src/de.cpp:245-270); the worker that finishes the half last sums the
//       splits, exchanges the sums with the peer GPUs through the peer-memory window, takes every MH decision of the half
//       (:397-463, :494-549); after half 1 it stores the thinned sample, draws the next iteration's migration decision,
//       announces phi and hands out the CLOSE items.  Half 0 is published by the worker that closes the last population
//       of the previous iteration, half 1 by the finisher of half 0.
//   CLOSE item (population p, group of 32 chains; hierarchy only)
//       a population whose half 0 was complete before phi was: its parked MH decisions, then the half's `done` flag.
//       (The last SUBJECT worker of a population and the phi finisher settle with one compare-and-swap each who closes
//       the population, so no population is closed twice and none is forgotten.)
// A fit with thousands of populations never parks much: phi is finished long before most of half 0 is handed out.  A fit
// with a few dozen populations per GPU has phi on its critical path (half 0 -> phi -> CLOSE -> half 1): the phi items
// overtake everything that is queued, which a single first-in first-out queue cannot give them.
//
// Every dependency points to work that is earlier in the order  half 0 (t) < PHI (t) < CLOSE (t) < half 1 (t) < half 0 (t+1)
// and a worker that waits for a flag keeps taking urgent work (which is always earlier than what it waits for), so the
// earliest unfinished piece of work can always proceed, whatever the number of resident warps (no co-residency
// assumption).  A wait that exceeds spin_ns raises `abort`; every worker then leaves and the host reports the error.
//
// All state other workers rewrite while the kernel runs is read with ld.global.cg (ldm()); flags are published with
// release semantics and read relaxed from L2 (see "Flag traffic" below).
#pragma once
#include "gg_kernels.cuh"

namespace gg {

constexpr int kPopFlagStride = 32; // ints: every population's flags live in a 128-byte line of their own

struct SamplerSync {
    unsigned int *queue;           // next work item of the launch
    unsigned int *exit_ctr;        // workers that have left the launch (the last one resets the queue)
    unsigned int *pop_flags;       // [npop][kPopFlagStride]: [0] done = 2 t + h + 1 once half h of iteration t is closed,
                                   //                          [1] arrive = finished items of the population's current half,
                                   //                          [2] mode, [3] mig_n of the iteration (copies of Level::mode / mig_n that
                                   //                          arrive with the done flag in ONE 16-byte load),
                                   //                          [4] 1 = half 0 complete and waiting for phi (whoever swaps it back closes)
    unsigned long long *urgent;    // [0] PHI items, [1] CLOSE items: the current batch as ONE word  batch << 48 | n << 24 | taken.
                                   // A worker takes an item with one fetch-and-add of 1 and reads batch, n and its own number
                                   // from the returned word (taken >= n: nothing left); the publisher of the next batch swaps
                                   // the whole word.  PHI batch b of a launch = half b & 1 of iteration t_begin + b / 2.
    unsigned int *close_list;      // [npop * groups]: CLOSE item i of the current batch is (population, chain group) close_list[i]
    unsigned int *chain_arrive;    // [npop][C] finished trial chunks of the chain's current proposal (nsplit > 1)
    unsigned long long *all_done;  // (population, iteration) pairs completed since iteration 1
    unsigned int *phi_arrive;      // finished PHI items of the current half
    unsigned int *phi_done;        // 2 t + h + 1
    int *abort;                    // != 0: a wait timed out (1 peer exchange, 2 local flag); every worker leaves
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
    double *hpart, *hsum;
    uint32_t *d_iter;   // set to t_end by the last worker to leave (the multi-launch kernels read it)
    uint32_t t_begin, t_end; // iterations [t_begin, t_end)
    int hier;           // 1: phi level present (run), 0: independent subjects (run_subject)
    int use_p2p;        // phi sums are exchanged with peer GPUs
    int decide_once;    // run_chains draws the migration decision once per iteration (src/de.cpp:210)
    int stage_bytes, warp_bytes; // shared memory: model tables staged once per CTA, then one region per warp
    unsigned int per_iter;  // SUBJECT items of one iteration = 2 npop nslot nsplit
    // diagnostics (GGDMC_B200_ITEMTRACE): 8 stamps per item of the launch's first trace_cap items --
    // taken, dependency met, proposal made, table built, trial loop done, finished (ns); SM id; item kind
    unsigned long long *trace;
    unsigned long long trace_cap;
};

// Flag traffic.  A PTX acquire (ld.acquire, and __threadfence's acquire half) makes the SM drop its whole L1
// (SASS: CCTL.IVALL) -- with 24 workers per SM taking several flags per item the L1 would never hold anything.  It is not
// needed here: every load of state that other workers rewrite is an ld.global.cg (ldm()), served by L2 where the
// writers' releases have already landed, and every such load is control-dependent on the flag value it follows.  So flags
// are READ relaxed (L2, no invalidate) and WRITTEN with release semantics (SASS: MEMBAR.ALL.GPU, then the store / atomic).
__device__ __forceinline__ unsigned int ld_flag_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_flag_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_flag_v4(const unsigned int *p) // 16 bytes = one access: a snapshot of the four words
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_release() { asm volatile("fence.release.gpu;" ::: "memory"); }
__device__ __forceinline__ void fence_sc() { asm volatile("fence.sc.gpu;" ::: "memory"); }
__device__ __forceinline__ ulonglong2 ld_flag_v2u64(const unsigned long long *p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
constexpr unsigned long long kUrgentMask = 0xffffffull; // 24-bit fields
__device__ __forceinline__ unsigned long long urgent_word(unsigned int batch, unsigned int n) { return ((unsigned long long)batch << 48) | ((unsigned long long)n << 24); }
__device__ __forceinline__ bool urgent_open(unsigned long long w) { return (w & kUrgentMask) < ((w >> 24) & kUrgentMask); }
__device__ __forceinline__ void urgent_publish(unsigned long long *q, unsigned int batch, unsigned int n)
{
    fence_release();
    atomicExch(q, urgent_word(batch, n));
}
__device__ __forceinline__ unsigned int atom_add_release_u32(unsigned int *p, unsigned int v)
{
    unsigned int old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// ONE thread waits until *flag >= need (polls back off to 256 ns); false: aborted / timed out
__device__ __forceinline__ bool spin_until_u32(const unsigned int *flag, unsigned int need, const SamplerSync &y)
{
    if (ld_flag_u32(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int ns = 32;
    for (;;) {
        __nanosleep(ns);
        if (ld_flag_u32(flag) >= need) return true;
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
    if (ld_flag_u64(flag) >= need) return true;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int ns = 32;
    for (;;) {
        __nanosleep(ns);
        if (ld_flag_u64(flag) >= need) return true;
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

// HIER: the fit has a phi level.  A compile-time switch, not a field of the arguments: without a phi level none of the urgent-work code exists,
// and the kernel of a run_subject fit is a third of the size -- its workers spend their time on short items spread over all
// phases, so instruction fetch (ncu: stall_no_instruction) is what they wait for most after memory.
template <int NACC, bool HIER>
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

    const unsigned int n_phi = HIER ? (unsigned int)(P.npop * C * A.H.nsplit) : 0u; // PHI items of one half
    const int n_grp = (C + 31) / 32;                                                  // CLOSE items per population
    const unsigned int total = A.per_iter * (A.t_end - A.t_begin);
    const unsigned int per_pop_half = (unsigned int)(nslot * nsplit);
    const unsigned int half_items = A.per_iter / 2;
    unsigned long long idle_since = 0;
    bool dead = false; // a wait timed out or a peer is gone: leave

    auto trace_slot = [&](bool urgent, unsigned int item) -> unsigned long long * {
        if (!A.trace || lane != 0) return nullptr;
        unsigned long long idx = item;
        if (urgent) idx = A.trace_cap / 2 + atomicAdd(A.trace + 8 * A.trace_cap, 1ull);
        else if (idx >= A.trace_cap / 2) return nullptr;
        if (idx >= A.trace_cap) return nullptr;
        unsigned long long *tr = A.trace + 8 * idx;
        unsigned int sm, wslot;
        asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
        asm volatile("mov.u32 %0, %warpid;" : "=r"(wslot));
        tr[0] = globaltimer_ns();
        tr[1] = tr[2] = tr[3] = tr[4] = 0;
        tr[6] = sm | ((unsigned long long)wslot << 16);
        return tr;
    };

    // the population's half is complete and decided: after half 1 store and draw the next sweep decision; raise the flag
    auto close_half = [&](int p, uint32_t t, int h) {
        unsigned int *pf = y.pop_flags + (size_t)p * kPopFlagStride;
        if (h == 1) {
            fence_release();
            store_pop(S, p, t, lane);
            __syncwarp();
            sweep_begin_pop<32>(S, p, t + 1, 0, A.decide_once, -1, sm_keys, ctl);
            __syncwarp();
            if (lane == 0) { // the next iteration's sweep decision, next to the flag that announces it
                pf[2] = (unsigned int)ctl[0];
                pf[3] = ctl[0] == 1 ? (unsigned int)ctl[1] : 0u;
            }
        }
        __syncwarp();
        if (lane == 0) {
            pf[1] = 0;
            pf[4] = 0;
            st_release_u32(pf, 2 * t + h + 1);
            if (h == 1) {
                unsigned long long old;
                asm volatile("atom.add.release.gpu.global.u64 %0, [%1], 1;" : "=l"(old) : "l"(y.all_done) : "memory");
                // the last population of the iteration: the next iteration's phi step can start
                if (HIER && old + 1 == (unsigned long long)S.npop * t && t + 1 < A.t_end) urgent_publish(y.urgent, 2 * (t + 1 - A.t_begin), n_phi);
            }
        }
    };

    // ---- urgent work ---------------------------------------------------------------------------------------------------
    auto run_phi = [&](unsigned int batch, unsigned int j) {
        const uint32_t t = A.t_begin + batch / 2;
        const int h = (int)(batch & 1u);
        unsigned long long *tr = trace_slot(true, 0);
        if (tr) { tr[7] = (unsigned long long)(2 + h) | ((unsigned long long)t << 8); tr[1] = tr[0]; }
        const int Hs = A.H.nsplit;
        const int split = (int)(j % Hs), rc = (int)(j / Hs);
        const int r = rc / C, c = rc - r * C;
        phi_half_part<32>(P, A.H, t, 0, h, r, c, split, A.hpart, reinterpret_cast<double *>(wsm), ctl);
        __syncwarp();
        if (tr) tr[4] = globaltimer_ns();
        int last = 0;
        if (lane == 0) last = atom_add_release_u32(y.phi_arrive, 1u) == n_phi - 1;
        if (__shfl_sync(FULL, last, 0)) {
            if (!phi_half_finish<32>(P, A.H, t, 0, A.hpart, A.hsum, A.w, A.use_p2p)) {
                if (lane == 0) atomicCAS(y.abort, 0, 1);
                dead = true;
                return;
            }
            __syncwarp();
            if (h == 0) {
                if (lane == 0) {
                    *y.phi_arrive = 0;
                    st_release_u32(y.phi_done, 2 * t + 1);
                    urgent_publish(y.urgent, batch + 1, n_phi); // half 1 may start
                }
            } else {
                fence_release();
                for (int r2 = 0; r2 < P.npop; ++r2) store_pop(P, r2, t, lane);
                for (int r2 = 0; r2 < P.npop; ++r2) {
                    __syncwarp();
                    sweep_begin_pop<32>(P, r2, t + 1, 0, 0, -1, sm_keys, ctl);
                }
                __syncwarp();
                if (lane == 0) {
                    *y.phi_arrive = 0;
                    st_release_u32(y.phi_done, 2 * t + 2);
                }
                // Populations whose half 0 was complete before this moment wait for somebody to close them: hand them out as
                // CLOSE items.  (Store phi_done, full fence, read the wait words -- the last SUBJECT worker of a population
                // stores its wait word, full fence, reads phi_done: at least one of the two sees the other.)
                fence_sc();
                unsigned int n_close = 0;
                for (int p0 = 0; p0 < S.npop; p0 += 32) {
                    const int p = p0 + lane;
                    bool mine = false;
                    if (p < S.npop) {
                        unsigned int *ws = y.pop_flags + (size_t)p * kPopFlagStride + 4;
                        if (ld_flag_u32(ws) == 1u) mine = atomicCAS(ws, 1u, 2u) == 1u;
                    }
                    const unsigned int m = __ballot_sync(FULL, mine);
                    if (mine) {
                        const unsigned int pos = n_close + __popc(m & ((1u << lane) - 1u));
                        for (int g = 0; g < n_grp; ++g) y.close_list[pos * n_grp + g] = (unsigned int)(p * n_grp + g);
                    }
                    n_close += __popc(m);
                }
                __syncwarp();
                if (lane == 0 && n_close) urgent_publish(y.urgent + 1, batch, n_close * n_grp);
            }
        }
        if (tr) tr[5] = globaltimer_ns();
    };

    auto run_close = [&](unsigned int ci) {
        unsigned int e = 0, done = 0;
        if (lane == 0) e = ldm(y.close_list + ci);
        e = __shfl_sync(FULL, e, 0);
        const int p = (int)(e / n_grp), g = (int)(e - (unsigned int)p * n_grp);
        unsigned int *pf = y.pop_flags + (size_t)p * kPopFlagStride;
        if (lane == 0) done = ld_flag_u32(pf);
        const uint32_t t = __shfl_sync(FULL, done, 0) / 2; // done = 2 t while half 0 of iteration t is open
        unsigned long long *tr = trace_slot(true, 0);
        if (tr) { tr[7] = 4ull | ((unsigned long long)t << 8); tr[1] = tr[0]; }
        accept_warp(S, p, 32 * g, min(C, 32 * g + 32), t, 0, A.ll_part, nsplit, lane);
        __syncwarp();
        int last = 0;
        if (lane == 0) last = atom_add_release_u32(pf + 1, 1u) == per_pop_half + (unsigned int)n_grp - 1;
        if (__shfl_sync(FULL, last, 0)) close_half(p, t, 0);
        if (tr) tr[5] = globaltimer_ns();
    };

    // take one piece of urgent work if there is any; uq: the two queue words as lane 0 just read them
    auto serve_urgent = [&](ulonglong2 uq) -> bool {
        int kind = 0;
        unsigned long long got = 0;
        if (lane == 0) {
            for (int q = 0; q < 2 && !kind; ++q)
                if (urgent_open(q == 0 ? uq.x : uq.y)) {
                    got = atomicAdd(y.urgent + q, 1ull);
                    if (urgent_open(got)) kind = q + 1;
                }
        }
        kind = __shfl_sync(FULL, kind, 0);
        if (!kind) return false;
        got = __shfl_sync(FULL, got, 0);
        if (kind == 1) run_phi((unsigned int)(got >> 48), (unsigned int)(got & kUrgentMask));
        else run_close((unsigned int)(got & kUrgentMask));
        __syncwarp();
        return true;
    };
    auto peek_urgent = [&]() -> ulonglong2 {
        ulonglong2 uq = make_ulonglong2(0ull, 0ull);
        if (lane == 0) uq = ld_flag_v2u64(y.urgent);
        return uq;
    };
    // lane 0: give up?  (another worker raised abort, or this wait is older than spin_ns)
    auto hopeless = [&](unsigned long long since) -> bool {
        if (*(volatile int *)y.abort) return true;
        if (globaltimer_ns() - since > y.spin_ns) {
            atomicCAS(y.abort, 0, 2);
            return true;
        }
        return false;
    };

    unsigned int next_item = 0;
    bool have_next = false;            // a SUBJECT item takes its successor in the same round trip as its arrival ...
    ulonglong2 uq = make_ulonglong2(0ull, 0ull);
    bool have_uq = false;              // ... and a look at the urgent queues
    while (!dead) {
        if (HIER) {
            if (!have_uq) uq = peek_urgent();
            have_uq = false;
            if (serve_urgent(uq)) continue;
        }
        unsigned int item = next_item;
        if (!have_next) {
            if (lane == 0) {
                item = atomicAdd(y.queue, 1u);
                if (*(volatile int *)y.abort) item = 0xffffffffu;
            }
            item = __shfl_sync(FULL, item, 0);
        }
        have_next = false;
        if (item >= total) {
            if (!HIER || item == 0xffffffffu) break;
            // no SUBJECT items left: stay for the urgent work of the last iteration until every population has finished it
            int fin = 0;
            if (lane == 0) {
                fin = ld_flag_u64(y.all_done) >= (unsigned long long)S.npop * (A.t_end - 1);
                if (!fin) {
                    if (!idle_since) idle_since = globaltimer_ns();
                    __nanosleep(1000);
                    if (hopeless(idle_since)) fin = 2;
                }
            }
            fin = __shfl_sync(FULL, fin, 0);
            if (fin) break;
            next_item = item;
            have_next = true;
            continue;
        }
        // ---------------------------------------------------------------------- SUBJECT item
        unsigned long long *tr = trace_slot(false, item);
        const uint32_t t = A.t_begin + item / A.per_iter;
        unsigned int j = item % A.per_iter;
        const int h = j >= half_items ? 1 : 0;
        j -= h * half_items;
        if (tr) tr[7] = (unsigned long long)h | ((unsigned long long)t << 8);
        const int split = (int)(j % nsplit);
        const int ps = (int)(j / nsplit);
        const int p = ps / nslot;
        const int slot = ps - p * nslot;
        unsigned int *pf = y.pop_flags + (size_t)p * kPopFlagStride;
        // one round trip: the population's flags with this iteration's sweep decision (lane 0), and whether phi is there (lane 1)
        uint4 f4 = make_uint4(0u, 0u, 0u, 0u);
        unsigned int phi_seen = 0;
        if (lane == 0) f4 = ld_flag_v4(pf);
        else if (lane == 1 && HIER) phi_seen = ld_flag_u32(y.phi_done);
        if (!__shfl_sync(FULL, f4.x >= 2 * t + h, 0)) {
            // the population's previous half is still open: wait for it, and look for urgent work every eighth poll
            const unsigned long long since = globaltimer_ns();
            unsigned int ns = 32;
            for (unsigned int poll = 1;; ++poll) {
                int st = 0; // 1 ready, 2 give up
                if (lane == 0) {
                    __nanosleep(ns);
                    if (ns < 256) ns += ns;
                    f4 = ld_flag_v4(pf);
                    st = f4.x >= 2 * t + h ? 1 : (hopeless(since) ? 2 : 0);
                }
                st = __shfl_sync(FULL, st, 0);
                if (st == 2) dead = true;
                if (st) break;
                if (HIER && (poll & 7u) == 0u && serve_urgent(peek_urgent()) && dead) break;
            }
            if (dead) break;
            if (lane == 1 && HIER) phi_seen = ld_flag_u32(y.phi_done);
        }
        if (tr) tr[1] = globaltimer_ns();
        const int mode = (int)__shfl_sync(FULL, f4.z, 0);
        const int nsteps = mode ? (int)__shfl_sync(FULL, f4.w, 0) : C;
        const bool phi_ready = !HIER || h == 1 || __shfl_sync(FULL, phi_seen, 1) >= 2 * t + 2;
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
                    unsigned int *ca = y.chain_arrive + (size_t)p * C + src;
                    decide = atom_add_release_u32(ca, 1u) == (unsigned int)(nsplit - 1);
                    if (decide) *ca = 0;
                }
                // half 0 of a hierarchy: the prior of theta is this iteration's phi -- decide now if it is there already
                if (decide && !phi_ready && ld_flag_u32(y.phi_done) < 2 * t + 2) decide = 2;
            }
            decide = __shfl_sync(FULL, decide, 0);
            __syncwarp();
            if (decide == 1) {
                accept_self(S, p, src, t, 0, sm_theta, lp, A.ll_part, nsplit, v, sm_scratch, lane);
            } else if (decide == 2) { // park it: whoever closes the half decides it
                double *pr = S.prop + ((size_t)p * C + src) * D;
                for (int d = lane; d < D; d += 32) pr[d] = sm_theta[d];
                if (lane == 0) S.target[p * C + src] = tgt;
            }
            __syncwarp(); // table, theta' and scratch are reused by a second sweep position
        }
        // one round trip: arrival (lane 0), the next item of the queue (lane 1), a look at the urgent queues (lane 2)
        unsigned int got = 0;
        if (lane < 2) got = atom_add_release_u32(lane == 0 ? pf + 1 : y.queue, 1u);
        ulonglong2 uq2 = make_ulonglong2(0ull, 0ull);
        if (lane == 2 && HIER) uq2 = ld_flag_v2u64(y.urgent);
        next_item = __shfl_sync(FULL, got, 1);
        if (*(volatile int *)y.abort) next_item = 0xffffffffu;
        have_next = true;
        uq.x = __shfl_sync(FULL, uq2.x, 2);
        uq.y = __shfl_sync(FULL, uq2.y, 2);
        have_uq = true;
        if (__shfl_sync(FULL, got, 0) == per_pop_half - 1) { // the population's half is complete
            if (h == 1 || !HIER) {
                if (mode != 0) { // migration: every decision of the sweep, now that every proposal is made
                    accept_warp(S, p, 0, C, t, 0, A.ll_part, nsplit, lane);
                    __syncwarp();
                }
                close_half(p, t, h);
            } else {
                // half 0 of a hierarchy: parked decisions need phi.  Wait word := 1, full fence, read phi_done; the phi
                // finisher does the mirror image, so at least one of us sees the other, and the swap decides who closes.
                int mine = 0;
                if (lane == 0) {
                    atomicExch(pf + 4, 1u);
                    fence_sc();
                    if (ld_flag_u32(y.phi_done) >= 2 * t + 2) mine = atomicCAS(pf + 4, 1u, 0u) == 1u;
                }
                if (__shfl_sync(FULL, mine, 0)) {
                    accept_warp(S, p, 0, C, t, 0, A.ll_part, nsplit, lane);
                    __syncwarp();
                    close_half(p, t, 0);
                }
            }
        }
        if (tr) tr[5] = globaltimer_ns();
    }

    // leave: the last worker out re-arms the queues for the next launch and publishes the iteration counter
    __syncwarp();
    if (lane == 0) {
        if (atom_add_release_u32(y.exit_ctr, 1u) == gridDim.x * (blockDim.x >> 5) - 1) {
            *y.exit_ctr = 0;
            *y.queue = 0u;
            if (HIER) {
                y.urgent[0] = urgent_word(0u, n_phi); // the next launch starts with a phi half 0 that may run
                y.urgent[1] = 0ull;
            }
            *A.d_iter = A.t_end;
        }
    }
}

} // namespace gg
