// ggdmc_b200 -- CUDA kernels of the DE-MCMC step (sm_100a).
//
//   K0 k_sweep_begin  : per population, migration-vs-crossover decision and the migration set
//                       (src/de.cpp:210, 296, 310, 348, 362 and get_subchains :62-78)
//   K1 k_propose      : crossover / migration proposals + their log prior
//                       (src/de.cpp:119-141, 164-184, 394-426, 488-518, 575-603, 627-652)
//   K2 k_like         : LBA sum-log-likelihood of every proposal (the FP64 hot kernel)
//                       (@hdr/likelihood.h:73-108, 272-292 + @hdr/lba.h)
//      k_like_ddm     : the same launch for model type "fastdm": DDM sum-log-likelihood
//                       (@hdr/likelihood.h:129-161, 295-305 + @hdr/ddm.h)
//   K3 k_accept       : Metropolis accept / commit (src/de.cpp:81-108); in a hierarchy also the log prior of the
//                       proposal under its phi chain (:599-604, 646-653)
//   K4 k_phi_half     : one phi half-sweep in one launch: proposal, hyper-likelihood partial sums over the local
//                       subjects (src/de.cpp:245-270), reduction, exchange with the peer GPUs, MH test
//                       (:397-400, 427-463, 494-500, 519-549); k_hyper, k_hyper_reduce(_exchange), k_phi_accept are
//                       the same steps as separate launches (in-place order, NCCL fallback)
//      k_phi_consts   : per phi chain constants of the truncated-normal prior of the subject level
//      k_store_advance: thinned sample storage (@hdr/theta.h:61-74) + the device-side iteration counter
//
// A "population" is one set of nchain chains: a (replicate, subject) pair at the subject level,
// a replicate at the phi level.  `step < 0` processes every chain of the sweep at once from the
// sweep-start state (PARALLEL schedule); `step >= 0` processes only sweep position `step`
// (REFERENCE schedule: the host walks step = 0 .. nchain-1 so chains are updated in place, one
// after another, exactly like the reference's loops).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gg_lba.cuh"
#include "gg_ddm.cuh"
#include "gg_rng.cuh"

namespace gg {

struct DevModel {
    int n_acc, n_cell, npar, n_const;
    const int *param_src;
    const double *const_val;
    const uint8_t *posdrift; // LBA: [n_acc].  DDM: [n_cell], non-zero = upper-boundary response (the host picks the kernels by model type)
    // LBA: the DISTINCT (cell, accumulator) rows of the model.  Entries that draw the same six parameters with the same
    // drift rule get one table row (B x v README model: 24 rows for 48 entries; 4-accumulator factorial model: 48 for 384).
    // A model whose st0 can be non-zero keeps one row per entry, because every entry then has its own `t0 + st0 U` draw.
    int n_row;
    const uint16_t *row_of;  // [n_cell][n_acc] row of entry (cell, accumulator)
    const int *row_src;      // [n_row][8]: sources of A, B, mean_v, sd_v, st0, t0 (>= 0 parameter, < 0 constant -1 - s), entry index
                             // of the row's first entry (addresses its st0 draw), positive-drift flag
};

// A "group" is the set of threads that works on one item together: a whole thread block (G = block size, or 0 = whatever
// the launch says) or ONE WARP of a larger block (G = 32: the persistent sampler kernel, whose warps are independent workers).
template <int G>
struct Grp {
    __device__ __forceinline__ static int tid() { return G == 32 ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
    __device__ __forceinline__ static int size() { return G > 0 ? G : (int)blockDim.x; }
    __device__ __forceinline__ static void sync()
    {
        if constexpr (G == 32) __syncwarp();
        else __syncthreads();
    }
};

struct DevPrior {
    int npar;
    const double *p0, *p1, *lower, *upper;
    const int *dist;
    const uint8_t *log_p;
};

// One level (subject or phi) of the sampler.
struct Level {
    int npop;        // populations on this device at this level
    int nchain;      // C
    int npar;        // length of one chain's vector
    int nmove;       // leading parameters perturbed by an unblocked sweep (src/de.cpp:592: half at the subject level of a hierarchy)
    int n_rep;        // replicates; populations are item-major: p = item * n_rep + replicate
                      // (item = local subject at the subject level, 0 at the phi level)
    int pop_id_base;  // global id of local item 0 (subject_begin), phi: unused
    int is_phi;       // population id is kPopPhi
    double gamma, rp, mig_prob;
    // state
    double *theta, *lp, *ll;          // [npop][C][npar], [npop][C]
    double *prop, *prop_lp;           // proposals
    int *target;                      // [npop][C] chain compared against / overwritten by the proposal of this chain, -1 = none
    int *mode;                        // [npop] 0 crossover, 1 migration
    int *mig_n;                       // [npop]
    int *mig_list;                    // [npop][C] sorted migration set
    int *para;                        // [npop] parameter moved by this sweep, -1 = all (blocked sweeps)
    int *mode0;                       // [npop] decision of sweep 0 (run_chains draws once per iteration, src/de.cpp:210)
    const uint64_t *seed;             // [n_replicate]
    // prior of the moved vector
    DevPrior prior;
    const double *prior_ovr;          // phi state [n_rep][C][2*npar] when the prior's p0/p1 come from phi chain (src/de.cpp:599-600), else null
    const double *ovr_consts;         // [n_rep][C][npar][2] = (1/sd, ln sqrt(2 pi) + ln sd + ln denom) of that phi state (k_phi_consts), or null
    // storage
    double *out_theta, *out_lp, *out_ll; // [npop][nmc][C][npar], [npop][nmc][C]
    int nmc, thin;
};

__device__ __forceinline__ uint32_t pop_global_id(const Level &L, int p)
{
    return L.is_phi ? kPopPhi : (uint32_t)(L.pop_id_base + p / L.n_rep);
}

__device__ __forceinline__ DrawAddr make_addr(const Level &L, int p, uint32_t iter, int sweep, int chain)
{
    DrawAddr a;
    a.seed = L.seed[p % L.n_rep];
    a.pop = pop_global_id(L, p);
    a.iter = iter;
    a.sweep = sweep < 0 ? 0u : (uint32_t)sweep;
    a.chain = (uint32_t)chain;
    return a;
}

// nmath runif(a, b) with the uniform already drawn (a == b handled by the callers' formulas:
// -rp + 2 rp u == -0 + 0 when rp == 0)
__device__ __forceinline__ double runif_from(double lo, double hi, double u)
{
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u));
}

// Load of MUTABLE sampler state (theta, lp, ll, proposals, targets, migration sets, phi constants, partial sums):
// L2 only.  Inside the persistent sampler kernel (gg_sampler.cuh) other CTAs rewrite this state while the kernel runs,
// and an L1 line filled in an earlier iteration would be stale; the multi-launch kernels lose nothing by it.
template <class T>
__device__ __forceinline__ T ldm(const T *p)
{
    return __ldcg(p);
}

// ------------------------------------------------------------------------------------------------
// block-wide sum of one double per thread (warp shuffles, then one shared-memory pass)
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v, double *scratch /* [BLOCK/32], not otherwise in use */)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if constexpr (BLOCK == 32) return v; // one warp: valid in lane 0, no shared memory, no barrier
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < BLOCK / 32; ++i) r += scratch[i];
    }
    return r; // valid in thread 0
}

// ------------------------------------------------------------------------------------------------
// K0: sweep prologue
// ------------------------------------------------------------------------------------------------
// decide_once = 0: a fresh migration/crossover draw for every sweep (run_hchains, src/de.cpp:283-319,
//                  :345-371); para_idx is the host's blocked-parameter index (-1 = all).
// decide_once = 1: run_chains (src/de.cpp:201-242): ONE draw per iteration (made in sweep 0); a
//                  migration is a single unblocked sweep, so migrating populations idle (mode 2)
//                  in the remaining blocked sweeps.
// block-wide; sm_keys: 2 * nchain ints of shared memory.  The caller provides a barrier before sm_keys is reused.
template <int G = 0>
__device__ __forceinline__ void sweep_begin_pop(const Level &L, int p, uint32_t iter, int sweep, int decide_once, int para_idx, int *sm_keys,
                                                int *s_mode_n /* 2 ints of shared memory */, bool clear_targets = true)
{
    const int C = L.nchain, tid = Grp<G>::tid(), nthr = Grp<G>::size();
    DrawAddr a = make_addr(L, p, iter, decide_once ? 0 : sweep, 0);
    if (tid == 0) {
        int mode;
        if (decide_once && sweep > 0) {
            mode = ldm(L.mode0 + p) ? 2 : 0;
        } else {
            double u = draw_uniform(a, U_DECIDE, 0);
            mode = (u < L.mig_prob) ? 1 : 0;
            L.mode0[p] = mode;
        }
        s_mode_n[0] = mode;
        L.mode[p] = mode;
        L.para[p] = (decide_once && mode == 1) ? -1 : para_idx;
        if (mode == 1) { // get_subchains, src/de.cpp:64-70
            double prop = draw_uniform(a, U_MIG_N, 0);
            unsigned n = (unsigned)ceil((double)C * prop);
            n = n < 2u ? 2u : n;
            n = n > (unsigned)C ? (unsigned)C : n;
            s_mode_n[1] = (int)n;
            L.mig_n[p] = (int)n;
        } else if (mode == 2) {
            L.mig_n[p] = 0;
        }
    }
    // no proposal is pending between sweeps (every MH test consumes its own); the launch that runs beside the last half's
    // MH tests must not touch them
    if (clear_targets)
        for (int c = tid; c < C; c += nthr) L.target[p * C + c] = -1;
    Grp<G>::sync();
    if (s_mode_n[0] != 1) return;
    // arma::shuffle keys for all chains, then the n smallest keys (ties by position), sorted by index
    for (int blk = tid; blk * 4 < C; blk += nthr) {
        U4 w = draw_block(a, U_MIG_KEYS, (uint32_t)blk);
        int j = blk * 4;
        sm_keys[j] = shuffle_key(word_to_uniform(w.x));
        if (j + 1 < C) sm_keys[j + 1] = shuffle_key(word_to_uniform(w.y));
        if (j + 2 < C) sm_keys[j + 2] = shuffle_key(word_to_uniform(w.z));
        if (j + 3 < C) sm_keys[j + 3] = shuffle_key(word_to_uniform(w.w));
    }
    Grp<G>::sync();
    const int n = s_mode_n[1];
    int *sm_rank = sm_keys + C;
    for (int j = tid; j < C; j += nthr) {
        const int kj = sm_keys[j];
        int rank = 0;
        for (int k = 0; k < C; ++k) {
            const int kk = sm_keys[k];
            rank += (kk < kj) || (kk == kj && k < j);
        }
        sm_rank[j] = rank;
    }
    Grp<G>::sync();
    for (int j = tid; j < C; j += nthr) {
        if (sm_rank[j] < n) { // selected: position among the selected chains in ascending index order
            int pos = 0;
            for (int k = 0; k < j; ++k) pos += sm_rank[k] < n;
            L.mig_list[p * C + pos] = j;
        }
    }
}

// iter_ofs = 1: the decisions of the NEXT iteration, drawn at the end of this one (off the next iteration's critical path)
// own / own_done (may be null): the launch counts iterations by itself instead of reading the engine's counter -- it then
// may run while the end-of-iteration kernel advances that one.  Every block reads *own before it signals completion; the
// last block to finish increments it.
__global__ void k_sweep_begin(Level L, const uint32_t *d_iter, int sweep, int decide_once, int para_idx, int iter_ofs = 0,
                              uint32_t *own = nullptr, unsigned int *own_done = nullptr)
{
    extern __shared__ int sm_keys[]; // keys [C], ranks [C]
    __shared__ int s_mode_n[2];
    const uint32_t base = own ? *own : *d_iter;
    sweep_begin_pop(L, blockIdx.x, base + (uint32_t)iter_ofs, sweep, decide_once, para_idx, sm_keys, s_mode_n, own == nullptr);
    if (own) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(own_done, 1u) == gridDim.x - 1) {
                *own_done = 0;
                *own = base + 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K1: proposals + log prior of the proposal.  One block per population, one warp per chain.
// ------------------------------------------------------------------------------------------------

// lexicographic (key, pos) minimum across the warp
__device__ __forceinline__ void warp_min_keypos(int &key, int &pos)
{
    // keys and positions are non-negative: two warp-wide integer minima (REDUX) instead of five shuffle rounds
    const int mk = __reduce_min_sync(0xffffffffu, key);
    const int mp = __reduce_min_sync(0xffffffffu, key == mk ? pos : 0x7fffffff);
    key = mk;
    pos = mp;
}

// get_chains(i, 2), src/de.cpp:54-60: the two smallest shuffle keys among the candidate chains.
// half < 0: candidates = every chain but i (the reference's rule); half = 0 / 1: candidates = the
// chains of the OTHER parity (two-half PARALLEL schedule).
// `pre`: this lane's partner block (index = lane) when the caller drew partner and noise blocks in one pass, else null
__device__ __forceinline__ void pick_partners(const DrawAddr &a, int C, int i, int half, int lane, int &c0, int &c1, const U4 *pre = nullptr)
{
    const int ncand = half < 0 ? C - 1 : (C + half) / 2;
    const int INTMAX = 0x7fffffff;
    int k1 = INTMAX, p1 = INTMAX, k2 = INTMAX, p2 = INTMAX; // lane-local best two
    for (int blk = lane; blk * 4 < ncand; blk += 32) {
        U4 w = pre ? *pre : draw_block(a, U_PARTNER, (uint32_t)blk);
        uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int j = blk * 4 + q;
            if (j < ncand) {
                int k = shuffle_key(word_to_uniform(ws[q]));
                if (k < k1 || (k == k1 && j < p1)) { k2 = k1; p2 = p1; k1 = k; p1 = j; }
                else if (k < k2 || (k == k2 && j < p2)) { k2 = k; p2 = j; }
            }
        }
    }
    int bk = k1, bp = p1;
    warp_min_keypos(bk, bp);
    // second: each lane offers its best candidate that is not the global minimum
    int sk = (k1 == bk && p1 == bp) ? k2 : k1;
    int sp = (k1 == bk && p1 == bp) ? p2 : p1;
    warp_min_keypos(sk, sp);
    if (half < 0) {
        c0 = bp + (bp >= i);
        c1 = sp + (sp >= i);
    } else {
        c0 = 2 * bp + (1 - half);
        c1 = 2 * sp + (1 - half);
    }
}

// prior_class::sumlogprior with arma::accu's two-accumulator order (@hdr/prior.h:469-476)
__device__ __forceinline__ double sum_arma_order(const double *v, int n)
{
    double a1 = 0.0, a2 = 0.0;
    int i = 0;
    for (; i + 1 < n; i += 2) { a1 += v[i]; a2 += v[i + 1]; }
    if (i < n) a1 += v[i];
    return a1 + a2;
}

// Per (replicate, phi chain, parameter) constants of the phi-driven truncated-normal prior, so the
// subject level does not redo 2 Phi + 2 log per parameter per subject (the reference re-runs
// tnorm_class::set_parameters for every element, @hdr/prior.h:349).  NaN marks "use the generic path".
__global__ void k_phi_consts(Level P, DevPrior like, int D, double *consts)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = P.npop * P.nchain * D;
    if (i >= n) return;
    const int d = i % D, rc = i / D;
    const double *phi = P.theta + (size_t)rc * 2 * D;
    const double m = phi[d], sd = phi[D + d];
    double inv = 0.0, K = NAN;
    if (like.dist[d] == 1 && like.log_p[d] != 0 && sd > 0.0 && isfinite(sd) && isfinite(m)) {
        const double den = pnorm5(like.upper[d], m, sd, true) - pnorm5(like.lower[d], m, sd, true);
        inv = 1.0 / sd;
        K = kLnSqrt2Pi + log(sd) + log(den);
    }
    consts[2 * (size_t)i] = inv;
    consts[2 * (size_t)i + 1] = K;
}

// log prior density of one proposed parameter; ovr/oc: the phi chain that drives the prior, or null
__device__ __forceinline__ double prior_term(const Level &L, int d, double x, const double *ovr, const double *oc)
{
    const int D = L.npar;
    const double lo = L.prior.lower[d], up = L.prior.upper[d];
    const double K = oc ? ldm(oc + 2 * d + 1) : NAN;
    if (K == K) { // phi-driven truncated normal, log scale: -(ln sqrt(2 pi) + z^2/2 + ln sd) - ln denom
        const double z = (x - ldm(ovr + d)) * ldm(oc + 2 * d);
        return (x < lo || x > up) ? -INFINITY : -(0.5 * z * z + K);
    }
    const double q0 = ovr ? ldm(ovr + d) : L.prior.p0[d], q1 = ovr ? ldm(ovr + D + d) : L.prior.p1[d];
    return dprior1(L.prior.dist[d], x, q0, q1, lo, up, L.prior.log_p[d] != 0);
}

// prior_class::sumlogprior of the proposal made from chain src under phi chain src (src/de.cpp:599-604, 646-653)
__device__ __forceinline__ double deferred_prop_lp(const Level &L, int p, int src)
{
    const int C = L.nchain, D = L.npar;
    const size_t rc = (size_t)(p % L.n_rep) * C + src;
    const double *ovr = L.prior_ovr + rc * 2 * D;
    const double *oc = L.ovr_consts ? L.ovr_consts + rc * 2 * D : nullptr;
    const double *pr = L.prop + ((size_t)p * C + src) * D;
    double a1 = 0.0, a2 = 0.0; // arma::accu order, as sum_arma_order
    int d = 0;
    if (oc) { // the usual case -- every term a phi-driven truncated normal: a pass without calls, whose loads can travel together
        bool all_fast = true;
#pragma unroll 4
        for (int e = 0; e < D; ++e) {
            const double x = ldm(pr + e), K = ldm(oc + 2 * e + 1), m = ldm(ovr + e), iv = ldm(oc + 2 * e);
            const double z = (x - m) * iv;
            const double term = (x < L.prior.lower[e] || x > L.prior.upper[e]) ? -INFINITY : -(0.5 * z * z + K);
            all_fast = all_fast && (K == K);
            if (e & 1) a2 += term; else a1 += term;
        }
        if (all_fast) return a1 + a2;
        a1 = a2 = 0.0;
    }
    for (; d + 1 < D; d += 2) {
        a1 += prior_term(L, d, ldm(pr + d), ovr, oc);
        a2 += prior_term(L, d + 1, ldm(pr + d + 1), ovr, oc);
    }
    if (d < D) a1 += prior_term(L, d, ldm(pr + d), ovr, oc);
    return a1 + a2;
}

// The proposal at sweep position k of population p, made by one warp (src/de.cpp:111-141 crossover,
// :157-185 migration, and their phi / hierarchy twins).  `out` = where the proposed vector goes (null: its row
// of L.prop).  src / tgt: the chain proposed from and the chain it is compared with; lp (lane 0): log prior of
// the proposal unless the prior is phi-driven (then k_accept evaluates it).
__device__ __forceinline__ void propose_position(const Level &L, int p, int k, int mode, int nsteps, int para_idx, uint32_t iter,
                                                 int sweep, int half, int lane, double *scratch, double *out, int &src, int &tgt,
                                                 double &lp)
{
    const int C = L.nchain, D = L.npar;
    int c0 = 0, c1 = 0;
    if (mode) {
        src = ldm(L.mig_list + p * C + k);
        tgt = ldm(L.mig_list + p * C + ((k + 1 == nsteps) ? 0 : k + 1));
    } else {
        src = k;
        tgt = k;
    }
    DrawAddr a = make_addr(L, p, iter, sweep, src);
    // One Philox pass per warp when everything fits: lanes [0, nb) draw the partner-key blocks, lanes [nb, nb + nn) the
    // noise blocks (a warp instruction costs the same with 10 or with 14 active lanes); lane d then fetches its noise
    // word by shuffle.  Same addressed draws as the two-pass form below, which stays for populations of > 100 chains.
    const int nb = mode ? 0 : (((half < 0 ? C - 1 : (C + half) / 2) + 3) >> 2), nn = (D + 3) >> 2;
    const bool one_pass = nb + nn <= 32 && D <= 32;
    U4 wb = {0u, 0u, 0u, 0u};
    if (one_pass && lane < nb + nn) // purpose and block index are data, so both kinds of lanes run the same Philox code together
        wb = draw_block(a, lane < nb ? U_PARTNER : U_NOISE, (uint32_t)(lane < nb ? lane : lane - nb));
    if (!mode) pick_partners(a, C, src, half, lane, c0, c1, one_pass ? &wb : nullptr);
    const double *th = L.theta + ((size_t)p * C + src) * D;
    const double *t0 = L.theta + ((size_t)p * C + c0) * D;
    const double *t1 = L.theta + ((size_t)p * C + c1) * D;
    double *pr = out ? out : L.prop + ((size_t)p * C + src) * D;
    // a phi-driven prior is evaluated in k_accept instead: the proposal and its likelihood do not need
    // this iteration's phi, so they can run while the phi sweep is still in flight
    const bool defer = L.prior_ovr != nullptr;
    uint32_t my_word = 0;
    if (one_pass) { // lane d needs word d % 4 of noise block d / 4, held by lane nb + d / 4
        const int from = nb + (lane >> 2);
        const uint32_t x = __shfl_sync(0xffffffffu, wb.x, from & 31), y = __shfl_sync(0xffffffffu, wb.y, from & 31);
        const uint32_t z = __shfl_sync(0xffffffffu, wb.z, from & 31), w4 = __shfl_sync(0xffffffffu, wb.w, from & 31);
        const int q = lane & 3;
        my_word = q == 0 ? x : (q == 1 ? y : (q == 2 ? z : w4));
    }
    for (int d = lane; d < D; d += 32) {
        double x = ldm(th + d);
        const bool moved = para_idx >= 0 ? (d == para_idx) : (d < L.nmove);
        if (moved) {
            double u = one_pass ? word_to_uniform(my_word) : draw_uniform(a, U_NOISE, (uint32_t)d);
            double noise = runif_from(-L.rp, L.rp, u);
            double inc = mode ? noise : __dadd_rn(noise, __dmul_rn(L.gamma, __dsub_rn(ldm(t0 + d), ldm(t1 + d))));
            x = __dadd_rn(x, inc);
        }
        pr[d] = x;
        if (!defer) scratch[d] = prior_term(L, d, x, nullptr, nullptr);
    }
    __syncwarp();
    lp = 0.0;
    if (lane == 0 && !defer) lp = sum_arma_order(scratch, D);
    __syncwarp();
}

// One warp per (population, sweep position).  MINB: resident blocks per SM the register budget is cut for -- 10 (48 registers,
// ~120 B of spills) when the launch is several waves of warps, so that more of its latency chains are in flight at once
// (1 024 subjects: 26 -> 13 us); 0 (no cap) for small launches, where the spills would only add latency.
template <int WARPS, int MINB = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_propose(Level L, const uint32_t *d_iter, int sweep, int step, int half)
{
    extern __shared__ double sm_prop[]; // [WARPS][npar] prior terms scratch
    const int C = L.nchain, D = L.npar;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS + w;
    int p, k, k2 = -1; // sweep positions this warp handles (k2: second one for migrating populations in half mode)
    if (step >= 0) {
        p = g;
        k = step;
    } else if (half < 0) {
        p = g / C;
        k = g - p * C;
    } else { // half-sweep: (C + 1) / 2 slots per population; crossover -> chain 2 slot + half, migration (half 0) -> slot, slot + nslot
        const int nslot = (C + 1) / 2;
        p = g / nslot;
        k = g - p * nslot;
        if (p < L.npop) {
            if (L.mode[p] == 0) k = 2 * k + half;
            else if (half == 0) k2 = k + nslot;
            else return;
        }
    }
    if (p >= L.npop) return;
    const uint32_t iter = *d_iter;
    const int mode = L.mode[p];
    const int para_idx = L.para[p];
    const int nsteps = mode ? L.mig_n[p] : C;
    for (; k >= 0; k = k2, k2 = -1) {
        if (k >= nsteps) continue;
        int src, tgt;
        double lp;
        propose_position(L, p, k, mode, nsteps, para_idx, iter, sweep, half, lane, sm_prop + w * D, nullptr, src, tgt, lp);
        if (lane == 0) {
            if (!L.prior_ovr) L.prop_lp[p * C + src] = lp;
            L.target[p * C + src] = tgt;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// K2: LBA sum-log-likelihood
// ------------------------------------------------------------------------------------------------
struct TrialData {
    const double *rt;        // [ntot_padded] grouped by cell within subject
    const uint16_t *cell;    // [ntot_padded]; 0xFFFF = padding
    const int64_t *offset;   // [S] start of each subject (multiple of 8)
    const int *count;        // [S] trials of each subject
    int chunk;               // trials per block (multiple of 8)
    int nsplit;              // blocks per (population, chain)
    unsigned long long *counter; // trial-likelihoods evaluated so far (one atomicAdd per block)
    double zero_floor;           // > 0: densities <= 0 are replaced by this value (R-side init rule, R/phi.R:3-13); 0: off
    unsigned long long *btrace;  // diagnostics (GGDMC_B200_BLOCKTRACE): per block {start ns, end ns, SM id}, normally null
};

struct BlockTrace { // thread 0 of a block stamps start, end, SM id, end of table build, end of trial loop into btrace[5 * block]
    unsigned long long *slot;
    __device__ __forceinline__ static unsigned long long now()
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        return t;
    }
    __device__ __forceinline__ explicit BlockTrace(unsigned long long *base) : slot(nullptr)
    {
        if (base && threadIdx.x == 0) {
            slot = base + 5 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
            unsigned int sm;
            asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
            slot[0] = now();
            slot[2] = sm;
        }
    }
    __device__ __forceinline__ ~BlockTrace()
    {
        if (slot) slot[1] = now();
    }
    __device__ __forceinline__ void mark(int i) const
    {
        if (slot) slot[i] = now();
    }
};

constexpr double kLn2Hi = 6.93147180369123816490e-01, kLn2Lo = 1.90821492927058770002e-10;

// running product of densities as (mantissa in [1,2), exponent) -- replaces one log() per trial
// (@hdr/likelihood.h:284-288 sums log(density) trial by trial) by one multiply + integer work.
struct LogProd {
    double m;
    int e;
    double extra; // log of factors that are 0, subnormal, inf or NaN (rare)
    __device__ __forceinline__ void init() { m = 1.0; e = 0; extra = 0.0; }
    // x >= 0 (fast-path densities): a zero factor makes the product 0.  A factor that is not a positive normal number after
    // all -- inf or NaN from an overflowing intermediate (dt or sd_v near 1e-290), a negative value -- must not be folded
    // in as if its bits were a mantissa and an exponent: it goes through log() like in mul(), so the sum becomes inf / NaN
    // and the MH test rejects it like the reference's NaN ratio (src/de.cpp:83-87).  One unsigned compare covers all.
    __device__ __forceinline__ void mul_fast(double x)
    {
        const long long b = __double_as_longlong(x);
        const int ex = (int)(b >> 52);
        if ((unsigned)(ex - 1) >= 0x7feu) {
            extra += (b << 1) == 0 ? -INFINITY : log(x); // zero (either sign) : subnormal, inf, NaN, negative
        } else {
            m *= __longlong_as_double((b & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll);
            const long long mb = __double_as_longlong(m);
            e += ex - 1023 + (int)(mb >> 52) - 1023;
            m = __longlong_as_double((mb & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll);
        }
    }
    // any x
    __device__ __forceinline__ void mul(double x)
    {
        const long long b = __double_as_longlong(x);
        const int ex = (int)((b >> 52) & 0x7ff);
        if ((unsigned)(ex - 1) < 0x7feu && b > 0) mul_fast(x);
        else extra += log(x);
    }
    __device__ __forceinline__ double value() const { return fma((double)e, kLn2Hi, fma((double)e, kLn2Lo, log(m))) + extra; }
};

// Shared memory of one likelihood evaluation by a group of `group` threads (bytes): the row table, the reduction scratch
// of block_sum, the class of every cell and of every row.
__host__ __device__ inline size_t like_smem_bytes(int n_row, int n_cell, int group)
{
    size_t b = (size_t)n_row * sizeof(CellAcc) + (size_t)(group / 32) * 8 + (size_t)n_cell + (size_t)n_row;
    return (b + 15) & ~(size_t)15;
}

// Builds the row table of one parameter vector in shared memory: thread k fills row k -- b = A + B, t0 + st0 U,
// max(Phi(mean_v / sd_v), 1e-10) and the reciprocals the trial loop multiplies by -- and classifies it
// (@hdr/lba.h:121-146: a cell is invalid when ANY accumulator has A, b, sd_v, st0 or t0 < 0 or b < A); then one thread per
// cell combines the classes of the cell's rows.  Two group barriers.
template <int G>
__device__ __forceinline__ void build_rows(const DevModel &M, const double *theta, CellAcc *rows, uint8_t *cell_bad, uint8_t *row_cls,
                                           const DrawAddr &addr)
{
    const int na = M.n_acc, tid = Grp<G>::tid();
    for (int r = tid; r < M.n_row; r += Grp<G>::size()) {
        const int *src = M.row_src + 8 * r;
        double v[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int s = src[i];
            v[i] = s >= 0 ? theta[s] : M.const_val[-1 - s];
        }
        double u = 0.0;
        if (v[4] != 0.0) u = draw_uniform(addr, U_ST0, (uint32_t)src[6]);
        cellacc_build(rows[r], v[0], v[1], v[2], v[3], v[4], v[5], src[7] != 0, u);
        row_cls[r] = cell_class_update(kCellRegular, v[0], v[1], v[2], v[3], v[4], v[5]);
    }
    Grp<G>::sync();
    for (int c = tid; c < M.n_cell; c += Grp<G>::size()) { // a cell is invalid if any row is, else generic if any row is
        uint8_t cls = kCellRegular;
        for (int j = 0; j < na; ++j) {
            const uint8_t r = row_cls[M.row_of[c * na + j]];
            cls = (r == kCellInvalid || cls == kCellInvalid) ? (uint8_t)kCellInvalid : (r != kCellRegular ? r : cls);
        }
        cell_bad[c] = cls;
    }
    Grp<G>::sync();
}

// The trials the hot loops of like_eval leave behind -- invalid / generic cells, rt <= t0, a last odd trial -- with the
// reference's full branch structure.  Not inlined: rare, and long enough to matter for the instruction cache of the
// persistent kernel.  PAIRS: the hot loop worked on pairs of trials (a pair is left behind as a whole).
template <int NACC, int G, bool PAIRS, bool TRACE>
__device__ __noinline__ void like_cold(const CellAcc *rows, const uint16_t *row_of, const uint8_t *bad, int na, const double *rt,
                                       const uint16_t *cl, int t_begin, int t_end, double zf, LogProd &acc, double *trace_out)
{
    const int tid = Grp<G>::tid();
    for (int t = t_begin + 2 * tid; t < t_end; t += 2 * G) {
        const int nh = (t + 1 < t_end) ? 2 : 1;
        if (PAIRS && nh == 2) {
            const int c0 = cl[t], c1 = cl[t + 1];
            if (bad[c0] == kCellRegular && bad[c1] == kCellRegular && n1pdf_fast_ok<NACC>(rt[t], RowRef{rows, row_of + c0 * na}, na) &&
                n1pdf_fast_ok<NACC>(rt[t + 1], RowRef{rows, row_of + c1 * na}, na))
                continue; // done in the hot loop
        }
        for (int h = 0; h < nh; ++h) {
            const int c = cl[t + h];
            const double r = rt[t + h];
            const RowRef e{rows, row_of + c * na};
            const uint8_t cls = bad[c];
            if (!PAIRS && cls == kCellRegular && n1pdf_fast_ok<NACC>(r, e, na)) continue; // done in the hot loop
            double pdf = cls == kCellInvalid ? kFloor : ((cls == kCellRegular && n1pdf_fast_ok<NACC>(r, e, na)) ? n1pdf_fast<NACC>(r, e, na) : n1pdf_generic_body<NACC>(r, e, na));
            if (zf > 0.0 && pdf <= 0.0) pdf = zf;
            acc.mul(pdf);
            if constexpr (TRACE) trace_out[t + h] = log(pdf);
        }
    }
}

// sum-log-likelihood of ONE parameter vector `th` (global or shared memory) for local subject s over trial chunk `split`, by
// one group (a block, or with G = 32 one warp); addr addresses the draws of `t0 + st0 U`.  The sum is returned in thread 0
// of the group (0 for an empty chunk).  sm_raw: like_smem_bytes() of shared memory; the caller provides a barrier before
// it is reused.  stamp (diagnostics, may be null): thread 0 writes the time the table was finished to stamp[0].
// TRACE (parity entry point ggdmc_b200_trial_logdens_hot only): every density the loops fold into the running product is
// also written, as its log, to trace_out[trial] -- the production loops, the production trial functions.
// EXPAND: after the distinct rows are built they are copied out into one row per (cell, accumulator) right behind the
// like_smem_bytes() region (n_cell n_acc more rows of shared memory), and the trial loop indexes that table by cell.
template <int NACC, int G, bool TRACE = false, bool EXPAND = false>
__device__ __forceinline__ double like_eval(const DevModel &M, const TrialData &T, const double *th, const DrawAddr &addr, int s, int split,
                                            unsigned char *sm_raw, double *trace_out = nullptr, unsigned long long *stamp = nullptr)
{
    const int na = M.n_acc, tid = Grp<G>::tid();
    const int ntr = T.count[s]; // these two loads are only needed after the table is built: they travel meanwhile
    const int64_t t_off = T.offset[s];
    const int t_begin = split * T.chunk;
    CellAcc *rows = reinterpret_cast<CellAcc *>(sm_raw);
    double *red = reinterpret_cast<double *>(rows + M.n_row);
    uint8_t *bad = reinterpret_cast<uint8_t *>(red + G / 32);
    build_rows<G>(M, th, rows, bad, bad + M.n_cell, addr);
    const CellAcc *ent = nullptr;
    if constexpr (EXPAND) {
        CellAcc *e = reinterpret_cast<CellAcc *>(sm_raw + like_smem_bytes(M.n_row, M.n_cell, G));
        const double *src = reinterpret_cast<const double *>(rows);
        double *dst = reinterpret_cast<double *>(e);
        for (int i = tid; i < M.n_cell * na * 8; i += G) dst[i] = src[(int)M.row_of[i >> 3] * 8 + (i & 7)];
        Grp<G>::sync();
        ent = e;
    }
    if (stamp && tid == 0) stamp[0] = BlockTrace::now();
    if (t_begin >= ntr) return 0.0; // empty chunk (a subject with fewer trials than the longest one)

    const int t_end = min(ntr, t_begin + T.chunk);
    const double *rt = T.rt + t_off;
    const uint16_t *cl = T.cell + t_off;
    const uint16_t *row_of = M.row_of;
    LogProd acc;
    acc.init();
    const double zf = T.zero_floor;
    // Hot loop.  Two trials per thread per pass: one 16-byte RT load + one 4-byte cell load (subjects are
    // padded to a multiple of 8 trials with cell = 0xFFFF); the pair is processed by a rolled loop so the
    // loop body stays small.  Only fast-path trials (regular cell, rt > t0) are evaluated here; anything
    // else is left to the cold loop below, which keeps every rare branch -- and its registers -- out of
    // the code the FP64 pipe spends its time in.
    bool leftovers = false;
    if constexpr (NACC == 2) {
        // two trials of a thread advance together (n1pdf_fast2); with more accumulators the second trial's
        // state no longer fits the register budget of 24 warps per SM and the one-trial loop below is faster
        for (int t = t_begin + 2 * tid; t < t_end; t += 2 * G) {
            const double2 r2 = __ldg(reinterpret_cast<const double2 *>(rt + t));
            const ushort2 c2 = __ldg(reinterpret_cast<const ushort2 *>(cl + t));
            const int c0 = c2.x, c1 = (t + 1 < t_end) ? c2.y : c2.x; // the partner of a last odd trial is padding
            bool pair_ok = (t + 1 < t_end) && bad[c0] == kCellRegular && bad[c1] == kCellRegular;
            double p0 = 0.0, p1 = 0.0;
            if constexpr (EXPAND) {
                const CellAcc *e0 = ent + c0 * na, *e1 = ent + c1 * na;
                pair_ok = pair_ok && n1pdf_fast_ok<NACC>(r2.x, e0, na) && n1pdf_fast_ok<NACC>(r2.y, e1, na);
                if (pair_ok) n1pdf_fast2<NACC>(r2.x, e0, r2.y, e1, na, p0, p1);
            } else {
                const RowRef e0{rows, row_of + c0 * na}, e1{rows, row_of + c1 * na};
                pair_ok = pair_ok && n1pdf_fast_ok<NACC>(r2.x, e0, na) && n1pdf_fast_ok<NACC>(r2.y, e1, na);
                if (pair_ok) n1pdf_fast2<NACC>(r2.x, e0, r2.y, e1, na, p0, p1);
            }
            if (pair_ok) {
                if (zf > 0.0) {
                    if (p0 <= 0.0) p0 = zf;
                    if (p1 <= 0.0) p1 = zf;
                }
                acc.mul_fast(p0);
                acc.mul_fast(p1);
                if constexpr (TRACE) {
                    trace_out[t] = log(p0);
                    trace_out[t + 1] = log(p1);
                }
            } else
                leftovers = true;
        }
        if (leftovers) like_cold<NACC, G, true, TRACE>(rows, row_of, bad, na, rt, cl, t_begin, t_end, zf, acc, trace_out);
    } else {
        for (int t = t_begin + 2 * tid; t < t_end; t += 2 * G) {
            const double2 r2 = __ldg(reinterpret_cast<const double2 *>(rt + t));
            const ushort2 c2 = __ldg(reinterpret_cast<const ushort2 *>(cl + t));
            const int nh = (t + 1 < t_end) ? 2 : 1;
#pragma unroll 1
            for (int h = 0; h < nh; ++h) {
                const int c = h ? c2.y : c2.x;
                const double r = h ? r2.y : r2.x;
                bool ok = bad[c] == kCellRegular;
                double pdf = 0.0;
                if constexpr (EXPAND) {
                    const CellAcc *e = ent + c * na;
                    ok = ok && n1pdf_fast_ok<NACC>(r, e, na);
                    if (ok) pdf = n1pdf_fast<NACC>(r, e, na);
                } else {
                    const RowRef e{rows, row_of + c * na};
                    ok = ok && n1pdf_fast_ok<NACC>(r, e, na);
                    if (ok) pdf = n1pdf_fast<NACC>(r, e, na);
                }
                if (ok) {
                    if (zf > 0.0 && pdf <= 0.0) pdf = zf;
                    acc.mul_fast(pdf);
                    if constexpr (TRACE) trace_out[t + h] = log(pdf);
                } else
                    leftovers = true;
            }
        }
        if (leftovers) like_cold<NACC, G, false, TRACE>(rows, row_of, bad, na, rt, cl, t_begin, t_end, zf, acc, trace_out);
    }
    if (stamp && tid == 0) stamp[1] = BlockTrace::now();
    double v = block_sum<G>(acc.value(), red);
    if (tid == 0 && T.counter) atomicAdd(T.counter, (unsigned long long)(t_end - t_begin));
    return v;
}

// sum-log-likelihood of ONE proposal (population p, chain) over one trial chunk, by the whole block
template <int NACC, int BLOCK, bool EXPAND>
__device__ __forceinline__ void like_one(const Level &L, const DevModel &M, const TrialData &T, uint32_t iter, int sweep, int p,
                                         int chain, int split, double *ll_part, unsigned char *sm_raw)
{
    const int C = L.nchain, D = L.npar;
    unsigned long long *bslot = (T.btrace && threadIdx.x == 0) ? T.btrace + 5 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    const double v = like_eval<NACC, BLOCK, false, EXPAND>(M, T, L.prop + ((size_t)p * C + chain) * D, make_addr(L, p, iter, sweep, chain), p / L.n_rep,
                                                           split, sm_raw, nullptr, bslot ? bslot + 3 : nullptr);
    if (threadIdx.x == 0) ll_part[((size_t)p * C + chain) * T.nsplit + split] = v;
}

// DDM twin of like_one (model type "fastdm"): sum over one trial chunk of log(max(g(rt), DBL_MIN))
// (@hdr/likelihood.h:295-305) for ONE proposal.  The per-(chain, cell) table holds what ddm_class::set_parameters,
// set_precision and validate_parameters produce (gg_ddm.cuh); a thread evaluates one trial at a time -- the series
// length and the two midpoint rules make the cost of a trial data dependent, so trials are dealt round-robin.
template <int BLOCK>
__device__ __forceinline__ void like_one_ddm(const Level &L, const DevModel &M, const TrialData &T, int p, int chain, int split,
                                             double *ll_part, unsigned char *sm_raw)
{
    const int C = L.nchain, D = L.npar, na = M.n_acc;
    const int s = p / L.n_rep;
    const int ntr = T.count[s];
    const int64_t t_off = T.offset[s];
    const int t_begin = split * T.chunk;
    double *part = ll_part + ((size_t)p * C + chain) * T.nsplit + split;
    DdmCell *ent = reinterpret_cast<DdmCell *>(sm_raw);
    double *red = reinterpret_cast<double *>(ent + M.n_cell);
    const double *th = L.prop + ((size_t)p * C + chain) * D;
    for (int c = threadIdx.x; c < M.n_cell; c += BLOCK) {
        const int *src = M.param_src + (size_t)c * kDdmRows * na; // column 0 of every row (@hdr/ddm.h:194-214)
        double P[kDdmRows];
#pragma unroll
        for (int r = 0; r < kDdmRows; ++r) {
            const int k = src[r * na];
            P[r] = k >= 0 ? th[k] : M.const_val[-1 - k];
        }
        ddmcell_build(ent[c], P, M.posdrift[c] != 0);
    }
    __syncthreads();
    if (t_begin >= ntr) {
        if (threadIdx.x == 0) *part = 0.0;
        return;
    }
    const int t_end = min(ntr, t_begin + T.chunk);
    const double *rt = T.rt + t_off;
    const uint16_t *cl = T.cell + t_off;
    LogProd acc;
    acc.init();
    const double zf = T.zero_floor;
    for (int t = t_begin + threadIdx.x; t < t_end; t += BLOCK) {
        double pdf = ddm_density(ent[cl[t]], rt[t]);
        if (zf > 0.0 && pdf <= 0.0) pdf = zf;
        acc.mul(ddm_floor(pdf));
    }
    double v = block_sum<BLOCK>(acc.value(), red);
    if (threadIdx.x == 0) {
        *part = v;
        if (T.counter) atomicAdd(T.counter, (unsigned long long)(t_end - t_begin));
    }
}

// Grid.  step >= 0 (REFERENCE schedule): block x = population, its chain is sweep position `step`.
// step < 0, half < 0: block x = (population, chain).  step < 0, half = 0 / 1 (PARALLEL schedule): block x =
// (population, slot) with (nchain + 1) / 2 slots; a crossover population evaluates chain 2 slot + half, a
// migrating population (whole migration in half 0) its chains slot and slot + nslots.
template <int NACC, int BLOCK, int MINB, bool EXPAND>
__global__ void __launch_bounds__(BLOCK, MINB) k_like(Level L, DevModel M, TrialData T, const uint32_t *d_iter, int sweep, int step,
                                                       int half, double *ll_part /* [npop][C][nsplit] */)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int C = L.nchain;
    const uint32_t iter = *d_iter;
    BlockTrace bt(T.btrace);
    int p, chain0, stride = C, nchain_blk = 1; // this block evaluates chains chain0, chain0 + stride, ... (nchain_blk of them)
    if (step >= 0) {
        p = blockIdx.x;
        chain0 = step;
        if (L.mode[p]) {
            if (step >= L.mig_n[p]) return;
            chain0 = L.mig_list[p * C + step];
        }
    } else if (half < 0) {
        p = blockIdx.x / C;
        chain0 = blockIdx.x - p * C;
    } else {
        const int nslot = (C + 1) / 2;
        p = blockIdx.x / nslot;
        const int slot = blockIdx.x - p * nslot;
        if (L.mode[p] == 0) {
            chain0 = 2 * slot + half;
        } else {
            if (half != 0) return;
            chain0 = slot;
            stride = nslot;
            nchain_blk = 2;
        }
    }
    DevModel Ms = M;
    if constexpr (!EXPAND) {
        // the cell -> row map next to the table: the trial loop reads it once per accumulator per trial (first used after
        // build_rows' first barrier)
        uint16_t *s_row_of = reinterpret_cast<uint16_t *>(sm_raw + like_smem_bytes(M.n_row, M.n_cell, BLOCK));
        for (int i = threadIdx.x; i < M.n_cell * M.n_acc; i += BLOCK) s_row_of[i] = M.row_of[i];
        Ms.row_of = s_row_of;
    }
    for (int i = 0, chain = chain0; i < nchain_blk && chain < C; ++i, chain += stride) {
        if (L.target[p * C + chain] >= 0) like_one<NACC, BLOCK, EXPAND>(L, Ms, T, iter, sweep, p, chain, blockIdx.y, ll_part, sm_raw);
        if (nchain_blk > 1) __syncthreads(); // shared table and reduction scratch are reused by the next chain
    }
}

// k_like for model type "fastdm": same grid, same arguments, same block -> (population, chain) mapping
template <int BLOCK, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_like_ddm(Level L, DevModel M, TrialData T, const uint32_t *d_iter, int sweep, int step,
                                                           int half, double *ll_part /* [npop][C][nsplit] */)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int C = L.nchain;
    (void)d_iter; // the DDM density consumes no draws (@hdr/ddm.h has no Rf_runif), so neither iteration nor sweep is needed
    (void)sweep;
    int p, chain0, stride = C, nchain_blk = 1;
    if (step >= 0) {
        p = blockIdx.x;
        chain0 = step;
        if (L.mode[p]) {
            if (step >= L.mig_n[p]) return;
            chain0 = L.mig_list[p * C + step];
        }
    } else if (half < 0) {
        p = blockIdx.x / C;
        chain0 = blockIdx.x - p * C;
    } else {
        const int nslot = (C + 1) / 2;
        p = blockIdx.x / nslot;
        const int slot = blockIdx.x - p * nslot;
        if (L.mode[p] == 0) {
            chain0 = 2 * slot + half;
        } else {
            if (half != 0) return;
            chain0 = slot;
            stride = nslot;
            nchain_blk = 2;
        }
    }
    for (int i = 0, chain = chain0; i < nchain_blk && chain < C; ++i, chain += stride) {
        if (L.target[p * C + chain] >= 0) like_one_ddm<BLOCK>(L, M, T, p, chain, blockIdx.y, ll_part, sm_raw);
        if (nchain_blk > 1) __syncthreads();
    }
}

// per-trial log densities of one subject for n_theta parameter vectors (parity / init entry point)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_trial_logdens(DevModel M, const double *rt, const uint16_t *cl, int ntr,
                                                         const double *theta, double *out)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int na = M.n_acc, D = M.npar, k = blockIdx.x;
    CellAcc *rows = reinterpret_cast<CellAcc *>(sm_raw);
    uint8_t *bad = reinterpret_cast<uint8_t *>(rows + M.n_row);
    DrawAddr addr = {0, 0, 0, 0, 0};
    build_rows<BLOCK>(M, theta + (size_t)k * D, rows, bad, bad + M.n_cell, addr);
    for (int t = blockIdx.y * BLOCK + threadIdx.x; t < ntr; t += gridDim.y * BLOCK) {
        const int c = cl[t];
        out[(size_t)k * ntr + t] = log(n1pdf_any<0>(bad[c], rt[t], RowRef{rows, M.row_of + c * na}, na));
    }
}

// Per-trial log densities through the PRODUCTION trial loops (like_eval with TRACE): block (theta k, subject 0, chunk y).
// sums[k * nsplit + y] also receives the chunk's sum as the sampler would see it.
template <int NACC, int BLOCK, bool EXPAND>
__global__ void __launch_bounds__(BLOCK) k_trial_logdens_hot(DevModel M, TrialData T, const double *theta, int ntr, uint64_t seed, uint32_t pop,
                                                             uint32_t iter, double *out, double *sums)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int k = blockIdx.x;
    DrawAddr addr = {seed, pop, iter, 0u, (uint32_t)k};
    const double v = like_eval<NACC, BLOCK, true, EXPAND>(M, T, theta + (size_t)k * M.npar, addr, 0, blockIdx.y, sm_raw, out + (size_t)k * ntr);
    if (threadIdx.x == 0) sums[(size_t)k * T.nsplit + blockIdx.y] = v;
}

// the same for the DDM: log(max(density, DBL_MIN)) of every trial (@hdr/likelihood.h:303)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_trial_logdens_ddm(DevModel M, const double *rt, const uint16_t *cl, int ntr,
                                                             const double *theta, double *out)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int na = M.n_acc, D = M.npar, k = blockIdx.x;
    DdmCell *ent = reinterpret_cast<DdmCell *>(sm_raw);
    const double *th = theta + (size_t)k * D;
    for (int c = threadIdx.x; c < M.n_cell; c += BLOCK) {
        const int *src = M.param_src + (size_t)c * kDdmRows * na;
        double P[kDdmRows];
#pragma unroll
        for (int r = 0; r < kDdmRows; ++r) {
            const int i = src[r * na];
            P[r] = i >= 0 ? th[i] : M.const_val[-1 - i];
        }
        ddmcell_build(ent[c], P, M.posdrift[c] != 0);
    }
    __syncthreads();
    for (int t = blockIdx.y * BLOCK + threadIdx.x; t < ntr; t += gridDim.y * BLOCK)
        out[(size_t)k * ntr + t] = log(ddm_floor(ddm_density(ent[cl[t]], rt[t])));
}

// ------------------------------------------------------------------------------------------------
// K3: Metropolis accept / commit at the subject level (update_theta, src/de.cpp:81-108)
// ------------------------------------------------------------------------------------------------
// MH tests of the pending proposals made from chains [c_begin, c_end) of population p by ONE WARP: lane = chain for the
// decision (update_theta, src/de.cpp:81-108), then the whole warp copies every accepted vector (a lane copying its own
// vector alone pays one memory round trip per element: the loads may alias the stores).
__device__ __forceinline__ void accept_warp(const Level &L, int p, int c_begin, int c_end, uint32_t iter, int sweep, const double *ll_part, int nsplit,
                                            int lane)
{
    const int C = L.nchain, D = L.npar;
    for (int base = c_begin; base < c_end; base += 32) {
        const int src = base + lane;
        int tgt = -1, acc = 0;
        double tmp_lp = 0.0, tmp_ll = 0.0;
        if (src < c_end) tgt = ldm(L.target + p * C + src);
        if (tgt >= 0) {
            const double *part = ll_part + ((size_t)p * C + src) * nsplit;
            for (int k = 0; k < nsplit; ++k) tmp_ll += ldm(part + k);
            tmp_lp = L.prior_ovr ? deferred_prop_lp(L, p, src) : ldm(L.prop_lp + p * C + src);
            const double cur = ldm(L.lp + p * C + tgt) + ldm(L.ll + p * C + tgt); // src/de.cpp:121 / :189-190 / :577 / :656-657
            const double mh = exp((tmp_lp + tmp_ll) - cur);                       // :147
            L.target[p * C + src] = -1;                                            // proposal consumed
            if (!isnan(mh)) {                                                      // :83-87, no draw
                DrawAddr a = make_addr(L, p, iter, sweep, src);
                acc = draw_uniform(a, U_ACCEPT, 0) < mh;                           // :88
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, acc);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int s2 = base + l, t2 = __shfl_sync(0xffffffffu, tgt, l);
            const double *pr = L.prop + ((size_t)p * C + s2) * D;
            double *th = L.theta + ((size_t)p * C + t2) * D;
            for (int d = lane; d < D; d += 32) th[d] = ldm(pr + d);
        }
        if (acc) {
            L.lp[p * C + tgt] = tmp_lp;
            L.ll[p * C + tgt] = tmp_ll;
        }
    }
}

// The same MH test by ONE WARP for a crossover proposal that is still in shared memory (prop_sm; made from chain src, which
// is also the chain it challenges): lane d evaluates prior term d, so the dependent loads of the thirteen terms travel
// together instead of one after the other.  No other item of the half-sweep reads chain src (difference partners come
// from the other parity), so the chain can be overwritten at once.  lp_lane0: log prior of the proposal (lane 0) when the
// prior is fixed; ll_lane0: its log-likelihood (lane 0) when there is one trial chunk.  scratch: D doubles of shared memory.
__device__ __forceinline__ void accept_self(const Level &L, int p, int src, uint32_t iter, int sweep, const double *prop_sm, double lp_lane0,
                                            const double *ll_part, int nsplit, double ll_lane0, double *scratch, int lane, int tgt = -1)
{
    const int C = L.nchain, D = L.npar;
    if (tgt < 0) tgt = src;
    // every global load of the decision is issued up front: they travel together
    double cur = 0.0, tmp_ll = ll_lane0;
    if (lane == 0) {
        cur = ldm(L.lp + p * C + tgt) + ldm(L.ll + p * C + tgt); // src/de.cpp:121 / :189-190 / :577 / :656-657
        if (nsplit > 1) {
            const double *part = ll_part + ((size_t)p * C + src) * nsplit;
            tmp_ll = 0.0;
            for (int k = 0; k < nsplit; ++k) tmp_ll += ldm(part + k);
        }
    }
    if (L.prior_ovr) {
        const size_t rc = (size_t)(p % L.n_rep) * C + src;
        const double *ovr = L.prior_ovr + rc * 2 * D;
        const double *oc = L.ovr_consts ? L.ovr_consts + rc * 2 * D : nullptr;
        bool generic = oc == nullptr;
        if (oc && D <= 32) { // phi-driven truncated normals (the hierarchical fits of the reference): no dependent loads
            double K = 0.0, m = 0.0, iv = 0.0, x = 0.0, lo = 0.0, up = 0.0;
            if (lane < D) {
                K = ldm(oc + 2 * lane + 1); m = ldm(ovr + lane); iv = ldm(oc + 2 * lane);
                lo = L.prior.lower[lane]; up = L.prior.upper[lane];
                x = prop_sm[lane];
                const double z = (x - m) * iv;
                scratch[lane] = (x < lo || x > up) ? -INFINITY : -(0.5 * z * z + K);
            }
            generic = __any_sync(0xffffffffu, lane < D && !(K == K));
        }
        if (generic)
            for (int d = lane; d < D; d += 32) scratch[d] = prior_term(L, d, prop_sm[d], ovr, oc);
        __syncwarp();
    }
    int acc = 0;
    double tmp_lp = 0.0;
    if (lane == 0) {
        tmp_lp = L.prior_ovr ? sum_arma_order(scratch, D) : lp_lane0;
        const double mh = exp((tmp_lp + tmp_ll) - cur); // :147
        if (!isnan(mh)) {                               // :83-87, no draw
            DrawAddr a = make_addr(L, p, iter, sweep, src);
            acc = draw_uniform(a, U_ACCEPT, 0) < mh;    // :88
        }
    }
    if (__shfl_sync(0xffffffffu, acc, 0)) {
        double *th = L.theta + ((size_t)p * C + tgt) * D;
        for (int d = lane; d < D; d += 32) th[d] = prop_sm[d];
        if (lane == 0) {
            L.lp[p * C + tgt] = tmp_lp;
            L.ll[p * C + tgt] = tmp_ll;
        }
    }
}

// One WARP per pending proposal (lane = parameter for the log prior under phi and for the copy): the decision's loads are
// one round trip wide instead of one per parameter.  Dynamic shared memory: (warps per block) x npar doubles.
// half = 0 / 1 (PARALLEL schedule): one warp per (population, slot).  Half 1: chain 2 slot + 1 (only crossover populations
// propose there).  Half 0: chains 2 slot and 2 slot + 1 -- a crossover population has nothing pending on the odd one, a
// migrating population may have on both.  The kernel does not read the sweep decisions (L.mode ...): the next iteration's
// may be drawn while it runs.
template <int WARPS, int MINB = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB) k_accept(Level L, const uint32_t *d_iter, int sweep, int step, const double *ll_part, int nsplit,
                                                        int half = -1)
{
    extern __shared__ double sm_acc[];
    const int C = L.nchain, D = L.npar, lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int g = blockIdx.x * WARPS + wl;
    int p, src, src2 = -1;
    if (step < 0 && half >= 0) {
        const int nslot = (C + 1) / 2;
        if (g >= L.npop * nslot) return;
        p = g / nslot;
        const int slot = g - p * nslot;
        src = 2 * slot + half;
        if (half == 0) src2 = 2 * slot + 1;
    } else if (step < 0) {
        if (g >= L.npop * C) return;
        p = g / C;
        src = g - p * C;
    } else {
        p = g;
        if (p >= L.npop) return;
        const int mode = L.mode[p];
        if (mode) {
            if (step >= L.mig_n[p]) return;
            src = L.mig_list[p * C + step];
        } else
            src = step;
    }
    for (; src >= 0; src = src2, src2 = -1) {
        if (src >= C) continue;
        const int tgt = L.target[p * C + src];
        if (tgt < 0) continue;
        const double *pr = L.prop + ((size_t)p * C + src) * D;
        double *scratch = sm_acc + (size_t)wl * D;
        double lp0 = 0.0, ll0 = 0.0;
        __syncwarp();
        if (lane == 0) {
            L.target[p * C + src] = -1; // proposal consumed
            if (!L.prior_ovr) lp0 = L.prop_lp[p * C + src];
            if (nsplit == 1) ll0 = ll_part[(size_t)p * C + src];
        }
        accept_self(L, p, src, *d_iter, sweep, pr, lp0, ll_part, nsplit, ll0, scratch, lane, tgt);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// K4: hyper-likelihood partial sums.  Block (population r, chain c, subject split).
//   sums[..][0] = sum_s log p(x_s | phi_c)   (current phi of chain c; src/de.cpp:397-398, 494-500)
//   sums[..][1] = sum_s log p(x_s | phi'_c)  (proposal made FROM chain c; :427, :519-520)
// x_s = theta of subject s, chain c (hierarchy) or row s of the data matrix (run_hyper,
// @hdr/likelihood.h:257-271).
// ------------------------------------------------------------------------------------------------
struct HyperArgs {
    DevPrior like;            // p_prior: lower/upper/dist/log_p of the subject-level parameters
    const double *x;          // subject thetas [n_rep][S][C][D] or data [S][D]
    int x_rep_stride, x_subj_stride, x_chain_stride; // in doubles
    int S;                    // local subjects
    int D;                    // subject-level npar (phi vector is 2 D)
    int subj_per_block, nsplit;
    int need_cur;             // refresh current hyper-likelihood (hierarchy) or not (run_hyper)
    // Constants of the phi-driven subject prior (what k_phi_consts computes) of every PROPOSED phi vector, written by the
    // proposing item and copied over the target chain's constants when the proposal is accepted -- the two Phi and the
    // logarithm are needed for the hyper-likelihood anyway.  Null: the constants are refreshed by k_phi_consts instead.
    double *prop_consts;      // [n_rep][C][D][2], indexed by the chain proposed FROM
    double *consts;           // [n_rep][C][D][2] = Level::ovr_consts of the subject level
};

// hyper-likelihood terms of one block: the subjects [s_begin, s_end) of (replicate r, chain c) under the
// current phi (cm, cs, cl) and the proposed phi (pm, ps, pl); sm_h layout as in k_hyper
// phi_c: global memory (mutable state); phi_p: the proposal, in shared memory (prop_in_smem) or global memory, read only
// when has_prop
template <int BLOCK>
__device__ __forceinline__ void hyper_block(const HyperArgs &H, int r, int c, int split, const double *phi_c, const double *phi_p,
                                            bool has_prop, bool prop_in_smem, double *sm_h, double &vc, double &vp, double *prop_consts = nullptr)
{
    const int D = H.D, tid = Grp<BLOCK>::tid();
    double *cm = sm_h, *cs = cm + D, *cl = cs + D, *pm = cl + D, *ps = pm + D, *pl = ps + D, *cls = pl + D, *pls = cls + D, *red = pls + D;
    auto prep_cur = [&](int d) {
        const double m = ldm(phi_c + d), s = ldm(phi_c + D + d);
        cm[d] = m; cs[d] = s;
        if (H.need_cur) {
            cl[d] = tnorm_logmass(H.like.lower[d], H.like.upper[d], m, s); // tnorm_class::set_parameters, @hdr/tnorm.h:59-67
            cls[d] = log(s); // dnorm4's log(sigma): once per parameter instead of once per subject
        }
    };
    auto prep_prop = [&](int d) {
        const double m = prop_in_smem ? phi_p[d] : ldm(phi_p + d);
        const double s = prop_in_smem ? phi_p[D + d] : ldm(phi_p + D + d);
        pm[d] = m; ps[d] = s;
        pl[d] = tnorm_logmass(H.like.lower[d], H.like.upper[d], m, s);
        const double ls = log(s);
        pls[d] = ls;
        if (prop_consts) { // the same expressions as k_phi_consts
            double inv = 0.0, K = NAN;
            if (H.like.dist[d] == 1 && H.like.log_p[d] != 0 && s > 0.0 && isfinite(s) && isfinite(m)) {
                inv = 1.0 / s;
                K = kLnSqrt2Pi + ls + pl[d];
            }
            prop_consts[2 * d] = inv;
            prop_consts[2 * d + 1] = K;
        }
    };
    if (BLOCK == 32 && D <= 16) { // one warp: the current phi on lanes 0-15, the proposed one on lanes 16-31, side by side
        const int d = tid & 15;
        if (d < D) {
            if (tid < 16) prep_cur(d);
            else if (has_prop) prep_prop(d);
        }
    } else {
        for (int d = tid; d < D; d += BLOCK) {
            prep_cur(d);
            if (has_prop) prep_prop(d);
        }
    }
    Grp<BLOCK>::sync();
    const int s_begin = split * H.subj_per_block, s_end = min(H.S, s_begin + H.subj_per_block);
    const int n_el = (s_end - s_begin) * D;
    const double *xbase = H.x + (size_t)r * H.x_rep_stride + (size_t)c * H.x_chain_stride;
    double sum_c = 0.0, sum_p = 0.0;
    for (int e0 = tid; e0 < n_el; e0 += 4 * BLOCK) { // four loads in flight per thread; the order of the additions is unchanged
        double xs[4];
        int ds[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = e0 + k * BLOCK;
            const int si = e / D;
            ds[k] = e - si * D;
            xs[k] = e < n_el ? ldm(xbase + (size_t)(s_begin + si) * H.x_subj_stride + ds[k]) : 0.0;
        }
        // log density of the truncated normal: regular arguments (finite, sd > 0, |z| far from overflow -- every element
        // of an ordinary fit) take dnorm4's main line directly, so the eight divisions of a pass are independent
        // instructions instead of eight branchy calls; anything else goes through dnorm4's full rules
        double tc[4], tp[4];
        bool slow = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int d = ds[k];
            const double x = xs[k];
            const double zc = (x - cm[d]) / cs[d], zp = has_prop ? (x - pm[d]) / ps[d] : 0.0;
            tc[k] = -(kLnSqrt2Pi + 0.5 * zc * zc + cls[d]) - cl[d];
            tp[k] = -(kLnSqrt2Pi + 0.5 * zp * zp + pls[d]) - pl[d];
            const bool reg = H.like.dist[d] == 1 && H.like.log_p[d] != 0 && fabs(zc) < 1e150 && fabs(zp) < 1e150 && cs[d] > 0.0 && cs[d] < 1e300 &&
                             (!has_prop || (ps[d] > 0.0 && ps[d] < 1e300));
            slow = slow || (!reg && e0 + k * BLOCK < n_el);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (e0 + k * BLOCK >= n_el) break;
            const int d = ds[k];
            const double x = xs[k];
            const double lo = H.like.lower[d], up = H.like.upper[d];
            if (!slow) {
                const bool outside = (x < lo) || (x > up);
                if (H.need_cur) sum_c += outside ? -INFINITY : tc[k];
                if (has_prop) sum_p += outside ? -INFINITY : tp[k];
                continue;
            }
            const int dist = H.like.dist[d];
            const bool lg = H.like.log_p[d] != 0;
            if (dist == 1 && lg) { // TNORM, log scale: the hierarchical fits of the reference
                const bool outside = (x < lo) || (x > up);
                if (H.need_cur) sum_c += outside ? -INFINITY : dnorm4_log_pre(x, cm[d], cs[d], cls[d]) - cl[d];
                if (has_prop) sum_p += outside ? -INFINITY : dnorm4_log_pre(x, pm[d], ps[d], pls[d]) - pl[d];
            } else {
                if (H.need_cur) sum_c += dprior1(dist, x, cm[d], cs[d], lo, up, lg);
                if (has_prop) sum_p += dprior1(dist, x, pm[d], ps[d], lo, up, lg);
            }
        }
    }
    vc = block_sum<BLOCK>(sum_c, red);
    vp = block_sum<BLOCK>(sum_p, red + BLOCK / 32);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_hyper(Level L, HyperArgs H, int step, double *hpart /* [npop][C][2][nsplit] */)
{
    extern __shared__ double sm_h[]; // cur: mean[D] sd[D] logden[D]; prop: same; log sd of both [2 D]; then reduction scratch
    const int C = L.nchain, D = H.D;
    int r, c;
    if (step < 0) {
        r = blockIdx.x / C;
        c = blockIdx.x - r * C;
    } else {
        r = blockIdx.x;
        const int mode = L.mode[r];
        if (mode) { // in-place migration step: both the source chain (z = 0) and the chain it is compared with (z = 1)
            const int n = L.mig_n[r];
            if (step >= n) return;
            c = blockIdx.z == 0 ? L.mig_list[r * C + step] : L.mig_list[r * C + ((step + 1 == n) ? 0 : step + 1)];
        } else {
            if (blockIdx.z != 0) return;
            c = step;
        }
    }
    const int split = blockIdx.y;
    double vc, vp;
    hyper_block<BLOCK>(H, r, c, split, L.theta + ((size_t)r * C + c) * 2 * D, L.prop + ((size_t)r * C + c) * 2 * D,
                       L.target[r * C + c] >= 0, false, sm_h, vc, vp);
    if (threadIdx.x == 0) {
        double *o = hpart + (((size_t)r * C + c) * 2) * H.nsplit + split;
        o[0] = vc;
        o[H.nsplit] = vp;
    }
}

// [npop*C*2][nsplit] -> [npop*C*2]
__global__ void k_hyper_reduce(const double *hpart, int n, int nsplit, double *hsum)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = 0.0;
    for (int k = 0; k < nsplit; ++k) v += hpart[(size_t)i * nsplit + k];
    hsum[i] = v;
}

// ------------------------------------------------------------------------------------------------
// The one exchange of the path, fused with its reduction: sum over the LOCAL subjects (the nsplit
// partials of k_hyper) and, in the same kernel, sum over the GPUs through peer memory on NVLink.
//
// Every rank owns a window  slots[2][n_rank][kP2PMaxN] doubles + flags[2][n_rank] u64  in its own
// HBM, mapped into every peer with CUDA IPC.  Exchange number `seq` (parity b = seq & 1):
//   1. thread i reduces value i and STORES it into slot [b][my_rank][i] of every rank's window
//      (remote stores over NVLink, posted -- nobody waits on a load round trip)
//   2. system-scope fence, then one flag store per peer: flags[b][my_rank] = seq
//   3. wait until the local flags[b][*] all read `seq` (every peer's data has landed)
//   4. thread i sums slots[b][0..n_rank)[i] in rank order -> identical bits on every rank
// Parity double-buffering is enough: a rank can only start exchange seq + 2 after every peer has
// raised its seq + 1 flag, i.e. after every peer finished reading exchange seq.
// A bounded spin (spin_ns) sets *status instead of hanging the GPU if a peer never arrives; the function then returns
// false in every thread and the callers do no further work on the exchanged sums (no MH test on partial sums): the
// host finds the status word set at its next synchronisation and reports GGDMC_ERR_COMM.  Once set, the status word
// makes every later exchange of the process return false immediately until ggdmc_b200_comm_finalize().
// ------------------------------------------------------------------------------------------------
constexpr int kP2PMaxN = 4096;   // doubles per exchange (2 * nchain * n_replicate)
constexpr int kP2PMaxRanks = 16;

struct P2PWindow {
    double *slots[kP2PMaxRanks];               // slots base of every rank's window (index = rank; own entry = local pointer)
    unsigned long long *flags[kP2PMaxRanks];   // flags base of every rank's window
    unsigned long long *seq;                   // local exchange counter (device memory)
    int *status;                               // local: set to 1 on timeout
    int n_rank, rank;
    unsigned long long spin_ns;                // how long a rank waits for its peers (GGDMC_B200_PEER_TIMEOUT_S, default 120 s)
};

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// sum of the nsplit partial sums of one value, in index order (bit-reproducible), eight loads in flight at a time: the
// caller sits on the critical path of the phi step and every load is an L2 round trip
__device__ __forceinline__ double sum_partials(const double *hp, int nsplit)
{
    double v = 0.0;
    int q = 0;
    for (; q + 8 <= nsplit; q += 8) {
        double x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = ldm(hp + q + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) v += x[k];
    }
    for (; q + 2 <= nsplit; q += 2) {
        const double a = ldm(hp + q), b = ldm(hp + q + 1);
        v += a;
        v += b;
    }
    if (q < nsplit) v += ldm(hp + q);
    return v;
}

// group-wide: reduce the nsplit partials of every value, exchange with the peers, leave the all-rank sums in hsum
template <int G = 0>
__device__ __forceinline__ bool reduce_exchange_block(const double *hpart, int n, int nsplit, double *hsum, const P2PWindow &w)
{
    const int tid = Grp<G>::tid(), nthr = Grp<G>::size();
    const unsigned long long seq = ldm(w.seq) + 1;
    const int b = (int)(seq & 1ull);
    for (int i = tid; i < n; i += nthr) {
        const double v = sum_partials(hpart + (size_t)i * nsplit, nsplit);
        for (int q = 0; q < w.n_rank; ++q) w.slots[q][((size_t)b * w.n_rank + w.rank) * kP2PMaxN + i] = v;
    }
    __threadfence_system();
    Grp<G>::sync();
    if (tid < w.n_rank) {
        volatile unsigned long long *f = w.flags[tid] + (size_t)b * kP2PMaxRanks + w.rank;
        *f = seq;
    }
    if (tid < w.n_rank) {
        volatile unsigned long long *mine = w.flags[w.rank] + (size_t)b * kP2PMaxRanks + tid;
        const unsigned long long t0 = globaltimer_ns();
        while (*mine < seq) {
            if (*(volatile int *)w.status) break; // an earlier exchange already timed out: do not wait again
            __nanosleep(64);
            if (globaltimer_ns() - t0 > w.spin_ns) { // a peer is gone
                *(volatile int *)w.status = 1;
                break;
            }
        }
    }
    __threadfence_system();
    Grp<G>::sync();
    if (*(volatile int *)w.status) return false; // group-uniform: read after the barrier, never cleared while kernels run
    const double *loc = w.slots[w.rank] + (size_t)b * w.n_rank * kP2PMaxN;
    for (int i = tid; i < n; i += nthr) {
        double v = 0.0;
        for (int q = 0; q < w.n_rank; ++q) v += *(volatile const double *)(loc + (size_t)q * kP2PMaxN + i);
        hsum[i] = v;
    }
    Grp<G>::sync();
    if (tid == 0) *w.seq = seq;
    return true;
}

__global__ void __launch_bounds__(256) k_hyper_reduce_exchange(const double *hpart, int n, int nsplit, double *hsum, P2PWindow w)
{
    reduce_exchange_block(hpart, n, nsplit, hsum, w);
}

// all ranks arrive (an exchange of zero values): lines the ranks up at the end of engine construction and at the start
// of every iterate call (the exchange inside the sampler assumes lock step within spin_ns), and outside timed regions
__global__ void k_peer_barrier(P2PWindow w) { reduce_exchange_block(nullptr, 0, 0, nullptr, w); }

// phi-level accept of the proposal made from chain src (src/de.cpp:397-400, 427-463 and :494-500, 519-549)
// prop_consts / consts (HyperArgs, may be null): on accept the constants of the proposal replace those of the target chain
__device__ __forceinline__ void phi_accept_one(const Level &L, int r, int src, uint32_t iter, int sweep, bool in_place_migration,
                                               const double *hsum, int need_cur, const double *prop_consts = nullptr, double *consts = nullptr)
{
    const int C = L.nchain, D = L.npar;
    const int tgt = ldm(L.target + r * C + src);
    if (tgt < 0) return;
    const double tmp_ll = ldm(hsum + ((size_t)r * C + src) * 2 + 1);
    const double tmp_lp = ldm(L.prop_lp + r * C + src);
    double cur_ll = ldm(L.ll + r * C + tgt);
    if (need_cur) {
        cur_ll = ldm(hsum + ((size_t)r * C + tgt) * 2 + 0);
        if (in_place_migration) L.ll[r * C + src] = ldm(hsum + ((size_t)r * C + src) * 2 + 0); // :494-496 (in place order only)
        L.ll[r * C + tgt] = cur_ll; // :397-398 / :498-500
    }
    const double cur = ldm(L.lp + r * C + tgt) + cur_ll;
    const double mh = exp((tmp_lp + tmp_ll) - cur);
    L.target[r * C + src] = -1; // proposal consumed
    if (isnan(mh)) return;
    DrawAddr a = make_addr(L, r, iter, sweep, src);
    if (draw_uniform(a, U_ACCEPT, 0) < mh) {
        const double *pr = L.prop + ((size_t)r * C + src) * D;
        double *th = L.theta + ((size_t)r * C + tgt) * D;
        for (int d = 0; d < D; ++d) th[d] = ldm(pr + d);
        L.lp[r * C + tgt] = tmp_lp;
        L.ll[r * C + tgt] = tmp_ll;
        if (prop_consts) { // D = 2 x (subject npar): [npar][2] doubles per chain
            const double *pc = prop_consts + ((size_t)r * C + src) * D;
            double *cc = consts + ((size_t)r * C + tgt) * D;
            for (int d = 0; d < D; ++d) cc[d] = ldm(pc + d);
        }
    }
}

// Every MH decision of a phi half-sweep by ONE WARP: lane = chain for the decision (phi_accept_one's rules), then the whole
// warp copies every accepted phi vector and the prior constants that belong to it.
// base0 / stride: the groups of 32 chains this warp takes (one warp: 0 / 32; warp w of a block of W warps: 32 w / 32 W)
__device__ __forceinline__ void phi_accept_warp(const Level &L, int n_rc, uint32_t iter, int sweep, const double *hsum, int need_cur,
                                                const double *prop_consts, double *consts, int lane, int base0 = 0, int stride = 32)
{
    const int C = L.nchain, D = L.npar;
    for (int base = base0; base < n_rc; base += stride) {
        const int g = base + lane; // = r * C + src
        int tgt = -1, acc = 0;
        double tmp_lp = 0.0, tmp_ll = 0.0;
        if (g < n_rc) tgt = ldm(L.target + g);
        const int r = g / C;
        if (tgt >= 0) {
            tmp_ll = ldm(hsum + (size_t)g * 2 + 1);
            tmp_lp = ldm(L.prop_lp + g);
            double cur_ll = ldm(L.ll + r * C + tgt);
            if (need_cur) {
                cur_ll = ldm(hsum + ((size_t)r * C + tgt) * 2 + 0);
                L.ll[r * C + tgt] = cur_ll; // :397-398 / :498-500
            }
            const double cur = ldm(L.lp + r * C + tgt) + cur_ll;
            const double mh = exp((tmp_lp + tmp_ll) - cur);
            L.target[g] = -1; // proposal consumed
            if (!isnan(mh)) {
                DrawAddr a = make_addr(L, r, iter, sweep, g - r * C);
                acc = draw_uniform(a, U_ACCEPT, 0) < mh;
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, acc);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const int g2 = base + l, t2 = __shfl_sync(0xffffffffu, tgt, l);
            const size_t to = ((size_t)(g2 / C) * C + t2) * D;
            const double *pr = L.prop + (size_t)g2 * D;
            for (int d = lane; d < D; d += 32) L.theta[to + d] = ldm(pr + d);
            if (prop_consts)
                for (int d = lane; d < D; d += 32) consts[to + d] = ldm(prop_consts + (size_t)g2 * D + d);
        }
        if (acc) {
            L.lp[r * C + tgt] = tmp_lp;
            L.ll[r * C + tgt] = tmp_ll;
        }
    }
}

// status: the peer window's status word (or null); set = the exchange before this launch timed out, hsum is not valid
__global__ void k_phi_accept(Level L, const uint32_t *d_iter, int sweep, int step, const double *hsum, int need_cur, const int *status)
{
    const int C = L.nchain;
    int r, src;
    if (status && *status) return;
    if (step < 0) {
        int g = blockIdx.x * blockDim.x + threadIdx.x;
        if (g >= L.npop * C) return;
        r = g / C;
        src = g - r * C;
    } else {
        r = blockIdx.x * blockDim.x + threadIdx.x;
        if (r >= L.npop) return;
        const int mode = L.mode[r];
        if (mode) {
            if (step >= L.mig_n[r]) return;
            src = L.mig_list[r * C + step];
        } else
            src = step;
    }
    phi_accept_one(L, r, src, *d_iter, sweep, step >= 0 && L.mode[r] != 0, hsum, need_cur);
}

// ------------------------------------------------------------------------------------------------
// One half-sweep (or one whole snapshot sweep, half < 0) of the phi level in ONE launch:
//   every block (replicate r, chain c, subject split): the proposal made from chain c (recomputed by each
//   split -- it is a pure function of the counter-addressed draws), then its share of the two hyper-likelihood
//   sums; the block that finishes last (ticket): sum over the splits, exchange with the peer GPUs through the
//   peer-memory window (multi-GPU), MH test of every proposal.
// Replaces k_propose + k_hyper + k_hyper_reduce(_exchange) + k_phi_accept on the phi critical path.
// ------------------------------------------------------------------------------------------------
// Part 1, every block (replicate r, chain c, subject split): proposal + the block's share of the two sums -> hpart.
// sm_h: 8 D + 2 BLOCK/32 doubles (hyper_block), then proposal [2 D], prior scratch [2 D]; s_k: one int of shared memory.
// Returns (block-uniform) the sweep position at which chain c proposes in this half, -1 if it does not.
template <int BLOCK>
__device__ __forceinline__ int phi_half_part(const Level &L, const HyperArgs &H, uint32_t iter, int sweep, int half, int r, int c, int split,
                                             double *hpart, double *sm_h, int *s_k)
{
    const int C = L.nchain, D = H.D, D2 = 2 * D, tid = Grp<BLOCK>::tid();
    const int mode = ldm(L.mode + r), para_idx = ldm(L.para + r);
    const int nsteps = mode ? ldm(L.mig_n + r) : C;
    double *sprop = sm_h + 8 * D + 2 * (BLOCK / 32), *scratch = sprop + D2;
    if (tid == 0) { // sweep position at which chain c proposes in this launch, -1: it does not
        int k = -1;
        if (mode == 0) {
            if (half < 0 || (c & 1) == half) k = c;
        } else if (mode == 1 && half <= 0) {
            for (int j = 0; j < nsteps; ++j)
                if (ldm(L.mig_list + r * C + j) == c) { k = j; break; }
        }
        *s_k = k;
    }
    Grp<BLOCK>::sync();
    const int k = *s_k;
    if (k >= 0 && tid < 32) {
        int src, tgt;
        double lp;
        propose_position(L, r, k, mode, nsteps, para_idx, iter, sweep, half, tid, scratch, sprop, src, tgt, lp);
        if (split == 0) {
            double *pr = L.prop + ((size_t)r * C + c) * D2;
            for (int d = tid; d < D2; d += 32) pr[d] = sprop[d];
            if (tid == 0) {
                L.prop_lp[r * C + c] = lp;
                L.target[r * C + c] = tgt;
            }
        }
    }
    Grp<BLOCK>::sync();
    double vc = 0.0, vp = 0.0;
    // A chain that does not propose in this half is nobody's target either (crossover: target = source; migration: the
    // targets are the proposing set), so neither of its two sums is ever read.
    if (k >= 0)
        hyper_block<BLOCK>(H, r, c, split, L.theta + ((size_t)r * C + c) * D2, sprop, true, true, sm_h, vc, vp,
                           (H.prop_consts && split == 0) ? H.prop_consts + ((size_t)r * C + c) * D2 : nullptr);
    if (tid == 0) {
        double *o = hpart + (((size_t)r * C + c) * 2) * H.nsplit + split;
        o[0] = vc;
        o[H.nsplit] = vp;
    }
    return k;
}

// Part 2, the group that finishes last: sum over the splits, exchange with the peer GPUs, MH test of every proposal.
// Returns false (group-uniform) when the exchange timed out: nothing was accepted.
template <int BLOCK>
__device__ __forceinline__ bool phi_half_finish(const Level &L, const HyperArgs &H, uint32_t iter, int sweep, const double *hpart, double *hsum,
                                                const P2PWindow &w, int use_p2p)
{
    const int C = L.nchain, tid = Grp<BLOCK>::tid();
    const int n = L.npop * C * 2;
    if (use_p2p) {
        if (!reduce_exchange_block<BLOCK>(hpart, n, H.nsplit, hsum, w)) return false;
    } else {
        for (int i = tid; i < n; i += BLOCK) hsum[i] = sum_partials(hpart + (size_t)i * H.nsplit, H.nsplit);
    }
    __threadfence();
    Grp<BLOCK>::sync();
    // lane = chain for the decision, the whole warp for the copies of an accepted vector (phi_accept_warp)
    if constexpr (BLOCK == 32) phi_accept_warp(L, L.npop * C, iter, sweep, hsum, H.need_cur, H.prop_consts, H.consts, tid);
    else phi_accept_warp(L, L.npop * C, iter, sweep, hsum, H.need_cur, H.prop_consts, H.consts, tid & 31, 32 * (tid >> 5), BLOCK);
    return true;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_phi_half(Level L, HyperArgs H, const uint32_t *d_iter, int sweep, int half, double *hpart,
                                                    double *hsum, unsigned int *ticket, P2PWindow w, int use_p2p)
{
    extern __shared__ double sm_h[]; // hyper_block's 8 D + 2 BLOCK/32, then proposal [2 D], prior scratch [2 D]
    __shared__ int s_k, s_last;
    const int C = L.nchain;
    const int r = blockIdx.x / C, c = blockIdx.x - r * C, split = blockIdx.y;
    const uint32_t iter = *d_iter;
    phi_half_part<BLOCK>(L, H, iter, sweep, half, r, c, split, hpart, sm_h, &s_k);
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    phi_half_finish<BLOCK>(L, H, iter, sweep, hpart, hsum, w, use_p2p);
    if (threadIdx.x == 0) *ticket = 0;
}

// ------------------------------------------------------------------------------------------------
// storage (theta_phi::store, @hdr/theta.h:61-74) and the iteration counter
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_level(const Level &L, uint32_t iter)
{
    if (iter % (uint32_t)L.thin != 0) return;
    const uint32_t slot = iter / (uint32_t)L.thin;
    if (slot >= (uint32_t)L.nmc) return;
    const size_t CD = (size_t)L.nchain * L.npar, C = L.nchain;
    const size_t total = (size_t)L.npop * CD;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / CD, r = i - p * CD;
        L.out_theta[(p * L.nmc + slot) * CD + r] = L.theta[i];
        if (r < C) {
            L.out_lp[(p * L.nmc + slot) * C + r] = L.lp[p * C + r];
            L.out_ll[(p * L.nmc + slot) * C + r] = L.ll[p * C + r];
        }
    }
}

__global__ void k_store(Level L, const uint32_t *d_iter) { store_level(L, *d_iter); }

// End of an iteration: store both levels (has_b: second level present) and advance the device-side
// iteration counter.  Every block reads the counter before it signals completion; the last block to
// finish performs the increment, so no block can see the new value.
__global__ void k_store_advance(Level A, Level Bv, int has_b, uint32_t *d_iter, unsigned int *done)
{
    const uint32_t iter = *d_iter;
    store_level(A, iter);
    if (has_b) store_level(Bv, iter);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0;
            *d_iter = iter + 1;
        }
    }
}

// the sweep decisions of a level into the persistent kernel's per-population flag lines (words 2 and 3 of `stride` ints)
__global__ void k_flags_init(Level L, unsigned int *flags, int stride)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= L.npop) return;
    flags[(size_t)p * stride + 2] = (unsigned int)L.mode[p];
    flags[(size_t)p * stride + 3] = L.mode[p] == 1 ? (unsigned int)L.mig_n[p] : 0u;
}
__global__ void k_stamp(unsigned long long *t) { *t = BlockTrace::now(); }

// ------------------------------------------------------------------------------------------------
// test / utility kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_sumlogprior(DevPrior P, const double *x, const double *p0, const double *p1, int n, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int D = P.npar;
    double a1 = 0.0, a2 = 0.0;
    for (int d = 0; d < D; ++d) {
        const double q0 = p0 ? p0[(size_t)i * D + d] : P.p0[d], q1 = p1 ? p1[(size_t)i * D + d] : P.p1[d];
        double v = dprior1(P.dist[d], x[(size_t)i * D + d], q0, q1, P.lower[d], P.upper[d], P.log_p[d] != 0);
        if (d & 1) a2 += v; else a1 += v;
    }
    out[i] = a1 + a2;
}

// get_chains / get_subchains from explicit uniforms (bit-exact index-selection check)
__global__ void k_select_chains(int C, int n, const int *k, const double *u_partner, int *out_partner, const double *u_mig,
                                int *out_mig, int *out_nmig)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (u_partner) {
        const double *u = u_partner + (size_t)i * (C - 1);
        int k1 = 0x7fffffff, p1 = 0x7fffffff, k2 = 0x7fffffff, p2 = 0x7fffffff;
        for (int j = 0; j < C - 1; ++j) {
            int key = shuffle_key(u[j]);
            if (key < k1 || (key == k1 && j < p1)) { k2 = k1; p2 = p1; k1 = key; p1 = j; }
            else if (key < k2 || (key == k2 && j < p2)) { k2 = key; p2 = j; }
        }
        out_partner[2 * i] = p1 + (p1 >= k[i]);
        out_partner[2 * i + 1] = p2 + (p2 >= k[i]);
    }
    if (u_mig) {
        const double *u = u_mig + (size_t)i * (C + 1);
        unsigned nn = (unsigned)ceil((double)C * u[0]);
        nn = nn < 2u ? 2u : nn;
        nn = nn > (unsigned)C ? (unsigned)C : nn;
        out_nmig[i] = (int)nn;
        int pos = 0;
        for (int j = 0; j < C; ++j) {
            int kj = shuffle_key(u[1 + j]), rank = 0;
            for (int q = 0; q < C; ++q) {
                int kq = shuffle_key(u[1 + q]);
                rank += (kq < kj) || (kq == kj && q < j);
            }
            if (rank < (int)nn) out_mig[(size_t)i * C + pos++] = j;
        }
        for (; pos < C; ++pos) out_mig[(size_t)i * C + pos] = -1;
    }
}

__global__ void k_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    U4 c = {ctr[0], ctr[1], ctr[2], ctr[3]};
    U4 r = philox4x32_10(c, key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// FP64 peak: 8 independent DFMA chains per thread
__global__ void k_dfma_peak(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

} // namespace gg
