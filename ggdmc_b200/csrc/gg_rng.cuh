// ggdmc_b200 -- counter-based uniform source (Philox4x32-10).
//
// The reference draws every uniform from R's global Mersenne-Twister stream through Rf_runif
// (src/RcppExports.cpp:19, src/de.cpp:65,88,131,...), which is inherently sequential.  Here every
// draw has an ADDRESS instead: (replicate seed, population, iteration, sweep, chain, purpose, slot)
// -> one 32-bit word of a Philox block, so any thread can produce any draw and all GPUs of a
// sharded fit produce identical phi-level draws without communicating.
//
//   key  = 64-bit seed of the replicate (config@seed)
//   ctr0 = slot >> 2            (word slot & 3 of the block is used)
//   ctr1 = purpose << 28 | (sweep & 0xFFF) << 16 | (chain & 0xFFFF)
//   ctr2 = population id (global subject index, 0xFFFFFFFF for phi)
//   ctr3 = iteration (1-based, like the loop variable of run_chains / run_hchains)
//   u    = (word + 0.5) * 2^-32  in (0, 1), 32-bit resolution like R's unif_rand
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GG_HD2 __device__ __forceinline__
#else
#define GG_HD2 inline
#endif

namespace gg {

enum : uint32_t { U_DECIDE = 0, U_PARTNER = 1, U_NOISE = 2, U_ST0 = 3, U_ACCEPT = 4, U_MIG_N = 5, U_MIG_KEYS = 6 };
constexpr uint32_t kPopPhi = 0xFFFFFFFFu;

struct U4 { uint32_t x, y, z, w; };

GG_HD2 uint32_t mulhi32(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

GG_HD2 U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 n;
        n.x = hi1 ^ c.y ^ k0;
        n.y = lo1;
        n.z = hi0 ^ c.w ^ k1;
        n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}

struct DrawAddr {
    uint64_t seed;
    uint32_t pop, iter, sweep, chain;
};

GG_HD2 double word_to_uniform(uint32_t w) { return ((double)w + 0.5) * (1.0 / 4294967296.0); }

// the Philox block holding slots 4*blk .. 4*blk+3 of one purpose
GG_HD2 U4 draw_block(const DrawAddr &a, uint32_t purpose, uint32_t blk)
{
    U4 c;
    c.x = blk;
    c.y = (purpose << 28) | ((a.sweep & 0xFFFu) << 16) | (a.chain & 0xFFFFu);
    c.z = a.pop;
    c.w = a.iter;
    return philox4x32_10(c, (uint32_t)(a.seed & 0xFFFFFFFFu), (uint32_t)(a.seed >> 32));
}

GG_HD2 uint32_t pick_word(const U4 &b, uint32_t i)
{
    return i == 0 ? b.x : (i == 1 ? b.y : (i == 2 ? b.z : b.w));
}

GG_HD2 double draw_uniform(const DrawAddr &a, uint32_t purpose, uint32_t slot)
{
    U4 b = draw_block(a, purpose, slot >> 2);
    return word_to_uniform(pick_word(b, slot & 3u));
}

// shuffle key of arma::shuffle as compiled into the reference (RcppArmadillo Alt_R_RNG.h:75):
// (int) Rf_runif(0, 2147483647)
GG_HD2 int shuffle_key(double u) { return (int)(2147483647.0 * u); }

} // namespace gg
