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

namespace {

using namespace gg;

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                                         \
    do {                                                                                                         \
        cudaError_t e_ = (expr);                                                                                 \
        if (e_ != cudaSuccess)                                                                                   \
            throw Error(GGDMC_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #expr);    \
    } while (0)

void require(bool ok, const char *msg)
{
    if (!ok) throw Error(GGDMC_ERR_ARG, msg);
}

// Device buffers come from the device's default stream-ordered memory pool with an unlimited release
// threshold: the first run* call pays for the allocations, later calls in the same process reuse the
// pooled memory (the reference re-creates all of its C++ objects on every .Call as well, but malloc is
// cheap there; cudaMalloc / cudaFree are not).  All pool operations are ordered on the legacy default
// stream; engines synchronise it once after construction.
inline void pool_setup(int device)
{
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}

template <class T>
struct DBuf { // device buffer
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
    void release()
    {
        if (p) cudaFreeAsync(p, 0);
        p = nullptr;
    }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (count) CUDA_CHECK(cudaMallocAsync(&p, count * sizeof(T), 0));
    }
    void zero() { if (n) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), 0)); }
    void upload(const T *h, size_t count)
    {
        alloc(count);
        if (count) CUDA_CHECK(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, 0));
    }
    void upload(const std::vector<T> &h)
    {
        upload(h.data(), h.size());
        CUDA_CHECK(cudaStreamSynchronize(0)); // the vector may die right after this call
    }
};

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen: single-GPU use has no NCCL dependency at all
// ---------------------------------------------------------------------------------------------
struct Nccl {
    typedef struct { char internal[128]; } UniqueId;
    typedef void *Comm;
    void *lib = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int, UniqueId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, Comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    Comm comm = nullptr;
    int n_rank = 1, rank = 0;

    void load()
    {
        if (lib) return;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) throw Error(GGDMC_ERR_COMM, std::string("cannot load libnccl: ") + dlerror());
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy)
            throw Error(GGDMC_ERR_COMM, "libnccl lacks required symbols");
    }
    void check(int r, const char *what)
    {
        if (r != 0)
            throw Error(GGDMC_ERR_COMM, std::string("NCCL error in ") + what + ": " + (GetErrorString ? GetErrorString(r) : "?"));
    }
};
Nccl g_nccl;

// Peer-memory window for the fused reduce + exchange kernel (k_hyper_reduce_exchange).  Set up once per
// communicator: every rank cudaMallocs a window, the CUDA IPC handles travel through one ncclAllGather,
// every rank maps its peers' windows.  GGDMC_B200_NO_P2P=1 keeps the plain NCCL all-reduce instead.
struct P2P {
    bool ready = false;
    void *base = nullptr;                 // local window
    void *peer_base[kP2PMaxRanks] = {};   // mapped peer windows (own entry = base)
    unsigned long long *seq = nullptr;
    int *status = nullptr;
    P2PWindow win{};
    static size_t slots_bytes(int n_rank) { return (size_t)2 * n_rank * kP2PMaxN * sizeof(double); }
    static size_t window_bytes(int n_rank) { return slots_bytes(n_rank) + (size_t)2 * kP2PMaxRanks * sizeof(unsigned long long); }

    void setup(Nccl &nc)
    {
        if (std::getenv("GGDMC_B200_NO_P2P") || nc.n_rank > kP2PMaxRanks || !nc.AllGather) return;
        const int n = nc.n_rank;
        const size_t bytes = window_bytes(n);
        if (cudaMalloc(&base, bytes) != cudaSuccess) { cudaGetLastError(); base = nullptr; return; }
        cudaMemset(base, 0, bytes);
        cudaIpcMemHandle_t mine;
        int ok = cudaIpcGetMemHandle(&mine, base) == cudaSuccess ? 1 : 0;
        // gather (ok flag + handle) of every rank
        struct Msg { int ok; cudaIpcMemHandle_t h; };
        Msg m{ok, mine};
        Msg *d_in = nullptr, *d_all = nullptr;
        std::vector<Msg> all(n);
        CUDA_CHECK(cudaMalloc(&d_in, sizeof(Msg)));
        CUDA_CHECK(cudaMalloc(&d_all, sizeof(Msg) * n));
        CUDA_CHECK(cudaMemcpy(d_in, &m, sizeof(Msg), cudaMemcpyHostToDevice));
        nc.check(nc.AllGather(d_in, d_all, sizeof(Msg), /*ncclInt8*/ 0, nc.comm, 0), "ncclAllGather");
        CUDA_CHECK(cudaStreamSynchronize(0));
        CUDA_CHECK(cudaMemcpy(all.data(), d_all, sizeof(Msg) * n, cudaMemcpyDeviceToHost));
        cudaFree(d_in);
        cudaFree(d_all);
        bool good = true;
        for (int r = 0; r < n; ++r) good = good && all[r].ok;
        if (good) {
            for (int r = 0; r < n && good; ++r) {
                if (r == nc.rank) { peer_base[r] = base; continue; }
                if (cudaIpcOpenMemHandle(&peer_base[r], all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    good = false;
                }
            }
        }
        // everybody must agree, otherwise some ranks would wait on flags nobody raises
        int *d_flag = nullptr, *d_flags = nullptr;
        int mine_ok = good ? 1 : 0;
        std::vector<int> oks(n);
        CUDA_CHECK(cudaMalloc(&d_flag, sizeof(int)));
        CUDA_CHECK(cudaMalloc(&d_flags, sizeof(int) * n));
        CUDA_CHECK(cudaMemcpy(d_flag, &mine_ok, sizeof(int), cudaMemcpyHostToDevice));
        nc.check(nc.AllGather(d_flag, d_flags, sizeof(int), 0, nc.comm, 0), "ncclAllGather");
        CUDA_CHECK(cudaStreamSynchronize(0));
        CUDA_CHECK(cudaMemcpy(oks.data(), d_flags, sizeof(int) * n, cudaMemcpyDeviceToHost));
        cudaFree(d_flag);
        cudaFree(d_flags);
        for (int r = 0; r < n; ++r) good = good && oks[r];
        if (!good) { teardown(nc.rank, n); return; }
        CUDA_CHECK(cudaMalloc(&seq, sizeof(unsigned long long)));
        CUDA_CHECK(cudaMalloc(&status, sizeof(int)));
        CUDA_CHECK(cudaMemset(seq, 0, sizeof(unsigned long long)));
        CUDA_CHECK(cudaMemset(status, 0, sizeof(int)));
        for (int r = 0; r < n; ++r) {
            win.slots[r] = reinterpret_cast<double *>(peer_base[r]);
            win.flags[r] = reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(peer_base[r]) + slots_bytes(n));
        }
        win.seq = seq; win.status = status; win.n_rank = n; win.rank = nc.rank;
        win.spin_ns = peer_timeout_ns();
        ready = true;
    }
    void teardown(int rank, int n)
    {
        for (int r = 0; r < n; ++r)
            if (r != rank && peer_base[r]) cudaIpcCloseMemHandle(peer_base[r]);
        for (auto &p : peer_base) p = nullptr;
        if (base) cudaFree(base);
        if (seq) cudaFree(seq);
        if (status) cudaFree(status);
        base = nullptr; seq = nullptr; status = nullptr;
        ready = false;
    }
    int timed_out()
    {
        int v = 0;
        if (status) cudaMemcpy(&v, status, sizeof(int), cudaMemcpyDeviceToHost);
        return v;
    }
    // how long a rank waits for its peers inside an exchange before it gives up (seconds, GGDMC_B200_PEER_TIMEOUT_S)
    static unsigned long long peer_timeout_ns()
    {
        double sec = 120.0;
        if (const char *e = std::getenv("GGDMC_B200_PEER_TIMEOUT_S")) sec = std::max(0.001, std::atof(e));
        return (unsigned long long)(sec * 1e9);
    }
};
P2P g_p2p;

// ---------------------------------------------------------------------------------------------
// uploads
// ---------------------------------------------------------------------------------------------
struct ModelDev {
    DBuf<int> param_src, row_src;
    DBuf<double> const_val;
    DBuf<uint8_t> posdrift;
    DBuf<uint16_t> row_of;
    DevModel d{};
    int type = GGDMC_MODEL_LBA; // enum ggdmc_model_type: which likelihood kernels the host launches
    void upload(const ggdmc_model_t *m)
    {
        require(m && m->n_acc >= 1 && m->n_acc <= 16 && m->n_cell >= 1 && m->npar >= 1, "bad model dimensions");
        require(m->n_cell < 65535, "too many cells");
        require(m->type == GGDMC_MODEL_LBA || m->type == GGDMC_MODEL_DDM, "Undefined model type"); // @hdr/likelihood.h:312
        const bool ddm = m->type == GGDMC_MODEL_DDM;
        const size_t n = (size_t)m->n_cell * (ddm ? GGDMC_DDM_ROWS : GGDMC_LBA_ROWS) * m->n_acc;
        for (size_t i = 0; i < n; ++i) {
            const int s = m->param_src[i];
            require(s >= 0 ? s < m->npar : (-1 - s) < m->n_const, "param_src out of range");
        }
        param_src.upload(m->param_src, n);
        std::vector<double> cv(m->const_val, m->const_val + std::max(m->n_const, 0));
        if (cv.empty()) cv.push_back(0.0);
        const_val.upload(cv);
        posdrift.upload(m->posdrift, ddm ? m->n_cell : m->n_acc); // the DDM path indexes it by cell (@hdr/likelihood.h:142)
        d.n_acc = m->n_acc; d.n_cell = m->n_cell; d.npar = m->npar; d.n_const = m->n_const;
        d.param_src = param_src.p; d.const_val = const_val.p; d.posdrift = posdrift.p;
        type = m->type;
        if (!ddm) build_rows_table(m);
    }
    // The distinct (cell, accumulator) rows of an LBA model: entries with the same six parameter sources and the same
    // drift rule share one row of the likelihood kernels' table.  If st0 can be non-zero every entry draws its own
    // uniform (`t0 + st0 U`, @hdr/lba.h:117) and keeps its own row.
    void build_rows_table(const ggdmc_model_t *m)
    {
        const int na = m->n_acc, n_ent = m->n_cell * na;
        bool st0_zero = true;
        for (int c = 0; c < m->n_cell && st0_zero; ++c)
            for (int j = 0; j < na; ++j) {
                const int s = m->param_src[((size_t)c * GGDMC_LBA_ROWS + 4) * na + j];
                if (s >= 0 || m->const_val[-1 - s] != 0.0) { st0_zero = false; break; }
            }
        std::vector<uint16_t> h_row_of((size_t)n_ent);
        std::vector<int> h_row_src;
        for (int c = 0; c < m->n_cell; ++c)
            for (int j = 0; j < na; ++j) {
                int key[8];
                for (int r = 0; r < 6; ++r) key[r] = m->param_src[((size_t)c * GGDMC_LBA_ROWS + r) * na + j];
                key[6] = c * na + j;
                key[7] = m->posdrift[j] != 0;
                int found = -1;
                const int n_row = (int)h_row_src.size() / 8;
                if (st0_zero)
                    for (int q = 0; q < n_row && found < 0; ++q) {
                        const int *o = &h_row_src[(size_t)q * 8];
                        bool same = o[7] == key[7];
                        for (int r = 0; r < 6 && same; ++r) same = o[r] == key[r];
                        if (same) found = q;
                    }
                if (found < 0) {
                    found = n_row;
                    h_row_src.insert(h_row_src.end(), key, key + 8);
                }
                h_row_of[(size_t)c * na + j] = (uint16_t)found;
            }
        require(h_row_src.size() / 8 <= 65535, "too many table rows");
        row_of.upload(h_row_of);
        row_src.upload(h_row_src);
        d.n_row = (int)h_row_src.size() / 8;
        d.row_of = row_of.p;
        d.row_src = row_src.p;
    }
};

struct PriorDev {
    DBuf<double> p0, p1, lower, upper;
    DBuf<int> dist;
    DBuf<uint8_t> log_p;
    DevPrior d{};
    void upload(const ggdmc_prior_t *p)
    {
        require(p && p->npar >= 1, "bad prior");
        p0.upload(p->p0, p->npar); p1.upload(p->p1, p->npar);
        lower.upload(p->lower, p->npar); upper.upload(p->upper, p->npar);
        dist.upload(p->dist, p->npar); log_p.upload(p->log_p, p->npar);
        d.npar = p->npar; d.p0 = p0.p; d.p1 = p1.p; d.lower = lower.p; d.upper = upper.p; d.dist = dist.p; d.log_p = log_p.p;
    }
};

// Trials of all local subjects: grouped by cell (stable), each subject padded to a multiple of 8
// trials with cell = 0xFFFF so that 16-byte vector loads never cross into the next subject.
struct TrialsDev {
    DBuf<double> rt;
    DBuf<uint16_t> cell;
    DBuf<int64_t> offset;
    DBuf<int> count;
    DBuf<unsigned long long> counter;
    std::vector<int> h_count;
    std::vector<std::vector<int>> order; // per subject: position in the grouped array -> caller's trial index
    int S = 0, max_count = 0;
    int64_t total = 0;
    TrialData d{};
    // Already grouped by cell, every subject a multiple of 8 trials: the caller's arrays ARE the device layout
    // (one validation pass, then two copies straight from the caller's memory, no staging).
    bool upload_direct(const ggdmc_trials_t *t, int n_cell, std::vector<int64_t> &off)
    {
        const int64_t base = t->subject_offset[0];
        int mx = 0;
        for (int s = 0; s < S; ++s) {
            const int64_t b = t->subject_offset[s], e = t->subject_offset[s + 1];
            if (e < b || e - b >= ((int64_t)1 << 31) || ((e - b) & 7) != 0) return false;
            const uint16_t *c = t->cell + b;
            const int n = (int)(e - b);
            unsigned prev = 0, bad = 0;
            for (int i = 0; i < n; ++i) {
                bad |= (unsigned)(c[i] >= n_cell) | (unsigned)(c[i] < prev);
                prev = c[i];
            }
            if (bad) return false; // out of range (reported by the general path) or not grouped
            off[s] = b - base;
            h_count[s] = n;
            mx = std::max(mx, n);
        }
        max_count = mx;
        total = t->subject_offset[S] - base;
        rt.upload(t->rt + base, (size_t)total);
        cell.upload(t->cell + base, (size_t)total);
        offset.upload(off); count.upload(h_count);
        counter.alloc(1); counter.zero();
        d.rt = rt.p; d.cell = cell.p; d.offset = offset.p; d.count = count.p; d.counter = counter.p;
        return true;
    }

    // sort_rt (model type "fastdm"): within a cell the trials are additionally ordered by response time, so that the
    // 32 trials of a warp need similar series lengths and take the same small-time / large-time branch
    void upload(const ggdmc_trials_t *t, int n_cell, bool keep_order, bool sort_rt = false)
    {
        require(t && t->n_subject >= 1, "no subjects");
        S = t->n_subject;
        std::vector<int64_t> off(S);
        h_count.resize(S);
        if (!keep_order && !sort_rt && upload_direct(t, n_cell, off)) return;
        std::vector<double> hrt;
        std::vector<uint16_t> hcl;
        if (keep_order) order.resize(S);
        {
            const int64_t ntot = t->subject_offset[S] - t->subject_offset[0];
            hrt.reserve((size_t)ntot + 8 * (size_t)S);
            hcl.reserve((size_t)ntot + 8 * (size_t)S);
        }
        int64_t pos = 0;
        for (int s = 0; s < S; ++s) {
            const int64_t b = t->subject_offset[s], e = t->subject_offset[s + 1];
            require(e >= b && e - b < (int64_t)1 << 31, "bad subject_offset");
            const int n = (int)(e - b);
            // group by cell: dmi@data usually arrives grouped already (then it is a straight copy), otherwise a
            // stable counting sort
            bool sorted = true;
            for (int i = 0; i < n; ++i) {
                require(t->cell[b + i] < n_cell, "cell index out of range");
                if (i > 0 && t->cell[b + i] < t->cell[b + i - 1]) sorted = false;
            }
            std::vector<int> idx;
            if (!sorted || keep_order || sort_rt) {
                idx.resize(n);
                std::vector<int> start((size_t)n_cell + 1, 0);
                for (int i = 0; i < n; ++i) ++start[t->cell[b + i] + 1];
                for (int c = 0; c < n_cell; ++c) start[c + 1] += start[c];
                const std::vector<int> first(start);
                for (int i = 0; i < n; ++i) idx[start[t->cell[b + i]]++] = i;
                if (sort_rt) {
                    const double *r = t->rt + b;
                    for (int c = 0; c < n_cell; ++c)
                        std::stable_sort(idx.begin() + first[c], idx.begin() + first[c + 1], [r](int x, int y) { return r[x] < r[y]; });
                }
            }
            off[s] = pos;
            h_count[s] = n;
            max_count = std::max(max_count, n);
            const int npad = (n + 7) & ~7;
            hrt.resize(pos + npad, 0.0);
            hcl.resize(pos + npad, 0xFFFF);
            if (idx.empty()) {
                std::memcpy(&hrt[pos], t->rt + b, sizeof(double) * (size_t)n);
                std::memcpy(&hcl[pos], t->cell + b, sizeof(uint16_t) * (size_t)n);
            } else {
                for (int i = 0; i < n; ++i) {
                    hrt[pos + i] = t->rt[b + idx[i]];
                    hcl[pos + i] = t->cell[b + idx[i]];
                }
            }
            if (keep_order) order[s] = idx;
            pos += npad;
            total += n;
        }
        rt.upload(hrt); cell.upload(hcl); offset.upload(off); count.upload(h_count);
        counter.alloc(1); counter.zero();
        d.rt = rt.p; d.cell = cell.p; d.offset = offset.p; d.count = count.p; d.counter = counter.p;
    }
    void set_chunking(int64_t blocks_per_split_unit)
    {
        // enough blocks to fill 148 SMs a few times over, at least 256 trials per block
        int want = (int)std::max<int64_t>(1, (4 * 148 + blocks_per_split_unit - 1) / blocks_per_split_unit);
        int max_split = std::max(1, (max_count + 255) / 256);
        int nsplit = std::min(want, max_split);
        if (const char *e = std::getenv("GGDMC_B200_NSPLIT")) nsplit = std::max(1, std::min(std::atoi(e), std::max(1, max_count / 8))); // experiments
        if (max_count > 8192) nsplit = std::max(nsplit, (max_count + 4095) / 4096);
        int chunk = ((std::max(1, (max_count + nsplit - 1) / nsplit)) + 7) & ~7;
        nsplit = std::max(1, (max_count + chunk - 1) / chunk);
        d.chunk = chunk;
        d.nsplit = nsplit;
    }
};

// ---------------------------------------------------------------------------------------------
// one level of the sampler on the device
// ---------------------------------------------------------------------------------------------
struct LevelDev {
    DBuf<double> theta, lp, ll, prop, prop_lp, out_theta, out_lp, out_ll;
    DBuf<int> target, mode, mig_n, mig_list, para, mode0;
    Level L{};
    int n_rep = 1;
    void create(int npop, int n_rep_, int C, int D, int nmc, int thin)
    {
        n_rep = n_rep_;
        const size_t PC = (size_t)npop * C;
        theta.alloc(PC * D); lp.alloc(PC); ll.alloc(PC); prop.alloc(PC * D); prop_lp.alloc(PC);
        prop.zero(); prop_lp.zero();
        target.alloc(PC); mode.alloc(npop); mig_n.alloc(npop); mig_list.alloc(PC); para.alloc(npop); mode0.alloc(npop);
        CUDA_CHECK(cudaMemset(target.p, 0xFF, PC * sizeof(int)));
        mode.zero(); mig_n.zero(); mig_list.zero(); para.zero(); mode0.zero();
        out_theta.alloc(PC * D * nmc); out_lp.alloc(PC * nmc); out_ll.alloc(PC * nmc);
        L.npop = npop; L.nchain = C; L.npar = D; L.nmc = nmc; L.thin = thin;
        L.theta = theta.p; L.lp = lp.p; L.ll = ll.p; L.prop = prop.p; L.prop_lp = prop_lp.p;
        L.target = target.p; L.mode = mode.p; L.mig_n = mig_n.p; L.mig_list = mig_list.p; L.para = para.p; L.mode0 = mode0.p;
        L.out_theta = out_theta.p; L.out_lp = out_lp.p; L.out_ll = out_ll.p;
    }
};

int pick_device(int requested)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        throw Error(GGDMC_ERR_CUDA, "no CUDA device: ggdmc_b200 has no CPU fallback");
    int cur = requested;
    if (requested >= 0) {
        require(requested < n, "device ordinal out of range");
        CUDA_CHECK(cudaSetDevice(requested));
    } else {
        CUDA_CHECK(cudaGetDevice(&cur));
    }
    pool_setup(cur);
    return cur;
}

struct PhaseTimer { // GGDMC_B200_TIMING=1 prints host wall time per phase of a run* call to stderr
    bool on;
    std::chrono::steady_clock::time_point t;
    PhaseTimer() : on(std::getenv("GGDMC_B200_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what)
    {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[ggdmc_b200] %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};
// Streams are kept for the life of the process: creating one is a call into the kernel-mode driver (a channel allocation),
// which costs a millisecond on a quiet box and tens of milliseconds when anything else talks to the driver (a monitoring
// tool polling clocks is enough) -- measured inside ggdmc_b200_run, whose engine lives for one call.
struct StreamCache {
    struct Item { int device, prio; cudaStream_t s; };
    std::mutex mu;
    std::vector<Item> idle;
    cudaStream_t get(int device, int prio)
    {
        {
            std::lock_guard<std::mutex> g(mu);
            for (size_t i = 0; i < idle.size(); ++i)
                if (idle[i].device == device && idle[i].prio == prio) {
                    cudaStream_t s = idle[i].s;
                    idle.erase(idle.begin() + (long)i);
                    return s;
                }
        }
        cudaStream_t s = nullptr;
        CUDA_CHECK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio));
        return s;
    }
    void put(int device, int prio, cudaStream_t s)
    {
        if (!s) return;
        cudaStreamSynchronize(s);
        std::lock_guard<std::mutex> g(mu);
        idle.push_back(Item{device, prio, s});
    }
};
StreamCache g_streams;

template <class K>
void allow_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

constexpr int kHyperBlock = 256;
constexpr int kProposeWarps = 4;
constexpr int kAcceptWarps = 4;

// Launch shape of the likelihood kernel: 64 threads per block, 12 resident blocks per SM (80 registers) -- picked by
// measurement on B200 among (128, 6), (128, 8), (64, 8 / 10 / 12 / 16), (32, 24 / 32), (256, 3) in round 1
// (profiles/r01_k_like.md); the other shapes are no longer compiled into the library.
constexpr int kLikeBlock = 64, kLikeMinBlocks = 12;

size_t like_smem(const DevModel &M, int block) { return like_smem_bytes(M.n_row, M.n_cell, block); }

// The trial loop reaches a cell's rows either directly -- the distinct rows are expanded into one row per (cell, accumulator)
// after they are built -- or through the cell's row indices.  Expanded is one dependent shared-memory load shorter per
// accumulator and trial (2 % of the launch on the README model); indexed keeps the table small (the 96-cell, 4-accumulator
// model: 3 KB instead of 27 KB per block, 12 instead of 9 resident blocks).  Expanded while 12 blocks' tables stay below 64 KB.
bool like_expand(const DevModel &M) { return (size_t)(M.n_row + M.n_cell * M.n_acc) * sizeof(CellAcc) * kLikeMinBlocks <= 64 * 1024; }
size_t like_launch_smem(const DevModel &M, int block, bool expand)
{
    return like_smem(M, block) + (expand ? (size_t)M.n_cell * M.n_acc * sizeof(CellAcc) : (((size_t)M.n_cell * M.n_acc * sizeof(uint16_t) + 15) & ~(size_t)15));
}

template <int NACC, bool EXPAND>
void launch_like_t(const Level &L, const DevModel &M, const TrialData &T, const uint32_t *d_iter, int sweep, int step, int half,
                   double *ll_part, cudaStream_t st, const int *prio)
{
    constexpr int BLOCK = kLikeBlock, MINB = kLikeMinBlocks;
    const int per_pop = step >= 0 ? 1 : (half < 0 ? L.nchain : (L.nchain + 1) / 2);
    dim3 grid(L.npop * per_pop, T.nsplit);
    const size_t sm = like_launch_smem(M, BLOCK, EXPAND);
    require(sm <= 220 * 1024, "cell table does not fit in shared memory");
    allow_smem(k_like<NACC, BLOCK, MINB, EXPAND>, sm);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority; // dispatch order among the likelihood launches of concurrent subject groups
    at[0].val.priority = prio ? *prio : 0;
    cfg.attrs = at;
    cfg.numAttrs = prio ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_like<NACC, BLOCK, MINB, EXPAND>, L, M, T, d_iter, sweep, step, half, ll_part));
}

template <int NACC>
void launch_like_n(const Level &L, const DevModel &M, const TrialData &T, const uint32_t *d_iter, int sweep, int step, int half,
                   double *ll_part, cudaStream_t st, const int *prio)
{
    if (like_expand(M)) launch_like_t<NACC, true>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio);
    else launch_like_t<NACC, false>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio);
}

// model type "fastdm": same grid and arguments as k_like.  Launch shape (threads per block, minimum resident blocks per
// SM) picked by measurement (profiles/r01_k_like_ddm.md).
template <int BLOCK, int MINB>
void launch_like_ddm_t(const Level &L, const DevModel &M, const TrialData &T, const uint32_t *d_iter, int sweep, int step, int half,
                       double *ll_part, cudaStream_t st, const int *prio)
{
    const int per_pop = step >= 0 ? 1 : (half < 0 ? L.nchain : (L.nchain + 1) / 2);
    dim3 grid(L.npop * per_pop, T.nsplit);
    const size_t sm = ((size_t)M.n_cell * sizeof(DdmCell) + (size_t)(BLOCK / 32) * 8 + 15) & ~(size_t)15;
    require(sm <= 220 * 1024, "cell table does not fit in shared memory");
    allow_smem(k_like_ddm<BLOCK, MINB>, sm);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(BLOCK); cfg.dynamicSmemBytes = sm; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributePriority;
    at[0].val.priority = prio ? *prio : 0;
    cfg.attrs = at;
    cfg.numAttrs = prio ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_like_ddm<BLOCK, MINB>, L, M, T, d_iter, sweep, step, half, ll_part));
}

// launch shape of the DDM kernel: 128 threads x 6 blocks per SM (80 registers), picked by measurement among eight shapes in round 1
// (profiles/r01_k_like_ddm.md: +20 % over 4 blocks; 8 blocks = 64 registers spill too much)
void launch_like_ddm(const Level &L, const DevModel &M, const TrialData &T, const uint32_t *d_iter, int sweep, int step, int half,
                     double *ll_part, cudaStream_t st, const int *prio)
{
    launch_like_ddm_t<128, 6>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio);
}

void launch_like(const Level &L, const ModelDev &MD, const TrialData &T, const uint32_t *d_iter, int sweep, int step, int half,
                 double *ll_part, cudaStream_t st, const int *prio = nullptr)
{
    const DevModel &M = MD.d;
    if (MD.type == GGDMC_MODEL_DDM) {
        launch_like_ddm(L, M, T, d_iter, sweep, step, half, ll_part, st, prio);
        return;
    }
    switch (M.n_acc) {
    case 2: launch_like_n<2>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio); break;
    case 3: launch_like_n<3>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio); break;
    case 4: launch_like_n<4>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio); break;
    default: launch_like_n<0>(L, M, T, d_iter, sweep, step, half, ll_part, st, prio);
    }
}

// the parity probe runs the trial loop the sampler would run for this model: expanded or indexed table (like_expand)
template <int NACC, bool EXPAND>
void launch_trial_logdens_hot_t(const DevModel &M, const TrialData &T, const double *theta, int n_theta, int ntr, uint64_t seed, uint32_t pop,
                                uint32_t iter, double *out, double *sums)
{
    const size_t sm = like_launch_smem(M, 64, EXPAND);
    require(sm <= 220 * 1024, "cell table does not fit in shared memory");
    allow_smem(k_trial_logdens_hot<NACC, 64, EXPAND>, sm);
    k_trial_logdens_hot<NACC, 64, EXPAND><<<dim3(n_theta, T.nsplit), 64, sm>>>(M, T, theta, ntr, seed, pop, iter, out, sums);
    CUDA_CHECK(cudaGetLastError());
}
template <int NACC>
void launch_trial_logdens_hot(const DevModel &M, const TrialData &T, const double *theta, int n_theta, int ntr, uint64_t seed, uint32_t pop,
                              uint32_t iter, double *out, double *sums)
{
    if (like_expand(M)) launch_trial_logdens_hot_t<NACC, true>(M, T, theta, n_theta, ntr, seed, pop, iter, out, sums);
    else launch_trial_logdens_hot_t<NACC, false>(M, T, theta, n_theta, ntr, seed, pop, iter, out, sums);
}
} // namespace

// GGDMC_B200_TRACE=1: every launch of an iteration is bracketed by CUDA events on its own stream and the
// last iteration's timeline (start, duration, stream) goes to stderr -- a diagnostic, never a bench path.
struct Tracer {
    struct Rec { const char *name; int side; cudaEvent_t a, b; };
    bool on = std::getenv("GGDMC_B200_TRACE") != nullptr;
    std::vector<Rec> recs;
    size_t used = 0;
    void reset() { used = 0; }
    void open(const char *name, cudaStream_t st, int is_side)
    {
        if (!on) return;
        if (used == recs.size()) {
            Rec r{name, 0, nullptr, nullptr};
            cudaEventCreate(&r.a);
            cudaEventCreate(&r.b);
            recs.push_back(r);
        }
        recs[used].name = name;
        recs[used].side = is_side;
        cudaEventRecord(recs[used].a, st);
    }
    void close(cudaStream_t st)
    {
        if (!on) return;
        cudaEventRecord(recs[used].b, st);
        ++used;
    }
    void dump(int rank)
    {
        if (!on || used == 0) return;
        cudaDeviceSynchronize();
        std::fprintf(stderr, "[ggdmc_b200 trace] rank %d, last iteration: start_us dur_us stream kernel\n", rank);
        for (size_t i = 0; i < used; ++i) {
            float t0 = 0.f, d = 0.f;
            cudaEventElapsedTime(&t0, recs[0].a, recs[i].a);
            cudaEventElapsedTime(&d, recs[i].a, recs[i].b);
            std::fprintf(stderr, "[ggdmc_b200 trace] %9.1f %8.1f %s %s\n", t0 * 1e3, d * 1e3, recs[i].side == 1 ? "side" : recs[i].side == 0 ? "main" : "grp ", recs[i].name);
        }
    }
    ~Tracer()
    {
        for (Rec &r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    }
};
#define TR(name, st, ...) do { trace.open(name, st, stream_tag(st)); __VA_ARGS__; trace.close(st); } while (0)

// ---------------------------------------------------------------------------------------------
// the engine
// ---------------------------------------------------------------------------------------------
struct ggdmc_engine {
    Tracer trace;
    // kind: 0 independent subjects (run_subject), 1 hyper only (run_hyper), 2 hierarchy (run)
    int kind = 0;
    int device = 0;
    int R = 1, S = 0, C = 0, D = 0, D2 = 0, nmc = 0, thin = 1;
    int schedule = GGDMC_SCHEDULE_PARALLEL;
    int is_hblocked = 0, is_pblocked = 0;
    int subject_begin = 0;
    uint32_t h_iter = 0;
    int64_t launches = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // Hierarchy: the phi sweep of an iteration does not feed the subjects' proposals or likelihoods
    // (only their MH test, through the prior), so it runs on a high-priority side stream next to the
    // first likelihood launch and joins before the first k_accept.  GGDMC_B200_NO_OVERLAP=1 serialises.
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool overlap = std::getenv("GGDMC_B200_NO_OVERLAP") == nullptr;
    bool fuse_phi = std::getenv("GGDMC_B200_NO_FUSED_PHI") == nullptr;
    // The subjects are independent given phi, so they run as groups on their own streams: one group's
    // proposal / MH kernels and the drain of its likelihood launch overlap another group's likelihood.
    // GGDMC_B200_GROUPS=n overrides the group count (1 = one launch over all subjects).
    static constexpr int kMaxGroups = 8;
    struct SubjGroup { Level L; TrialData T; double *ll_part; int index; };
    std::vector<SubjGroup> groups;
    cudaStream_t gstream[kMaxGroups] = {};
    cudaEvent_t ev_gdone[kMaxGroups] = {}, ev_prop[kMaxGroups] = {}, ev_swept[kMaxGroups] = {}, ev_sb = nullptr;
    DBuf<uint32_t> sb_iter;     // [kMaxGroups] iteration counters of the groups' decision launches on the side stream
    DBuf<unsigned int> sb_done; // [kMaxGroups]
    // optional per-launch timing of the likelihood kernel (bench.py roofline)
    // one DE-MCMC iteration captured as a CUDA graph (fixed launch sequence: every data-dependent
    // decision is taken on the device); GGDMC_B200_NO_GRAPH=1 falls back to plain stream launches
    bool use_graph = std::getenv("GGDMC_B200_NO_GRAPH") == nullptr && std::getenv("GGDMC_B200_TRACE") == nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    int64_t launches_per_iter = 0;
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;
    size_t prof_used = 0;
    double like_ms = 0.0;
    int64_t like_launches = 0;

    ModelDev model;
    PriorDev p_prior, h_prior;
    TrialsDev trials;
    LevelDev subj, phi;
    DBuf<uint64_t> seeds;
    DBuf<uint32_t> d_iter;
    DBuf<unsigned int> done_ctr, phi_ticket;
    DBuf<double> ll_part, hpart, hsum, hyper_data, phi_consts, prop_consts;
    HyperArgs H{};

    ~ggdmc_engine()
    {
        PhaseTimer pt;
        if (stream) cudaStreamSynchronize(stream); // buffers go back to the pool right after this
        pt.lap("  ~sync");
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (graph) cudaGraphDestroy(graph);
        pt.lap("  ~graph");
        for (cudaEvent_t e : prof_ev) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        for (cudaEvent_t e : slot_ev) if (e) cudaEventDestroy(e);
        g_streams.put(device, prio_lo, copy_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_sb) cudaEventDestroy(ev_sb);
        for (cudaEvent_t e : ev_swept) if (e) cudaEventDestroy(e);
        for (int g = 1; g < kMaxGroups; ++g) {
            if (ev_gdone[g]) cudaEventDestroy(ev_gdone[g]);
            if (ev_prop[g - 1]) cudaEventDestroy(ev_prop[g - 1]);
            g_streams.put(device, prio_lo, gstream[g]);
        }
        g_streams.put(device, prio_hi, side);
        g_streams.put(device, prio_lo, stream);
        pt.lap("  ~stream");
    }

    void common_init(const ggdmc_config_t *cfg)
    {
        require(cfg != nullptr, "null config");
        if (cfg->nchain <= 2) throw Error(GGDMC_ERR_CHAINS, "Require three or more chains."); // src/de.cpp:7-10
        require(cfg->nchain <= 65535, "nchain too large");
        require(cfg->nmc >= 1 && cfg->thin >= 1, "nmc and thin must be >= 1");
        require(cfg->n_replicate >= 1 && cfg->seed != nullptr, "need n_replicate >= 1 seeds");
        require(cfg->schedule >= GGDMC_SCHEDULE_REFERENCE && cfg->schedule <= GGDMC_SCHEDULE_SIMULTANEOUS, "bad schedule");
        require(cfg->nparameter >= 1, "de_input nparameter must be >= 1");
        PhaseTimer pt;
        device = pick_device(cfg->device);
        pt.lap("   device");
        R = cfg->n_replicate; C = cfg->nchain; nmc = cfg->nmc; thin = cfg->thin;
        schedule = (cfg->schedule == GGDMC_SCHEDULE_PARALLEL && cfg->nchain < 4) ? GGDMC_SCHEDULE_REFERENCE : cfg->schedule;
        is_hblocked = cfg->is_hblocked; is_pblocked = cfg->is_pblocked;
        subject_begin = cfg->subject_begin;
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        stream = g_streams.get(device, prio_lo);
        side = g_streams.get(device, prio_hi);
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&ev_sb, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreate(&ev0));
        CUDA_CHECK(cudaEventCreate(&ev1));
        pt.lap("   streams");
        seeds.upload(cfg->seed, R);
        uint32_t z = 0;
        d_iter.upload(&z, 1); // 0 while the start state is stored in slot 0, then 1 = first iteration
        done_ctr.alloc(1);
        done_ctr.zero();
        phi_ticket.alloc(1);
        phi_ticket.zero();
        pt.lap("   counters");
    }

    void init_level_state(LevelDev &lv, const ggdmc_start_t *starts, int n_items, int D_, bool pool_synced = false)
    {
        // starts[i] holds [R][C][D_] for item i (subject or phi); device population p = i * R + r,
        // so item i's block is one contiguous copy
        const size_t CD = (size_t)C * D_, blk = (size_t)R * CD, blk1 = (size_t)R * C;
        if (!pool_synced) CUDA_CHECK(cudaStreamSynchronize(0)); // pool allocations (ordered on the default stream) are now usable on `stream`
        bool adjacent = true;
        for (int i = 0; i < n_items; ++i) {
            require(starts[i].theta && starts[i].lp && starts[i].ll, "null start state");
            if (i > 0 && (starts[i].theta != starts[i - 1].theta + blk || starts[i].lp != starts[i - 1].lp + blk1 ||
                          starts[i].ll != starts[i - 1].ll + blk1))
                adjacent = false;
        }
        if (adjacent) { // the caller's arrays are the device layout (the Python binding and the R glue allocate them that way)
            CUDA_CHECK(cudaMemcpyAsync(lv.theta.p, starts[0].theta, (size_t)n_items * blk * 8, cudaMemcpyHostToDevice, stream));
            CUDA_CHECK(cudaMemcpyAsync(lv.lp.p, starts[0].lp, (size_t)n_items * blk1 * 8, cudaMemcpyHostToDevice, stream));
            CUDA_CHECK(cudaMemcpyAsync(lv.ll.p, starts[0].ll, (size_t)n_items * blk1 * 8, cudaMemcpyHostToDevice, stream));
            store(lv);
            return;
        }
        std::vector<double> th((size_t)n_items * blk), lp((size_t)n_items * blk1), ll(lp.size());
        for (int i = 0; i < n_items; ++i) {
            std::memcpy(&th[i * blk], starts[i].theta, blk * 8);
            std::memcpy(&lp[i * blk1], starts[i].lp, blk1 * 8);
            std::memcpy(&ll[i * blk1], starts[i].ll, blk1 * 8);
        }
        CUDA_CHECK(cudaMemcpyAsync(lv.theta.p, th.data(), th.size() * 8, cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaMemcpyAsync(lv.lp.p, lp.data(), lp.size() * 8, cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaMemcpyAsync(lv.ll.p, ll.data(), ll.size() * 8, cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream)); // the staging vectors die here
        store(lv); // d_iter == 0: slot 0 of the storage = start state (@hdr/theta.h: slot 1 in R)
    }

    // ---- construction for the three run kinds ------------------------------------------------
    void create_lba(const ggdmc_model_t *m, const ggdmc_trials_t *t, const ggdmc_prior_t *pp, const ggdmc_prior_t *hp,
                    const ggdmc_config_t *cfg, const ggdmc_start_t *phi_start, const ggdmc_start_t *subj_start)
    {
        PhaseTimer pt;
        common_init(cfg);
        pt.lap("  init");
        kind = hp ? 2 : 0;
        model.upload(m);
        D = m->npar;
        require(pp && pp->npar == D, "p_prior length != model npar");
        p_prior.upload(pp);
        pt.lap("  model");
        require(t && t->n_subject >= 1, "no subjects");
        S = t->n_subject;
        // The two large uploads of a call -- the trials and the subjects' start state, both from pageable memory -- go side by
        // side: a second host thread stages the trials (default stream) while this one sends the start state (engine stream).
        subj.create(R * S, R, C, D, nmc, thin);
        CUDA_CHECK(cudaStreamSynchronize(0)); // pool allocations are usable on `stream` from here on
        pt.lap("  alloc");
        std::exception_ptr trials_err;
        std::thread trials_thread([&] {
            try {
                CUDA_CHECK(cudaSetDevice(device));
                trials.upload(t, m->n_cell, false, m->type == GGDMC_MODEL_DDM);
            } catch (...) {
                trials_err = std::current_exception();
            }
        });
        try {
            init_level_state(subj, subj_start, S, D, true);
        } catch (...) {
            trials_thread.join();
            throw;
        }
        trials_thread.join();
        if (trials_err) std::rethrow_exception(trials_err);
        pt.lap("  uploads");
        const bool want_persist = persist_planned = sampler_wanted(hp != nullptr) && m->type == GGDMC_MODEL_LBA && schedule == GGDMC_SCHEDULE_PARALLEL && !is_hblocked &&
                                  !is_pblocked && (!(hp && g_nccl.comm && g_nccl.n_rank > 1) || (g_p2p.ready && R * C * 2 <= kP2PMaxN)) &&
                                  sampler_fits(m->npar, hp != nullptr);
        if (want_persist) sampler_chunking((int64_t)R * S * ((C + 1) / 2));
        else trials.set_chunking((int64_t)R * S * C);
        Level &L = subj.L;
        L.n_rep = R; L.pop_id_base = subject_begin; L.is_phi = 0;
        L.gamma = cfg->gamma_precursor / std::sqrt(2.0 * cfg->nparameter); // src/de.cpp:12,24
        L.rp = cfg->rp; L.mig_prob = cfg->sub_migration_prob;
        L.seed = seeds.p; L.prior = p_prior.d; L.prior_ovr = nullptr;
        L.nmove = std::min(D, kind == 2 ? cfg->nparameter / 2 : cfg->nparameter); // src/de.cpp:136 / :592
        ll_part.alloc((size_t)R * S * C * trials.d.nsplit);
        ll_part.zero();
        if (kind == 2) {
            D2 = 2 * D;
            require(hp->npar == D2, "h_prior length != 2 * npar");
            h_prior.upload(hp);
            phi.create(R, R, C, D2, nmc, thin);
            Level &P = phi.L;
            P.n_rep = R; P.pop_id_base = 0; P.is_phi = 1;
            P.gamma = L.gamma; P.rp = cfg->rp; P.mig_prob = cfg->pop_migration_prob;
            P.seed = seeds.p; P.prior = h_prior.d; P.prior_ovr = nullptr;
            P.nmove = std::min(D2, cfg->nparameter);
            init_level_state(phi, phi_start, 1, D2);
            L.prior_ovr = phi.theta.p; // src/de.cpp:599-600, 646-649
            phi_consts.alloc((size_t)R * C * D * 2);
            L.ovr_consts = phi_consts.p;
            setup_hyper(subj.theta.p, C * D, R * C * D, D, 1);
            // The fused phi half-sweep (k_phi_half, and the persistent kernel) computes the prior constants of every proposed phi
            // vector on the way and copies them over the target chain's on accept: no k_phi_consts launch per iteration.
            const bool multi_rank = g_nccl.comm && g_nccl.n_rank > 1;
            consts_travel = want_persist || (schedule != GGDMC_SCHEDULE_REFERENCE && fuse_phi && !is_hblocked && (!multi_rank || (g_p2p.ready && R * C * 2 <= kP2PMaxN)));
            if (consts_travel) {
                prop_consts.alloc((size_t)R * C * D * 2);
                prop_consts.zero();
                H.prop_consts = prop_consts.p;
                H.consts = phi_consts.p;
            }
        }
        make_groups();
        start_counter();
        if (want_persist) setup_sampler();
        else if (consts_travel) phi_constants(stream); // the constants of the start state; later ones travel with accepted proposals
        if (kind == 2 && g_nccl.comm && g_nccl.n_rank > 1 && g_p2p.ready && g_p2p.timed_out())
            throw Error(GGDMC_ERR_COMM, "the communicator is in an error state (an earlier exchange timed out): call ggdmc_b200_comm_finalize and initialise it again");
        peer_barrier();
        pt.lap("  phi");
    }

    void create_hyper(const ggdmc_prior_t *pp, const ggdmc_prior_t *hp, const double *data_theta, int n_subject,
                      const ggdmc_config_t *cfg, const ggdmc_start_t *start)
    {
        common_init(cfg);
        kind = 1;
        require(pp && hp && data_theta && n_subject >= 1, "bad run_hyper arguments");
        D = pp->npar; D2 = 2 * D; S = n_subject;
        require(hp->npar == D2, "h_prior length != 2 * npar");
        p_prior.upload(pp);
        h_prior.upload(hp);
        hyper_data.upload(data_theta, (size_t)S * D);
        phi.create(R, R, C, D2, nmc, thin);
        Level &P = phi.L;
        P.n_rep = R; P.pop_id_base = 0; P.is_phi = 1;
        P.gamma = cfg->gamma_precursor / std::sqrt(2.0 * cfg->nparameter);
        P.rp = cfg->rp; P.mig_prob = cfg->sub_migration_prob; // run_chains uses m_sub_migration_prob, src/de.cpp:205-206
        P.seed = seeds.p; P.prior = h_prior.d; P.prior_ovr = nullptr;
        P.nmove = std::min(D2, cfg->nparameter);
        init_level_state(phi, start, 1, D2);
        setup_hyper(hyper_data.p, 0, D, 0, 0);
        start_counter();
    }

    // group g = local subjects [S g / G, S (g + 1) / G): views of the subject level, its trials and its partial sums
    void make_groups()
    {
        // two groups pay off once each group's likelihood launch fills the GPU by itself (measured: 32 subjects x 39 proposals
        // run 5 % faster as one group, 128 subjects 2 % faster as two)
        int G = (int64_t)R * S * ((C + 1) / 2) >= 2 * (int64_t)sm_count() * 12 ? 2 : 1;
        if (const char *e = std::getenv("GGDMC_B200_GROUPS")) G = std::atoi(e);
        G = std::max(1, std::min(std::min(G, S), kMaxGroups));
        gstream[0] = nullptr; // group 0 runs on `stream`
        groups.clear();
        for (int g = 0; g < G; ++g) {
            const int i0 = (int)((int64_t)S * g / G), i1 = (int)((int64_t)S * (g + 1) / G);
            const size_t p0 = (size_t)i0 * R;
            SubjGroup sg{subj.L, trials.d, ll_part.p + p0 * C * trials.d.nsplit, g};
            Level &L = sg.L;
            L.npop = (i1 - i0) * R;
            L.pop_id_base += i0;
            L.theta += p0 * C * D; L.prop += p0 * C * D;
            L.lp += p0 * C; L.ll += p0 * C; L.prop_lp += p0 * C; L.target += p0 * C; L.mig_list += p0 * C;
            L.mode += p0; L.mig_n += p0; L.para += p0; L.mode0 += p0;
            sg.T.offset += i0; sg.T.count += i0;
            groups.push_back(sg);
            CUDA_CHECK(cudaEventCreateWithFlags(&ev_swept[g], cudaEventDisableTiming));
            if (g > 0) {
                gstream[g] = g_streams.get(device, prio_lo);
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_gdone[g], cudaEventDisableTiming));
                CUDA_CHECK(cudaEventCreateWithFlags(&ev_prop[g - 1], cudaEventDisableTiming));
            }
        }
    }

    void start_counter()
    {
        CUDA_CHECK(cudaStreamSynchronize(stream)); // slot-0 stores (which read iteration 0) are done
        const uint32_t one = 1;
        CUDA_CHECK(cudaMemcpy(d_iter.p, &one, sizeof(one), cudaMemcpyHostToDevice));
        const std::vector<uint32_t> ones(kMaxGroups, 1u);
        sb_iter.upload(ones);
        sb_done.alloc(kMaxGroups);
        sb_done.zero();
        CUDA_CHECK(cudaStreamSynchronize(0));
    }

    void setup_hyper(const double *x, int rep_stride, int subj_stride, int chain_stride, int need_cur)
    {
        H.like = p_prior.d;
        H.x = x; H.x_rep_stride = rep_stride; H.x_subj_stride = subj_stride; H.x_chain_stride = chain_stride;
        H.S = S; H.D = D; H.need_cur = need_cur;
        // split subjects over blocks so that the phi kernels fill the GPU (R*C blocks alone would not) in ONE wave
        int per_sm = 4, n_sm = 148;
        const size_t sm_bytes = (size_t)(8 * D + 2 * (kHyperBlock / 32) + 4 * D) * 8;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_phi_half<kHyperBlock>, kHyperBlock, sm_bytes);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        int want = std::max(1, (std::max(per_sm, 1) * n_sm) / (R * C));
        int spb = std::max(64, (S + want - 1) / want); // >= 3 terms per thread: the per-block setup (proposal, 4 Phi + 2 log per parameter) is not free
        if (persist_planned) // one WARP per item in the sampler kernel, the phi half-sweep on the critical path of a small fit and its fixed
            spb = std::max(8, (S + 15) / 16); // cost (proposal, 4 Phi + 4 log per parameter) paid per item: at most 16 items per chain
        H.subj_per_block = spb;
        H.nsplit = (S + spb - 1) / spb;
        hpart.alloc((size_t)R * C * 2 * H.nsplit);
        hpart.zero();
        hsum.alloc((size_t)R * C * 2);
        hsum.zero();
    }

    // likelihood launch, optionally bracketed by CUDA events on the launching stream
    // Launch with the highest dispatch priority whatever the stream's own: the short proposal / MH kernels of a
    // subject group must not queue behind the not-yet-dispatched blocks of another group's likelihood launch.
    int prio_hi = 0, prio_lo = 0;
    template <typename... KArgs, typename... Args>
    void launch_hi(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributePriority;
        at[0].val.priority = prio_hi;
        cfg.attrs = at;
        cfg.numAttrs = hi_small ? 1 : 0;
        CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
    }
    bool hi_small = std::getenv("GGDMC_B200_NO_HI_SMALL") == nullptr;

    int stream_tag(cudaStream_t st) const { return st == side ? 1 : st == this->stream ? 0 : 2; }

    void timed_like(const SubjGroup &G, cudaStream_t stream, int sweep, int step, int half)
    {
        const Level &L = G.L;
        if (!profile) {
            // concurrent groups: likelihood launches are dispatched in pipeline order (group 0 half 0, group 1 half 0,
            // group 0 half 1, ...) instead of sharing the SMs in lock-step, so one group's short kernels and launch
            // ramp / drain fall under another group's likelihood
            int prio = std::min(prio_lo, prio_hi + 1 + std::max(half, 0) * (int)groups.size() + G.index);
            const bool staged = hi_small && groups.size() > 1 && stream_tag(stream) != 1;
            TR("k_like", stream, launch_like(L, model, G.T, d_iter.p, sweep, step, half, G.ll_part, stream, staged ? &prio : nullptr));
            return;
        }
        if (prof_used + 2 > prof_ev.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e;
                CUDA_CHECK(cudaEventCreate(&e));
                prof_ev.push_back(e);
            }
        }
        CUDA_CHECK(cudaEventRecord(prof_ev[prof_used], stream));
        launch_like(L, model, G.T, d_iter.p, sweep, step, half, G.ll_part, stream);
        CUDA_CHECK(cudaEventRecord(prof_ev[prof_used + 1], stream));
        prof_used += 2;
    }
    void collect_profile()
    {
        for (size_t i = 0; i + 1 < prof_used; i += 2) {
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, prof_ev[i], prof_ev[i + 1]));
            like_ms += ms;
            ++like_launches;
        }
        prof_used = 0;
    }

    // ---- one sweep at each level --------------------------------------------------------------
    // wait_first / rec_first: the groups' FIRST proposal kernels of an iteration run one after the other instead of side by
    // side, so that group 0's likelihood launch -- the first thing able to fill the GPU -- starts as early as possible
    void sweep_lba(const SubjGroup &G, cudaStream_t stream, int sweep, int decide_once, int para_idx, cudaEvent_t join = nullptr,
                   cudaEvent_t wait_first = nullptr, cudaEvent_t rec_first = nullptr, bool sb_aside = false)
    {
        const Level &L = G.L;
        const size_t prop_sm = (size_t)kProposeWarps * D * 8;
        // An unblocked hierarchical sweep draws its migration decision one iteration ahead: at the end of the previous
        // iteration's sweep of this group (below), where it overlaps other groups' likelihood launches, instead of in front
        // of this iteration's first proposal kernel.  Only the very first iteration draws its own.
        const bool ahead = sweep_ahead();
        if (!ahead || h_iter <= 1) {
            TR("k_sweep_begin", stream, launch_hi(k_sweep_begin, L.npop, 128, (size_t)2 * C * sizeof(int), stream, L, d_iter.p, sweep, decide_once, para_idx, 0, (uint32_t *)nullptr, (unsigned int *)nullptr));
            ++launches;
        }
        const int nslot_warps = L.npop * ((C + 1) / 2);
        if (schedule != GGDMC_SCHEDULE_REFERENCE) {
            const int n = L.npop * C;
            const int nhalf = schedule == GGDMC_SCHEDULE_PARALLEL ? 2 : 1;
            for (int h = 0; h < nhalf; ++h) {
                const int half = nhalf == 2 ? h : -1;
                const int nw = half < 0 ? n : nslot_warps; // warps: one per (population, chain) or per (population, slot)
                if (h == 0 && wait_first) CUDA_CHECK(cudaStreamWaitEvent(stream, wait_first, 0));
                TR("k_propose", stream, launch_hi(k_propose<kProposeWarps>, (nw + kProposeWarps - 1) / kProposeWarps, kProposeWarps * 32, prop_sm, stream, L, d_iter.p, sweep, -1, half));
                if (h == 0 && rec_first) CUDA_CHECK(cudaEventRecord(rec_first, stream));
                ++launches;
                timed_like(G, stream, sweep, -1, half);
                if (sb_aside && h + 1 == nhalf) {
                    // The next iteration's decisions, drawn on the side stream beside this half's MH tests instead of behind them: the
                    // likelihood launch was the last reader of this iteration's.  The launch counts iterations by itself (sb_iter),
                    // because the end-of-iteration kernel may advance the engine's counter while it runs.
                    CUDA_CHECK(cudaEventRecord(ev_swept[G.index], stream));
                    CUDA_CHECK(cudaStreamWaitEvent(side, ev_swept[G.index], 0));
                    TR("k_sweep_begin", side, launch_hi(k_sweep_begin, L.npop, 128, (size_t)2 * C * sizeof(int), side, L, d_iter.p, sweep, decide_once, para_idx, 1, sb_iter.p + G.index, sb_done.p + G.index));
                    ++launches;
                }
                if (join && h == 0) CUDA_CHECK(cudaStreamWaitEvent(stream, join, 0)); // the MH test needs this iteration's phi
                TR("k_accept", stream, launch_hi(k_accept<kAcceptWarps>, (nw + kAcceptWarps - 1) / kAcceptWarps, kAcceptWarps * 32, (size_t)kAcceptWarps * D * 8, stream, L, d_iter.p, sweep, -1, (const double *)G.ll_part, G.T.nsplit, half));
                launches += 2;
            }
        } else {
            for (int step = 0; step < C; ++step) {
                TR("k_propose", stream, launch_hi(k_propose<kProposeWarps>, (L.npop + kProposeWarps - 1) / kProposeWarps, kProposeWarps * 32, prop_sm, stream, L, d_iter.p, sweep, step, -1));
                timed_like(G, stream, sweep, step, -1);
                if (join && step == 0) CUDA_CHECK(cudaStreamWaitEvent(stream, join, 0));
                TR("k_accept", stream, launch_hi(k_accept<kAcceptWarps>, (L.npop + kAcceptWarps - 1) / kAcceptWarps, kAcceptWarps * 32, (size_t)kAcceptWarps * D * 8, stream, L, d_iter.p, sweep, step, (const double *)G.ll_part, G.T.nsplit, -1));
                launches += 3;
            }
        }
        if (ahead && !sb_aside) {
            TR("k_sweep_begin", stream, launch_hi(k_sweep_begin, L.npop, 128, (size_t)2 * C * sizeof(int), stream, L, d_iter.p, sweep, decide_once, para_idx, 1, (uint32_t *)nullptr, (unsigned int *)nullptr));
            ++launches;
        }
        CUDA_CHECK(cudaGetLastError());
    }
    // sweep decisions one iteration ahead: hierarchical fits without per-parameter sweeps (GGDMC_B200_NO_SWEEP_AHEAD=1: off)
    bool sweep_ahead() const { return kind == 2 && !is_pblocked && !is_hblocked && sweep_ahead_ok; }
    bool sweep_ahead_ok = std::getenv("GGDMC_B200_NO_SWEEP_AHEAD") == nullptr;

    // the subject-level sweep(s) of one iteration, group by group (ev_fork has been recorded on `stream`)
    void sweep_groups(int decide_once, cudaEvent_t join, bool conc)
    {
        const int nsweep = is_pblocked ? (kind == 2 ? D : subj.L.nmove) : 1;
        for (size_t g = 0; g < groups.size(); ++g) {
            cudaStream_t st = (conc && g > 0) ? gstream[g] : stream;
            if (st != stream) CUDA_CHECK(cudaStreamWaitEvent(st, ev_fork, 0));
            for (int p = 0; p < nsweep; ++p) {
                const bool first = p == 0 && conc && groups.size() > 1;
                sweep_lba(groups[g], st, p, decide_once, is_pblocked ? p : -1, p == 0 ? join : nullptr,
                          first && g > 0 ? ev_prop[g - 1] : nullptr, first && g + 1 < groups.size() ? ev_prop[g] : nullptr, sb_aside(conc));
            }
            if (st != stream) {
                CUDA_CHECK(cudaEventRecord(ev_gdone[g], st));
                CUDA_CHECK(cudaStreamWaitEvent(stream, ev_gdone[g], 0));
            }
        }
    }
    // the groups' next-iteration decisions run on the side stream (sweep_lba): the PARALLEL schedule of an unblocked hierarchy
    bool sb_aside(bool conc) const { return conc && sweep_ahead() && schedule == GGDMC_SCHEDULE_PARALLEL && sb_aside_ok; }
    bool sb_aside_ok = std::getenv("GGDMC_B200_NO_SB_ASIDE") == nullptr;
    void join_groups(bool conc)
    {
        if (!sb_aside(conc)) return;
        CUDA_CHECK(cudaEventRecord(ev_sb, side));
        CUDA_CHECK(cudaStreamWaitEvent(stream, ev_sb, 0));
    }

    void hyper_eval(int step, cudaStream_t st)
    {
        Level &P = phi.L;
        const size_t sm = (size_t)(8 * D + 2 * (kHyperBlock / 32)) * 8;
        dim3 grid(step < 0 ? R * C : R, H.nsplit, step < 0 ? 1 : 2);
        TR("k_hyper", st, k_hyper<kHyperBlock><<<grid, kHyperBlock, sm, st>>>(P, H, step, hpart.p));
        const int n = R * C * 2;
        const bool multi = kind == 2 && g_nccl.comm && g_nccl.n_rank > 1;
        if (multi && g_p2p.ready && n <= kP2PMaxN) {
            // the one exchange of the path, fused with the local reduction (peer-memory stores over NVLink)
            TR("k_hyper_reduce_exchange", st, k_hyper_reduce_exchange<<<1, 256, 0, st>>>(hpart.p, n, H.nsplit, hsum.p, g_p2p.win));
            launches += 2;
        } else {
            TR("k_hyper_reduce", st, k_hyper_reduce<<<(n + 127) / 128, 128, 0, st>>>(hpart.p, n, H.nsplit, hsum.p));
            launches += 2;
            if (multi) // fallback: partial sums over the local subjects -> sums over all subjects by NCCL
                g_nccl.check(g_nccl.AllReduce(hsum.p, hsum.p, (size_t)n, /*ncclDouble*/ 8, /*ncclSum*/ 0, g_nccl.comm, st),
                             "ncclAllReduce");
        }
    }

    void sweep_phi(int sweep, int decide_once, int para_idx, cudaStream_t st)
    {
        Level &P = phi.L;
        const size_t prop_sm = (size_t)kProposeWarps * D2 * 8;
        const int need_cur = H.need_cur;
        const bool ahead = sweep_ahead();
        if (!ahead || h_iter <= 1) {
            TR("k_sweep_begin", st, k_sweep_begin<<<R, 128, (size_t)2 * C * sizeof(int), st>>>(P, d_iter.p, sweep, decide_once, para_idx, 0));
            ++launches;
        }
        const bool multi = kind == 2 && g_nccl.comm && g_nccl.n_rank > 1;
        const bool p2p = multi && g_p2p.ready && R * C * 2 <= kP2PMaxN;
        if (schedule != GGDMC_SCHEDULE_REFERENCE && fuse_phi && (!multi || p2p)) {
            // one launch per half-sweep: proposal + hyper-likelihood + reduction (+ peer exchange) + MH test
            const int nhalf = schedule == GGDMC_SCHEDULE_PARALLEL ? 2 : 1;
            const size_t sm = (size_t)(8 * D + 2 * (kHyperBlock / 32) + 2 * D2) * 8;
            for (int h = 0; h < nhalf; ++h) {
                dim3 grid(R * C, H.nsplit);
                TR("k_phi_half", st, k_phi_half<kHyperBlock><<<grid, kHyperBlock, sm, st>>>(P, H, d_iter.p, sweep, nhalf == 2 ? h : -1, hpart.p,
                                                                                           hsum.p, phi_ticket.p, g_p2p.win, p2p ? 1 : 0));
                ++launches;
            }
        } else if (schedule != GGDMC_SCHEDULE_REFERENCE) {
            const int n = R * C;
            const int nhalf = schedule == GGDMC_SCHEDULE_PARALLEL ? 2 : 1;
            for (int h = 0; h < nhalf; ++h) {
                const int half = nhalf == 2 ? h : -1;
                const int nw = half < 0 ? n : R * ((C + 1) / 2);
                TR("k_propose", st, k_propose<kProposeWarps><<<(nw + kProposeWarps - 1) / kProposeWarps, kProposeWarps * 32, prop_sm, st>>>(P, d_iter.p, sweep, -1, half));
                hyper_eval(-1, st);
                TR("k_phi_accept", st, k_phi_accept<<<(n + 127) / 128, 128, 0, st>>>(P, d_iter.p, sweep, -1, hsum.p, need_cur, p2p_status()));
                launches += 2;
            }
        } else {
            for (int step = 0; step < C; ++step) {
                TR("k_propose", st, k_propose<kProposeWarps><<<(R + kProposeWarps - 1) / kProposeWarps, kProposeWarps * 32, prop_sm, st>>>(P, d_iter.p, sweep, step, -1));
                hyper_eval(step, st);
                TR("k_phi_accept", st, k_phi_accept<<<(R + 127) / 128, 128, 0, st>>>(P, d_iter.p, sweep, step, hsum.p, need_cur, p2p_status()));
                launches += 2;
            }
        }
        if (ahead) { // the next iteration's decision, behind this iteration's phi step instead of in front of the next one's
            TR("k_sweep_begin", st, k_sweep_begin<<<R, 128, (size_t)2 * C * sizeof(int), st>>>(P, d_iter.p, sweep, decide_once, para_idx, 1));
            ++launches;
        }
        CUDA_CHECK(cudaGetLastError());
    }

    const int *p2p_status() const { return (kind == 2 && g_nccl.comm && g_nccl.n_rank > 1 && g_p2p.ready) ? g_p2p.status : nullptr; }

    void phi_constants(cudaStream_t st)
    {
        const int n = R * C * D;
        TR("k_phi_consts", st, k_phi_consts<<<(n + 127) / 128, 128, 0, st>>>(phi.L, p_prior.d, D, phi_consts.p));
        ++launches;
    }

    void store(LevelDev &lv)
    {
        const size_t total = (size_t)lv.L.npop * C * lv.L.npar;
        int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
        TR("k_store", stream, k_store<<<blocks, 256, 0, stream>>>(lv.L, d_iter.p));
        ++launches;
    }

    // end of an iteration: thinned storage of every level + device-side iteration counter advance, one kernel
    void store_and_advance(LevelDev &a, LevelDev *b)
    {
        size_t total = (size_t)a.L.npop * C * a.L.npar;
        if (b) total = std::max(total, (size_t)b->L.npop * C * b->L.npar);
        const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
        TR("k_store_advance", stream, k_store_advance<<<blocks, 256, 0, stream>>>(a.L, b ? b->L : a.L, b ? 1 : 0, d_iter.p, done_ctr.p));
        ++launches;
    }

    // ---- persistent sampler kernel (gg_sampler.cuh): the PARALLEL schedule of an LBA fit, whole iterations per launch -----
    // GGDMC_B200_NO_PERSIST=1 keeps the multi-launch path (also used by the other schedules, per-parameter sweeps and the DDM).
    bool persist = false, persist_planned = false, consts_travel = false;
    SamplerArgs SA{};
    int sampler_grid = 0, sampler_threads = 0, sampler_nacc = 0, sampler_max_batch = 64;
    size_t sampler_smem = 0;
    DBuf<unsigned long long> sy_all_done, sy_trace, sy_urgent;
    DBuf<unsigned int> sy_close_list, sy_queue, sy_exit, sy_pop_flags, sy_chain_arrive, sy_phi_arrive, sy_phi_done;
    DBuf<int> sy_abort;

    template <int NACC>
    int sampler_blocks_per_sm()
    {
        auto kern = SA.hier ? k_sampler<NACC, true> : k_sampler<NACC, false>;
        allow_smem(kern, sampler_smem);
        int per_sm = 0;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, sampler_threads, sampler_smem));
        return per_sm;
    }
    template <int NACC>
    void sampler_launch()
    {
        if (SA.hier) CUDA_CHECK(cudaLaunchKernelEx(&sampler_cfg, k_sampler<NACC, true>, SA));
        else CUDA_CHECK(cudaLaunchKernelEx(&sampler_cfg, k_sampler<NACC, false>, SA));
    }
    cudaLaunchConfig_t sampler_cfg{};

    // The persistent kernel is the default where it is the faster path on B200 (profiles/r02_sampler.md): fits without a phi
    // level (run_subject: a 3-replicate README fit takes 133 ms instead of 205 ms).  For a hierarchy the launch sequence
    // still wins at every measured size -- its proposal / MH kernels hide their memory latency behind tens of thousands of
    // warps, a persistent worker pays it item by item -- so there it is opt-in: GGDMC_B200_PERSIST=1.
    // GGDMC_B200_NO_PERSIST=1 forces the launch sequence everywhere.
    static bool sampler_wanted(bool hier)
    {
        if (std::getenv("GGDMC_B200_NO_PERSIST")) return false;
        if (const char *e = std::getenv("GGDMC_B200_PERSIST")) return std::atoi(e) != 0;
        return !hier;
    }
    int sm_count() const
    {
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
        return n_sm;
    }
    size_t sampler_cta_bytes(int D_, bool hier, int warps) const
    {
        return sampler_stage_bytes(model.d.n_cell, model.d.n_acc, model.d.n_row, model.d.n_const) +
               (size_t)warps * sampler_warp_bytes(model.d.n_cell, model.d.n_row, D_, C, hier ? 1 : 0);
    }
    // The kernel wants its 24 warps per SM; a model whose row table leaves room for fewer than 16 stays on the multi-launch path.
    bool sampler_fits(int D_, bool hier) const { return 2 * sampler_cta_bytes(D_, hier, 8) <= 220 * 1024; }

    // Trial chunks per proposal: a warp per (proposal, chunk).  Large fits: one chunk (the table build is paid once per
    // proposal).  Small fits: as many chunks as it takes to give every resident warp of the GPU an item in each half-sweep,
    // down to 64 trials (one pass of a warp) per chunk.
    void sampler_chunking(int64_t proposals_per_half)
    {
        const int64_t cap = (int64_t)sm_count() * 24;
        int nsplit = (int)std::max<int64_t>(1, cap / std::max<int64_t>(1, proposals_per_half));
        nsplit = std::min(nsplit, std::max(1, trials.max_count / 64));
        if (const char *e = std::getenv("GGDMC_B200_NSPLIT")) nsplit = std::max(1, std::min(std::atoi(e), std::max(1, trials.max_count / 8)));
        if (trials.max_count > 8192) nsplit = std::max(nsplit, (trials.max_count + 4095) / 4096);
        const int chunk = ((std::max(1, (trials.max_count + nsplit - 1) / nsplit)) + 7) & ~7;
        trials.d.chunk = chunk;
        trials.d.nsplit = std::max(1, (trials.max_count + chunk - 1) / chunk);
    }

    void setup_sampler()
    {
        const bool multi = kind == 2 && g_nccl.comm && g_nccl.n_rank > 1;
        const bool p2p = multi && g_p2p.ready && R * C * 2 <= kP2PMaxN;
        const int npop = R * S;
        sy_queue.alloc(1); sy_queue.zero();
        sy_all_done.alloc(1); sy_all_done.zero();
        sy_exit.alloc(1); sy_exit.zero();
        sy_chain_arrive.alloc((size_t)npop * C); sy_chain_arrive.zero();
        sy_phi_arrive.alloc(1); sy_phi_arrive.zero();
        sy_abort.alloc(1); sy_abort.zero();
        std::vector<unsigned int> flags((size_t)npop * kPopFlagStride, 0u);
        for (int p = 0; p < npop; ++p) flags[(size_t)p * kPopFlagStride] = 2u; // "half 1 of iteration 0 is closed"
        sy_pop_flags.upload(flags);
        sy_phi_done.upload(flags.data(), 1);
        CUDA_CHECK(cudaStreamSynchronize(0));
        SA.S = subj.L;
        if (kind == 2) SA.P = phi.L;
        SA.M = model.d;
        SA.T = trials.d;
        SA.H = H;
        SA.w = g_p2p.win;
        SA.y.queue = sy_queue.p; SA.y.exit_ctr = sy_exit.p; SA.y.pop_flags = sy_pop_flags.p; SA.y.chain_arrive = sy_chain_arrive.p;
        SA.y.all_done = sy_all_done.p; SA.y.phi_arrive = sy_phi_arrive.p; SA.y.phi_done = sy_phi_done.p; SA.y.abort = sy_abort.p;
        double sec = 20.0; // a local wait is bounded by the longest item chain of an iteration; peers are waited for inside the exchange
        if (const char *e = std::getenv("GGDMC_B200_SPIN_TIMEOUT_S")) sec = std::max(0.001, std::atof(e));
        SA.y.spin_ns = (unsigned long long)(sec * 1e9) + (multi ? g_p2p.win.spin_ns : 0ull);
        SA.ll_part = ll_part.p; SA.hpart = hpart.p; SA.hsum = hsum.p;
        SA.d_iter = d_iter.p;
        SA.hier = kind == 2; SA.use_p2p = p2p ? 1 : 0; SA.decide_once = kind == 0;
        // items of one iteration, and the launch shape: every warp is a worker; CTAs of 8 warps share one copy of the model's
        // tables, small fits use smaller CTAs so that their few workers spread over all SMs
        const unsigned long long n_sub = (unsigned long long)npop * ((C + 1) / 2) * trials.d.nsplit;
        const unsigned long long n_phi = SA.hier ? (unsigned long long)R * C * H.nsplit : 0ull;
        const unsigned long long per_iter = 2 * n_sub; // SUBJECT items; the phi level's items are published as they become runnable
        require(per_iter < 0x7fffffffull && n_phi < 0xffffffull && (unsigned long long)npop * ((C + 31) / 32) < 0xffffffull, "too many work items per iteration");
        SA.per_iter = (unsigned int)per_iter;
        {   // urgent queues: a phi half 0 that may run; no CLOSE items
            std::vector<unsigned long long> uq = {n_phi << 24, 0ull}; // batch 0 of the phi level with n_phi items (gg_sampler.cuh urgent_word)
            sy_urgent.upload(uq);
            sy_close_list.alloc((size_t)npop * ((C + 31) / 32));
            sy_close_list.zero();
            SA.y.urgent = sy_urgent.p;
            SA.y.close_list = sy_close_list.p;
        }
        const int n_sm = sm_count();
        int warps = 8;
        while (warps > 1 && per_iter < (unsigned long long)n_sm * 24 && per_iter < (unsigned long long)n_sm * warps * 3) warps >>= 1;
        if (const char *e = std::getenv("GGDMC_B200_SAMPLER_WARPS")) warps = std::max(1, std::min(8, std::atoi(e)));
        sampler_threads = warps * 32;
        SA.stage_bytes = (int)sampler_stage_bytes(model.d.n_cell, model.d.n_acc, model.d.n_row, model.d.n_const);
        SA.warp_bytes = (int)sampler_warp_bytes(model.d.n_cell, model.d.n_row, D, C, SA.hier);
        sampler_smem = (size_t)SA.stage_bytes + (size_t)warps * SA.warp_bytes;
        require(sampler_smem <= 220 * 1024, "row table does not fit in shared memory");
        sampler_nacc = model.d.n_acc;
        int per_sm = 0;
        switch (sampler_nacc) {
        case 2: per_sm = sampler_blocks_per_sm<2>(); break;
        case 3: per_sm = sampler_blocks_per_sm<3>(); break;
        case 4: per_sm = sampler_blocks_per_sm<4>(); break;
        default: per_sm = sampler_blocks_per_sm<0>();
        }
        require(per_sm >= 1, "sampler kernel does not fit on an SM (row table too large)");
        per_sm = std::min(per_sm, 24 / warps);
        sampler_grid = (int)std::min<unsigned long long>((unsigned long long)per_sm * n_sm, (per_iter + 2 * n_phi + warps - 1) / warps);
        if (const char *e = std::getenv("GGDMC_B200_BATCH")) sampler_max_batch = std::max(1, std::atoi(e));
        if (const char *e = std::getenv("GGDMC_B200_ITEMTRACE")) { // diagnostics: stamps of the first items of every launch
            (void)e;
            unsigned long long cap = 400000;
            if (const char *c = std::getenv("GGDMC_B200_ITEMTRACE_CAP")) cap = std::strtoull(c, nullptr, 10);
            sy_trace.alloc((size_t)cap * 8 + 8); // + the counter of the urgent items' slots
            sy_trace.zero();
            CUDA_CHECK(cudaStreamSynchronize(0));
            SA.trace = sy_trace.p;
            SA.trace_cap = cap;
        }
        // the migration decisions of iteration 1 (later ones are drawn inside the kernel at the end of the previous iteration),
        // and the constants of the subject prior under the start state of phi (later ones travel with accepted proposals)
        k_sweep_begin<<<npop, 128, (size_t)2 * C * sizeof(int), stream>>>(subj.L, d_iter.p, 0, SA.decide_once, -1);
        k_flags_init<<<(npop + 127) / 128, 128, 0, stream>>>(subj.L, sy_pop_flags.p, kPopFlagStride);
        if (kind == 2) {
            k_sweep_begin<<<R, 128, (size_t)2 * C * sizeof(int), stream>>>(phi.L, d_iter.p, 0, 0, -1);
            phi_constants(stream);
        }
        CUDA_CHECK(cudaGetLastError());
        launches += kind == 2 ? 2 : 1;
        persist = true;
    }

    // iterations [h_iter + 1, h_iter + n] in one launch
    void run_persist(int n)
    {
        require((unsigned long long)n * SA.per_iter < 0xfff00000ull, "too many work items for one launch (lower GGDMC_B200_BATCH)");
        if (SA.trace) CUDA_CHECK(cudaMemsetAsync(sy_trace.p, 0, ((size_t)SA.trace_cap * 8 + 8) * 8, stream));
        SA.t_begin = h_iter + 1;
        SA.t_end = h_iter + 1 + (uint32_t)n;
        sampler_cfg = cudaLaunchConfig_t{};
        sampler_cfg.gridDim = dim3(sampler_grid); sampler_cfg.blockDim = dim3(sampler_threads); sampler_cfg.dynamicSmemBytes = sampler_smem;
        sampler_cfg.stream = stream;
        cudaEvent_t ea = nullptr, eb = nullptr;
        if (profile) {
            if (prof_used + 2 > prof_ev.size()) {
                for (int i = 0; i < 2; ++i) {
                    cudaEvent_t e;
                    CUDA_CHECK(cudaEventCreate(&e));
                    prof_ev.push_back(e);
                }
            }
            ea = prof_ev[prof_used]; eb = prof_ev[prof_used + 1];
            prof_used += 2;
            CUDA_CHECK(cudaEventRecord(ea, stream));
        }
        switch (sampler_nacc) {
        case 2: sampler_launch<2>(); break;
        case 3: sampler_launch<3>(); break;
        case 4: sampler_launch<4>(); break;
        default: sampler_launch<0>();
        }
        if (profile) CUDA_CHECK(cudaEventRecord(eb, stream));
        h_iter += (uint32_t)n;
        ++launches;
    }

    void check_sampler_status()
    {
        if (!persist) return;
        int v = 0;
        CUDA_CHECK(cudaMemcpy(&v, sy_abort.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (v == 1) throw Error(GGDMC_ERR_COMM, "peer exchange timed out: a rank did not arrive");
        if (v != 0) throw Error(GGDMC_ERR_CUDA, "sampler kernel: a dependency wait timed out");
        if (SA.trace) dump_item_trace();
    }
    // GGDMC_B200_ITEMTRACE=<file>: the stamps of the LAST launch (tools/exp_itemtrace.py reads them)
    void dump_item_trace()
    {
        const char *path = std::getenv("GGDMC_B200_ITEMTRACE");
        if (!path || !*path) return;
        std::vector<unsigned long long> h((size_t)SA.trace_cap * 8);
        CUDA_CHECK(cudaMemcpy(h.data(), sy_trace.p, h.size() * 8, cudaMemcpyDeviceToHost));
        if (FILE *f = std::fopen(path, "wb")) {
            const unsigned long long hdr[8] = {SA.trace_cap, (unsigned long long)R * S, (unsigned long long)((C + 1) / 2), (unsigned long long)trials.d.nsplit,
                                               SA.hier ? (unsigned long long)R * C * H.nsplit : 0ull, (unsigned long long)sampler_grid, (unsigned long long)sampler_threads, (unsigned long long)SA.per_iter};
            std::fwrite(hdr, 8, 8, f);
            std::fwrite(h.data(), 8, h.size(), f);
            std::fclose(f);
        }
    }

    // all ranks of a sharded fit arrive before anybody iterates (the exchange assumes lock step within its timeout)
    void peer_barrier()
    {
        if (kind == 2 && g_nccl.comm && g_nccl.n_rank > 1 && g_p2p.ready) {
            k_peer_barrier<<<1, 32, 0, stream>>>(g_p2p.win);
            ++launches;
        }
    }

    // one DE-MCMC iteration: run_chains body (src/de.cpp:208-240) or run_hchains body (:281-381).
    // The iteration number (1-based, like the reference's loop variable) lives in device memory.
    void iteration()
    {
        ++h_iter;
        trace.reset();
        const bool conc = overlap && !profile; // per-launch timing (bench.py roofline pass) wants one launch at a time
        if (kind == 2) {
            cudaStream_t ps = conc ? side : stream;
            if (conc) {
                CUDA_CHECK(cudaEventRecord(ev_fork, stream));
                CUDA_CHECK(cudaStreamWaitEvent(side, ev_fork, 0));
            }
            if (is_hblocked)
                for (int p = 0; p < D2; ++p) sweep_phi(p, 0, p, ps);
            else
                sweep_phi(0, 0, -1, ps);
            if (!consts_travel) phi_constants(ps);
            cudaEvent_t join = nullptr;
            if (conc) {
                CUDA_CHECK(cudaEventRecord(ev_join, side));
                join = ev_join;
            }
            sweep_groups(0, join, conc);
            store_and_advance(subj, &phi);
            join_groups(conc);
        } else if (kind == 0) {
            if (conc && groups.size() > 1) CUDA_CHECK(cudaEventRecord(ev_fork, stream));
            sweep_groups(1, nullptr, conc);
            store_and_advance(subj, nullptr);
        } else {
            if (is_pblocked)
                for (int p = 0; p < phi.L.nmove; ++p) sweep_phi(p, 1, p, stream);
            else
                sweep_phi(0, 1, -1, stream);
            store_and_advance(phi, nullptr);
        }
    }

    // iteration() either as plain launches or as one graph launch
    bool short_call = false; // a one-shot run* call of a few dozen iterations: capturing and instantiating the graph costs more than it saves
    void step_once()
    {
        if (!use_graph || profile || h_iter == 0 || (short_call && !graph_exec)) { // the very first iteration runs uncaptured (one-off kernel attribute calls)
            iteration();
            return;
        }
        if (!graph_exec) {
            const int64_t l0 = launches;
            const uint32_t h0 = h_iter;
            CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
            iteration();
            CUDA_CHECK(cudaStreamEndCapture(stream, &graph));
            CUDA_CHECK(cudaGraphInstantiate(&graph_exec, graph, 0));
            launches_per_iter = launches - l0;
            launches = l0;
            h_iter = h0;
        }
        CUDA_CHECK(cudaGraphLaunch(graph_exec, stream));
        launches += launches_per_iter;
        ++h_iter;
    }

    // ---- streamed results: stored slot k goes to the caller's arrays while later iterations run -------
    struct OutSink { LevelDev *lv; int n_items; ggdmc_samples_t *outs; };
    std::vector<OutSink> sinks;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> slot_ev;
    int slots_sent = 0;

    void stream_results_to(LevelDev &lv, int n_items, ggdmc_samples_t *outs)
    {
        for (int i = 0; i < n_items; ++i) {
            require(outs[i].theta && outs[i].lp && outs[i].ll, "null output arrays");
            outs[i].npar = lv.L.npar; outs[i].nchain = C; outs[i].nmc = nmc;
        }
        sinks.push_back(OutSink{&lv, n_items, outs});
        if (!copy_stream) copy_stream = g_streams.get(device, prio_lo);
    }
    // copy slots [slots_sent, upto): one strided copy per array when the per-item arrays are adjacent
    void send_slots(int upto)
    {
        for (; slots_sent < upto; ++slots_sent) {
            const int k = slots_sent;
            if (k > 0) CUDA_CHECK(cudaStreamWaitEvent(copy_stream, slot_ev[k], 0));
            for (OutSink &o : sinks) {
                const int D_ = o.lv->L.npar;
                const size_t row = (size_t)C * D_ * 8, row1 = (size_t)C * 8;
                const size_t blk = (size_t)R * nmc * C * D_, blk1 = (size_t)R * nmc * C;
                bool adjacent = true;
                for (int i = 1; i < o.n_items; ++i)
                    if (o.outs[i].theta != o.outs[i - 1].theta + blk || o.outs[i].lp != o.outs[i - 1].lp + blk1 ||
                        o.outs[i].ll != o.outs[i - 1].ll + blk1)
                        adjacent = false;
                const int n_copy = adjacent ? 1 : o.n_items, rows = adjacent ? o.n_items * R : R;
                for (int i = 0; i < n_copy; ++i) {
                    CUDA_CHECK(cudaMemcpy2DAsync(o.outs[i].theta + (size_t)k * C * D_, row * nmc, o.lv->out_theta.p + i * blk + (size_t)k * C * D_,
                                                 row * nmc, row, rows, cudaMemcpyDeviceToHost, copy_stream));
                    CUDA_CHECK(cudaMemcpy2DAsync(o.outs[i].lp + (size_t)k * C, row1 * nmc, o.lv->out_lp.p + i * blk1 + (size_t)k * C, row1 * nmc,
                                                 row1, rows, cudaMemcpyDeviceToHost, copy_stream));
                    CUDA_CHECK(cudaMemcpy2DAsync(o.outs[i].ll + (size_t)k * C, row1 * nmc, o.lv->out_ll.p + i * blk1 + (size_t)k * C, row1 * nmc,
                                                 row1, rows, cudaMemcpyDeviceToHost, copy_stream));
                }
            }
        }
    }

    void iterate(int n_iter, float *elapsed_ms, ggdmc_progress_fn progress, void *user, int report_length)
    {
        CUDA_CHECK(cudaSetDevice(device));
        peer_barrier(); // ranks that enter seconds apart (uploads, host work) meet here, not inside the first exchange
        CUDA_CHECK(cudaEventRecord(ev0, stream));
        const bool streaming = !sinks.empty();
        if (streaming) {
            CUDA_CHECK(cudaStreamSynchronize(stream)); // slot 0 (the start state) is stored
            slot_ev.resize((size_t)nmc, nullptr);
        }
        const bool per_slot = streaming || (progress && report_length > 0);
        for (int i = 0; i < n_iter;) {
            if (persist) {
                // whole iterations per launch: up to the next stored sample when results are streamed, else up to the batch limit
                int n = std::min(n_iter - i, (int)std::min<unsigned long long>((unsigned long long)sampler_max_batch, std::max<unsigned long long>(1ull, 0xfff00000ull / SA.per_iter - 1)));
                if (per_slot) n = std::min(n, thin - (int)(h_iter % (uint32_t)thin));
                run_persist(n);
                i += n;
            } else {
                step_once();
                ++i;
            }
            if (streaming && h_iter % (uint32_t)thin == 0 && h_iter / (uint32_t)thin < (uint32_t)nmc) {
                // slot k is complete once this iteration is; it is sent one slot late, so that the (host-blocking, for
                // pageable arrays) copy runs while the device already works on the iterations of the next slot
                const int k = (int)(h_iter / (uint32_t)thin);
                if (!slot_ev[k]) CUDA_CHECK(cudaEventCreateWithFlags(&slot_ev[k], cudaEventDisableTiming));
                CUDA_CHECK(cudaEventRecord(slot_ev[k], stream));
                send_slots(k);
            }
            if (progress && report_length > 0 && h_iter % (uint32_t)thin == 0) {
                uint32_t stored = h_iter / (uint32_t)thin; // theta_phi::print_progress, @hdr/theta.h:76-85
                if ((stored + 1) % (uint32_t)report_length == 0) progress((int32_t)(stored + 1), user);
            }
        }
        CUDA_CHECK(cudaEventRecord(ev1, stream));
        if (streaming) {
            send_slots((int)std::min<uint32_t>((uint32_t)nmc, h_iter / (uint32_t)thin + 1));
            CUDA_CHECK(cudaStreamSynchronize(copy_stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        if (elapsed_ms) CUDA_CHECK(cudaEventElapsedTime(elapsed_ms, ev0, ev1));
        if (profile) collect_profile();
        trace.dump(g_nccl.comm ? g_nccl.rank : 0);
        check_sampler_status();
        if (kind == 2 && g_nccl.comm && g_nccl.n_rank > 1 && g_p2p.ready && g_p2p.timed_out())
            throw Error(GGDMC_ERR_COMM, "peer exchange timed out: a rank did not arrive");
    }

    // Timed iterations with an L2 flush (a memset larger than L2) before each one; only the iterations
    // themselves are inside the event brackets.  Returns the summed per-iteration time.
    void iterate_flushed(int n_iter, size_t flush_bytes, float *elapsed_ms)
    {
        CUDA_CHECK(cudaSetDevice(device));
        DBuf<unsigned char> flush;
        flush.alloc(flush_bytes);
        CUDA_CHECK(cudaStreamSynchronize(0));
        std::vector<cudaEvent_t> ev((size_t)2 * n_iter);
        for (auto &e : ev) CUDA_CHECK(cudaEventCreate(&e));
        for (int i = 0; i < n_iter; ++i) {
            CUDA_CHECK(cudaMemsetAsync(flush.p, i & 0xff, flush_bytes, stream));
            // the flush de-synchronises the ranks (a fit keeps them in lock-step through its exchanges): line them
            // up again before the bracket opens, so that the skew of the memsets is not booked as exchange wait
            if (kind == 2 && g_nccl.comm && g_nccl.n_rank > 1 && g_p2p.ready) k_peer_barrier<<<1, 32, 0, stream>>>(g_p2p.win);
            CUDA_CHECK(cudaEventRecord(ev[2 * i], stream));
            if (persist) run_persist(1);
            else step_once();
            CUDA_CHECK(cudaEventRecord(ev[2 * i + 1], stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        double total = 0.0;
        for (int i = 0; i < n_iter; ++i) {
            float ms = 0.f;
            CUDA_CHECK(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
            total += ms;
        }
        for (auto &e : ev) cudaEventDestroy(e);
        if (elapsed_ms) *elapsed_ms = (float)total;
        if (profile) collect_profile();
        check_sampler_status();
    }

};

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
