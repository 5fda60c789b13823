// ggdmc_b200 -- host-side building blocks of the engine (one translation unit: included by gg_engine.cu only):
// errors, the device memory pool and buffers, NCCL through dlopen and the peer-memory window, the device images of a
// model / a prior / the trials / one sampler level, the stream cache, the likelihood launchers, the launch tracer.
#pragma once

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
    // half_sweep_blocks > 0 (the sampler's PARALLEL half-sweeps of an LBA fit): a likelihood launch of less than one wave of
    // blocks ends with its slowest block, and blocks of one wave differ by a factor of two (the warp schedulers favour some
    // slots); two half-size blocks per proposal end sooner although each builds its own table (32 subjects x 39 proposals:
    // 39 -> 36 us per launch; three chunks 37, four 44)
    void set_chunking(int64_t blocks_per_split_unit, int64_t half_sweep_blocks = 0)
    {
        // enough blocks to fill 148 SMs a few times over, at least 256 trials per block
        int want = (int)std::max<int64_t>(1, (4 * 148 + blocks_per_split_unit - 1) / blocks_per_split_unit);
        int max_split = std::max(1, (max_count + 255) / 256);
        if (half_sweep_blocks > 0 && half_sweep_blocks < 148 * 12) want = std::max(want, 2);
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
// k_propose / k_accept launches of more warps than this run the build capped at 48 registers (10 blocks of 4 warps per SM)
constexpr int kShortKernelMinBlocks = 10, kShortKernelWaveWarps = 148 * 24;

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

