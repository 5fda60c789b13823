// ggdmc_b200 -- struct ggdmc_engine: the device state of one fit and the launch sequences of an iteration
// (de_class::run_chains / run_hchains, src/de.cpp:201-242, 272-383), the persistent sampler's set-up, result streaming.
// Included by gg_engine.cu only, after gg_host.cuh.
#pragma once

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
        else trials.set_chunking((int64_t)R * S * C, m->type == GGDMC_MODEL_LBA && schedule == GGDMC_SCHEDULE_PARALLEL ? (int64_t)R * S * ((C + 1) / 2) : 0);
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
                const bool waves = nw > short_wave_warps; // several waves of warps: the register-capped build of the short kernels
                TR("k_propose", stream, launch_hi(waves ? k_propose<kProposeWarps, kShortKernelMinBlocks> : k_propose<kProposeWarps, 0>, (nw + kProposeWarps - 1) / kProposeWarps, kProposeWarps * 32, prop_sm, stream, L, d_iter.p, sweep, -1, half));
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
                TR("k_accept", stream, launch_hi(waves ? k_accept<kAcceptWarps, kShortKernelMinBlocks> : k_accept<kAcceptWarps, 0>, (nw + kAcceptWarps - 1) / kAcceptWarps, kAcceptWarps * 32, (size_t)kAcceptWarps * D * 8, stream, L, d_iter.p, sweep, -1, (const double *)G.ll_part, G.T.nsplit, half));
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
            // the group's own iteration counter advances here too (it equals the engine's at this point), so that a fit may
            // switch between this order and the side-stream one (per-launch profiling on / off) at any iteration
            TR("k_sweep_begin", stream, launch_hi(k_sweep_begin, L.npop, 128, (size_t)2 * C * sizeof(int), stream, L, d_iter.p, sweep, decide_once, para_idx, 1, sb_iter.p + G.index, sb_done.p + G.index));
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
    // GGDMC_B200_SHORT_WAVE_WARPS=n: launch size (warps) from which k_propose / k_accept run their register-capped build (tests: 0)
    int short_wave_warps = std::getenv("GGDMC_B200_SHORT_WAVE_WARPS") ? std::atoi(std::getenv("GGDMC_B200_SHORT_WAVE_WARPS")) : kShortKernelWaveWarps;
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

