// ge_context.cuh — device context of libgeneevolve_b200.so: growable device buffers, per-population
// generation state (double-buffered), genome layout, scratch and the error plumbing shared by the
// translation unit (ge_api.cu includes the kernel headers around this file).
#pragma once
#include "../../include/geneevolve_b200.h"
#include "ge_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace gek;

static thread_local std::string g_err;
static int fail(int code, const std::string &m) { g_err = m; return code; }

#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) return fail(GE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define GE_TRY(expr) do { int rc_ = (expr); if (rc_ != GE_OK) return rc_; } while (0)

static inline unsigned nblk(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------------
struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    template <class T> T *as() const { return static_cast<T *>(p); }
};

struct SegState {   // founder segments of one generation (GE_REP_SEGMENTS): CSR over slots (i*n_chr + c)*2 + h
    Buf off, seg;   // off: uint64 [n_slots+1]; seg: uint4 {st, en, hap_index, root_population}
    uint64_t n_seg = 0;             // host mirror of StepState::n_seg[k] (exact after seg_finish_all)
    bool valid = false;
};
// The four sorts of assortative mating (males and females by mating value, the two template columns) are independent
// and, at <= N/2 keys each, pure launch latency (8 radix passes of ~10 us): they run side by side on four lanes.
struct SortLane {
    cudaStream_t s = nullptr;
    cudaEvent_t done = nullptr;
    Buf keys_in, vals_in, keys_out, vals_out, tmp;
};
constexpr int N_SORT_LANES = 4;

struct MateScratch {  // scratch of the mating kernels (ge_mating.cuh)
    Buf fam_off, keep, keep_off, keys_a, keys_b, idx_a, idx_b, list_m, list_f, t1, t2, rank1, rank2, tmp_sort;
};

struct Scheme { double va = 0, vd = 0, ve = 0, vc = 0, vf = 0, omega = 0, beta = 0, lambda = 0; };

struct CvHost {  // one (phen, chr) block of one population
    std::vector<uint64_t> bp;
    std::vector<double> a, d;
    std::vector<uint8_t> val;
    uint64_t nhap = 0;
};

struct GenState {  // one generation of one population on the device
    uint64_t n = 0;                     // host mirror of *d_n (exact after ge_ctx::pull_state)
    uint64_t *d_n = nullptr;            // this buffer's size on the device: &StepState::n[k]
    uint64_t *d_n_hm = nullptr;         // &StepState::n_hm[k]
    Buf hap, cv_allele /* bit plane [2n][Wcv] */, cv_root /* byte plane [2n][n_cv_tot], n_pop > 1 only */, ids, sex, A, D, G, C, E, F, P, mv, sv, svf;
    const uint32_t *rowmap = nullptr;   // after a migration: physical row pair of every individual (PopDev::rowmap_buf); identity otherwise
    Buf hm_off, hm_bp;  // per-haplotype mutation lists (CSR over slots)
    uint64_t n_hm = 0;
    bool has_hm = false;
    SegState seg;       // founder segments (GE_REP_SEGMENTS)
};

// Draws of one reproduce call.  Double-buffered: the bulk stream may still be copying haplotype rows of
// generation g from set g&1 while the control stream already samples generation g+1 into the other set.
struct DrawSet {
    Buf father, mother, couple_of, xo_off, xo_bp, flips, start_hap;
    cudaEvent_t bulk_done = nullptr;   // recorded on the bulk stream after the propagation that reads this set
    bool bulk_pending = false;
};

struct PopDev {
    bool avoid_inbreeding = false, RM = false, has_mut = false;
    double MM = 0;
    std::vector<std::vector<uint64_t>> rmap_bp, mutmap_bp;
    std::vector<std::vector<double>> recom_prob, mutmap_rate;
    std::vector<uint64_t> bp_dist;
    std::vector<std::vector<CvHost>> cv;  // [phen][chr]
    std::vector<Scheme> scheme;
    std::vector<std::vector<uint8_t>> panel;  // host copy until generation 0 is built
    std::vector<std::vector<uint32_t>> panel_packed;  // alternative: bit-packed by the host
    uint64_t n_founder_haps = 0;
    Buf founder_rows, founder_cv;  // kept on the device for segment materialisation / ras_find_cv (GE_REP_SEGMENTS)
    Buf founder_root;              // after a re-base: generation-0 root population of every (founder haplotype, CV) — populations with different effect tables only
    // re-basing history (ge_rebase_founders with keep_history): the lists of the generation that became the founder panel, oldest first
    struct SegSnapshot { Buf off, seg; uint64_t n = 0, n_seg = 0; };
    std::vector<SegSnapshot> history;
    // device maps
    Buf d_row_off, d_bp, d_T, d_bp_dist, d_mrow_off, d_mbp, d_mT, d_cov_lo, d_cov_hi;
    Buf d_vb, d_vb_off, d_vb_scale, d_mvb, d_mvb_off, d_mvb_scale;   // value indexes of the survival tables
    Buf d_omega, d_lambda, d_vd_zero;
    GenState st[2];
    int cur = 0;
    Buf prev_P, prev_F;
    Buf mig_pop[1], mig_idx[1];                 // gather lists of a migration (control stream only)
    Buf rowmap_buf[2];                          // row maps of the last two migrations (by generation parity; the bulk stream reads them late)
    uint64_t prev_n = 0;
    cudaEvent_t ev_ready = nullptr;   // control stream -> bulk stream: this population's draws are complete
    // device-resident sizes of the generation step (ge_kernels.cuh) and the host's copy of them
    StepState *d_ss = nullptr;
    StepState hs{};
    // couples
    Buf c_male, c_female, c_inbreed, c_noff;
    uint64_t n_couples = 0, couples_cap = 0;
    // draws of the last reproduce
    DrawSet ds[2];
    int dcur = 0;
    DrawSet &draws() { return ds[dcur]; }
    Buf mut_off, mut_bp, mut_gam, e_raw, cnt32;
    uint64_t n_off = 0, n_xo = 0, n_mut = 0;
    bool have_couple_of = false, have_e_raw = false;
    // constants
    std::vector<double> var_a0, var_d0;
    double sv_mean0 = 0, sv_var0 = 0;
    Buf d_sv0;  // [2] mean, var on device
    MateScratch mate;
};

struct KernelStat { double ms = 0; uint64_t launches = 0, bytes = 0; };

struct ge_ctx {
    ge_config cfg;
    cudaStream_t stream = nullptr;  // control stream (high priority): mating, sampling, CV planes, phenotypes
    cudaStream_t bulk = nullptr;    // bulk stream (low priority): bit-packed haplotype propagation, one generation behind at most
    cudaEvent_t ev_ready = nullptr, ev_join = nullptr, ev_fork = nullptr;
    SortLane lane[N_SORT_LANES];
    int fork_lanes() {  // the lanes start after everything queued on the control stream so far
        CUDA_TRY(cudaEventRecord(ev_fork, stream));
        for (SortLane &l : lane) CUDA_TRY(cudaStreamWaitEvent(l.s, ev_fork, 0));
        return GE_OK;
    }
    int join_lanes() {  // the control stream continues after every lane
        for (SortLane &l : lane) { CUDA_TRY(cudaEventRecord(l.done, l.s)); CUDA_TRY(cudaStreamWaitEvent(stream, l.done, 0)); }
        return GE_OK;
    }
    bool serial = false;            // GE_FLAG_SERIAL: the bulk copy is queued on the control stream (no overlap; measurements)
    bool gen0_done = false;         // ge_init_generation0 has run: maps, the segment format and the draw buffers are fixed
    bool needs_prev = false;        // some phenotype has vertical transmission (vf > 0): keep the previous generation's P and F
    bool use_graph = true;          // replay the control chain of a generation as a CUDA graph where possible (GE_FLAG_NO_GRAPH)
    uint64_t n_loci_total = 0;      // loci of this context over all its chromosomes
    int n_rebase = 0;               // ge_rebase_founders calls so far
    std::vector<uint64_t> gen0_off, gen0_seg;   // lists composed by ge_get_segment_count_gen0, handed out by ge_download_segments_gen0
    int gen0_pop = -1, gen0_chr = -1;
    int thin = 8;   // CTAs per SM the heavy control-stream kernels may take while a bulk copy is in flight (0 = no limit)
    bool bulk_busy = false;
    Buf seg_desc, seg_iv_off;       // copy descriptor and output offset of every interval (seg_plan_kernel -> seg_gather_kernel)
    Buf seg_cnt, seg_flags, seg_verb;   // scratch of the segment path (its own: it may run on the bulk stream)
    bool seg_packed = false;        // 8-byte parts {st, hap_index | root_population << 27} (decided in seg_init_gen0); else the reference's 16-byte parts
    bool seg_wide = false;          // GE_FLAG_SEG_WIDE_PARTS: never pack
    size_t seg_esz() const { return seg_packed ? 8 : 16; }
    bool seg_per_thread = false;    // a genetic map with rows closer than bp_dist_in_rmap was given, or GE_FLAG_SEG_VERBATIM: the reference's loop, one thread per slot
    bool cv_from_segments = false;  // GE_FLAG_CV_FROM_SEGMENTS: ge_compute_AD rescans the segment lists every generation like the reference
    double thin_min_bytes = 4e9;   // bytes moved by one bulk launch above which the control kernels go thin (below, the control chain is the critical path)
    int thin_now = 0;              // CTAs per SM in force for the copy in flight
    // Measured (scripts/emulate_rank.py, one rank of a 1/2/4/8-way shard of config 3): copies of >= 30 GB want 8 CTAs per SM for the
    // control kernels, the 6-25 GB copies of a shard 4 (2-3 % faster than 8 or no limit), copies of ~1 GB (config 2) no limit.
    void note_bulk(double bytes) {
        bulk_busy = !serial && bytes > thin_min_bytes;
        thin_now = bytes >= 30e9 ? thin : std::max(1, thin / 2);   // (re-measured after the control kernels were rebuilt: profiles/r2u_thin_grid_sweep.txt)
    }
    // grid of a grid-stride control kernel: full width, or thin while it shares the GPU with the bulk copy,
    // so that the high-priority control stream displaces only a fraction of the bulk kernel's resident CTAs
    unsigned ctrl_grid(uint64_t n_threads, unsigned block) const {
        uint64_t full = std::max<uint64_t>(1, (n_threads + block - 1) / block);
        if (!bulk_busy || serial || thin <= 0) return (unsigned)std::min<uint64_t>(full, 1u << 30);
        return (unsigned)std::min<uint64_t>(full, (uint64_t)n_sm * thin_now);
    }
    std::vector<PopDev> pop;
    std::vector<std::vector<uint64_t>> loci;  // host positions per chromosome
    std::vector<double> gamma;
    std::vector<std::vector<uint64_t>> mig_sample;  // fixed-draw mode: migrants per source population
    std::vector<uint32_t> chr_ids;                  // global chromosome index of each local chromosome
    Buf d_chr_ids, ar_scratch;
    Buf mig_lists[2][5], mig_stage;             // row moves of the last two migrations (the bulk stream reads them late) + staging rows
    cudaEvent_t mig_done[2] = {nullptr, nullptr};
    bool mig_pending[2] = {false, false};
    ge_allreduce_fn allreduce = nullptr;
    void *allreduce_user = nullptr;
    // ge_set_allreduce_nccl: ncclAllReduce(sendbuff, recvbuff, count, ncclDataType_t, ncclRedOp_t, comm, stream) called directly
    typedef int (*nccl_all_reduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    nccl_all_reduce_fn nccl_all_reduce = nullptr;
    void *nccl_comm = nullptr;
    bool sharded() const { return allreduce || nccl_all_reduce; }
    Stream rng;
    // genome layout
    std::vector<uint32_t> chr_word_off, chr_nloci, locus_off;
    uint32_t W = 0;
    Buf d_chr_word_off, d_chr_nloci, d_locus_off, d_pos, d_bkt_off, d_bkt_shift, d_bkt;
    bool genome_ready = false;
    // tile table
    Buf d_tile_chr, d_tile_chunk0, d_tile_nchunk;
    uint32_t n_tiles = 0;
    // causal-variant set
    std::vector<uint32_t> cv_block_off;  // [n_phen*n_chr+1]
    std::vector<uint32_t> cv_word_off;   // [n_phen*n_chr+1]
    uint32_t n_cv_tot = 0, Wcv = 4;
    bool cv_sorted = true;
    bool use_root = false;   // populations with different effect tables: carry the root population of every CV allele
    Buf d_LA /* double2 [n_cv][3] */, d_LG /* double2 [Wcv*8][256]: sums per group of four CVs */, d_cv_bitpos, xo_stash;
    Buf d_cv_word_off, d_cv_word_blk, d_cv_block_off, d_cv_bp, d_cv_chr, d_a_eff, d_d_eff, d_cv_count;
    bool cv_ready = false;
    // scratch
    Buf scan_blocks, bulk_scan_blocks, partial, scalars;
    Buf d_ss_all;                               // StepState [n_pop]
    StepState *h_ss_all = nullptr;              // pinned read-back buffer of the same
    uint64_t graph_epoch = 0;                   // bumped whenever a device buffer is (re)allocated: captured graphs hold raw pointers
    int tmp_depth = 0;                          // > 0 inside a download: its scratch buffers come and go without touching any graph
    struct TmpScope { ge_ctx *c; explicit TmpScope(ge_ctx *ctx) : c(ctx) { c->tmp_depth++; } ~TmpScope() { c->tmp_depth--; } };
    void buffers_moved() { if (tmp_depth == 0) graph_epoch++; }
    int n_sm = 148;
    int xo_ctas_per_sm = 10;   // resident CTAs of sample_xo_kernel per SM (occupancy query at creation)
    unsigned prop_threads = 256;   // threads of a propagate_bits_kernel CTA (one offspring): by row length, build_genome
    // stats
    bool profiling = false;        // CUDA events around the dominant kernel (propagate_bits / the segment passes) on its own stream
    bool phase_timing = false;     // ... and around the phases of the control chain (ge_set_profiling level 2; disables graph replay)
    KernelStat kstat[GE_KERNEL_COUNT];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // ge_timer_start / ge_timer_stop
    // bytes_per_offspring != 0: the launch's size was only known on the device; pull_state multiplies by the offspring count it reads back
    // count_src != nullptr: bytes = *count_src * count_scale, a pinned word the launch's stream fills behind the kernels (resolved lazily)
    struct EvPair { cudaEvent_t a, b; int kernel; uint64_t bytes; uint64_t bytes_per_offspring = 0; int pop = 0; const uint64_t *count_src = nullptr; uint64_t count_scale = 0; };
    std::vector<uint64_t *> pinned_chunks;   // 512 words each
    size_t pinned_used = 0;
    uint64_t *pinned_slot() {
        if (pinned_chunks.empty() || pinned_used == 512) { uint64_t *c = nullptr; if (cudaMallocHost(&c, 512 * 8) != cudaSuccess) return nullptr; pinned_chunks.push_back(c); pinned_used = 0; }
        return pinned_chunks.back() + pinned_used++;
    }
    // ---- a generation's control chain as a CUDA graph (ge_api.cu: ge_step_generation) ----
    struct StepGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        std::vector<cudaGraphNode_t> begin_nodes;      // the step_begin_kernel node of every population: the one parameter update per replay
        std::vector<std::function<int()>> bulk;        // what follows every launch on the bulk stream (it runs a generation behind)
        uint64_t launches = 0, epoch = 0;
        int warm = 0;                                  // eager runs of this key since the buffers last moved (negative: after a failed recording)
        int fails = 0;
    };
    std::map<std::string, StepGraph> graphs;
    uint64_t plain_epoch = ~0ull;                      // graph_epoch at which plain_gens generations were queued kernel by kernel
    int plain_gens = 0;
    bool capturing = false;
    std::vector<std::function<int()>> deferred;
    void drop_graphs();
    // Hand-off control stream -> bulk stream: fn (launches on the bulk stream) runs behind everything queued on the control stream so
    // far.  Queued directly: event record, wait, fn.  While a generation is being captured: an external event-record node goes into
    // the graph, fn is kept and run (wait + launches) after every launch of that graph.
    int to_bulk(cudaEvent_t ev, std::function<int()> fn) {
        cudaStream_t b = serial ? stream : bulk;
        if (!capturing) {
            CUDA_TRY(cudaEventRecord(ev, stream));
            CUDA_TRY(cudaStreamWaitEvent(b, ev, 0));
            return fn();
        }
        CUDA_TRY(cudaEventRecordWithFlags(ev, stream, cudaEventRecordExternal));
        deferred.push_back([=]() -> int { CUDA_TRY(cudaStreamWaitEvent(b, ev, 0)); return fn(); });
        return GE_OK;
    }
    // the control stream must not overwrite a draw set the bulk stream still reads
    int wait_bulk_done(DrawSet &D) {
        if (capturing) { CUDA_TRY(cudaStreamWaitEvent(stream, D.bulk_done, cudaEventWaitExternal)); return GE_OK; }
        if (D.bulk_pending) { CUDA_TRY(cudaStreamWaitEvent(stream, D.bulk_done, 0)); D.bulk_pending = false; }
        return GE_OK;
    }
    std::vector<EvPair> ev_pending;            // per-launch events of profiled kernels, resolved lazily (no sync in the loop)
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t get_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void resolve_events() {
        for (auto &p : ev_pending) {
            cudaEventSynchronize(p.b);
            float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
            if (p.count_src) p.bytes = *p.count_src * p.count_scale;
            kstat[p.kernel].ms += ms; kstat[p.kernel].launches++; kstat[p.kernel].bytes += p.bytes;
            ev_pool.push_back(p.a); ev_pool.push_back(p.b);
        }
        ev_pending.clear();
    }
    // CUDA events around a phase of the control chain (only while profiling)
    struct PhaseTimer {
        ge_ctx *c; EvPair p;
        PhaseTimer(ge_ctx *ctx, int id) : c(ctx), p{nullptr, nullptr, id, 0, 0, 0} {
            if (c->phase_timing) { p.a = c->get_event(); p.b = c->get_event(); cudaEventRecord(p.a, c->stream); }
        }
        ~PhaseTimer() { if (p.a) { cudaEventRecord(p.b, c->stream); c->ev_pending.push_back(p); } }
    };
    uint64_t launches = 0, graph_replays = 0;
    size_t mem_now = 0, mem_peak = 0;

    bool bits() const { return cfg.representation & GE_REP_BITS; }
    bool segs() const { return cfg.representation & GE_REP_SEGMENTS; }

    int ensure(Buf &b, size_t bytes) {
        if (bytes <= b.cap && b.p) return GE_OK;
        if (bytes == 0) bytes = 16;
        if (b.p) { cudaFree(b.p); mem_now -= b.cap; b.p = nullptr; b.cap = 0; }
        size_t want = bytes + (bytes >> 3);  // a little slack so growing buffers do not reallocate every generation
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.p, want); }
        if (e != cudaSuccess) return fail(GE_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
        b.cap = want; mem_now += want; mem_peak = std::max(mem_peak, mem_now);
        buffers_moved();
        return GE_OK;
    }
    int ensure_exact(Buf &b, size_t bytes) {
        if (bytes <= b.cap && b.p) return GE_OK;
        if (b.p) { cudaFree(b.p); mem_now -= b.cap; b.p = nullptr; b.cap = 0; }
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) return fail(GE_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
        b.cap = bytes; mem_now += bytes; mem_peak = std::max(mem_peak, mem_now);
        buffers_moved();
        return GE_OK;
    }
    void release(Buf &b) { if (b.p) { cudaFree(b.p); mem_now -= b.cap; buffers_moved(); } b.p = nullptr; b.cap = 0; }
    template <class T> int upload(Buf &b, const std::vector<T> &v) {
        GE_TRY(ensure(b, v.size() * sizeof(T)));
        if (!v.empty()) CUDA_TRY(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));  // v may be a temporary
        return GE_OK;
    }
    Genome genome() const {
        Genome g;
        g.n_chr = cfg.n_chr; g.W = W;
        g.chr_word_off = d_chr_word_off.as<uint32_t>(); g.chr_nloci = d_chr_nloci.as<uint32_t>();
        g.locus_off = d_locus_off.as<uint32_t>(); g.pos = d_pos.as<uint32_t>();
        g.bkt_off = d_bkt_off.as<uint32_t>(); g.bkt_shift = d_bkt_shift.as<uint32_t>(); g.bkt = d_bkt.as<uint32_t>();
        return g;
    }
    CvSet cvset() const {
        CvSet c;
        c.n_chr = cfg.n_chr; c.n_phen = cfg.n_phen; c.n_cv_tot = n_cv_tot; c.Wcv = Wcv; c.sorted = cv_sorted;
        c.word_off = d_cv_word_off.as<uint32_t>(); c.word_blk = d_cv_word_blk.as<uint32_t>();
        c.block_off = d_cv_block_off.as<uint32_t>(); c.bp = d_cv_bp.as<uint32_t>(); c.chr_of = d_cv_chr.as<uint32_t>();
        return c;
    }
    TileTable tiles() const {
        TileTable t;
        t.n_items = n_tiles; t.chr = d_tile_chr.as<uint32_t>(); t.chunk0 = d_tile_chunk0.as<uint32_t>(); t.nchunk = d_tile_nchunk.as<uint32_t>();
        return t;
    }
    MapDev rmap(const PopDev &P) const {
        MapDev m; m.row_off = P.d_row_off.as<uint32_t>(); m.bp = P.d_bp.as<uint32_t>(); m.T = P.d_T.as<double>(); m.bp_dist = P.d_bp_dist.as<uint32_t>();
        m.vb = P.d_vb.as<uint32_t>(); m.vb_off = P.d_vb_off.as<uint32_t>(); m.vb_scale = P.d_vb_scale.as<double>();
        m.chr_id = d_chr_ids.as<uint32_t>();
        return m;
    }
    MapDev mmap(const PopDev &P) const {
        MapDev m; m.row_off = P.d_mrow_off.as<uint32_t>(); m.bp = P.d_mbp.as<uint32_t>(); m.T = P.d_mT.as<double>(); m.bp_dist = nullptr;
        m.vb = P.d_mvb.as<uint32_t>(); m.vb_off = P.d_mvb_off.as<uint32_t>(); m.vb_scale = P.d_mvb_scale.as<double>();
        m.chr_id = d_chr_ids.as<uint32_t>();
        return m;
    }
    // the control stream waits for everything queued on the bulk stream (before it touches haplotype rows)
    int join_bulk() {
        CUDA_TRY(cudaEventRecord(ev_join, bulk));
        CUDA_TRY(cudaStreamWaitEvent(stream, ev_join, 0));
        return GE_OK;
    }
    int check_launch(const char *what) {
        launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(GE_ERR_CUDA, std::string(what) + " launch: " + cudaGetErrorString(e));
        return GE_OK;
    }
    // grid of a grid-stride kernel over at most `bound` elements
    unsigned grid_for(uint64_t bound, unsigned block) const { return (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (bound + block - 1) / block), 1u << 20); }
    // device exclusive scan: out[0..n] (n + 1 entries, out[n] = grand total); n lives on the device, n_bound (host) sizes the
    // grid; on_total runs in the thread that stores out[n].  One launch for small arrays, three otherwise.
    template <class In, class OnTotal>
    int scan_in(cudaStream_t st, In in, DevN n, uint64_t n_bound, uint64_t *out, OnTotal on_total) {
        Buf &blocks = st == bulk ? bulk_scan_blocks : scan_blocks;
        n = limited(n, n_bound);
        if (n_bound <= 2 * SCAN1_CHUNK) {
            scan_one_cta_kernel<In, OnTotal><<<1, SCAN1_THREADS, 0, st>>>(in, n, out, on_total);
            return check_launch("scan_one_cta");
        }
        uint32_t nb = (uint32_t)(n_bound / SCAN_TILE) + 1;
        GE_TRY(ensure(blocks, (size_t)nb * 8));
        scan_block_sums_kernel<In><<<nb, SCAN_THREADS, 0, st>>>(in, n, blocks.as<uint64_t>());
        GE_TRY(check_launch("scan_block_sums"));
        scan_single_block_kernel<<<1, SCAN_THREADS, 0, st>>>(blocks.as<uint64_t>(), n);
        GE_TRY(check_launch("scan_single_block"));
        scan_final_kernel<In, OnTotal><<<nb, SCAN_THREADS, 0, st>>>(in, n, blocks.as<uint64_t>(), out, on_total);
        return check_launch("scan_final");
    }
    template <class T, class OnTotal>
    int scan(cudaStream_t st, const T *in, DevN n, uint64_t n_bound, uint64_t *out, OnTotal on_total) {
        return scan_in(st, PtrIn<T>{in}, n, n_bound, out, on_total);
    }
    // host-known n (setup, migration, downloads); the total is also returned to the host when asked (one sync)
    int exclusive_scan(const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *host_total) {
        GE_TRY(scan(stream, in, hostn(n), n, out, NoTotal{}));
        if (host_total) {
            CUDA_TRY(cudaMemcpyAsync(host_total, out + n, 8, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
        }
        return GE_OK;
    }
    // mean (denominator n) or variance (two-pass, n-1) of a device column into a device scalar
    unsigned moment_grid(uint64_t n_bound) const { return (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (n_bound + 255) / 256), MOMENT_MAX_BLOCKS); }
    int ensure_partial() {
        if (partial.p) return GE_OK;
        GE_TRY(ensure(partial, MOMENT_MAX_BLOCKS * 8 + 16));
        CUDA_TRY(cudaMemsetAsync(partial.p, 0, MOMENT_MAX_BLOCKS * 8 + 16, stream));
        return GE_OK;
    }
    int d_mean(const double *x, DevN n, uint64_t n_bound, double *out) {
        GE_TRY(ensure_partial());
        n = limited(n, n_bound);
        moment_kernel<<<moment_grid(n_bound), 256, 0, stream>>>(x, n, nullptr, 0, 0, partial.as<double>(), out);
        return check_launch("moment<mean>");
    }
    int d_var(const double *x, DevN n, uint64_t n_bound, double *out /* device */, double *mean_scratch /* device */) {
        GE_TRY(d_mean(x, n, n_bound, mean_scratch));
        moment_kernel<<<moment_grid(n_bound), 256, 0, stream>>>(x, limited(n, n_bound), mean_scratch, 1, 1, partial.as<double>(), out);
        return check_launch("moment<var>");
    }
    int h_var(const double *x, uint64_t n, double *host_out, double *host_mean = nullptr) {
        GE_TRY(ensure(scalars, 64 * 8));
        double *s = scalars.as<double>();
        GE_TRY(d_var(x, hostn(n), n, s + 0, s + 1));
        double h[2];
        CUDA_TRY(cudaMemcpyAsync(h, s, 16, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        *host_out = h[0];
        if (host_mean) *host_mean = h[1];
        return GE_OK;
    }
    // ---- device-resident step state <-> host ----
    // host copy -> device (setup, replayed draws, migration: whenever the host decides a size)
    int push_state(PopDev &P, size_t offset = 0, size_t bytes = sizeof(StepState)) {
        CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<char *>(P.d_ss) + offset, reinterpret_cast<const char *>(&P.hs) + offset, bytes, cudaMemcpyHostToDevice, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));   // P.hs is pageable and changes under the caller
        return GE_OK;
    }
    // device -> host: THE host synchronisation of a generation.  Refreshes every host mirror of a device-side size and turns
    // the error bits the kernels left into the reference's errors (message, code); the bits are cleared on the device.
    int pull_state(const char *where);
};

#define CHECK_CTX(ctx) if (!(ctx)) return fail(GE_ERR_INVALID, "null context")
#define CHECK_POP(ctx, p) CHECK_CTX(ctx); if ((p) < 0 || (p) >= (ctx)->cfg.n_pop) return fail(GE_ERR_INVALID, "bad population index")
#define CHECK_CHR(ctx, c) if ((c) < 0 || (c) >= (ctx)->cfg.n_chr) return fail(GE_ERR_INVALID, "bad chromosome index")
#define CHECK_PHEN(ctx, f) if ((f) < 0 || (f) >= (ctx)->cfg.n_phen) return fail(GE_ERR_INVALID, "bad phenotype index")

