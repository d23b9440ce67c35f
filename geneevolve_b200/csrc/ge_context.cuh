// ge_context.cuh — device context of libgeneevolve_b200.so: growable device buffers, per-population
// generation state (double-buffered), genome layout, scratch and the error plumbing shared by the
// translation unit (ge_api.cu includes the kernel headers around this file).
#pragma once
#include "../../include/geneevolve_b200.h"
#include "ge_kernels.cuh"

#include <algorithm>
#include <cmath>
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
    uint64_t n_seg = 0;
    bool valid = false;
    // asynchronous form (bulk stream, seg_capacity given): n_seg arrives in pinned host memory behind `ready`
    bool pending = false;
    cudaEvent_t ready = nullptr;
    uint64_t *h_total = nullptr;              // pinned
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // profiling events of the two passes, accounted when n_seg is known
};
// The four sorts of assortative mating (males and females by mating value, the two template columns) are independent
// and, at <= N/2 keys each, pure launch latency (8 radix passes of ~10 us): they run side by side on four lanes.
struct SortLane {
    cudaStream_t s = nullptr;
    cudaEvent_t done = nullptr;
    Buf keys_in, keys_out, vals_out, tmp;
};
constexpr int N_SORT_LANES = 4;

struct MateScratch {  // scratch of the mating kernels (ge_mating.cuh)
    Buf fam_off, keep, keys_a, keys_b, idx_a, idx_b, list_m, list_f, t1, t2, rank1, rank2, tmp_sort, counters, mv_m, mv_f;
};

struct Scheme { double va = 0, vd = 0, ve = 0, vc = 0, vf = 0, omega = 0, beta = 0, lambda = 0; };

struct CvHost {  // one (phen, chr) block of one population
    std::vector<uint64_t> bp;
    std::vector<double> a, d;
    std::vector<uint8_t> val;
    uint64_t nhap = 0;
};

struct GenState {  // one generation of one population on the device
    uint64_t n = 0;
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
    // couples
    Buf c_male, c_female, c_inbreed, c_noff;
    uint64_t n_couples = 0;
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
    bool serial = false;
    bool cub_sorts = false;         // GE_CUB_SORTS: cub::DeviceRadixSort for every sort of the mating chain (A/B against the small sorts)
    int thin = 8;   // CTAs per SM the heavy control-stream kernels may take while a bulk copy is in flight (0 = no limit)
    bool bulk_busy = false;
    int prop_depth = 4;
    int seg_group = 0;              // GE_SEG_GROUP: force 1, 8 or 32 lanes per slot in the segment recombination (0 = by list length)
    bool seg_walk = false;          // GE_SEG_WALK: the two walk passes (seg_recombine_warp_kernel) instead of plan + gather
    Buf seg_desc, seg_iv_off;       // copy descriptor and output offset of every interval (seg_plan_kernel -> seg_gather_kernel)
    Buf seg_cnt, seg_scan_blocks, seg_scan_total, seg_flags, seg_verb;   // scratch of the segment path (its own: it may run on the bulk stream)
    double seg_plan_min_parts = 0;  // GE_SEG_PLAN_MIN: parts per parental list below which the thread-per-slot walk is used (measured: plan + gather wins from generation 1)
    bool seg_packed = false;        // 8-byte parts {st, hap_index | root_population << 27} (decided in seg_init_gen0); else the reference's 16-byte parts
    int seg_depth = 4;              // GE_SEG_DEPTH: independent loads per thread in the packed gather (4 or 8; 8 measured slower: 53 registers)
    bool seg_wide = false;          // GE_SEG_FORMAT=16: never pack
    size_t seg_esz() const { return seg_packed ? 8 : 16; }
    bool seg_sync_mode = false;     // GE_SEG_SYNC: never queue the segment path on the bulk stream
    bool seg_per_thread = false;    // a genetic map with rows closer than bp_dist_in_rmap was given, or GE_SEG_PER_THREAD is set
    bool cv_from_segments = false;  // GE_CV_FROM_SEGMENTS: ge_compute_AD rescans the segment lists every generation like the reference
    bool use_tma = false, tma_attr_set = false;
    double thin_min_bytes = 4e9;   // bytes moved by one bulk launch above which the control kernels go thin (below, the control chain is the critical path)
    int thin_now = 0;              // CTAs per SM in force for the copy in flight
    // Measured (scripts/emulate_rank.py, one rank of a 1/2/4/8-way shard of config 3): copies of >= 30 GB want 8 CTAs per SM for the
    // control kernels, the 6-25 GB copies of a shard 4 (2-3 % faster than 8 or no limit), copies of ~1 GB (config 2) no limit.
    void note_bulk(double bytes) {
        bulk_busy = !serial && bytes > thin_min_bytes;
        thin_now = bytes >= 30e9 ? thin : std::max(1, thin / 2);
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
    Buf d_LA /* double2 [n_cv][3] */, d_cv_bitpos, xo_stash;
    Buf d_cv_word_off, d_cv_word_blk, d_cv_block_off, d_cv_bp, d_cv_chr, d_a_eff, d_d_eff, d_cv_count;
    bool cv_ready = false;
    // scratch
    Buf scan_blocks, scan_total, partial, scalars, flags;
    int n_sm = 148;
    // stats
    bool profiling = false;
    KernelStat kstat[GE_KERNEL_COUNT];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // ge_timer_start / ge_timer_stop
    struct EvPair { cudaEvent_t a, b; int kernel; uint64_t bytes; };
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
            kstat[p.kernel].ms += ms; kstat[p.kernel].launches++; kstat[p.kernel].bytes += p.bytes;
            ev_pool.push_back(p.a); ev_pool.push_back(p.b);
        }
        ev_pending.clear();
    }
    // CUDA events around a phase of the control chain (only while profiling)
    struct PhaseTimer {
        ge_ctx *c; EvPair p;
        PhaseTimer(ge_ctx *ctx, int id) : c(ctx), p{nullptr, nullptr, id, 0} {
            if (c->profiling) { p.a = c->get_event(); p.b = c->get_event(); cudaEventRecord(p.a, c->stream); }
        }
        ~PhaseTimer() { if (p.a) { cudaEventRecord(p.b, c->stream); c->ev_pending.push_back(p); } }
    };
    uint64_t launches = 0;
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
        return GE_OK;
    }
    int ensure_exact(Buf &b, size_t bytes) {
        if (bytes <= b.cap && b.p) return GE_OK;
        if (b.p) { cudaFree(b.p); mem_now -= b.cap; b.p = nullptr; b.cap = 0; }
        if (bytes == 0) bytes = 16;
        cudaError_t e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) return fail(GE_ERR_CUDA, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
        b.cap = bytes; mem_now += bytes; mem_peak = std::max(mem_peak, mem_now);
        return GE_OK;
    }
    void release(Buf &b) { if (b.p) { cudaFree(b.p); mem_now -= b.cap; } b.p = nullptr; b.cap = 0; }
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
    // device exclusive scan: out[0..n] (n+1 entries), grand total also returned to the host when asked
    int exclusive_scan(const uint32_t *in, uint64_t n, uint64_t *out, uint64_t *host_total) {
        GE_TRY(exclusive_scan_on(stream, scan_blocks, scan_total, in, n, out));
        if (host_total) {
            if (n == 0) { *host_total = 0; return GE_OK; }
            CUDA_TRY(cudaMemcpyAsync(host_total, scan_total.p, 8, cudaMemcpyDeviceToHost, stream));
            CUDA_TRY(cudaStreamSynchronize(stream));
        }
        return GE_OK;
    }
    // the same on any stream with its own scratch (the segment path scans on the bulk stream); the total stays in total.p
    int exclusive_scan_on(cudaStream_t st, Buf &blocks, Buf &total, const uint32_t *in, uint64_t n, uint64_t *out) {
        GE_TRY(ensure(total, 8));
        if (n == 0) {
            CUDA_TRY(cudaMemsetAsync(out, 0, 8, st));
            CUDA_TRY(cudaMemsetAsync(total.p, 0, 8, st));
            return GE_OK;
        }
        uint32_t nb = nblk(n, SCAN_THREADS * SCAN_ITEMS);
        GE_TRY(ensure(blocks, (size_t)nb * 8));
        scan_block_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, n, blocks.as<uint64_t>());
        GE_TRY(check_launch("scan_block_sums"));
        scan_single_block_kernel<<<1, SCAN_THREADS, 0, st>>>(blocks.as<uint64_t>(), nb, total.as<uint64_t>());
        GE_TRY(check_launch("scan_single_block"));
        scan_final_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, n, blocks.as<uint64_t>(), out);
        GE_TRY(check_launch("scan_final"));
        return GE_OK;
    }
    // mean (denominator n) or variance (two-pass, n-1) of a device column into a device scalar
    int d_mean(const double *x, uint64_t n, double *out) {
        int nb = (int)std::min<uint64_t>(std::max<uint64_t>(1, nblk(n, 256)), 1024);
        GE_TRY(ensure(partial, 1024 * 8));
        moment_partial_kernel<<<nb, 256, 0, stream>>>(x, n, nullptr, 0, partial.as<double>());
        GE_TRY(check_launch("moment_partial"));
        moment_final_kernel<<<1, 32, 0, stream>>>(partial.as<double>(), nb, (double)n, out);
        return check_launch("moment_final");
    }
    int d_var(const double *x, uint64_t n, double *out /* device */, double *mean_scratch /* device */) {
        if (n <= 1) { CUDA_TRY(cudaMemsetAsync(out, 0, 8, stream)); return GE_OK; }
        GE_TRY(d_mean(x, n, mean_scratch));
        int nb = (int)std::min<uint64_t>(std::max<uint64_t>(1, nblk(n, 256)), 1024);
        moment_partial_kernel<<<nb, 256, 0, stream>>>(x, n, mean_scratch, 1, partial.as<double>());
        GE_TRY(check_launch("moment_partial"));
        moment_final_kernel<<<1, 32, 0, stream>>>(partial.as<double>(), nb, (double)(n - 1), out);
        return check_launch("moment_final");
    }
    int h_var(const double *x, uint64_t n, double *host_out, double *host_mean = nullptr) {
        GE_TRY(ensure(scalars, 64 * 8));
        double *s = scalars.as<double>();
        GE_TRY(d_var(x, n, s + 0, s + 1));
        if (n <= 1) GE_TRY(d_mean(x, std::max<uint64_t>(n, 1), s + 1));
        double h[2];
        CUDA_TRY(cudaMemcpyAsync(h, s, 16, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        *host_out = h[0];
        if (host_mean) *host_mean = h[1];
        return GE_OK;
    }
    int check_flags(const char *where) {
        int h[4];
        CUDA_TRY(cudaMemcpyAsync(h, flags.p, sizeof(h), cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        if (h[0]) return fail(GE_ERR_NAN, std::string("Error: A or D is nan (") + where + ")");
        if (h[1]) return fail(GE_ERR_INVALID, "parent ID outside the previous generation (the reference reads out of bounds here, :3118-3133)");
        return GE_OK;
    }
};

#define CHECK_CTX(ctx) if (!(ctx)) return fail(GE_ERR_INVALID, "null context")
#define CHECK_POP(ctx, p) CHECK_CTX(ctx); if ((p) < 0 || (p) >= (ctx)->cfg.n_pop) return fail(GE_ERR_INVALID, "bad population index")
#define CHECK_CHR(ctx, c) if ((c) < 0 || (c) >= (ctx)->cfg.n_chr) return fail(GE_ERR_INVALID, "bad chromosome index")
#define CHECK_PHEN(ctx, f) if ((f) < 0 || (f) >= (ctx)->cfg.n_phen) return fail(GE_ERR_INVALID, "bad phenotype index")

