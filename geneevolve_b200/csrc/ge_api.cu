// ge_api.cu — context, device memory and the C-ABI (include/geneevolve_b200.h) of libgeneevolve_b200.so.
//
// The library owns the generation state on ONE GPU: bit-packed haplotype rows (double-buffered: parents /
// offspring), causal-variant planes, per-individual fp64 columns, pedigree, couples and draws.  The host
// (the reference's unchanged C++ front end, or the ctypes veneer in geneevolve_b200/capi.py) passes parsed
// flat arrays in and pulls `.info` / `.hap` / `.int` content out on demand.  There is no CPU compute path.
// Reference lines cited as :N are src/Simulation.cpp:N.
#include "ge_context.cuh"
#include <cstdlib>
#include "ge_segments.cuh"
#include "ge_mating.cuh"

// ------------------------------------------------------------------------------------------------
// setup helpers
// ------------------------------------------------------------------------------------------------
static int alloc_gen_state(ge_ctx *ctx, GenState &s) {
    uint64_t cap = ctx->cfg.capacity;
    int nf = ctx->cfg.n_phen;
    if (ctx->bits()) GE_TRY(ctx->ensure_exact(s.hap, (size_t)cap * 2 * ctx->W * 4));
    GE_TRY(ctx->ensure_exact(s.cv_allele, (size_t)cap * 2 * ctx->Wcv * 4));
    CUDA_TRY(cudaMemsetAsync(s.cv_allele.p, 0, (size_t)cap * 2 * ctx->Wcv * 4, ctx->stream));   // padding words stay zero
    if (ctx->use_root) GE_TRY(ctx->ensure_exact(s.cv_root, (size_t)cap * 2 * std::max<uint32_t>(ctx->n_cv_tot, 1)));
    GE_TRY(ctx->ensure_exact(s.ids, (size_t)cap * 7 * 8));
    GE_TRY(ctx->ensure_exact(s.sex, (size_t)cap));
    for (Buf *b : {&s.A, &s.D, &s.G, &s.C, &s.E, &s.F, &s.P}) {
        GE_TRY(ctx->ensure_exact(*b, (size_t)cap * nf * 8));
        CUDA_TRY(cudaMemsetAsync(b->p, 0, (size_t)cap * nf * 8, ctx->stream));
    }
    for (Buf *b : {&s.mv, &s.sv, &s.svf}) GE_TRY(ctx->ensure_exact(*b, (size_t)cap * 8));
    return GE_OK;
}

// genome layout + tile table, once all loci are known
static int build_genome(ge_ctx *ctx) {
    if (ctx->genome_ready) return GE_OK;
    int C = ctx->cfg.n_chr;
    ctx->chr_word_off.assign(C, 0); ctx->chr_nloci.assign(C, 0); ctx->locus_off.assign(C + 1, 0);
    std::vector<uint32_t> pos;
    uint32_t w = 0;
    for (int c = 0; c < C; c++) {
        const auto &L = ctx->loci[c];
        for (size_t s = 0; s < L.size(); s++) {
            if (L[s] > 0xFFFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "locus position does not fit 32 bits");
            if (s && L[s] < L[s - 1]) return fail(GE_ERR_UNSUPPORTED, "ge_set_loci: positions must be sorted ascending");
            pos.push_back((uint32_t)L[s]);
        }
        ctx->chr_word_off[c] = w; ctx->chr_nloci[c] = (uint32_t)L.size(); ctx->locus_off[c + 1] = (uint32_t)pos.size();
        uint32_t nw = (uint32_t)((L.size() + 31) / 32);
        w += (nw + 3) & ~3u;  // chromosomes start on 16-byte boundaries
    }
    ctx->W = std::max<uint32_t>((w + 31) & ~31u, 32);  // rows are 128-byte aligned
    GE_TRY(ctx->upload(ctx->d_chr_word_off, ctx->chr_word_off));
    GE_TRY(ctx->upload(ctx->d_chr_nloci, ctx->chr_nloci));
    GE_TRY(ctx->upload(ctx->d_locus_off, ctx->locus_off));
    GE_TRY(ctx->upload(ctx->d_pos, pos));
    GE_TRY(ctx->upload(ctx->d_chr_ids, ctx->chr_ids));
    // coarse position index (locus_lower_bound): about two buckets per locus, at least 1024 per chromosome
    std::vector<uint32_t> bkt_off(C + 1, 0), bkt_shift(C, 0), bkt;
    for (int c = 0; c < C; c++) {
        const auto &L = ctx->loci[c];
        uint64_t maxpos = L.empty() ? 0 : L.back();
        uint32_t sh = 0;
        while ((maxpos >> sh) + 1 > 2 * L.size() + 1024) sh++;
        bkt_shift[c] = sh;
        uint64_t nb = (maxpos >> sh) + 1;
        size_t s0 = 0;
        for (uint64_t b = 0; b < nb; b++) {
            while (s0 < L.size() && L[s0] < (b << sh)) s0++;
            bkt.push_back((uint32_t)s0);
        }
        bkt.push_back((uint32_t)L.size());
        bkt_off[c + 1] = (uint32_t)bkt.size();
    }
    GE_TRY(ctx->upload(ctx->d_bkt_off, bkt_off)); GE_TRY(ctx->upload(ctx->d_bkt_shift, bkt_shift)); GE_TRY(ctx->upload(ctx->d_bkt, bkt));
    // tile table: (chromosome, first chunk, chunk count), longest first so the warps of a CTA balance
    // 16-byte chunks per work item; smaller when this context owns few chromosomes (a shard of a multi-GPU run), so that
    // the 8 warps of a CTA still find a dozen items per offspring
    uint64_t chunks_per_gamete = 0;
    for (int c = 0; c < C; c++) chunks_per_gamete += ((ctx->chr_nloci[c] + 31) / 32 + 3) / 4;
    uint32_t TILE = 1024;  // 16 KB; measured 0.7 % better than 8 KB on the whole genome, 4 KB and 2 KB are 2 % worse
    while (TILE > 64 && chunks_per_gamete / TILE < 6) TILE >>= 1;
    if (const char *t = std::getenv("GE_TILE")) TILE = (uint32_t)std::max(16, std::atoi(t));  // measurement aid
    struct Item { uint32_t c, q0, nq; };
    std::vector<Item> items;
    for (int c = 0; c < C; c++) {
        uint32_t nw = (ctx->chr_nloci[c] + 31) / 32, nq = (nw + 3) / 4;
        for (uint32_t q = 0; q < nq; q += TILE) items.push_back({(uint32_t)c, q, std::min(TILE, nq - q)});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.nq > b.nq; });
    std::vector<uint32_t> tc, t0, tn;
    for (auto &it : items) { tc.push_back(it.c); t0.push_back(it.q0); tn.push_back(it.nq); }
    ctx->n_tiles = (uint32_t)items.size();
    GE_TRY(ctx->upload(ctx->d_tile_chr, tc));
    GE_TRY(ctx->upload(ctx->d_tile_chunk0, t0));
    GE_TRY(ctx->upload(ctx->d_tile_nchunk, tn));
    ctx->genome_ready = true;
    return GE_OK;
}

static std::vector<double> survival_table(const std::vector<double> &p, size_t first) {
    std::vector<double> T(p.size() + 1, 1.0);
    for (size_t k = 0; k < p.size(); k++) {
        double q = k < first ? 0.0 : p[k];
        if (q < 0) q = 0;
        if (q > 1) q = 1;
        T[k + 1] = T[k] * (1.0 - q);
    }
    return T;
}

// value index of one survival table for next_success (ge_kernels.cuh): entry b = first row k with T[k+1] < 1 - b/scale
static void value_index(const std::vector<double> &T, std::vector<uint32_t> &vb, double &scale) {
    const size_t R = T.size() - 1;
    const size_t B = std::max<size_t>(16, 2 * R);
    const double span = 1.0 - T[R];
    scale = span > 0 ? (double)B / span : 0.0;
    size_t k = 0;
    for (size_t b = 0; b <= B; b++) {
        const double v_hi = scale > 0 ? 1.0 - (double)b / scale : 1.0;
        while (k < R && !(T[k + 1] < v_hi)) k++;
        vb.push_back((uint32_t)std::min(k, R > 0 ? R - 1 : 0));
    }
}

static int build_maps(ge_ctx *ctx, PopDev &P) {
    int C = ctx->cfg.n_chr;
    std::vector<uint32_t> row_off(C + 1, 0), bp, dist, cov_lo(C), cov_hi(C);
    std::vector<double> T, vscale(C, 0.0);
    std::vector<uint32_t> vb, vb_off(C + 1, 0);
    for (int c = 0; c < C; c++) {
        if (P.rmap_bp[c].size() < 2) return fail(GE_ERR_INVALID, "genetic map of a chromosome is missing (ge_set_genetic_map)");
        for (uint64_t v : P.rmap_bp[c]) { if (v > 0xFFFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "map position does not fit 32 bits"); bp.push_back((uint32_t)v); }
        row_off[c + 1] = (uint32_t)bp.size();
        std::vector<double> t = survival_table(P.recom_prob[c], 0);
        T.insert(T.end(), t.begin(), t.end());
        value_index(t, vb, vscale[c]);
        vb_off[c + 1] = (uint32_t)vb.size();
        dist.push_back((uint32_t)P.bp_dist[c]);
        cov_lo[c] = (uint32_t)P.rmap_bp[c].front(); cov_hi[c] = (uint32_t)P.rmap_bp[c].back();
    }
    GE_TRY(ctx->upload(P.d_row_off, row_off)); GE_TRY(ctx->upload(P.d_bp, bp)); GE_TRY(ctx->upload(P.d_T, T));
    GE_TRY(ctx->upload(P.d_vb, vb)); GE_TRY(ctx->upload(P.d_vb_off, vb_off)); GE_TRY(ctx->upload(P.d_vb_scale, vscale));
    GE_TRY(ctx->upload(P.d_bp_dist, dist)); GE_TRY(ctx->upload(P.d_cov_lo, cov_lo)); GE_TRY(ctx->upload(P.d_cov_hi, cov_hi));
    if (P.has_mut) {
        std::vector<uint32_t> mro(C + 1, 0), mbp, mvb, mvb_off(C + 1, 0); std::vector<double> mT, mvscale(C, 0.0);
        for (int c = 0; c < C; c++) {
            for (uint64_t v : P.mutmap_bp[c]) mbp.push_back((uint32_t)v);
            mro[c + 1] = (uint32_t)mbp.size();
            std::vector<double> t = survival_table(P.mutmap_rate[c], 1);
            mT.insert(mT.end(), t.begin(), t.end());
            value_index(t, mvb, mvscale[c]);
            mvb_off[c + 1] = (uint32_t)mvb.size();
        }
        GE_TRY(ctx->upload(P.d_mrow_off, mro)); GE_TRY(ctx->upload(P.d_mbp, mbp)); GE_TRY(ctx->upload(P.d_mT, mT));
        GE_TRY(ctx->upload(P.d_mvb, mvb)); GE_TRY(ctx->upload(P.d_mvb_off, mvb_off)); GE_TRY(ctx->upload(P.d_mvb_scale, mvscale));
    }
    return GE_OK;
}

// causal-variant set shared by all populations (positions must agree; effect sizes are per root population)
static int build_cvset(ge_ctx *ctx) {
    if (ctx->cv_ready) return GE_OK;
    int C = ctx->cfg.n_chr, nf = ctx->cfg.n_phen, np = ctx->cfg.n_pop;
    ctx->cv_block_off.assign((size_t)nf * C + 1, 0);
    std::vector<uint32_t> bp, chr_of;
    for (int f = 0; f < nf; f++)
        for (int c = 0; c < C; c++) {
            const CvHost &h = ctx->pop[0].cv[f][c];
            for (int p = 1; p < np; p++)
                if (ctx->pop[p].cv[f][c].bp != h.bp) return fail(GE_ERR_UNSUPPORTED, "all populations must list the same causal-variant positions (ras_find_cv indexes every root population's cv_info by the same icv, :2762)");
            for (uint64_t v : h.bp) { bp.push_back((uint32_t)v); chr_of.push_back((uint32_t)c); }
            ctx->cv_block_off[(size_t)f * C + c + 1] = (uint32_t)bp.size();
        }
    ctx->n_cv_tot = (uint32_t)bp.size();
    // bit-plane layout: every (phenotype, chromosome) block starts on a word boundary
    ctx->cv_word_off.assign((size_t)nf * C + 1, 0);
    std::vector<uint32_t> word_blk;
    for (int b = 0; b < nf * C; b++) {
        uint32_t nw = (ctx->cv_block_off[b + 1] - ctx->cv_block_off[b] + 31) / 32;
        ctx->cv_word_off[b + 1] = ctx->cv_word_off[b] + nw;
        word_blk.insert(word_blk.end(), nw, (uint32_t)b);
    }
    ctx->Wcv = std::max<uint32_t>((ctx->cv_word_off.back() + 3) & ~3u, 4);
    word_blk.resize(ctx->Wcv, 0xFFFFFFFFu);
    ctx->cv_sorted = true;
    for (int b = 0; b < nf * C; b++)
        for (uint32_t k = ctx->cv_block_off[b] + 1; k < ctx->cv_block_off[b + 1]; k++) if (bp[k] < bp[k - 1]) ctx->cv_sorted = false;
    GE_TRY(ctx->ensure(ctx->d_LA, (size_t)std::max<uint32_t>(ctx->n_cv_tot, 1) * 48));
    std::vector<uint32_t> bitpos(ctx->n_cv_tot);
    for (int b = 0; b < nf * C; b++)
        for (uint32_t k = ctx->cv_block_off[b]; k < ctx->cv_block_off[b + 1]; k++) bitpos[k] = ctx->cv_word_off[b] * 32 + (k - ctx->cv_block_off[b]);
    GE_TRY(ctx->upload(ctx->d_cv_bitpos, bitpos));
    GE_TRY(ctx->upload(ctx->d_cv_word_off, ctx->cv_word_off)); GE_TRY(ctx->upload(ctx->d_cv_word_blk, word_blk));
    std::vector<double> a_eff((size_t)np * ctx->n_cv_tot), d_eff((size_t)np * ctx->n_cv_tot);
    for (int p = 0; p < np; p++)
        for (int f = 0; f < nf; f++)
            for (int c = 0; c < C; c++) {
                const CvHost &h = ctx->pop[p].cv[f][c];
                uint32_t b0 = ctx->cv_block_off[(size_t)f * C + c];
                uint64_t lo = ctx->pop[p].rmap_bp[c].front(), hi = ctx->pop[p].rmap_bp[c].back();
                for (size_t k = 0; k < h.bp.size(); k++) {
                    bool cov = h.bp[k] >= lo && h.bp[k] < hi;  // uncovered CVs keep a = d = 0 (Human_CV ctor, src/Population.h:96-108)
                    a_eff[(size_t)p * ctx->n_cv_tot + b0 + k] = cov ? h.a[k] : 0.0;
                    d_eff[(size_t)p * ctx->n_cv_tot + b0 + k] = cov ? h.d[k] : 0.0;
                }
            }
    GE_TRY(ctx->upload(ctx->d_cv_block_off, ctx->cv_block_off));
    GE_TRY(ctx->upload(ctx->d_cv_bp, bp)); GE_TRY(ctx->upload(ctx->d_cv_chr, chr_of));
    GE_TRY(ctx->upload(ctx->d_a_eff, a_eff)); GE_TRY(ctx->upload(ctx->d_d_eff, d_eff));
    // ras_find_cv takes a, d from the ROOT population of each allele (:2776-2786); that only matters when the populations'
    // effect tables differ — otherwise no root plane is carried and the tabulated single-population path applies
    ctx->use_root = false;
    for (int p = 1; p < np && !ctx->use_root; p++)
        for (uint32_t k = 0; k < ctx->n_cv_tot; k++)
            if (a_eff[(size_t)p * ctx->n_cv_tot + k] != a_eff[k] || d_eff[(size_t)p * ctx->n_cv_tot + k] != d_eff[k]) { ctx->use_root = true; break; }
    GE_TRY(ctx->ensure(ctx->d_cv_count, (size_t)std::max<uint32_t>(ctx->n_cv_tot, 1) * 8));
    ctx->cv_ready = true;
    return GE_OK;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char *ge_last_error(void) { return g_err.c_str(); }
int ge_version(void) { return 100; }

int ge_create(const ge_config *cfg, ge_ctx **out) {
    if (!cfg || !out) return fail(GE_ERR_INVALID, "null argument");
    if (cfg->n_pop < 1 || cfg->n_pop > 15 || cfg->n_chr < 1 || cfg->n_phen < 1) return fail(GE_ERR_INVALID, "bad n_pop/n_chr/n_phen");
    if (cfg->capacity == 0) return fail(GE_ERR_INVALID, "capacity must be > 0");
    if (cfg->rng_mode != GE_RNG_PHILOX && cfg->rng_mode != GE_RNG_REPLAY) return fail(GE_ERR_INVALID, "rng_mode must be GE_RNG_PHILOX or GE_RNG_REPLAY");
    if (!(cfg->representation & (GE_REP_BITS | GE_REP_SEGMENTS))) return fail(GE_ERR_INVALID, "representation must include GE_REP_BITS and/or GE_REP_SEGMENTS");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(GE_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)");
    CUDA_TRY(cudaSetDevice(cfg->device));
    ge_ctx *c = new ge_ctx();
    c->cfg = *cfg;
    c->rng.k0 = (uint32_t)cfg->seed; c->rng.k1 = (uint32_t)(cfg->seed >> 32);
    c->pop.resize(cfg->n_pop);
    c->loci.resize(cfg->n_chr);
    c->chr_ids.resize(cfg->n_chr);
    for (int k = 0; k < cfg->n_chr; k++) c->chr_ids[k] = (uint32_t)k;
    for (PopDev &P : c->pop) {
        P.rmap_bp.resize(cfg->n_chr); P.recom_prob.resize(cfg->n_chr); P.bp_dist.assign(cfg->n_chr, 1);
        P.mutmap_bp.resize(cfg->n_chr); P.mutmap_rate.resize(cfg->n_chr);
        P.cv.assign(cfg->n_phen, std::vector<CvHost>(cfg->n_chr));
        P.scheme.resize(cfg->n_phen);
        P.panel.resize(cfg->n_chr);
        P.var_a0.assign(cfg->n_phen, 0); P.var_d0.assign(cfg->n_phen, 0);
    }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi);
    cudaStreamCreateWithPriority(&c->bulk, cudaStreamNonBlocking, prio_lo);
    c->serial = std::getenv("GE_SERIAL") != nullptr;
    c->cub_sorts = std::getenv("GE_CUB_SORTS") != nullptr;
    if (const char *t = std::getenv("GE_THIN")) c->thin = std::atoi(t);
    c->cv_from_segments = std::getenv("GE_CV_FROM_SEGMENTS") != nullptr;
    c->seg_per_thread = std::getenv("GE_SEG_PER_THREAD") != nullptr;
    if (const char *t = std::getenv("GE_SEG_GROUP")) c->seg_group = std::atoi(t);
    c->seg_walk = std::getenv("GE_SEG_WALK") != nullptr;
    c->seg_sync_mode = std::getenv("GE_SEG_SYNC") != nullptr;
    if (const char *t = std::getenv("GE_SEG_FORMAT")) c->seg_wide = std::atoi(t) == 16;
    if (const char *t = std::getenv("GE_SEG_DEPTH")) c->seg_depth = std::atoi(t);
    if (const char *t = std::getenv("GE_SEG_PLAN_MIN")) c->seg_plan_min_parts = std::atof(t);
    if (const char *t = std::getenv("GE_PROP_DEPTH")) c->prop_depth = std::atoi(t);
    if (const char *t = std::getenv("GE_PROP")) c->use_tma = std::string(t) == "tma";
    if (const char *t = std::getenv("GE_THIN_MIN_GB")) c->thin_min_bytes = std::atof(t) * 1e9;  // measurement aid: queue the bulk kernel on the control stream (no overlap)
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    for (PopDev &P : c->pop) for (DrawSet &D : P.ds) cudaEventCreateWithFlags(&D.bulk_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
    for (SortLane &l : c->lane) { cudaStreamCreateWithPriority(&l.s, cudaStreamNonBlocking, prio_hi); cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming); }
    cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, cfg->device);
    if (c->ensure(c->flags, 64) != GE_OK) { delete c; return GE_ERR_CUDA; }
    cudaMemsetAsync(c->flags.p, 0, 64, c->stream);
    *out = c;
    return GE_OK;
}

int ge_destroy(ge_ctx *ctx) {
    if (!ctx) return GE_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->bulk);
    cudaStreamSynchronize(ctx->stream);
    for (SortLane &l : ctx->lane) cudaStreamSynchronize(l.s);
    auto freeb = [&](Buf &b) { ctx->release(b); };
    for (PopDev &P : ctx->pop) {
        for (Buf *b : {&P.d_row_off, &P.d_bp, &P.d_T, &P.d_bp_dist, &P.d_mrow_off, &P.d_mbp, &P.d_mT, &P.d_vb, &P.d_vb_off, &P.d_vb_scale, &P.d_mvb, &P.d_mvb_off, &P.d_mvb_scale, &P.d_cov_lo, &P.d_cov_hi, &P.d_omega,
                       &P.d_lambda, &P.d_vd_zero, &P.prev_P, &P.prev_F, &P.c_male, &P.c_female, &P.c_inbreed, &P.c_noff, &P.mut_off, &P.mut_bp, &P.mut_gam, &P.e_raw, &P.cnt32, &P.d_sv0, &P.founder_rows, &P.founder_cv})
            freeb(*b);
        freeb(P.mig_pop[0]); freeb(P.mig_idx[0]); freeb(P.rowmap_buf[0]); freeb(P.rowmap_buf[1]);
        for (DrawSet &D : P.ds) {
            for (Buf *b : {&D.father, &D.mother, &D.couple_of, &D.xo_off, &D.xo_bp, &D.flips, &D.start_hap}) freeb(*b);
            cudaEventDestroy(D.bulk_done);
        }
        for (GenState &s : P.st) {
            for (Buf *b : {&s.hap, &s.cv_allele, &s.cv_root, &s.ids, &s.sex, &s.A, &s.D, &s.G, &s.C, &s.E, &s.F, &s.P, &s.mv, &s.sv, &s.svf, &s.hm_off, &s.hm_bp}) freeb(*b);
            for (cudaEvent_t e : s.seg.ev) if (e) cudaEventDestroy(e);
            seg_release(s.seg);
        }
        mate_release(P.mate);
    }
    for (Buf *b : {&ctx->d_chr_word_off, &ctx->d_chr_nloci, &ctx->d_locus_off, &ctx->d_pos, &ctx->d_bkt_off, &ctx->d_bkt_shift, &ctx->d_bkt, &ctx->d_LA, &ctx->d_cv_bitpos, &ctx->xo_stash, &ctx->d_tile_chr, &ctx->d_tile_chunk0, &ctx->d_tile_nchunk,
                   &ctx->d_cv_block_off, &ctx->d_cv_word_off, &ctx->d_cv_word_blk, &ctx->d_cv_bp, &ctx->d_cv_chr, &ctx->d_a_eff, &ctx->d_d_eff, &ctx->d_cv_count, &ctx->scan_blocks,
                   &ctx->scan_total, &ctx->partial, &ctx->scalars, &ctx->flags, &ctx->d_chr_ids, &ctx->ar_scratch, &ctx->seg_desc, &ctx->seg_iv_off, &ctx->seg_cnt,
                   &ctx->seg_scan_blocks, &ctx->seg_scan_total, &ctx->seg_flags, &ctx->seg_verb})
        freeb(*b);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev_ready); cudaEventDestroy(ctx->ev_join);
    for (auto &e : ctx->ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (SortLane &l : ctx->lane) {
        for (Buf *b : {&l.keys_in, &l.keys_out, &l.vals_out, &l.tmp}) freeb(*b);
        cudaEventDestroy(l.done); cudaStreamDestroy(l.s);
    }
    cudaEventDestroy(ctx->ev_fork);
    for (int q = 0; q < 2; q++) { for (Buf &b : ctx->mig_lists[q]) freeb(b); if (ctx->mig_done[q]) cudaEventDestroy(ctx->mig_done[q]); }
    freeb(ctx->mig_stage);
    cudaStreamDestroy(ctx->bulk);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return GE_OK;
}

int ge_set_population(ge_ctx *ctx, int pop, int avoid_inbreeding, int random_mating, double mm) {
    CHECK_POP(ctx, pop);
    PopDev &P = ctx->pop[pop];
    P.avoid_inbreeding = avoid_inbreeding; P.RM = random_mating; P.MM = mm;
    return GE_OK;
}
int ge_set_genetic_map(ge_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rp, uint64_t n, uint64_t bp_dist) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    if (!bp || !rp || n < 2 || bp_dist == 0 || bp_dist > 0xFFFFFFFFull) return fail(GE_ERR_INVALID, "ge_set_genetic_map: need >= 2 rows and 0 < bp_dist < 2^32");
    // the bit-packed representation needs monotone crossover lists: row j's crossover lies in
    // [bp[j], bp[j]+bp_dist) (:2989), so rows must be at least bp_dist apart (true for the uniform b37 maps)
    for (uint64_t j = 0; j + 2 < n; j++)
        if (bp[j + 1] < bp[j] + bp_dist && rp[j] > 0) {
            if (ctx->cfg.representation & GE_REP_BITS)
                return fail(GE_ERR_UNSUPPORTED, "genetic map rows closer than bp_dist_in_rmap give non-monotone crossover lists (only GE_REP_SEGMENTS follows the reference there)");
            ctx->seg_per_thread = true;  // segment lists may become unsorted: keep the reference's scan verbatim
        }
    PopDev &P = ctx->pop[pop];
    P.rmap_bp[chr].assign(bp, bp + n); P.recom_prob[chr].assign(rp, rp + n); P.bp_dist[chr] = bp_dist;
    return GE_OK;
}
int ge_set_mutation_map(ge_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rate, uint64_t n) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    if (!bp || !rate || n < 2) return fail(GE_ERR_INVALID, "ge_set_mutation_map: need >= 2 rows");
    for (uint64_t k = 0; k < n; k++) if (bp[k] > 0xFFFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "mutation-map position does not fit 32 bits");
    PopDev &P = ctx->pop[pop];
    P.mutmap_bp[chr].assign(bp, bp + n); P.mutmap_rate[chr].assign(rate, rate + n); P.has_mut = true;
    return GE_OK;
}
int ge_set_loci(ge_ctx *ctx, int chr, const uint64_t *pos, uint64_t n) {
    CHECK_CTX(ctx); CHECK_CHR(ctx, chr);
    if (ctx->genome_ready) return fail(GE_ERR_INVALID, "ge_set_loci after the genome layout was frozen");
    ctx->loci[chr].assign(pos, pos + n);
    return GE_OK;
}
int ge_set_founder_panel(ge_ctx *ctx, int pop, int chr, const uint8_t *al, uint64_t nh) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    if (!al || nh == 0 || (nh & 1)) return fail(GE_ERR_INVALID, "ge_set_founder_panel: need an even, non-zero number of founder haplotypes");
    if (ctx->loci[chr].empty()) return fail(GE_ERR_INVALID, "ge_set_founder_panel: call ge_set_loci for this chromosome first");
    PopDev &P = ctx->pop[pop];
    P.panel[chr].assign(al, al + nh * ctx->loci[chr].size());
    P.n_founder_haps = nh;
    return GE_OK;
}
int ge_set_founder_panel_packed(ge_ctx *ctx, int pop, int chr, const uint32_t *words, uint64_t nh) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    if (!words || nh == 0 || (nh & 1)) return fail(GE_ERR_INVALID, "ge_set_founder_panel_packed: need an even, non-zero number of founder haplotypes");
    if (ctx->loci[chr].empty()) return fail(GE_ERR_INVALID, "ge_set_founder_panel_packed: call ge_set_loci for this chromosome first");
    PopDev &P = ctx->pop[pop];
    if (P.panel_packed.empty()) P.panel_packed.resize(ctx->cfg.n_chr);
    uint64_t nw = (ctx->loci[chr].size() + 31) / 32;
    P.panel_packed[chr].assign(words, words + nh * nw);
    P.n_founder_haps = nh;
    return GE_OK;
}
int ge_set_cv(ge_ctx *ctx, int pop, int phen, int chr, const uint64_t *bp, const double *a, const double *d, uint64_t ncv, const uint8_t *val, uint64_t nh) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr); CHECK_PHEN(ctx, phen);
    if (ncv && (!bp || !a || !d || !val)) return fail(GE_ERR_INVALID, "ge_set_cv: null array");
    if (nh == 0 || (nh & 1)) return fail(GE_ERR_INVALID, "ge_set_cv: need an even, non-zero number of founder haplotypes");
    for (uint64_t k = 0; k < ncv; k++) if (bp[k] > 0xFFFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "causal-variant position does not fit 32 bits");
    CvHost &h = ctx->pop[pop].cv[phen][chr];
    h.bp.assign(bp, bp + ncv); h.a.assign(a, a + ncv); h.d.assign(d, d + ncv); h.val.assign(val, val + nh * ncv); h.nhap = nh;
    return GE_OK;
}
int ge_set_pheno_scheme(ge_ctx *ctx, int pop, int phen, double va, double vd, double ve, double vc, double vf, double omega, double beta, double lambda) {
    CHECK_POP(ctx, pop); CHECK_PHEN(ctx, phen);
    Scheme &S = ctx->pop[pop].scheme[phen];
    S.va = va; S.vd = vd; S.ve = ve; S.vc = vc; S.vf = vf; S.omega = omega; S.beta = beta; S.lambda = lambda;
    return GE_OK;
}
int ge_set_chromosome_ids(ge_ctx *ctx, const int32_t *ids) {
    CHECK_CTX(ctx);
    if (ctx->genome_ready) return fail(GE_ERR_INVALID, "ge_set_chromosome_ids after the genome layout was frozen");
    for (int k = 0; k < ctx->cfg.n_chr; k++) { if (ids[k] < 0 || ids[k] > 32767) return fail(GE_ERR_INVALID, "bad chromosome id"); ctx->chr_ids[k] = (uint32_t)ids[k]; }
    return GE_OK;
}
int ge_set_allreduce(ge_ctx *ctx, ge_allreduce_fn fn, void *user) { CHECK_CTX(ctx); ctx->allreduce = fn; ctx->allreduce_user = user; return GE_OK; }
int ge_set_gamma(ge_ctx *ctx, const double *g) { CHECK_CTX(ctx); ctx->gamma.assign(g, g + ctx->cfg.n_phen); return GE_OK; }

// ---------------- per-method entry points ----------------

int ge_compute_AD(ge_ctx *ctx, int pop, int gen) {  // ras_compute_AD :2624-2749
    CHECK_POP(ctx, pop);
    (void)gen;
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    if (S.n == 0) return fail(GE_ERR_INVALID, "empty population");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    ge_ctx::PhaseTimer timer(ctx, GE_PHASE_CV_AD);
    if (ctx->segs() && !ctx->bits() && ctx->cv_from_segments) GE_TRY(seg_find_cv(ctx, pop));  // ras_find_cv on the segment lists (verification mode)
    uint32_t ncv = ctx->n_cv_tot;
    if (ncv) {
        CUDA_TRY(cudaMemsetAsync(ctx->d_cv_count.p, 0, (size_t)ncv * 8, ctx->stream));
        dim3 grid(nblk(ctx->Wcv, 32), nblk(2 * S.n, 512));
        cv_count_bits_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(ctx->cvset(), S.cv_allele.as<uint32_t>(), 2 * S.n, ctx->d_cv_count.as<unsigned long long>());
        GE_TRY(ctx->check_launch("cv_count"));
    }
    uint64_t nw = S.n * ctx->cfg.n_phen;
    if (!ctx->use_root && ncv) {
        cv_tables_kernel<<<nblk(ncv, 128), 128, 0, ctx->stream>>>(ctx->cvset(), ctx->d_cv_count.as<unsigned long long>(), S.n, ctx->d_a_eff.as<double>(),
                                                                  ctx->d_d_eff.as<double>(), P.d_vd_zero.as<uint8_t>(), ctx->d_LA.as<double2>());
        GE_TRY(ctx->check_launch("cv_tables"));
        genetic_value_lut_kernel<<<ctx->ctrl_grid(nw * 32, 256), 256, 0, ctx->stream>>>(ctx->cvset(), S.cv_allele.as<uint32_t>(), ctx->d_cv_bitpos.as<uint32_t>(), ctx->d_LA.as<double2>(), S.n,
                                                                              S.A.as<double>(), S.D.as<double>(), S.G.as<double>(), ctx->flags.as<int>());
        GE_TRY(ctx->check_launch("genetic_value_lut"));
    } else {
        genetic_value_kernel<<<nblk(nw * 32, 256), 256, 0, ctx->stream>>>(
            ctx->cvset(), S.cv_allele.as<uint32_t>(), ctx->use_root ? S.cv_root.as<uint8_t>() : nullptr, ctx->d_cv_count.as<unsigned long long>(), S.n,
            ctx->d_a_eff.as<double>(), ctx->d_d_eff.as<double>(), P.d_vd_zero.as<uint8_t>(), S.n, S.A.as<double>(), S.D.as<double>(), S.G.as<double>(),
            ctx->flags.as<int>());
        GE_TRY(ctx->check_launch("genetic_value"));
    }
    if (ctx->allreduce) {
        // chromosome-sharded contexts hold partial sums over their own chromosomes: sum A, D, G over the ranks
        size_t nb = (size_t)S.n * ctx->cfg.n_phen * 8;
        GE_TRY(ctx->ensure(ctx->ar_scratch, 3 * nb));
        char *sc = ctx->ar_scratch.as<char>();
        CUDA_TRY(cudaMemcpyAsync(sc, S.A.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(sc + nb, S.D.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(sc + 2 * nb, S.G.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        int rc = ctx->allreduce(ctx->allreduce_user, reinterpret_cast<double *>(sc), 3 * nb / 8, (void *)ctx->stream);
        if (rc != 0) return fail(GE_ERR_INVALID, "allreduce hook failed");
        CUDA_TRY(cudaMemcpyAsync(S.A.p, sc, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(S.D.p, sc + nb, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(S.G.p, sc + 2 * nb, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return GE_OK;
}

static int scale_AD_compute_GEF_impl(ge_ctx *ctx, int pop, int gen, int f, const double *e_host, const double *f0_host) {  // :3075-3206
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    uint64_t n = S.n;
    Scheme &sc = P.scheme[f];
    GE_TRY(ctx->ensure(P.e_raw, (size_t)ctx->cfg.capacity * ctx->cfg.n_phen * 8));
    GE_TRY(ctx->ensure(ctx->scalars, 64 * 8));
    double *e = P.e_raw.as<double>() + (uint64_t)f * n;
    if (e_host) CUDA_TRY(cudaMemcpyAsync(e, e_host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    else if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
        enoise_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->rng, pop, gen, f, 0, n, e);
        GE_TRY(ctx->check_launch("enoise"));
    } else return fail(GE_ERR_INVALID, "replay mode needs e_raw");
    P.have_e_raw = true;
    double *var_e = ctx->scalars.as<double>() + 8, *tmp = ctx->scalars.as<double>() + 9;
    GE_TRY(ctx->d_var(e, n, var_e, tmp));
    double *f0 = nullptr;
    if (gen == 0 && sc.vf > 0) {
        f0 = S.F.as<double>() + (uint64_t)f * n;  // staged in place, the kernel reads f0[i] before writing F[i]
        if (f0_host) CUDA_TRY(cudaMemcpyAsync(f0, f0_host, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        else if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
            normal_scaled_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->rng, P_F0, pop, 0, f, 0, n, std::sqrt(sc.vf), f0);
            GE_TRY(ctx->check_launch("f0"));
        } else return fail(GE_ERR_INVALID, "replay mode needs parental0 for vf > 0");
    }
    PhenoArgs a;
    a.s_a = 1; if (sc.va > 0) a.s_a = std::sqrt(P.var_a0[f] / sc.va);
    a.s_d = 0; if (sc.vd > 0) a.s_d = std::sqrt(P.var_d0[f] / sc.vd); else if (sc.vd == -1) a.s_d = 1;
    a.ve = sc.ve; a.vf = sc.vf; a.beta = sc.beta; a.gen = gen; a.vt_type = ctx->cfg.vt_type; a.n = n; a.prev_n = P.prev_n;
    uint64_t o = (uint64_t)f * n;
    phenotype_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(
        a, e, var_e, S.A.as<double>() + o, S.D.as<double>() + o, S.G.as<double>() + o, S.C.as<double>() + o, S.E.as<double>() + o,
        S.F.as<double>() + o, S.P.as<double>() + o, S.ids.as<uint64_t>(), P.prev_P.as<double>() ? P.prev_P.as<double>() + (uint64_t)f * P.prev_n : nullptr,
        P.prev_F.as<double>() ? P.prev_F.as<double>() + (uint64_t)f * P.prev_n : nullptr, f0, ctx->flags.as<int>() + 1);
    return ctx->check_launch("phenotype");
}

int ge_scale_AD_compute_GEF(ge_ctx *ctx, int pop, int gen, int phen, const double *e_raw) {
    CHECK_POP(ctx, pop); CHECK_PHEN(ctx, phen);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(scale_AD_compute_GEF_impl(ctx, pop, gen, phen, e_raw, nullptr));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // e_raw is a caller buffer
    return GE_OK;
}

int ge_compute_mating_value_selection_value(ge_ctx *ctx, int pop, int gen, const ge_gen_params *gp) {  // :3300-3342
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    uint64_t n = S.n;
    GE_TRY(ctx->ensure(P.d_sv0, 16));
    // sv_raw staged in S.sv, standardised in place by selection_kernel
    mv_sv_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, ctx->cfg.n_phen, S.P.as<double>(), P.d_omega.as<double>(), P.d_lambda.as<double>(),
                                                        S.mv.as<double>(), S.sv.as<double>());
    GE_TRY(ctx->check_launch("mv_sv"));
    double *sv0 = P.d_sv0.as<double>();
    if (gen == 0) {
        GE_TRY(ctx->d_var(S.sv.as<double>(), n, sv0 + 1, sv0 + 0));
        if (n <= 1) GE_TRY(ctx->d_mean(S.sv.as<double>(), n, sv0));
        double h[2];
        CUDA_TRY(cudaMemcpyAsync(h, sv0, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        P.sv_mean0 = h[0]; P.sv_var0 = h[1];
    }
    ge_gen_params dummy{};
    if (!gp) gp = &dummy;
    selection_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, gen, gp->selection_func, gp->selection_par1, gp->selection_par2, sv0, sv0 + 1,
                                                            S.sv.as<double>(), S.sv.as<double>(), S.svf.as<double>());
    return ctx->check_launch("selection");
}

int ge_save_human_info_to_Pop_info_prev_gen(ge_ctx *ctx, int pop) {  // :3211-3236
    CHECK_POP(ctx, pop);
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    size_t bytes = (size_t)ctx->cfg.capacity * ctx->cfg.n_phen * 8;
    GE_TRY(ctx->ensure(P.prev_P, bytes)); GE_TRY(ctx->ensure(P.prev_F, bytes));
    size_t nb = (size_t)S.n * ctx->cfg.n_phen * 8;
    CUDA_TRY(cudaMemcpyAsync(P.prev_P.p, S.P.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(P.prev_F.p, S.F.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    P.prev_n = S.n;
    return GE_OK;
}

int ge_environmental_effects_specific_to_each_population(ge_ctx *ctx, int f) {  // :3345-3381
    CHECK_CTX(ctx); CHECK_PHEN(ctx, f);
    if (ctx->gamma.empty() || ctx->gamma[f] == 0) return GE_OK;
    int np = ctx->cfg.n_pop;
    if (np < 2) return fail(GE_ERR_UNSUPPORTED, "--gamma with one population divides by zero in the reference (:3269)");
    // pooled variance of P + a*s_i is quadratic in a; the moments come from the device, the scalar Newton
    // iteration (NewtonRaphson :44-63, central difference :35-39) runs on the host with the same start/tolerance
    std::vector<double> ni(np), mu(np), m2(np), s(np);
    double N = 0;
    for (int p = 0; p < np; p++) {
        GenState &S = ctx->pop[p].st[ctx->pop[p].cur];
        double v, m;
        GE_TRY(ctx->h_var(S.P.as<double>() + (uint64_t)f * S.n, S.n, &v, &m));
        ni[p] = (double)S.n; mu[p] = m; m2[p] = v * (double)(S.n > 1 ? S.n - 1 : 0);
        s[p] = (double)(2 * p / (np - 1) - 1);  // integer arithmetic as in the reference (:3269, :3289)
        N += ni[p];
    }
    double mx = 0, ms = 0;
    for (int p = 0; p < np; p++) { mx += ni[p] * mu[p]; ms += ni[p] * s[p]; }
    mx /= N; ms /= N;
    double Sxx = 0, Sxs = 0, Sss = 0;
    for (int p = 0; p < np; p++) {
        Sxx += m2[p] + ni[p] * (mu[p] - mx) * (mu[p] - mx);
        Sxs += ni[p] * (mu[p] - mx) * (s[p] - ms);
        Sss += ni[p] * (s[p] - ms) * (s[p] - ms);
    }
    double Vx = Sxx / (N - 1), Cxs = Sxs / (N - 1), Vs = Sss / (N - 1), gam = ctx->gamma[f];
    auto fn = [&](double a) { return (Vx + 2 * a * Cxs + a * a * Vs) - (1 + gam) * Vx; };
    double x0 = 10, x1 = x0;
    for (int it = 0; it < 200; it++) {
        const double dx = 0.001;
        double fp = (fn(x0 + dx) - fn(x0 - dx)) / (2 * dx);
        x1 = x0 - fn(x0) / fp;
        if (std::abs(fn(x1)) < 1e-4) break;
        x0 = x1;
    }
    for (int p = 0; p < np; p++) {
        GenState &S = ctx->pop[p].st[ctx->pop[p].cur];
        add_scalar_kernel<<<nblk(S.n, 256), 256, 0, ctx->stream>>>(S.P.as<double>() + (uint64_t)f * S.n, S.n, x1 * s[p]);
        GE_TRY(ctx->check_launch("add_scalar"));
    }
    return GE_OK;
}

// ---------------- generation 0 ----------------

static int init_pop_gen0(ge_ctx *ctx, int p, const ge_draws *d0) {  // ras_initial_human_gen0 :3000-3072
    PopDev &P = ctx->pop[p];
    int C = ctx->cfg.n_chr, nf = ctx->cfg.n_phen;
    uint64_t nhaps = P.cv[0][0].nhap, n = nhaps / 2;
    if (n == 0) return fail(GE_ERR_INVALID, "no founders: call ge_set_cv first (nhaps is read from the CV panel, :3009)");
    if (n > ctx->cfg.capacity) return fail(GE_ERR_CAPACITY, "founders exceed capacity");
    for (GenState &s : P.st) GE_TRY(alloc_gen_state(ctx, s));
    P.cur = 0;
    GenState &S = P.st[0];
    S.n = n;
    // scheme constants on the device
    std::vector<double> om(nf), la(nf); std::vector<uint8_t> vz(nf);
    for (int f = 0; f < nf; f++) { om[f] = P.scheme[f].omega; la[f] = P.scheme[f].lambda; vz[f] = P.scheme[f].vd == 0; }
    GE_TRY(ctx->upload(P.d_omega, om)); GE_TRY(ctx->upload(P.d_lambda, la)); GE_TRY(ctx->upload(P.d_vd_zero, vz));
    // bit-packed rows from the founder panel
    bool have_panel = !P.panel_packed.empty();
    for (int c = 0; c < C && !have_panel; c++) have_panel = !P.panel[c].empty();
    if (ctx->bits() && !have_panel) return fail(GE_ERR_INVALID, "GE_REP_BITS needs the founder panel (ge_set_founder_panel)");
    if (have_panel) {
        uint32_t *rows0 = S.hap.as<uint32_t>();
        if (ctx->segs()) { GE_TRY(ctx->ensure_exact(P.founder_rows, (size_t)n * 2 * ctx->W * 4)); rows0 = P.founder_rows.as<uint32_t>(); }
        CUDA_TRY(cudaMemsetAsync(rows0, 0, (size_t)n * 2 * ctx->W * 4, ctx->stream));
        for (int c = 0; c < C; c++) {
            uint32_t nl = ctx->chr_nloci[c];
            if (nl == 0) continue;
            bool packed = !P.panel_packed.empty() && !P.panel_packed[c].empty();
            uint32_t nwc = (nl + 31) / 32;
            if (packed) {
                if (P.panel_packed[c].size() != (size_t)2 * n * nwc) return fail(GE_ERR_INVALID, "packed founder panel of the wrong size");
                Buf tmp;
                GE_TRY(ctx->ensure_exact(tmp, P.panel_packed[c].size() * 4));
                CUDA_TRY(cudaMemcpyAsync(tmp.p, P.panel_packed[c].data(), P.panel_packed[c].size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                uint64_t tot = (uint64_t)2 * n * nwc;
                mask_packed_panel_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(tmp.as<uint32_t>(), (uint32_t)(2 * n), nl, ctx->d_pos.as<uint32_t>() + ctx->locus_off[c],
                                                                                  (uint32_t)P.rmap_bp[c].front(), (uint32_t)P.rmap_bp[c].back(), rows0, ctx->W, ctx->chr_word_off[c]);
                GE_TRY(ctx->check_launch("mask_packed_panel"));
                CUDA_TRY(cudaStreamSynchronize(ctx->stream));
                ctx->release(tmp);
                std::vector<uint32_t>().swap(P.panel_packed[c]);
                continue;
            }
            if (P.panel[c].size() != (size_t)2 * n * nl) return fail(GE_ERR_INVALID, "founder panel missing or of the wrong size (ge_set_founder_panel)");
            Buf tmp;
            GE_TRY(ctx->ensure_exact(tmp, P.panel[c].size()));
            CUDA_TRY(cudaMemcpyAsync(tmp.p, P.panel[c].data(), P.panel[c].size(), cudaMemcpyHostToDevice, ctx->stream));
            uint64_t tot = (uint64_t)2 * n * nwc;
            pack_panel_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(tmp.as<uint8_t>(), (uint32_t)(2 * n), nl, ctx->d_pos.as<uint32_t>() + ctx->locus_off[c],
                                                                       (uint32_t)P.rmap_bp[c].front(), (uint32_t)P.rmap_bp[c].back(), rows0, ctx->W, ctx->chr_word_off[c]);
            GE_TRY(ctx->check_launch("pack_panel"));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            ctx->release(tmp);
        }
        if (ctx->segs() && ctx->bits())
            CUDA_TRY(cudaMemcpyAsync(S.hap.p, P.founder_rows.p, (size_t)n * 2 * ctx->W * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    // causal-variant planes
    if (ctx->n_cv_tot) {
        std::vector<uint8_t> fcv((size_t)2 * n * ctx->n_cv_tot);
        for (int f = 0; f < nf; f++)
            for (int c = 0; c < C; c++) {
                const CvHost &h = P.cv[f][c];
                if (h.nhap != nhaps) return fail(GE_ERR_INVALID, "CV panels disagree on the number of founder haplotypes");
                uint32_t b0 = ctx->cv_block_off[(size_t)f * C + c], ncv = (uint32_t)h.bp.size();
                for (uint64_t r = 0; r < 2 * n; r++) std::memcpy(&fcv[r * ctx->n_cv_tot + b0], &h.val[r * ncv], ncv);
            }
        Buf tmp_local;
        Buf &tmp = ctx->segs() ? P.founder_cv : tmp_local;
        GE_TRY(ctx->ensure_exact(tmp, fcv.size()));
        CUDA_TRY(cudaMemcpyAsync(tmp.p, fcv.data(), fcv.size(), cudaMemcpyHostToDevice, ctx->stream));
        uint64_t tot = (uint64_t)2 * n * ctx->Wcv;
        cv_init_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(ctx->cvset(), tmp.as<uint8_t>(), (uint32_t)(2 * n), P.d_cov_lo.as<uint32_t>(), P.d_cov_hi.as<uint32_t>(),
                                                                (uint8_t)p, S.cv_allele.as<uint32_t>(), ctx->use_root ? S.cv_root.as<uint8_t>() : nullptr);
        GE_TRY(ctx->check_launch("cv_init"));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->release(tmp_local);
    }
    if (ctx->segs()) GE_TRY(seg_init_gen0(ctx, p, n));
    // pedigree, sex, sibling-common effect
    std::vector<uint64_t> ids(n * 7);
    for (uint64_t i = 0; i < n; i++) for (int k = 0; k < 7; k++) ids[i * 7 + k] = i;  // :3037-3043
    CUDA_TRY(cudaMemcpyAsync(S.ids.p, ids.data(), ids.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
        sex_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->rng, p, 0, 0, n, S.sex.as<uint8_t>());
        GE_TRY(ctx->check_launch("sex"));
    } else {
        if (!d0 || !d0[p].sex) return fail(GE_ERR_INVALID, "replay mode needs draws0[pop].sex");
        CUDA_TRY(cudaMemcpyAsync(S.sex.p, d0[p].sex, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(cudaMemsetAsync(S.C.p, 0, (size_t)n * nf * 8, ctx->stream));
    for (int f = 0; f < nf; f++) {
        if (P.scheme[f].vc > 0) {
            double *dst = S.C.as<double>() + (uint64_t)f * n;
            if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
                normal_scaled_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->rng, P_COMMON, p, 0, f, 0, n, std::sqrt(P.scheme[f].vc), dst);
                GE_TRY(ctx->check_launch("common0"));
            } else if (d0 && d0[p].common) CUDA_TRY(cudaMemcpyAsync(dst, d0[p].common + (uint64_t)f * n, n * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GE_OK;
}

int ge_init_generation0(ge_ctx *ctx, const ge_draws *d0) {  // ras_init_generation0 :529-679
    CHECK_CTX(ctx);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int nf = ctx->cfg.n_phen, np = ctx->cfg.n_pop;
    GE_TRY(build_genome(ctx));
    for (int p = 0; p < np; p++) GE_TRY(build_maps(ctx, ctx->pop[p]));
    if (ctx->bits()) {
        for (int p = 1; p < np; p++)
            for (int c = 0; c < ctx->cfg.n_chr; c++)
                if (ctx->pop[p].rmap_bp[c].front() != ctx->pop[0].rmap_bp[c].front() || ctx->pop[p].rmap_bp[c].back() != ctx->pop[0].rmap_bp[c].back())
                    return fail(GE_ERR_UNSUPPORTED, "bit-packed representation needs the same genetic-map range in every population");
    }
    GE_TRY(build_cvset(ctx));
    for (int p = 0; p < np; p++) {
        PopDev &P = ctx->pop[p];
        GE_TRY(init_pop_gen0(ctx, p, d0));
        GE_TRY(ge_compute_AD(ctx, p, 0));
        GenState &S = P.st[P.cur];
        // ras_fill_Pop_info_prev_gen_for_gen0_prev :3240-3251
        size_t bytes = (size_t)ctx->cfg.capacity * nf * 8;
        GE_TRY(ctx->ensure(P.prev_P, bytes)); GE_TRY(ctx->ensure(P.prev_F, bytes));
        CUDA_TRY(cudaMemsetAsync(P.prev_P.p, 0, bytes, ctx->stream)); CUDA_TRY(cudaMemsetAsync(P.prev_F.p, 0, bytes, ctx->stream));
        P.prev_n = S.n;
        for (int f = 0; f < nf; f++) {  // :555-566
            GE_TRY(ctx->h_var(S.A.as<double>() + (uint64_t)f * S.n, S.n, &P.var_a0[f]));
            GE_TRY(ctx->h_var(S.D.as<double>() + (uint64_t)f * S.n, S.n, &P.var_d0[f]));
            bool rp = ctx->cfg.rng_mode == GE_RNG_REPLAY && d0;
            GE_TRY(scale_AD_compute_GEF_impl(ctx, p, 0, f, (rp && d0[p].e_raw) ? d0[p].e_raw + (uint64_t)f * S.n : nullptr,
                                             (rp && d0[p].parental0) ? d0[p].parental0 + (uint64_t)f * S.n : nullptr));
        }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    for (int f = 0; f < nf; f++) GE_TRY(ge_environmental_effects_specific_to_each_population(ctx, f));
    for (int p = 0; p < np; p++) GE_TRY(ge_compute_mating_value_selection_value(ctx, p, 0, nullptr));
    for (int p = 0; p < np; p++) GE_TRY(ge_save_human_info_to_Pop_info_prev_gen(ctx, p));
    for (int p = 0; p < np; p++) {  // :645-654 beta adjustment
        PopDev &P = ctx->pop[p];
        GenState &S = P.st[P.cur];
        for (int f = 0; f < nf; f++) {
            double vP, vF;
            GE_TRY(ctx->h_var(S.P.as<double>() + (uint64_t)f * S.n, S.n, &vP));
            GE_TRY(ctx->h_var(S.F.as<double>() + (uint64_t)f * S.n, S.n, &vF));
            if (ctx->cfg.vt_type == 1) P.scheme[f].beta = std::sqrt(P.scheme[f].vf / (2 * vP));
            else if (ctx->cfg.vt_type == 2) { if (vF > 0) P.scheme[f].beta = std::sqrt(P.scheme[f].vf / (2 * vF)); }
        }
        std::vector<std::vector<uint8_t>>().swap(P.panel);  // the host copy of the panel is no longer needed
        P.panel.resize(ctx->cfg.n_chr);
    }
    GE_TRY(ctx->check_flags("generation 0"));
    return GE_OK;
}

// ---------------- mating ----------------

int ge_set_couples(ge_ctx *ctx, int pop, const uint64_t *m, const uint64_t *f, const uint8_t *inb, const int32_t *no, uint64_t n) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    std::vector<uint32_t> mm(n), ff(n);
    for (uint64_t k = 0; k < n; k++) { mm[k] = (uint32_t)m[k]; ff[k] = (uint32_t)f[k]; }
    GE_TRY(ctx->upload(P.c_male, mm)); GE_TRY(ctx->upload(P.c_female, ff));
    GE_TRY(ctx->upload(P.c_inbreed, std::vector<uint8_t>(inb, inb + n)));
    GE_TRY(ctx->upload(P.c_noff, std::vector<int32_t>(no, no + n)));
    P.n_couples = n;
    return GE_OK;
}
int ge_get_couples_count(ge_ctx *ctx, int pop, uint64_t *n) { CHECK_POP(ctx, pop); *n = ctx->pop[pop].n_couples; return GE_OK; }
int ge_get_couples(ge_ctx *ctx, int pop, uint64_t *m, uint64_t *f, uint8_t *inb, int32_t *no) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    uint64_t n = P.n_couples;
    std::vector<uint32_t> mm(n), ff(n);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (n) {
        CUDA_TRY(cudaMemcpy(mm.data(), P.c_male.p, n * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ff.data(), P.c_female.p, n * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(inb, P.c_inbreed.p, n, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(no, P.c_noff.p, n * 4, cudaMemcpyDeviceToHost));
    }
    for (uint64_t k = 0; k < n; k++) { m[k] = mm[k]; f[k] = ff[k]; }
    return GE_OK;
}

int ge_mate(ge_ctx *ctx, int pop, int gen, const ge_gen_params *gp) {  // random_mate :2090-2157 / assort_mate :2167-2360
    CHECK_POP(ctx, pop);
    if (!gp) return fail(GE_ERR_INVALID, "null params");
    if (ctx->cfg.rng_mode != GE_RNG_PHILOX) return fail(GE_ERR_INVALID, "replay mode: supply couples with ge_set_couples or offspring draws");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    ge_ctx::PhaseTimer timer(ctx, GE_PHASE_MATE);
    return mate_philox(ctx, pop, gen, *gp);
}

// ---------------- reproduce ----------------

static int upload_u64_as_u32(ge_ctx *ctx, Buf &b, const uint64_t *src, uint64_t n, std::vector<uint32_t> &tmp, const char *what) {
    tmp.resize(n);
    for (uint64_t k = 0; k < n; k++) {
        if (src[k] > 0xFFFFFFFFull) return fail(GE_ERR_UNSUPPORTED, std::string(what) + " does not fit 32 bits");
        tmp[k] = (uint32_t)src[k];
    }
    return ctx->upload(b, tmp);
}

int ge_reproduce(ge_ctx *ctx, int pop, int gen, const ge_draws *dr) {  // reproduce :2394-2493
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    int C = ctx->cfg.n_chr, nf = ctx->cfg.n_phen;
    GenState &par = P.st[P.cur], &off = P.st[P.cur ^ 1];
    if (par.n == 0) return fail(GE_ERR_INVALID, "ge_reproduce before ge_init_generation0");
    uint64_t n_off = 0;
    std::vector<uint32_t> tmp;
    cudaStream_t st = ctx->stream;
    if (dr) {  // validate the caller's draws before any state changes
        n_off = dr->n_offspring;
        if (n_off == 0) return fail(GE_ERR_NO_MATES, "no offspring");
        if (n_off > ctx->cfg.capacity) return fail(GE_ERR_CAPACITY, "offspring exceed capacity");
        if (!dr->father || !dr->mother || !dr->sex || !dr->xo_off || !dr->start_hap) return fail(GE_ERR_INVALID, "incomplete draws");
        for (uint64_t i = 0; i < n_off; i++) if (dr->father[i] >= par.n || dr->mother[i] >= par.n) return fail(GE_ERR_INVALID, "parent index out of range");
        const uint64_t ns = n_off * C * 2;
        for (uint64_t k = 0; k < ns; k++) if (dr->xo_off[k + 1] < dr->xo_off[k]) return fail(GE_ERR_INVALID, "xo_off must be non-decreasing");
        if (dr->xo_off[ns] && !dr->xo_bp) return fail(GE_ERR_INVALID, "incomplete draws: xo_bp");
    } else {
        if (ctx->cfg.rng_mode != GE_RNG_PHILOX) return fail(GE_ERR_INVALID, "replay mode needs draws");
        if (P.n_couples == 0) return fail(GE_ERR_INVALID, "no couples: call ge_mate or ge_set_couples first");
    }
    // the other draw set; the bulk stream may still read it for the generation before last
    P.dcur ^= 1;
    DrawSet &D = P.draws();
    if (D.bulk_pending) { CUDA_TRY(cudaStreamWaitEvent(st, D.bulk_done, 0)); D.bulk_pending = false; }
    if (dr) {
        GE_TRY(upload_u64_as_u32(ctx, D.father, dr->father, n_off, tmp, "father"));
        GE_TRY(upload_u64_as_u32(ctx, D.mother, dr->mother, n_off, tmp, "mother"));
        uint64_t n_slots = n_off * C * 2;
        P.n_xo = dr->xo_off[n_slots];
        GE_TRY(ctx->upload(D.xo_off, std::vector<uint64_t>(dr->xo_off, dr->xo_off + n_slots + 1)));
        GE_TRY(upload_u64_as_u32(ctx, D.xo_bp, dr->xo_bp, P.n_xo, tmp, "crossover position"));
        GE_TRY(ctx->upload(D.start_hap, std::vector<uint8_t>(dr->start_hap, dr->start_hap + n_slots)));
        CUDA_TRY(cudaMemcpyAsync(off.sex.p, dr->sex, n_off, cudaMemcpyHostToDevice, st));
        if (dr->mut_off) {
            P.n_mut = dr->mut_off[n_off * C];
            GE_TRY(ctx->upload(P.mut_off, std::vector<uint64_t>(dr->mut_off, dr->mut_off + n_off * C + 1)));
            GE_TRY(upload_u64_as_u32(ctx, P.mut_bp, dr->mut_bp, P.n_mut, tmp, "mutation position"));
            GE_TRY(ctx->upload(P.mut_gam, std::vector<uint8_t>(dr->mut_gam, dr->mut_gam + P.n_mut)));
        } else P.n_mut = 0;
        CUDA_TRY(cudaMemsetAsync(off.C.p, 0, (size_t)n_off * nf * 8, st));
        if (dr->common) CUDA_TRY(cudaMemcpyAsync(off.C.p, dr->common, (size_t)n_off * nf * 8, cudaMemcpyHostToDevice, st));
        P.have_couple_of = false;
    } else {
        ge_ctx::PhaseTimer timer(ctx, GE_PHASE_SAMPLE);
        // offspring offsets = exclusive scan of the family sizes of the couples that may marry (:2402-2406)
        GE_TRY(ctx->ensure(P.cnt32, (size_t)std::max<uint64_t>(P.n_couples, ctx->cfg.capacity * C * 2 + 1) * 4));
        GE_TRY(ctx->ensure(P.mate.fam_off, (P.n_couples + 1) * 8));
        family_size_kernel<<<nblk(P.n_couples, 256), 256, 0, st>>>(P.n_couples, P.c_inbreed.as<uint8_t>(), P.c_noff.as<int32_t>(), P.cnt32.as<uint32_t>());
        GE_TRY(ctx->check_launch("family_size"));
        GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), P.n_couples, P.mate.fam_off.as<uint64_t>(), &n_off));
        if (n_off == 0) return fail(GE_ERR_NO_MATES, "no offspring");
        if (n_off > ctx->cfg.capacity) return fail(GE_ERR_CAPACITY, "offspring (" + std::to_string(n_off) + ") exceed capacity");
        GE_TRY(ctx->ensure(D.father, n_off * 4)); GE_TRY(ctx->ensure(D.mother, n_off * 4)); GE_TRY(ctx->ensure(D.couple_of, n_off * 4));
        expand_couples_kernel<<<nblk(P.n_couples, 256), 256, 0, st>>>(P.n_couples, P.mate.fam_off.as<uint64_t>(), P.c_male.as<uint32_t>(), P.c_female.as<uint32_t>(),
                                                                      D.father.as<uint32_t>(), D.mother.as<uint32_t>(), D.couple_of.as<uint32_t>());
        GE_TRY(ctx->check_launch("expand_couples"));
        P.have_couple_of = true;
        // crossovers: count, scan, fill
        uint64_t n_slots = n_off * C * 2;
        GE_TRY(ctx->ensure(D.xo_off, (n_slots + 1) * 8)); GE_TRY(ctx->ensure(D.start_hap, n_slots));
        GE_TRY(ctx->ensure(ctx->xo_stash, n_slots * XO_STASH * 4));
        sample_xo_kernel<false><<<ctx->ctrl_grid(n_slots, 128), 128, 0, st>>>(ctx->rng, ctx->rmap(P), C, pop, gen, 0, n_slots, P.cnt32.as<uint32_t>(), nullptr, nullptr, D.start_hap.as<uint8_t>(),
                                                                    ctx->xo_stash.as<uint32_t>());
        GE_TRY(ctx->check_launch("sample_xo<count>"));
        GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), n_slots, D.xo_off.as<uint64_t>(), &P.n_xo));
        GE_TRY(ctx->ensure(D.xo_bp, std::max<uint64_t>(P.n_xo, 1) * 4));
        if (ctx->bits()) GE_TRY(ctx->ensure(D.flips, std::max<uint64_t>(P.n_xo, 1) * 4));
        xo_place_kernel<<<ctx->ctrl_grid(n_slots, 128), 128, 0, st>>>(ctx->rng, ctx->rmap(P), ctx->genome(), C, pop, gen, n_slots, D.xo_off.as<uint64_t>(), ctx->xo_stash.as<uint32_t>(),
                                                            D.xo_bp.as<uint32_t>(), ctx->bits() ? D.flips.as<uint32_t>() : nullptr);
        GE_TRY(ctx->check_launch("xo_place"));
        if (P.has_mut) {
            uint64_t n_items = n_off * C;
            GE_TRY(ctx->ensure(P.mut_off, (n_items + 1) * 8));
            sample_mut_kernel<false><<<nblk(n_items, 128), 128, 0, st>>>(ctx->rng, ctx->mmap(P), C, pop, gen, 0, n_items, P.cnt32.as<uint32_t>(), nullptr, nullptr, nullptr);
            GE_TRY(ctx->check_launch("sample_mut<count>"));
            GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), n_items, P.mut_off.as<uint64_t>(), &P.n_mut));
            GE_TRY(ctx->ensure(P.mut_bp, std::max<uint64_t>(P.n_mut, 1) * 4)); GE_TRY(ctx->ensure(P.mut_gam, std::max<uint64_t>(P.n_mut, 1)));
            sample_mut_kernel<true><<<nblk(n_items, 128), 128, 0, st>>>(ctx->rng, ctx->mmap(P), C, pop, gen, 0, n_items, nullptr, P.mut_off.as<uint64_t>(), P.mut_bp.as<uint32_t>(), P.mut_gam.as<uint8_t>());
            GE_TRY(ctx->check_launch("sample_mut<fill>"));
        } else P.n_mut = 0;
        sex_kernel<<<nblk(n_off, 256), 256, 0, st>>>(ctx->rng, pop, gen, 0, n_off, off.sex.as<uint8_t>());
        GE_TRY(ctx->check_launch("sex"));
        CUDA_TRY(cudaMemsetAsync(off.C.p, 0, (size_t)n_off * nf * 8, st));
        for (int f = 0; f < nf; f++)
            if (P.scheme[f].vc > 0) {
                common_from_couples_kernel<<<nblk(n_off, 256), 256, 0, st>>>(ctx->rng, pop, gen, f, std::sqrt(P.scheme[f].vc), 0, n_off, D.couple_of.as<uint32_t>(),
                                                                            off.C.as<double>() + (uint64_t)f * n_off);
                GE_TRY(ctx->check_launch("common"));
            }
    }
    P.n_off = n_off;
    uint64_t n_slots = n_off * C * 2;
    // ---- bit-packed propagation: the HBM-bound bulk of the generation, on the bulk stream.  Nothing later on the
    // control stream needs the rows (genetic values come from the causal-variant planes), so mating, sampling
    // and phenotypes of the NEXT generation overlap with this copy.
    bool bulk_launched = false;
    if (ctx->bits()) {
        if (dr) {  // replayed crossovers: positions -> locus indices (the Philox path did it in xo_place_kernel)
            GE_TRY(ctx->ensure(D.flips, std::max<uint64_t>(P.n_xo, 1) * 4));
            xo_to_flips_kernel<<<nblk(n_slots, 128), 128, 0, st>>>(ctx->genome(), n_slots, D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.flips.as<uint32_t>());
            GE_TRY(ctx->check_launch("xo_to_flips"));
        }
        cudaStream_t bulk = ctx->serial ? st : ctx->bulk;
        CUDA_TRY(cudaEventRecord(ctx->ev_ready, st));
        CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_ready, 0));
        ge_ctx::EvPair evp{nullptr, nullptr, GE_KERNEL_PROPAGATE_BITS, 0};
        if (ctx->profiling) { evp.a = ctx->get_event(); evp.b = ctx->get_event(); CUDA_TRY(cudaEventRecord(evp.a, bulk)); }
        unsigned grid = (unsigned)std::min<uint64_t>(n_off, 1u << 20);  // one short-lived CTA per offspring: control-stream kernels get SM slots quickly
        if (ctx->use_tma) {
            size_t sm = prop_tma_smem_bytes(C);
            if (!ctx->tma_attr_set) { CUDA_TRY(cudaFuncSetAttribute(propagate_bits_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); ctx->tma_attr_set = true; }
            propagate_bits_tma_kernel<<<grid, TMA_WARPS * 32 + TMA_MERGE_THREADS, sm, bulk>>>(ctx->genome(), ctx->tiles(), par.hap.as<uint32_t>(), par.rowmap, off.hap.as<uint32_t>(), D.father.as<uint32_t>(),
                                                                         D.mother.as<uint32_t>(), D.xo_off.as<uint64_t>(), D.flips.as<uint32_t>(), D.start_hap.as<uint8_t>(), 0, (uint32_t)n_off);
        } else if (ctx->prop_depth == 8)
        propagate_bits_kernel<8><<<grid, PROP_THREADS, prop_smem_bytes(C), bulk>>>(ctx->genome(), ctx->tiles(), par.hap.as<uint32_t>(), par.rowmap, off.hap.as<uint32_t>(), D.father.as<uint32_t>(),
                                                                    D.mother.as<uint32_t>(), D.xo_off.as<uint64_t>(), D.flips.as<uint32_t>(), D.start_hap.as<uint8_t>(), 0, (uint32_t)n_off);
        else
        propagate_bits_kernel<4><<<grid, PROP_THREADS, prop_smem_bytes(C), bulk>>>(ctx->genome(), ctx->tiles(), par.hap.as<uint32_t>(), par.rowmap, off.hap.as<uint32_t>(), D.father.as<uint32_t>(),
                                                                    D.mother.as<uint32_t>(), D.xo_off.as<uint64_t>(), D.flips.as<uint32_t>(), D.start_hap.as<uint8_t>(), 0, (uint32_t)n_off);
        GE_TRY(ctx->check_launch("propagate_bits"));
        if (ctx->profiling) {
            CUDA_TRY(cudaEventRecord(evp.b, bulk));
            uint64_t M = 0;
            for (uint32_t v : ctx->chr_nloci) M += v;
            evp.bytes = n_off * M / 2;  // 0.5 byte per individual-locus
            ctx->ev_pending.push_back(evp);
        }
        CUDA_TRY(cudaEventRecord(D.bulk_done, bulk));
        D.bulk_pending = true;
        // thin control kernels only pay off while the bulk copy is longer than the control chain (~0.75 ms at 100k individuals)
        ctx->note_bulk((double)n_off * ctx->W * 16.0);
        bulk_launched = true;
    }
    // ---- causal-variant planes
    if (ctx->n_cv_tot) {  // in every representation: crossover parity at the CV positions, never a rescan of the segment lists
        uint64_t tot = n_off * 2 * ctx->Wcv;
        cv_propagate_bits_kernel<<<ctx->ctrl_grid(tot, 256), 256, 0, st>>>(ctx->cvset(), par.cv_allele.as<uint32_t>(), off.cv_allele.as<uint32_t>(), D.father.as<uint32_t>(),
                                                                 D.mother.as<uint32_t>(), D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.start_hap.as<uint8_t>(), 0, n_off);
        GE_TRY(ctx->check_launch("cv_propagate_bits"));
        if (ctx->use_root) {
            uint64_t tr = n_off * 2 * ctx->n_cv_tot;
            cv_root_propagate_kernel<<<nblk(tr, 256), 256, 0, st>>>(ctx->cvset(), par.cv_root.as<uint8_t>(), off.cv_root.as<uint8_t>(), D.father.as<uint32_t>(), D.mother.as<uint32_t>(),
                                                                    D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.start_hap.as<uint8_t>(), 0, n_off);
            GE_TRY(ctx->check_launch("cv_root_propagate"));
        }
    }
    // ---- founder segments
    if (ctx->segs()) GE_TRY(seg_recombine(ctx, pop, n_off));
    // ---- mutation lists (inherit + this generation's hits)
    if (P.has_mut || par.has_hm) {
        MutArgs a;
        a.n_chr = C; a.off_first = 0; a.n_off = n_off; a.father = D.father.as<uint32_t>(); a.mother = D.mother.as<uint32_t>();
        a.xo_off = D.xo_off.as<uint64_t>(); a.xo_bp = D.xo_bp.as<uint32_t>(); a.start_hap = D.start_hap.as<uint8_t>();
        a.par_hm_off = par.has_hm ? par.hm_off.as<uint64_t>() : nullptr; a.par_hm_bp = par.hm_bp.as<uint32_t>();
        bool hits = P.has_mut && (dr ? dr->mut_off != nullptr : true);
        a.mut_off = hits ? P.mut_off.as<uint64_t>() : nullptr; a.mut_bp = P.mut_bp.as<uint32_t>(); a.mut_gam = P.mut_gam.as<uint8_t>();
        a.cov_lo = P.d_cov_lo.as<uint32_t>(); a.cov_hi = P.d_cov_hi.as<uint32_t>();
        GE_TRY(ctx->ensure(P.cnt32, (n_slots + 1) * 4));
        GE_TRY(ctx->ensure(off.hm_off, (n_slots + 1) * 8));
        mutation_lists_kernel<false><<<nblk(n_slots, 128), 128, 0, st>>>(a, ctx->genome(), ctx->cvset(), P.cnt32.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr);
        GE_TRY(ctx->check_launch("mutation_lists<count>"));
        GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), n_slots, off.hm_off.as<uint64_t>(), &off.n_hm));
        GE_TRY(ctx->ensure(off.hm_bp, std::max<uint64_t>(off.n_hm, 1) * 4));
        if (bulk_launched) GE_TRY(ctx->join_bulk());  // the fill pass toggles bits of the freshly propagated rows
        mutation_lists_kernel<true><<<nblk(n_slots, 128), 128, 0, st>>>(a, ctx->genome(), ctx->cvset(), nullptr, off.hm_off.as<uint64_t>(), off.hm_bp.as<uint32_t>(),
                                                                        ctx->bits() ? off.hap.as<uint32_t>() : nullptr, ctx->n_cv_tot ? off.cv_allele.as<uint32_t>() : nullptr);
        GE_TRY(ctx->check_launch("mutation_lists<fill>"));
        off.has_hm = true;
    } else off.has_hm = false;
    // ---- pedigree
    pedigree_kernel<<<nblk(n_off, 256), 256, 0, st>>>(0, n_off, D.father.as<uint32_t>(), D.mother.as<uint32_t>(), par.ids.as<uint64_t>(), off.ids.as<uint64_t>());
    GE_TRY(ctx->check_launch("pedigree"));
    off.n = n_off;
    off.rowmap = nullptr;   // a fresh generation is written in identity order
    P.cur ^= 1;
    if (dr) CUDA_TRY(cudaStreamSynchronize(st));  // caller buffers were read asynchronously
    return GE_OK;
}

int ge_set_migration_sample(ge_ctx *ctx, int src, const uint64_t *pos, uint64_t n) {
    CHECK_POP(ctx, src);
    ctx->mig_sample.resize(ctx->cfg.n_pop);
    ctx->mig_sample[src].assign(pos, pos + n);
    return GE_OK;
}
int ge_do_migration(ge_ctx *ctx, int gen, const double *row) {  // ras_do_migration :877-989
    CHECK_CTX(ctx);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(seg_finish_all(ctx));
    return migrate(ctx, gen, row);
}

int ge_step_generation(ge_ctx *ctx, int gen, const ge_gen_params *gp, const double *mig, const ge_draws *dr) {  // sim_next_generation :1890-2082
    CHECK_CTX(ctx);
    if (!gp) return fail(GE_ERR_INVALID, "null params");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int nf = ctx->cfg.n_phen, np = ctx->cfg.n_pop;
    for (int p = 0; p < np; p++) {
        if (!dr) GE_TRY(ge_mate(ctx, p, gen, &gp[p]));
        GE_TRY(ge_reproduce(ctx, p, gen, dr ? &dr[p] : nullptr));
        GE_TRY(ge_compute_AD(ctx, p, gen));
        uint64_t n = ctx->pop[p].st[ctx->pop[p].cur].n;
        {
            ge_ctx::PhaseTimer timer(ctx, GE_PHASE_PHENOTYPE);
            for (int f = 0; f < nf; f++)
                GE_TRY(scale_AD_compute_GEF_impl(ctx, p, gen, f, (dr && dr[p].e_raw) ? dr[p].e_raw + (uint64_t)f * n : nullptr, nullptr));
        }
    }
    for (int f = 0; f < nf; f++) GE_TRY(ge_environmental_effects_specific_to_each_population(ctx, f));
    for (int p = 0; p < np; p++) GE_TRY(ge_compute_mating_value_selection_value(ctx, p, gen, &gp[p]));
    if (np > 1 && mig) GE_TRY(ge_do_migration(ctx, gen, mig));
    for (int p = 0; p < np; p++) GE_TRY(ge_save_human_info_to_Pop_info_prev_gen(ctx, p));
    GE_TRY(ctx->check_flags("generation"));  // also the one host sync of the generation
    return GE_OK;
}

// ---------------- results ----------------

int ge_get_population_size(ge_ctx *ctx, int pop, uint64_t *n) { CHECK_POP(ctx, pop); *n = ctx->pop[pop].st[ctx->pop[pop].cur].n; return GE_OK; }

int ge_download_individuals(ge_ctx *ctx, int pop, ge_indiv_soa *o) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    uint64_t n = S.n; int nf = ctx->cfg.n_phen;
    cudaStream_t st = ctx->stream;
    auto cp = [&](void *dst, const Buf &src, size_t bytes) -> cudaError_t { return dst ? cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
    CUDA_TRY(cp(o->ids, S.ids, n * 56)); CUDA_TRY(cp(o->sex, S.sex, n));
    CUDA_TRY(cp(o->A, S.A, n * nf * 8)); CUDA_TRY(cp(o->D, S.D, n * nf * 8)); CUDA_TRY(cp(o->G, S.G, n * nf * 8)); CUDA_TRY(cp(o->C, S.C, n * nf * 8));
    CUDA_TRY(cp(o->E, S.E, n * nf * 8)); CUDA_TRY(cp(o->F, S.F, n * nf * 8)); CUDA_TRY(cp(o->P, S.P, n * nf * 8));
    CUDA_TRY(cp(o->mv, S.mv, n * 8)); CUDA_TRY(cp(o->sv, S.sv, n * 8)); CUDA_TRY(cp(o->svf, S.svf, n * 8));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GE_OK;
}

int ge_get_moments(ge_ctx *ctx, int pop, int f, ge_moments *m) {
    CHECK_POP(ctx, pop); CHECK_PHEN(ctx, f);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    uint64_t o = (uint64_t)f * S.n;
    GE_TRY(ctx->h_var(S.A.as<double>() + o, S.n, &m->var_A)); GE_TRY(ctx->h_var(S.D.as<double>() + o, S.n, &m->var_D));
    GE_TRY(ctx->h_var(S.G.as<double>() + o, S.n, &m->var_G)); GE_TRY(ctx->h_var(S.C.as<double>() + o, S.n, &m->var_C));
    GE_TRY(ctx->h_var(S.E.as<double>() + o, S.n, &m->var_E)); GE_TRY(ctx->h_var(S.F.as<double>() + o, S.n, &m->var_F));
    GE_TRY(ctx->h_var(S.P.as<double>() + o, S.n, &m->var_P));
    m->h2 = m->var_A / m->var_P;
    return GE_OK;
}
int ge_get_mv_sv_var(ge_ctx *ctx, int pop, double *vm, double *vs) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    GE_TRY(ctx->h_var(S.mv.as<double>(), S.n, vm));
    return ctx->h_var(S.sv.as<double>(), S.n, vs);
}
int ge_get_gen0_constants(ge_ctx *ctx, int pop, int f, double *va0, double *vd0, double *beta, double *m0, double *v0) {
    CHECK_POP(ctx, pop); CHECK_PHEN(ctx, f);
    PopDev &P = ctx->pop[pop];
    *va0 = P.var_a0[f]; *vd0 = P.var_d0[f]; *beta = P.scheme[f].beta; *m0 = P.sv_mean0; *v0 = P.sv_var0;
    return GE_OK;
}

int ge_download_haplotypes(ge_ctx *ctx, int pop, int c, uint8_t *al) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    uint32_t nl = ctx->chr_nloci[c];
    uint64_t tot = (uint64_t)2 * S.n * nl;
    if (tot == 0) return GE_OK;
    Buf tmp;
    GE_TRY(ctx->ensure_exact(tmp, tot));
    if (ctx->bits()) {
        GE_TRY(ctx->join_bulk());
        unpack_rows_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(S.hap.as<uint32_t>(), S.rowmap, ctx->W, ctx->chr_word_off[c], (uint32_t)(2 * S.n), nl,
                                                                    tmp.as<uint8_t>());
        GE_TRY(ctx->check_launch("unpack_rows"));
    } else GE_TRY(seg_materialise(ctx, pop, c, tmp.as<uint8_t>()));
    CUDA_TRY(cudaMemcpyAsync(al, tmp.p, tot, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tmp);
    return GE_OK;
}
int ge_download_haplotypes_packed(ge_ctx *ctx, int pop, int c, uint32_t *words) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    if (!ctx->bits()) return fail(GE_ERR_UNSUPPORTED, "packed download needs GE_REP_BITS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    uint32_t nw = (ctx->chr_nloci[c] + 31) / 32;
    uint64_t tot = (uint64_t)2 * S.n * nw;
    if (tot == 0) return GE_OK;
    Buf tmp;
    GE_TRY(ctx->ensure_exact(tmp, tot * 4));
    GE_TRY(ctx->join_bulk());
    gather_packed_chr_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(S.hap.as<uint32_t>(), S.rowmap, ctx->W, ctx->chr_word_off[c], (uint32_t)(2 * S.n), nw,
                                                                      tmp.as<uint32_t>());
    GE_TRY(ctx->check_launch("gather_packed"));
    CUDA_TRY(cudaMemcpyAsync(words, tmp.p, tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tmp);
    return GE_OK;
}

int ge_get_segment_count(ge_ctx *ctx, int pop, int c, uint64_t *ns, uint64_t *nm) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_count(ctx, pop, c, ns, nm);
}
int ge_download_segments(ge_ctx *ctx, int pop, int c, uint64_t *off, uint64_t *seg, uint64_t *moff, uint64_t *mbp) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_download(ctx, pop, c, off, seg, moff, mbp);
}

int ge_download_cv_alleles(ge_ctx *ctx, int pop, int f, int c, uint8_t *out) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c); CHECK_PHEN(ctx, f);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    uint32_t b0 = ctx->cv_block_off[(size_t)f * ctx->cfg.n_chr + c], b1 = ctx->cv_block_off[(size_t)f * ctx->cfg.n_chr + c + 1];
    if (b1 == b0 || S.n == 0) return GE_OK;
    uint64_t tot = 2 * S.n * (b1 - b0);
    Buf tmp;
    GE_TRY(ctx->ensure_exact(tmp, tot));
    cv_unpack_block_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(S.cv_allele.as<uint32_t>(), ctx->Wcv, ctx->cv_word_off[(size_t)f * ctx->cfg.n_chr + c], b1 - b0, 2 * S.n, tmp.as<uint8_t>());
    GE_TRY(ctx->check_launch("cv_unpack_block"));
    CUDA_TRY(cudaMemcpyAsync(out, tmp.p, tot, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tmp);
    return GE_OK;
}

int ge_compact_segments(ge_ctx *ctx, int pop, uint64_t *n_before, uint64_t *n_after) {
    CHECK_POP(ctx, pop);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_compact_segments needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_compact(ctx, pop, n_before, n_after);
}

int ge_recompute_cv_from_segments(ge_ctx *ctx, int pop) {  // ras_find_cv :2752-2815 literally: scan the parts of every haplotype
    CHECK_POP(ctx, pop);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_recompute_cv_from_segments needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_find_cv(ctx, pop);
}

int ge_ibd_sharing(ge_ctx *ctx, int pop, int chr, const uint64_t *ind_a, const uint64_t *ind_b, uint64_t n_pairs, uint64_t min_bp, uint64_t *shared_bp, uint32_t *n_runs) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_ibd_sharing needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_ibd(ctx, pop, chr, ind_a, ind_b, n_pairs, min_bp, shared_bp, n_runs);
}

int ge_get_segment_format(ge_ctx *ctx, int *bytes) {
    CHECK_CTX(ctx);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_get_segment_format needs GE_REP_SEGMENTS");
    *bytes = (int)ctx->seg_esz();
    return GE_OK;
}

int ge_get_draw_counts(ge_ctx *ctx, int pop, uint64_t *no, uint64_t *nx, uint64_t *nm) {
    CHECK_POP(ctx, pop);
    PopDev &P = ctx->pop[pop];
    *no = P.n_off; *nx = P.n_xo; *nm = P.n_mut;
    return GE_OK;
}
int ge_download_draws(ge_ctx *ctx, int pop, uint64_t *fa, uint64_t *mo, uint8_t *sex, uint64_t *xo_off, uint64_t *xo_bp, uint8_t *start,
                      uint64_t *mut_off, uint64_t *mut_bp, uint8_t *mut_gam) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    DrawSet &D = P.draws();
    int C = ctx->cfg.n_chr;
    uint64_t n = P.n_off, ns = n * C * 2;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> t;
    auto get32 = [&](const Buf &b, uint64_t cnt, uint64_t *dst) -> int {
        if (!dst || cnt == 0) return GE_OK;
        t.resize(cnt);
        CUDA_TRY(cudaMemcpy(t.data(), b.p, cnt * 4, cudaMemcpyDeviceToHost));
        for (uint64_t k = 0; k < cnt; k++) dst[k] = t[k];
        return GE_OK;
    };
    GE_TRY(get32(D.father, n, fa)); GE_TRY(get32(D.mother, n, mo));
    if (sex && n) CUDA_TRY(cudaMemcpy(sex, S.sex.p, n, cudaMemcpyDeviceToHost));
    if (xo_off) CUDA_TRY(cudaMemcpy(xo_off, D.xo_off.p, (ns + 1) * 8, cudaMemcpyDeviceToHost));
    GE_TRY(get32(D.xo_bp, P.n_xo, xo_bp));
    if (start && ns) CUDA_TRY(cudaMemcpy(start, D.start_hap.p, ns, cudaMemcpyDeviceToHost));
    if (mut_off) {
        if (P.has_mut && P.mut_off.p) CUDA_TRY(cudaMemcpy(mut_off, P.mut_off.p, (n * C + 1) * 8, cudaMemcpyDeviceToHost));
        else std::memset(mut_off, 0, (n * C + 1) * 8);
    }
    GE_TRY(get32(P.mut_bp, P.n_mut, mut_bp));
    if (mut_gam && P.n_mut) CUDA_TRY(cudaMemcpy(mut_gam, P.mut_gam.p, P.n_mut, cudaMemcpyDeviceToHost));
    return GE_OK;
}

// ---------------- measurement hooks ----------------
int ge_set_profiling(ge_ctx *ctx, int enabled) { CHECK_CTX(ctx); ctx->profiling = enabled != 0; return GE_OK; }
int ge_get_kernel_time(ge_ctx *ctx, int k, double *ms, uint64_t *launches, uint64_t *bytes) {
    CHECK_CTX(ctx);
    if (k < 0 || k >= GE_KERNEL_COUNT) return fail(GE_ERR_INVALID, "bad kernel id");
    GE_TRY(seg_finish_all(ctx));
    ctx->resolve_events();
    *ms = ctx->kstat[k].ms; *launches = ctx->kstat[k].launches; *bytes = ctx->kstat[k].bytes;
    return GE_OK;
}
int ge_reset_kernel_times(ge_ctx *ctx) { CHECK_CTX(ctx); GE_TRY(seg_finish_all(ctx)); ctx->resolve_events(); for (auto &k : ctx->kstat) k = KernelStat(); ctx->launches = 0; return GE_OK; }
int ge_get_launch_count(ge_ctx *ctx, uint64_t *n) { CHECK_CTX(ctx); *n = ctx->launches; return GE_OK; }
int ge_synchronize(ge_ctx *ctx) {
    CHECK_CTX(ctx);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(seg_finish_all(ctx));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->bulk));
    return GE_OK;
}
int ge_device_memory_bytes(ge_ctx *ctx, uint64_t *b) { CHECK_CTX(ctx); *b = ctx->mem_peak; return GE_OK; }
// both streams are joined on either side, so the region covers the bulk propagation of every step queued in between
int ge_timer_start(ge_ctx *ctx) { CHECK_CTX(ctx); CUDA_TRY(cudaSetDevice(ctx->cfg.device)); GE_TRY(seg_finish_all(ctx)); GE_TRY(ctx->join_bulk()); CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream)); return GE_OK; }
int ge_timer_stop(ge_ctx *ctx, double *ms) {
    CHECK_CTX(ctx);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(seg_finish_all(ctx));
    GE_TRY(ctx->join_bulk());
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(ctx->ev1));
    float f = 0;
    CUDA_TRY(cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    *ms = f;
    return GE_OK;
}

}  // extern "C"
