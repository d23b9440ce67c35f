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

// expected draws of one generation at full capacity (+ ten standard deviations): what the draw buffers are sized for once
static uint64_t draw_bound(double mean) { return (uint64_t)(mean + 10.0 * std::sqrt(mean + 1.0)) + 4096; }

void ge_ctx::drop_graphs() {
    for (auto &kv : graphs) {
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    graphs.clear();
}

// ge_ctx::pull_state — see ge_context.cuh
int ge_ctx::pull_state(const char *where) {
    const int np = cfg.n_pop;
    CUDA_TRY(cudaMemcpyAsync(h_ss_all, d_ss_all.p, sizeof(StepState) * np, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    int rc = GE_OK;
    for (int p = 0; p < np; p++) {
        PopDev &P = pop[p];
        P.hs = h_ss_all[p];
        const StepState &h = P.hs;
        for (int k = 0; k < 2; k++) { P.st[k].n = h.n[k]; P.st[k].n_hm = h.n_hm[k]; P.st[k].seg.n_seg = h.n_seg[k]; }
        P.n_couples = h.n_couples; P.n_off = h.n_off; P.n_xo = h.n_xo; P.n_mut = h.n_mut; P.prev_n = h.prev_n;
        for (EvPair &ev : ev_pending) if (ev.bytes_per_offspring && ev.pop == p) { ev.bytes = ev.bytes_per_offspring * h.n_off; ev.bytes_per_offspring = 0; }
        if (!h.err || rc != GE_OK) continue;
        const uint32_t e = h.err;
        const std::string lists = "num_males_mate=" + std::to_string(h.n_m) + ", num_females_mate=" + std::to_string(h.n_f);
        if (e & SE_NO_MATES_RM) rc = fail(GE_ERR_NO_MATES, "Error: No one can marry, " + lists);
        else if (e & SE_NO_COUPLES) rc = fail(GE_ERR_NO_MATES, "Error: couples=0, " + lists);
        else if (e & SE_ALL_INBRED) rc = fail(GE_ERR_NO_MATES, "every couple is inbred");
        else if (e & SE_CAP_COUPLES) rc = fail(GE_ERR_CAPACITY, "couples exceed the couple buffers");
        else if (e & SE_NO_OFFSPRING) rc = fail(GE_ERR_NO_MATES, "no offspring");
        else if (e & SE_CAP_OFFSPRING) rc = fail(GE_ERR_CAPACITY, "offspring (" + std::to_string(h.n_off) + ") exceed capacity");
        else if (e & SE_CAP_XO) rc = fail(GE_ERR_CAPACITY, "crossovers (" + std::to_string(h.n_xo) + ") exceed the draw buffers (" + std::to_string(h.xo_cap) + ", sized from the genetic map)");
        else if (e & SE_CAP_MUT) rc = fail(GE_ERR_CAPACITY, "mutation hits (" + std::to_string(h.n_mut) + ") exceed the draw buffers");
        else if (e & SE_CAP_HM) rc = fail(GE_ERR_CAPACITY, "per-haplotype mutation lists exceed their buffer");
        else if (e & SE_CAP_SEG) rc = fail(GE_ERR_CAPACITY, "segments exceed seg_capacity (" + std::to_string(std::max(h.n_seg[0], h.n_seg[1])) + " parts; reported by the first read-back after the generation that overflowed)");
        else if (e & SE_SEG_UNSORTED) rc = fail(GE_ERR_UNSUPPORTED, "gametes had crossover positions that do not ascend: the packed segment format cannot hold the pieces the reference emits for them "
                                                                    "(create the context with GE_FLAG_SEG_WIDE_PARTS to keep its 16-byte parts)");
        else if (e & SE_NAN) rc = fail(GE_ERR_NAN, std::string("Error: A or D is nan (") + where + ")");
        else if (e & SE_PARENT_ID) rc = fail(GE_ERR_INVALID, "parent ID outside the previous generation (the reference reads out of bounds here, :3118-3133)");
    }
    if (rc != GE_OK) {
        const std::string keep = g_err;
        for (int p = 0; p < np; p++) if (pop[p].hs.err) { pop[p].hs.err = 0; push_state(pop[p], offsetof(StepState, err), 4); }
        g_err = keep;
    }
    return rc;
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
    // the warps of a CTA (1-8 by row length, below) still find a few items each per offspring
    uint64_t chunks_per_gamete = 0;
    for (int c = 0; c < C; c++) chunks_per_gamete += ((ctx->chr_nloci[c] + 31) / 32 + 3) / 4;
    uint32_t TILE = 1024;  // 16 KB; measured 0.7 % better than 8 KB on the whole genome, 4 KB and 2 KB are 2 % worse
    // (one rank's share of config 3, 8 / 4 ranks, at least 1 / 3 / 6 / 12 items per gamete: 1.111 / 1.107 / 1.117 / 1.367 and 1.929 / 1.929 / 1.972 / 2.036 ms
    //  per generation — profiles/r2G_tile_rule.txt)
    while (TILE > 64 && chunks_per_gamete / TILE < 3) TILE >>= 1;
    struct Item { uint32_t c, q0, nq; };
    std::vector<Item> items;
    for (int c = 0; c < C; c++) {
        uint32_t nw = (ctx->chr_nloci[c] + 31) / 32, nq = (nw + 3) / 4;
        for (uint32_t q = 0; q < nq; q += TILE) items.push_back({(uint32_t)c, q, std::min(TILE, nq - q)});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item &a, const Item &b) { return a.nq > b.nq; });
    std::vector<uint32_t> tc, t0, tn;
    for (auto &it : items) { tc.push_back(it.c); t0.push_back(it.q0); tn.push_back(it.nq); }
    // warps per offspring CTA: about one per 16 KB of the offspring's two rows.  A shard of a multi-GPU run has short rows: fewer, longer-lived
    // warps per CTA and more offspring in flight per SM instead of eight warps that finish after one item.  Measured on one rank's share of
    // config 3 (scripts/emulate_rank.py; 8 / 4 / 2 / 1 warps): 31 KB per offspring (8 ranks) 1.250 / 1.106 / 1.066 / 1.228 ms per generation,
    // 62 KB (4 ranks) 2.129 / 1.985 / 2.034 / 2.330, 125 KB (2 ranks) 3.877 / 3.805 / 3.938, 250 KB (1 GPU) 7.32 / 7.50 / 7.94.
    {
        const uint64_t bytes_per_offspring = chunks_per_gamete * 16 * 2;
        ctx->prop_threads = 32 * (unsigned)std::min<uint64_t>(8, std::max<uint64_t>(1, (bytes_per_offspring + 8192) / 16384));
    }
    ctx->n_tiles = (uint32_t)items.size();
    ctx->n_loci_total = pos.size();
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
    GE_TRY(ctx->ensure(ctx->d_LG, (size_t)ctx->Wcv * 8 * 256 * 16));
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
    if (cfg->n_pop < 1 || cfg->n_pop > 15 || cfg->n_chr < 1 || cfg->n_phen < 1 || cfg->n_phen > 8) return fail(GE_ERR_INVALID, "bad n_pop/n_chr/n_phen (at most 15 populations, 8 phenotypes)");
    if (cfg->capacity == 0) return fail(GE_ERR_INVALID, "capacity must be > 0");
    if (cfg->capacity > 0x7FFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "capacity above 2^31 individuals (positions in a generation are 32-bit)");
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
    c->serial = (cfg->flags & GE_FLAG_SERIAL) != 0;
    c->seg_wide = (cfg->flags & GE_FLAG_SEG_WIDE_PARTS) != 0;
    c->seg_per_thread = (cfg->flags & GE_FLAG_SEG_VERBATIM) != 0;
    c->cv_from_segments = (cfg->flags & GE_FLAG_CV_FROM_SEGMENTS) != 0;
    c->use_graph = (cfg->flags & GE_FLAG_NO_GRAPH) == 0;
    // every CUDA object is created through one failure path: ge_destroy copes with whatever exists by then
    auto setup = [&]() -> int {
        int prio_lo = 0, prio_hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_TRY(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->xo_ctas_per_sm, sample_xo_kernel, 128, 0));   // one resident wave of the lane-refilling sampler
        c->xo_ctas_per_sm = std::max(1, c->xo_ctas_per_sm);
        if (!c->stream) CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
        if (!c->bulk) CUDA_TRY(cudaStreamCreateWithPriority(&c->bulk, cudaStreamNonBlocking, prio_lo));
        CUDA_TRY(cudaEventCreate(&c->ev0)); CUDA_TRY(cudaEventCreate(&c->ev1));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        for (PopDev &P : c->pop) for (DrawSet &D : P.ds) CUDA_TRY(cudaEventCreateWithFlags(&D.bulk_done, cudaEventDisableTiming));
        for (PopDev &P : c->pop) CUDA_TRY(cudaEventCreateWithFlags(&P.ev_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        for (SortLane &l : c->lane) { if (!l.s) CUDA_TRY(cudaStreamCreateWithPriority(&l.s, cudaStreamNonBlocking, prio_hi)); CUDA_TRY(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming)); }
        // the device-resident step state of every population, and its pinned read-back buffer
        GE_TRY(c->ensure(c->d_ss_all, sizeof(StepState) * cfg->n_pop));
        CUDA_TRY(cudaMemsetAsync(c->d_ss_all.p, 0, sizeof(StepState) * cfg->n_pop, c->stream));
        CUDA_TRY(cudaMallocHost(&c->h_ss_all, sizeof(StepState) * cfg->n_pop));
        for (int p = 0; p < cfg->n_pop; p++) {
            PopDev &P = c->pop[p];
            P.d_ss = c->d_ss_all.as<StepState>() + p;
            P.hs = StepState{};
            P.hs.cap = cfg->capacity;
            for (int k = 0; k < 2; k++) { P.st[k].d_n = &P.d_ss->n[k]; P.st[k].d_n_hm = &P.d_ss->n_hm[k]; }
            GE_TRY(c->push_state(P));
        }
        GE_TRY(c->ensure_partial());
        return GE_OK;
    };
    if (int rc = setup()) { const std::string keep = g_err; ge_destroy(c); g_err = keep; return rc; }
    *out = c;
    return GE_OK;
}

int ge_destroy(ge_ctx *ctx) {
    if (!ctx) return GE_OK;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->bulk) cudaStreamSynchronize(ctx->bulk);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (SortLane &l : ctx->lane) if (l.s) cudaStreamSynchronize(l.s);
    ctx->drop_graphs();
    auto freeb = [&](Buf &b) { ctx->release(b); };
    for (PopDev &P : ctx->pop) {
        for (Buf *b : {&P.d_row_off, &P.d_bp, &P.d_T, &P.d_bp_dist, &P.d_mrow_off, &P.d_mbp, &P.d_mT, &P.d_vb, &P.d_vb_off, &P.d_vb_scale, &P.d_mvb, &P.d_mvb_off, &P.d_mvb_scale, &P.d_cov_lo, &P.d_cov_hi, &P.d_omega,
                       &P.d_lambda, &P.d_vd_zero, &P.prev_P, &P.prev_F, &P.c_male, &P.c_female, &P.c_inbreed, &P.c_noff, &P.mut_off, &P.mut_bp, &P.mut_gam, &P.e_raw, &P.cnt32, &P.d_sv0, &P.founder_rows, &P.founder_cv})
            freeb(*b);
        freeb(P.mig_pop[0]); freeb(P.mig_idx[0]); freeb(P.rowmap_buf[0]); freeb(P.rowmap_buf[1]); freeb(P.founder_root);
        for (PopDev::SegSnapshot &H : P.history) { freeb(H.off); freeb(H.seg); }
        for (DrawSet &D : P.ds) {
            for (Buf *b : {&D.father, &D.mother, &D.couple_of, &D.xo_off, &D.xo_bp, &D.flips, &D.start_hap}) freeb(*b);
            if (D.bulk_done) cudaEventDestroy(D.bulk_done);
        }
        for (GenState &s : P.st) {
            for (Buf *b : {&s.hap, &s.cv_allele, &s.cv_root, &s.ids, &s.sex, &s.A, &s.D, &s.G, &s.C, &s.E, &s.F, &s.P, &s.mv, &s.sv, &s.svf, &s.hm_off, &s.hm_bp}) freeb(*b);
            seg_release(ctx, s.seg);
        }
        mate_release(ctx, P.mate);
        if (P.ev_ready) cudaEventDestroy(P.ev_ready);
    }
    for (Buf *b : {&ctx->d_chr_word_off, &ctx->d_chr_nloci, &ctx->d_locus_off, &ctx->d_pos, &ctx->d_bkt_off, &ctx->d_bkt_shift, &ctx->d_bkt, &ctx->d_LA, &ctx->d_LG, &ctx->d_cv_bitpos, &ctx->xo_stash, &ctx->d_tile_chr, &ctx->d_tile_chunk0, &ctx->d_tile_nchunk,
                   &ctx->d_cv_block_off, &ctx->d_cv_word_off, &ctx->d_cv_word_blk, &ctx->d_cv_bp, &ctx->d_cv_chr, &ctx->d_a_eff, &ctx->d_d_eff, &ctx->d_cv_count, &ctx->scan_blocks, &ctx->bulk_scan_blocks,
                   &ctx->partial, &ctx->scalars, &ctx->d_ss_all, &ctx->d_chr_ids, &ctx->ar_scratch, &ctx->seg_desc, &ctx->seg_iv_off, &ctx->seg_cnt, &ctx->seg_flags, &ctx->seg_verb})
        freeb(*b);
    if (ctx->h_ss_all) cudaFreeHost(ctx->h_ss_all);
    for (uint64_t *c : ctx->pinned_chunks) cudaFreeHost(c);
    for (cudaEvent_t e : {ctx->ev0, ctx->ev1, ctx->ev_ready, ctx->ev_join, ctx->ev_fork}) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_pending) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    for (SortLane &l : ctx->lane) {
        for (Buf *b : {&l.keys_in, &l.vals_in, &l.keys_out, &l.vals_out, &l.tmp}) freeb(*b);
        if (l.done) cudaEventDestroy(l.done);
        if (l.s) cudaStreamDestroy(l.s);
    }
    for (int q = 0; q < 2; q++) { for (Buf &b : ctx->mig_lists[q]) freeb(b); if (ctx->mig_done[q]) cudaEventDestroy(ctx->mig_done[q]); }
    freeb(ctx->mig_stage);
    if (ctx->bulk) cudaStreamDestroy(ctx->bulk);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
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
    if (ctx->gen0_done) return fail(GE_ERR_INVALID, "ge_set_genetic_map after ge_init_generation0 (the segment format and the draw buffers are fixed there)");
    // A row with probability >= 1 would end the survival product the skip sampler searches (T = 0 from there on): the reference
    // draws an independent Bernoulli per row (:2985-2990) and would go on.  Such maps (a 100 cM jump between two rows) are refused.
    for (uint64_t j = 0; j < n; j++) if (rp[j] >= 1.0) return fail(GE_ERR_UNSUPPORTED, "genetic map row with recombination probability >= 1 (a jump of 100 cM or more between two rows)");
    // the bit-packed representation needs monotone crossover lists: row j's crossover lies in
    // [bp[j], bp[j]+bp_dist) (:2989), so two rows that can both recombine must be at least bp_dist apart (true for the uniform b37 maps)
    bool close_rows = false;
    for (uint64_t j = 0; j + 1 < n; j++)
        if (bp[j + 1] < bp[j] + bp_dist && rp[j] > 0 && rp[j + 1] > 0) close_rows = true;
    if (close_rows) {
        if (ctx->cfg.representation & GE_REP_BITS)
            return fail(GE_ERR_UNSUPPORTED, "genetic map rows closer than bp_dist_in_rmap give non-monotone crossover lists (only GE_REP_SEGMENTS follows the reference there)");
        ctx->seg_per_thread = true;  // segment lists may become unsorted: keep the reference's scan verbatim (and its 16-byte parts)
    }
    PopDev &P = ctx->pop[pop];
    P.rmap_bp[chr].assign(bp, bp + n); P.recom_prob[chr].assign(rp, rp + n); P.bp_dist[chr] = bp_dist;
    return GE_OK;
}
int ge_set_mutation_map(ge_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rate, uint64_t n) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    if (!bp || !rate || n < 2) return fail(GE_ERR_INVALID, "ge_set_mutation_map: need >= 2 rows");
    if (ctx->gen0_done) return fail(GE_ERR_INVALID, "ge_set_mutation_map after ge_init_generation0");
    for (uint64_t k = 0; k < n; k++) {
        if (bp[k] > 0x7FFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "mutation-map position does not fit 31 bits");
        if (rate[k] >= 1.0 && k >= 1) return fail(GE_ERR_UNSUPPORTED, "mutation rate >= 1 in a map row (the skip sampler's survival product would end there)");
    }
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
int ge_set_allreduce(ge_ctx *ctx, ge_allreduce_fn fn, void *user) {
    CHECK_CTX(ctx);
    ctx->allreduce = fn; ctx->allreduce_user = user;
    ctx->nccl_all_reduce = nullptr; ctx->nccl_comm = nullptr;
    ctx->graph_epoch++;
    return GE_OK;
}
int ge_set_allreduce_nccl(ge_ctx *ctx, void *nccl_comm, void *nccl_all_reduce) {
    CHECK_CTX(ctx);
    if ((nccl_comm == nullptr) != (nccl_all_reduce == nullptr)) return fail(GE_ERR_INVALID, "ge_set_allreduce_nccl: communicator and entry point go together");
    ctx->nccl_comm = nccl_comm;
    ctx->nccl_all_reduce = reinterpret_cast<ge_ctx::nccl_all_reduce_fn>(nccl_all_reduce);
    ctx->allreduce = nullptr; ctx->allreduce_user = nullptr;
    ctx->graph_epoch++;
    return GE_OK;
}
int ge_set_gamma(ge_ctx *ctx, const double *g) { CHECK_CTX(ctx); ctx->gamma.assign(g, g + ctx->cfg.n_phen); return GE_OK; }

// ---------------- per-method entry points ----------------
// Every method has an enqueue_* form that only queues work (sizes are read from StepState on the device) and a public
// form that adds the host synchronisation (pull_state) the reference's `bool` return needs.  ge_step_generation chains the
// enqueue_* forms and synchronises once.

static int enqueue_AD(ge_ctx *ctx, int pop) {  // ras_compute_AD :2624-2749
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    const uint64_t cap = ctx->cfg.capacity;
    ge_ctx::PhaseTimer timer(ctx, GE_PHASE_CV_AD);
    if (ctx->segs() && !ctx->bits() && ctx->cv_from_segments) GE_TRY(seg_find_cv(ctx, pop));  // ras_find_cv on the segment lists (verification mode)
    uint32_t ncv = ctx->n_cv_tot;
    if (ncv) {
        CUDA_TRY(cudaMemsetAsync(ctx->d_cv_count.p, 0, (size_t)ncv * 8, ctx->stream));
        // rows per CTA: long enough to amortise the unpacking of the bit-sliced counters, short enough that a small population still spreads out
        const uint32_t rows_per_cta = (uint32_t)std::min<uint64_t>(1024, std::max<uint64_t>(64, ((2 * cap / ((uint64_t)ctx->n_sm * 4)) + 7) & ~7ull));
        dim3 grid(nblk(ctx->Wcv, 32), nblk(2 * cap, rows_per_cta));
        cv_count_bits_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(ctx->cvset(), S.cv_allele.as<uint32_t>(), S.d_n, rows_per_cta, ctx->d_cv_count.as<unsigned long long>());
        GE_TRY(ctx->check_launch("cv_count"));
    }
    const uint64_t nw = cap * ctx->cfg.n_phen;
    if (!ctx->use_root && ncv) {
        cv_tables_kernel<<<nblk(ncv, 128), 128, 0, ctx->stream>>>(ctx->cvset(), ctx->d_cv_count.as<unsigned long long>(), S.d_n, ctx->d_a_eff.as<double>(),
                                                                  ctx->d_d_eff.as<double>(), P.d_vd_zero.as<uint8_t>(), ctx->d_LA.as<double2>());
        GE_TRY(ctx->check_launch("cv_tables"));
        cv_group_tables_kernel<<<nblk((uint64_t)ctx->Wcv * 8 * 256, 256), 256, 0, ctx->stream>>>(ctx->cvset(), ctx->d_LA.as<double2>(), ctx->d_LG.as<double2>());
        GE_TRY(ctx->check_launch("cv_group_tables"));
        genetic_value_groups_kernel<<<ctx->ctrl_grid(cap * 32, 256), 256, 0, ctx->stream>>>(ctx->cvset(), S.cv_allele.as<uint32_t>(), ctx->d_LG.as<double2>(), S.d_n, cap, S.A.as<double>(),
                                                                                      S.D.as<double>(), S.G.as<double>(), &P.d_ss->err);
        GE_TRY(ctx->check_launch("genetic_value_groups"));
    } else {
        genetic_value_kernel<<<ctx->ctrl_grid(nw * 32, 256), 256, 0, ctx->stream>>>(
            ctx->cvset(), S.cv_allele.as<uint32_t>(), ctx->use_root ? S.cv_root.as<uint8_t>() : nullptr, ctx->d_cv_count.as<unsigned long long>(), S.d_n,
            ctx->d_a_eff.as<double>(), ctx->d_d_eff.as<double>(), P.d_vd_zero.as<uint8_t>(), cap, S.A.as<double>(), S.D.as<double>(), S.G.as<double>(), &P.d_ss->err);
        GE_TRY(ctx->check_launch("genetic_value"));
    }
    if (ctx->sharded()) {
        // sharded contexts hold partial sums over their own loci: sum A, D, G over the ranks.  The columns go whole (stride =
        // capacity; rows beyond the population hold stale finite values nobody reads), so the count does not depend on a device-side size.
        size_t nb = (size_t)cap * ctx->cfg.n_phen * 8;
        GE_TRY(ctx->ensure(ctx->ar_scratch, 3 * nb));
        char *sc = ctx->ar_scratch.as<char>();
        CUDA_TRY(cudaMemcpyAsync(sc, S.A.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(sc + nb, S.D.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(sc + 2 * nb, S.G.p, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        if (ctx->nccl_all_reduce) {   // ncclDouble = 8, ncclSum = 0 (nccl.h); in place, stream-ordered, capturable
            const int rc = ctx->nccl_all_reduce(sc, sc, 3 * nb / 8, 8, 0, ctx->nccl_comm, ctx->stream);
            if (rc != 0) return fail(GE_ERR_INVALID, "ncclAllReduce failed with ncclResult_t " + std::to_string(rc));
        } else {
            const int rc = ctx->allreduce(ctx->allreduce_user, reinterpret_cast<double *>(sc), 3 * nb / 8, (void *)ctx->stream);
            if (rc != 0) return fail(GE_ERR_INVALID, "allreduce hook failed");
        }
        CUDA_TRY(cudaMemcpyAsync(S.A.p, sc, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(S.D.p, sc + nb, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(S.G.p, sc + 2 * nb, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return GE_OK;
}
int ge_compute_AD(ge_ctx *ctx, int pop, int gen) {
    CHECK_POP(ctx, pop);
    (void)gen;
    if (ctx->pop[pop].st[ctx->pop[pop].cur].n == 0) return fail(GE_ERR_INVALID, "empty population");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(enqueue_AD(ctx, pop));
    return ctx->pull_state("ge_compute_AD");
}

// e_host / f0_host: replayed draws (host arrays of the population's current size), else Philox
static int enqueue_GEF(ge_ctx *ctx, int pop, int f, bool gen0, const double *e_host, const double *f0_host, bool with_mv_sv = false) {  // :3075-3206
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    const uint64_t cap = ctx->cfg.capacity;
    Scheme &sc = P.scheme[f];
    GE_TRY(ctx->ensure(P.e_raw, (size_t)cap * ctx->cfg.n_phen * 8));
    GE_TRY(ctx->ensure(ctx->scalars, 64 * 8));
    GE_TRY(ctx->ensure_partial());
    double *e = P.e_raw.as<double>() + (uint64_t)f * cap;
    double *var_e = ctx->scalars.as<double>() + 8, *mean_e = ctx->scalars.as<double>() + 9;
    const unsigned mg = ctx->moment_grid(cap);
    if (e_host) {
        CUDA_TRY(cudaMemcpyAsync(e, e_host, S.n * 8, cudaMemcpyHostToDevice, ctx->stream));
        GE_TRY(ctx->d_mean(e, devn(S.d_n), cap, mean_e));
    } else if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
        enoise_mean_kernel<<<mg, 256, 0, ctx->stream>>>(ctx->rng, P.d_ss, pop, f, S.d_n, e, ctx->partial.as<double>(), mean_e);
        GE_TRY(ctx->check_launch("enoise_mean"));
    } else return fail(GE_ERR_INVALID, "replay mode needs e_raw");
    P.have_e_raw = true;
    moment_kernel<<<mg, 256, 0, ctx->stream>>>(e, limited(devn(S.d_n), cap), mean_e, 1, 1, ctx->partial.as<double>(), var_e);
    GE_TRY(ctx->check_launch("moment<var>"));
    double *f0 = nullptr;
    if (gen0 && sc.vf > 0) {
        f0 = S.F.as<double>() + (uint64_t)f * cap;  // staged in place, the kernel reads f0[i] before writing F[i]
        if (f0_host) CUDA_TRY(cudaMemcpyAsync(f0, f0_host, S.n * 8, cudaMemcpyHostToDevice, ctx->stream));
        else if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
            normal_scaled_kernel<<<nblk(S.n, 256), 256, 0, ctx->stream>>>(ctx->rng, P_F0, pop, 0, f, S.n, std::sqrt(sc.vf), f0);
            GE_TRY(ctx->check_launch("f0"));
        } else return fail(GE_ERR_INVALID, "replay mode needs parental0 for vf > 0");
    }
    PhenoArgs a;
    a.s_a = 1; if (sc.va > 0) a.s_a = std::sqrt(P.var_a0[f] / sc.va);
    a.s_d = 0; if (sc.vd > 0) a.s_d = std::sqrt(P.var_d0[f] / sc.vd); else if (sc.vd == -1) a.s_d = 1;
    a.ve = sc.ve; a.vf = sc.vf; a.beta = sc.beta; a.vt_type = ctx->cfg.vt_type;
    a.omega = a.lambda = a.sv0 = nullptr; a.mv = a.sv = a.svf = nullptr;
    if (with_mv_sv) {
        a.omega = P.d_omega.as<double>(); a.lambda = P.d_lambda.as<double>(); a.sv0 = P.d_sv0.as<double>();
        a.mv = S.mv.as<double>(); a.sv = S.sv.as<double>(); a.svf = S.svf.as<double>();
    }
    uint64_t o = (uint64_t)f * cap;
    phenotype_kernel<<<ctx->grid_for(cap, 256), 256, 0, ctx->stream>>>(
        a, P.d_ss, S.d_n, &P.d_ss->prev_n, e, var_e, S.A.as<double>() + o, S.D.as<double>() + o, S.G.as<double>() + o, S.C.as<double>() + o, S.E.as<double>() + o,
        S.F.as<double>() + o, S.P.as<double>() + o, S.ids.as<uint64_t>(), P.prev_P.as<double>() ? P.prev_P.as<double>() + o : nullptr,
        P.prev_F.as<double>() ? P.prev_F.as<double>() + o : nullptr, f0, &P.d_ss->err);
    return ctx->check_launch("phenotype");
}

int ge_scale_AD_compute_GEF(ge_ctx *ctx, int pop, int gen, int phen, const double *e_raw) {
    CHECK_POP(ctx, pop); CHECK_PHEN(ctx, phen);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    if (P.hs.gen != gen) { P.hs.gen = gen; GE_TRY(ctx->push_state(P, offsetof(StepState, gen), 4)); }
    GE_TRY(enqueue_GEF(ctx, pop, phen, gen == 0, e_raw, nullptr));
    return ctx->pull_state("ge_scale_AD_compute_GEF");   // (also: e_raw is a caller buffer)
}

static int enqueue_mv_sv(ge_ctx *ctx, int pop) {  // :3300-3342, generations >= 1 (generation 0 also takes the moments, below)
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    const uint64_t cap = ctx->cfg.capacity;
    mv_sv_selection_kernel<<<ctx->grid_for(cap, 256), 256, 0, ctx->stream>>>(P.d_ss, S.d_n, ctx->cfg.n_phen, cap, S.P.as<double>(), P.d_omega.as<double>(), P.d_lambda.as<double>(),
                                                                             P.d_sv0.as<double>(), S.mv.as<double>(), S.sv.as<double>(), S.svf.as<double>());
    return ctx->check_launch("mv_sv_selection");
}
int ge_compute_mating_value_selection_value(ge_ctx *ctx, int pop, int gen, const ge_gen_params *gp) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    const uint64_t cap = ctx->cfg.capacity;
    GE_TRY(ctx->ensure(P.d_sv0, 16));
    ge_gen_params dummy{};
    if (!gp) gp = &dummy;
    P.hs.gen = gen; P.hs.sel_func = gp->selection_func; P.hs.sel_par1 = gp->selection_par1; P.hs.sel_par2 = gp->selection_par2;
    GE_TRY(ctx->push_state(P, 0, offsetof(StepState, n)));
    if (gen == 0) {   // the generation-0 mean and variance of the raw selection value standardise every later generation (:3325-3337)
        mv_sv_selection_kernel<<<ctx->grid_for(cap, 256), 256, 0, ctx->stream>>>(P.d_ss, S.d_n, ctx->cfg.n_phen, cap, S.P.as<double>(), P.d_omega.as<double>(), P.d_lambda.as<double>(),
                                                                                 nullptr, S.mv.as<double>(), S.sv.as<double>(), S.svf.as<double>());
        GE_TRY(ctx->check_launch("mv_sv_raw"));
        double *sv0 = P.d_sv0.as<double>();
        GE_TRY(ctx->d_var(S.sv.as<double>(), devn(S.d_n), cap, sv0 + 1, sv0 + 0));
        double h[2];
        CUDA_TRY(cudaMemcpyAsync(h, sv0, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        P.sv_mean0 = h[0]; P.sv_var0 = h[1];
    }
    GE_TRY(enqueue_mv_sv(ctx, pop));
    return ctx->pull_state("ge_compute_mating_value_selection_value");
}

static int enqueue_save_prev(ge_ctx *ctx, int pop) {  // :3211-3236
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    const uint64_t cap = ctx->cfg.capacity;
    size_t bytes = (size_t)cap * ctx->cfg.n_phen * 8;
    GE_TRY(ctx->ensure(P.prev_P, bytes)); GE_TRY(ctx->ensure(P.prev_F, bytes));
    save_prev_kernel<<<ctx->grid_for(cap * ctx->cfg.n_phen, 256), 256, 0, ctx->stream>>>(S.d_n, ctx->cfg.n_phen, cap, S.P.as<double>(), S.F.as<double>(), P.prev_P.as<double>(),
                                                                                       P.prev_F.as<double>(), &P.d_ss->prev_n);
    return ctx->check_launch("save_prev");
}
int ge_save_human_info_to_Pop_info_prev_gen(ge_ctx *ctx, int pop) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(enqueue_save_prev(ctx, pop));
    ctx->pop[pop].prev_n = ctx->pop[pop].st[ctx->pop[pop].cur].n;
    return GE_OK;
}

int ge_environmental_effects_specific_to_each_population(ge_ctx *ctx, int f) {  // :3345-3381
    CHECK_CTX(ctx); CHECK_PHEN(ctx, f);
    if (ctx->gamma.empty() || ctx->gamma[f] == 0) return GE_OK;
    int np = ctx->cfg.n_pop;
    if (np < 2) return fail(GE_ERR_UNSUPPORTED, "--gamma with one population divides by zero in the reference (:3269)");
    GE_TRY(ctx->pull_state("environmental effects"));   // population sizes (the moments below go through the host anyway)
    const uint64_t cap = ctx->cfg.capacity;
    // pooled variance of P + a*s_i is quadratic in a; the moments come from the device, the scalar Newton
    // iteration (NewtonRaphson :44-63, central difference :35-39) runs on the host with the same start/tolerance
    std::vector<double> ni(np), mu(np), m2(np), s(np);
    double N = 0;
    for (int p = 0; p < np; p++) {
        GenState &S = ctx->pop[p].st[ctx->pop[p].cur];
        double v, m;
        GE_TRY(ctx->h_var(S.P.as<double>() + (uint64_t)f * cap, S.n, &v, &m));
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
        add_scalar_kernel<<<nblk(S.n, 256), 256, 0, ctx->stream>>>(S.P.as<double>() + (uint64_t)f * cap, S.n, x1 * s[p]);
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
    // sizes the host decides go to the device-resident step state; the draw buffers are sized ONCE, from the maps: expected
    // crossovers / mutation hits of a generation at full capacity plus ten standard deviations
    {
        const uint64_t cap = ctx->cfg.capacity;
        double e_xo = 0, e_mut = 0;
        for (int c = 0; c < C; c++) {
            for (double q : P.recom_prob[c]) e_xo += std::min(std::max(q, 0.0), 1.0);
            for (size_t k = 1; k < P.mutmap_rate[c].size(); k++) e_mut += std::min(std::max(P.mutmap_rate[c][k], 0.0), 1.0);
        }
        P.hs.n[0] = n; P.hs.n[1] = 0; P.hs.n_hm[0] = P.hs.n_hm[1] = 0; P.hs.gen = 0; P.hs.prev_n = n;
        P.hs.xo_cap = draw_bound(2.0 * (double)cap * e_xo);
        P.hs.mut_cap = P.has_mut ? draw_bound((double)cap * e_mut) : 0;
        P.hs.hm_cap = 0;
        GE_TRY(ctx->push_state(P));
        const uint64_t slots = cap * C * 2;
        for (DrawSet &D : P.ds) {
            GE_TRY(ctx->ensure(D.father, cap * 4)); GE_TRY(ctx->ensure(D.mother, cap * 4)); GE_TRY(ctx->ensure(D.couple_of, cap * 4));
            GE_TRY(ctx->ensure(D.xo_off, (slots + 1) * 8)); GE_TRY(ctx->ensure(D.start_hap, slots));
            GE_TRY(ctx->ensure(D.xo_bp, P.hs.xo_cap * 4));
            if (ctx->bits()) GE_TRY(ctx->ensure(D.flips, P.hs.xo_cap * 4));
        }
        if (ctx->cfg.rng_mode == GE_RNG_PHILOX) GE_TRY(ctx->ensure(ctx->xo_stash, slots * XO_STASH * 4));
        GE_TRY(ctx->ensure(P.cnt32, (slots + 1) * 4));
        if (P.has_mut) {
            GE_TRY(ctx->ensure(P.mut_off, (cap * C + 1) * 8)); GE_TRY(ctx->ensure(P.mut_bp, P.hs.mut_cap * 4)); GE_TRY(ctx->ensure(P.mut_gam, P.hs.mut_cap));
        }
    }
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
        sex_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->rng, p, 0, n, S.sex.as<uint8_t>());
        GE_TRY(ctx->check_launch("sex"));
    } else {
        if (!d0 || !d0[p].sex) return fail(GE_ERR_INVALID, "replay mode needs draws0[pop].sex");
        CUDA_TRY(cudaMemcpyAsync(S.sex.p, d0[p].sex, n, cudaMemcpyHostToDevice, ctx->stream));
    }
    CUDA_TRY(cudaMemsetAsync(S.C.p, 0, (size_t)ctx->cfg.capacity * nf * 8, ctx->stream));
    for (int f = 0; f < nf; f++) {
        if (P.scheme[f].vc > 0) {
            double *dst = S.C.as<double>() + (uint64_t)f * ctx->cfg.capacity;
            if (ctx->cfg.rng_mode == GE_RNG_PHILOX) {
                normal_scaled_kernel<<<nblk(n, 256), 256, 0, ctx->stream>>>(ctx->rng, P_COMMON, p, 0, f, n, std::sqrt(P.scheme[f].vc), dst);
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
    if (ctx->gen0_done) return fail(GE_ERR_INVALID, "ge_init_generation0 called twice");
    int nf = ctx->cfg.n_phen, np = ctx->cfg.n_pop;
    const uint64_t cap = ctx->cfg.capacity;
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
        GE_TRY(enqueue_AD(ctx, p));
        GenState &S = P.st[P.cur];
        // ras_fill_Pop_info_prev_gen_for_gen0_prev :3240-3251
        size_t bytes = (size_t)cap * nf * 8;
        GE_TRY(ctx->ensure(P.prev_P, bytes)); GE_TRY(ctx->ensure(P.prev_F, bytes));
        CUDA_TRY(cudaMemsetAsync(P.prev_P.p, 0, bytes, ctx->stream)); CUDA_TRY(cudaMemsetAsync(P.prev_F.p, 0, bytes, ctx->stream));
        P.prev_n = S.n;
        for (int f = 0; f < nf; f++) {  // :555-566
            GE_TRY(ctx->h_var(S.A.as<double>() + (uint64_t)f * cap, S.n, &P.var_a0[f]));
            GE_TRY(ctx->h_var(S.D.as<double>() + (uint64_t)f * cap, S.n, &P.var_d0[f]));
            bool rp = ctx->cfg.rng_mode == GE_RNG_REPLAY && d0;
            GE_TRY(enqueue_GEF(ctx, p, f, true, (rp && d0[p].e_raw) ? d0[p].e_raw + (uint64_t)f * S.n : nullptr,
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
            GE_TRY(ctx->h_var(S.P.as<double>() + (uint64_t)f * cap, S.n, &vP));
            GE_TRY(ctx->h_var(S.F.as<double>() + (uint64_t)f * cap, S.n, &vF));
            if (ctx->cfg.vt_type == 1) P.scheme[f].beta = std::sqrt(P.scheme[f].vf / (2 * vP));
            else if (ctx->cfg.vt_type == 2) { if (vF > 0) P.scheme[f].beta = std::sqrt(P.scheme[f].vf / (2 * vF)); }
        }
        std::vector<std::vector<uint8_t>>().swap(P.panel);  // the host copy of the panel is no longer needed
        P.panel.resize(ctx->cfg.n_chr);
    }
    for (PopDev &P : ctx->pop) for (const Scheme &sc : P.scheme) if (sc.vf > 0) ctx->needs_prev = true;
    ctx->gen0_done = true;
    return ctx->pull_state("generation 0");
}

// ---------------- mating ----------------

int ge_set_couples(ge_ctx *ctx, int pop, const uint64_t *m, const uint64_t *f, const uint8_t *inb, const int32_t *no, uint64_t n) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    if (!ctx->gen0_done) return fail(GE_ERR_INVALID, "ge_set_couples before ge_init_generation0");
    if (n == 0 || !m || !f || !inb || !no) return fail(GE_ERR_INVALID, "ge_set_couples: null array or no couples");
    const uint64_t n_par = P.st[P.cur].n;
    std::vector<uint32_t> mm(n), ff(n);
    for (uint64_t k = 0; k < n; k++) {   // positions index the parent generation: validated before any kernel reads parental rows through them
        if (m[k] >= n_par || f[k] >= n_par) return fail(GE_ERR_INVALID, "ge_set_couples: couple " + std::to_string(k) + " names a position outside the current generation (" + std::to_string(n_par) + " individuals)");
        if (no[k] < 0) return fail(GE_ERR_INVALID, "ge_set_couples: negative num_offspring");
        mm[k] = (uint32_t)m[k]; ff[k] = (uint32_t)f[k];
    }
    GE_TRY(ensure_couples(ctx, P, n));
    CUDA_TRY(cudaMemcpyAsync(P.c_male.p, mm.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(P.c_female.p, ff.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(P.c_inbreed.p, inb, n, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(P.c_noff.p, no, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    P.n_couples = n;
    P.hs.n_couples = n;
    return ctx->push_state(P, offsetof(StepState, n_couples), 8);   // (synchronises: the caller's arrays and mm/ff are free again)
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

static StepRow step_row(int gen, const ge_gen_params &gp) {
    StepRow r;
    r.gen = gen; r.sel_func = gp.selection_func; r.offspring_dist = gp.offspring_dist; r.pad = 0; r.pop_size = gp.pop_size;
    r.mat_cor = gp.mat_cor; r.sel_par1 = gp.selection_par1; r.sel_par2 = gp.selection_par2;
    return r;
}
// the host's copy of the row follows the device's (replayed draws push the whole struct later in the step)
static void host_row(PopDev &P, const StepRow &r) {
    P.hs.gen = r.gen; P.hs.sel_func = r.sel_func; P.hs.offspring_dist = r.offspring_dist; P.hs.pop_size = r.pop_size; P.hs.mat_cor = r.mat_cor;
    P.hs.u11 = std::sqrt(1.0 - r.mat_cor * r.mat_cor); P.hs.sel_par1 = r.sel_par1; P.hs.sel_par2 = r.sel_par2; P.hs.n_inbreed = 0;
}
static int enqueue_step_begin(ge_ctx *ctx, int pop, int gen, const ge_gen_params &gp) {
    PopDev &P = ctx->pop[pop];
    const StepRow r = step_row(gen, gp);
    host_row(P, r);
    step_begin_kernel<<<1, 1, 0, ctx->stream>>>(P.d_ss, r);
    return ctx->check_launch("step_begin");
}

int ge_mate(ge_ctx *ctx, int pop, int gen, const ge_gen_params *gp) {  // random_mate :2090-2157 / assort_mate :2167-2360
    CHECK_POP(ctx, pop);
    if (!gp) return fail(GE_ERR_INVALID, "null params");
    if (ctx->cfg.rng_mode != GE_RNG_PHILOX) return fail(GE_ERR_INVALID, "replay mode: supply couples with ge_set_couples or offspring draws");
    if (ctx->pop[pop].st[ctx->pop[pop].cur].n == 0) return fail(GE_ERR_INVALID, "empty population");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(enqueue_step_begin(ctx, pop, gen, *gp));
    {
        ge_ctx::PhaseTimer timer(ctx, GE_PHASE_MATE);
        GE_TRY(enqueue_mate(ctx, pop, *gp));
    }
    return ctx->pull_state("ge_mate");
}

int ge_mate_replay(ge_ctx *ctx, int pop, int gen, const ge_gen_params *gp, const ge_mate_draws *md) {
    CHECK_POP(ctx, pop);
    if (!gp || !md) return fail(GE_ERR_INVALID, "null params / draws");
    PopDev &P = ctx->pop[pop];
    if (P.st[P.cur].n == 0) return fail(GE_ERR_INVALID, "empty population");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(enqueue_step_begin(ctx, pop, gen, *gp));
    GE_TRY(enqueue_mate(ctx, pop, *gp, md));
    GE_TRY(ctx->pull_state("ge_mate_replay"));   // (also: the caller's arrays were read asynchronously)
    // the draws must be the ones of THIS population state: list lengths and couple count follow from the thinning uniforms
    const StepState &h = P.hs;
    if (!P.RM) {
        if (h.n2 != md->n_couples) return fail(GE_ERR_INVALID, "ge_mate_replay: the thinning uniforms give " + std::to_string(h.n2) + " couples, the draws were made for " + std::to_string(md->n_couples));
        if (h.n_trim_list != (md->trim_order ? md->n_trim_order : 0)) return fail(GE_ERR_INVALID, "ge_mate_replay: trim_order does not have the length of the longer sex list");
    }
    return GE_OK;
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

// everything wrong with a caller's draws is found BEFORE any state changes (buffers are flipped only afterwards)
static int validate_draws(ge_ctx *ctx, PopDev &P, const ge_draws *dr, uint64_t n_par) {
    const int C = ctx->cfg.n_chr;
    const uint64_t n_off = dr->n_offspring;
    if (n_off == 0) return fail(GE_ERR_NO_MATES, "no offspring");
    if (n_off > ctx->cfg.capacity) return fail(GE_ERR_CAPACITY, "offspring exceed capacity");
    if (!dr->father || !dr->mother || !dr->sex || !dr->xo_off || !dr->start_hap) return fail(GE_ERR_INVALID, "incomplete draws");
    for (uint64_t i = 0; i < n_off; i++) if (dr->father[i] >= n_par || dr->mother[i] >= n_par) return fail(GE_ERR_INVALID, "parent index out of range");
    const uint64_t ns = n_off * C * 2;
    if (dr->xo_off[0] != 0) return fail(GE_ERR_INVALID, "xo_off[0] must be 0");
    for (uint64_t k = 0; k < ns; k++) if (dr->xo_off[k + 1] < dr->xo_off[k]) return fail(GE_ERR_INVALID, "xo_off must be non-decreasing");
    if (dr->xo_off[ns] && !dr->xo_bp) return fail(GE_ERR_INVALID, "incomplete draws: xo_bp");
    // the bit-packed rows (and packed 8-byte parts) need every slot's crossover positions in ascending order; the reference's own
    // lists are (:2983-2993) whenever the map's rows are at least bp_dist_in_rmap apart
    if (ctx->bits())
        for (uint64_t k = 0; k < ns; k++)
            for (uint64_t e = dr->xo_off[k] + 1; e < dr->xo_off[k + 1]; e++)
                if (dr->xo_bp[e] < dr->xo_bp[e - 1]) return fail(GE_ERR_INVALID, "crossover positions of a gamete must ascend (GE_REP_BITS)");
    if (dr->mut_off) {
        if (!P.has_mut) return fail(GE_ERR_INVALID, "draws carry mutation hits but the population has no mutation map (ge_set_mutation_map)");
        const uint64_t ni = n_off * C;
        if (dr->mut_off[0] != 0) return fail(GE_ERR_INVALID, "mut_off[0] must be 0");
        for (uint64_t k = 0; k < ni; k++) if (dr->mut_off[k + 1] < dr->mut_off[k]) return fail(GE_ERR_INVALID, "mut_off must be non-decreasing");
        if (dr->mut_off[ni] && (!dr->mut_bp || !dr->mut_gam)) return fail(GE_ERR_INVALID, "incomplete draws: mut_bp / mut_gam");
        for (uint64_t e = 0; e < dr->mut_off[ni]; e++) if (dr->mut_bp[e] > 0x7FFFFFFFull) return fail(GE_ERR_UNSUPPORTED, "mutation position does not fit 31 bits");
    }
    return GE_OK;
}

static int enqueue_reproduce(ge_ctx *ctx, int pop, const ge_draws *dr) {  // reproduce :2394-2493
    PopDev &P = ctx->pop[pop];
    const int C = ctx->cfg.n_chr, nf = ctx->cfg.n_phen;
    const uint64_t cap = ctx->cfg.capacity, slots = cap * C * 2;
    GenState &par = P.st[P.cur], &off = P.st[P.cur ^ 1];
    std::vector<uint32_t> tmp;
    cudaStream_t st = ctx->stream;
    StepState *ss = P.d_ss;
    // the other draw set; the bulk stream may still read it for the generation before last
    P.dcur ^= 1;
    DrawSet &D = P.draws();
    GE_TRY(ctx->wait_bulk_done(D));
    if (dr) {
        const uint64_t n_off = dr->n_offspring, n_slots = n_off * C * 2;
        GE_TRY(upload_u64_as_u32(ctx, D.father, dr->father, n_off, tmp, "father"));
        GE_TRY(upload_u64_as_u32(ctx, D.mother, dr->mother, n_off, tmp, "mother"));
        P.n_xo = dr->xo_off[n_slots];
        if (P.n_xo > P.hs.xo_cap) {   // replayed lists may be longer than anything the map would draw
            P.hs.xo_cap = P.n_xo + P.n_xo / 4;
            for (DrawSet &Q : P.ds) { GE_TRY(ctx->ensure(Q.xo_bp, P.hs.xo_cap * 4)); if (ctx->bits()) GE_TRY(ctx->ensure(Q.flips, P.hs.xo_cap * 4)); }
        }
        GE_TRY(ctx->upload(D.xo_off, std::vector<uint64_t>(dr->xo_off, dr->xo_off + n_slots + 1)));
        GE_TRY(upload_u64_as_u32(ctx, D.xo_bp, dr->xo_bp, P.n_xo, tmp, "crossover position"));
        GE_TRY(ctx->upload(D.start_hap, std::vector<uint8_t>(dr->start_hap, dr->start_hap + n_slots)));
        CUDA_TRY(cudaMemcpyAsync(off.sex.p, dr->sex, n_off, cudaMemcpyHostToDevice, st));
        if (dr->mut_off) {
            P.n_mut = dr->mut_off[n_off * C];
            if (P.n_mut > P.hs.mut_cap) { P.hs.mut_cap = P.n_mut + P.n_mut / 4; GE_TRY(ctx->ensure(P.mut_bp, P.hs.mut_cap * 4)); GE_TRY(ctx->ensure(P.mut_gam, P.hs.mut_cap)); }
            GE_TRY(ctx->upload(P.mut_off, std::vector<uint64_t>(dr->mut_off, dr->mut_off + n_off * C + 1)));
            GE_TRY(upload_u64_as_u32(ctx, P.mut_bp, dr->mut_bp, P.n_mut, tmp, "mutation position"));
            GE_TRY(ctx->upload(P.mut_gam, std::vector<uint8_t>(dr->mut_gam, dr->mut_gam + P.n_mut)));
        } else P.n_mut = 0;
        CUDA_TRY(cudaMemsetAsync(off.C.p, 0, (size_t)cap * nf * 8, st));
        if (dr->common)
            for (int f = 0; f < nf; f++) CUDA_TRY(cudaMemcpyAsync(off.C.as<double>() + (uint64_t)f * cap, dr->common + (uint64_t)f * n_off, n_off * 8, cudaMemcpyHostToDevice, st));
        P.have_couple_of = false;
        // the host decided every size of this step
        P.hs.n_off = n_off; P.hs.n_xo = P.n_xo; P.hs.n_mut = P.n_mut; P.hs.n_iv = P.n_xo + n_slots; P.hs.n[P.cur ^ 1] = n_off;
        P.hs.dc[P.dcur] = DrawCounts{n_off, P.n_xo, P.n_xo + n_slots, 0u, 0u};
        off.n = n_off; P.n_off = n_off;   // (host mirrors: the replayed environment draws are copied by this count before the step's read-back)
        GE_TRY(ctx->push_state(P));
        pedigree_kernel<<<nblk(n_off, 256), 256, 0, st>>>(n_off, D.father.as<uint32_t>(), D.mother.as<uint32_t>(), par.ids.as<uint64_t>(), off.ids.as<uint64_t>());
        GE_TRY(ctx->check_launch("pedigree"));
    } else {
        ge_ctx::PhaseTimer timer(ctx, GE_PHASE_SAMPLE);
        // offspring offsets = exclusive scan of the family sizes of the couples that may marry (:2402-2406); the couple threads then
        // write parents, sex, pedigree and the sibling-common effect of their children
        const uint64_t cb = P.couples_cap;
        GE_TRY(ctx->scan_in(st, FamilyIn{ss, P.c_inbreed.as<uint8_t>(), P.c_noff.as<int32_t>()}, devn(&ss->n_couples), cb, P.mate.fam_off.as<uint64_t>(), OffspringTotal{ss}));
        CommonArgs ca;
        ca.n_phen = nf; ca.stride = cap;
        for (int f = 0; f < 8; f++) ca.sd[f] = (f < nf && P.scheme[f].vc > 0) ? std::sqrt(P.scheme[f].vc) : 0.0;
        offspring_kernel<<<ctx->grid_for(cb, 128), 128, 0, st>>>(ctx->rng, ss, pop, P.mate.fam_off.as<uint64_t>(), P.c_male.as<uint32_t>(), P.c_female.as<uint32_t>(), par.ids.as<uint64_t>(), ca,
                                                                 D.father.as<uint32_t>(), D.mother.as<uint32_t>(), D.couple_of.as<uint32_t>(), off.sex.as<uint8_t>(), off.ids.as<uint64_t>(),
                                                                 off.C.as<double>(), off.d_n);
        GE_TRY(ctx->check_launch("offspring"));
        P.have_couple_of = true;
        // crossovers: count + stash, scan, place
        // (its warps refill their lanes from a chunk of slots: about one resident wave, so that the chunks are long)
        sample_xo_kernel<<<std::min<unsigned>(ctx->ctrl_grid(slots, 128), (unsigned)(ctx->n_sm * ctx->xo_ctas_per_sm)), 128, 0, st>>>(ctx->rng, ctx->rmap(P), ss, C, pop, P.cnt32.as<uint32_t>(), D.start_hap.as<uint8_t>(), ctx->xo_stash.as<uint32_t>());
        GE_TRY(ctx->check_launch("sample_xo"));
        GE_TRY(ctx->scan(st, P.cnt32.as<uint32_t>(), devn(&ss->n_off, (uint64_t)C * 2), slots, D.xo_off.as<uint64_t>(), XoTotal{ss, (uint64_t)C * 2, P.dcur}));
        xo_place_kernel<<<ctx->ctrl_grid(slots, 128), 128, 0, st>>>(ctx->rng, ctx->rmap(P), ctx->genome(), ss, C, pop, D.xo_off.as<uint64_t>(), ctx->xo_stash.as<uint32_t>(),
                                                                    D.xo_bp.as<uint32_t>(), ctx->bits() ? D.flips.as<uint32_t>() : nullptr);
        GE_TRY(ctx->check_launch("xo_place"));
        if (P.has_mut) {
            const uint64_t items = cap * C;
            sample_mut_kernel<false><<<ctx->grid_for(items, 128), 128, 0, st>>>(ctx->rng, ctx->mmap(P), ss, C, pop, P.cnt32.as<uint32_t>(), nullptr, nullptr, nullptr);
            GE_TRY(ctx->check_launch("sample_mut<count>"));
            GE_TRY(ctx->scan(st, P.cnt32.as<uint32_t>(), devn(&ss->n_off, (uint64_t)C), items, P.mut_off.as<uint64_t>(), StoreTotal{&ss->n_mut, &ss->err, &ss->mut_cap, SE_CAP_MUT}));
            sample_mut_kernel<true><<<ctx->grid_for(items, 128), 128, 0, st>>>(ctx->rng, ctx->mmap(P), ss, C, pop, nullptr, P.mut_off.as<uint64_t>(), P.mut_bp.as<uint32_t>(), P.mut_gam.as<uint8_t>());
            GE_TRY(ctx->check_launch("sample_mut<fill>"));
        }
    }
    // ---- bit-packed propagation: the HBM-bound bulk of the generation, on the bulk stream.  Nothing later on the
    // control stream needs the rows (genetic values come from the causal-variant planes), so mating, sampling
    // and phenotypes of the NEXT generation overlap with this copy.
    bool bulk_launched = false;
    if (ctx->bits()) {
        if (dr) {  // replayed crossovers: positions -> locus indices (the Philox path did it in xo_place_kernel)
            const uint64_t n_slots = dr->n_offspring * C * 2;
            xo_to_flips_kernel<<<nblk(n_slots, 128), 128, 0, st>>>(ctx->genome(), n_slots, D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.flips.as<uint32_t>());
            GE_TRY(ctx->check_launch("xo_to_flips"));
        }
        // one short-lived CTA per offspring (control-stream kernels get SM slots quickly); CTAs beyond the device-side count exit at once
        const unsigned grid = (unsigned)std::min<uint64_t>(dr ? dr->n_offspring : cap, 1u << 20);
        const Genome gnm = ctx->genome();
        const TileTable tiles = ctx->tiles();
        const DrawCounts *dc = &ss->dc[P.dcur];
        const uint32_t *par_rows = par.hap.as<uint32_t>(), *rowmap = par.rowmap, *fa = D.father.as<uint32_t>(), *mo = D.mother.as<uint32_t>(), *fl = D.flips.as<uint32_t>();
        uint32_t *off_rows = off.hap.as<uint32_t>();
        const uint64_t *xo_off = D.xo_off.as<uint64_t>();
        const uint8_t *start = D.start_hap.as<uint8_t>();
        const size_t smem = prop_smem_bytes(C);
        const unsigned prop_threads = ctx->prop_threads;
        DrawSet *Dp = &D;
        const double bulk_bytes = (double)cap * ctx->W * 16.0;
        GE_TRY(ctx->to_bulk(P.ev_ready, [=]() -> int {
            cudaStream_t bulk = ctx->serial ? ctx->stream : ctx->bulk;
            ge_ctx::EvPair evp{nullptr, nullptr, GE_KERNEL_PROPAGATE_BITS, 0, 0, 0};
            if (ctx->profiling) { evp.a = ctx->get_event(); evp.b = ctx->get_event(); CUDA_TRY(cudaEventRecord(evp.a, bulk)); }
            propagate_bits_kernel<<<grid, prop_threads, smem, bulk>>>(gnm, tiles, dc, par_rows, rowmap, off_rows, fa, mo, xo_off, fl, start);
            GE_TRY(ctx->check_launch("propagate_bits"));
            if (ctx->profiling) {
                CUDA_TRY(cudaEventRecord(evp.b, bulk));
                evp.bytes_per_offspring = ctx->n_loci_total / 2;  // 0.5 byte per individual-locus; the offspring count comes with the step's read-back
                evp.pop = pop;
                ctx->ev_pending.push_back(evp);
            }
            CUDA_TRY(cudaEventRecord(Dp->bulk_done, bulk));
            Dp->bulk_pending = true;
            // thin control kernels only pay off while the bulk copy is longer than the control chain (~0.75 ms at 100k individuals)
            ctx->note_bulk(bulk_bytes);
            return GE_OK;
        }));
        bulk_launched = true;
    }
    // ---- causal-variant planes
    if (ctx->n_cv_tot) {  // in every representation: crossover parity at the CV positions, never a rescan of the segment lists
        uint64_t tot = cap * 2 * ctx->Wcv;
        const CvSet cvs = ctx->cvset();
        if (cvs.sorted && (size_t)ctx->Wcv * 4 * 10 <= 48 * 1024)   // a warp per gamete row, toggle words in shared memory (about one resident wave: the CTAs keep their tables)
            cv_propagate_rows_kernel<<<std::min<unsigned>(ctx->ctrl_grid(cap * 2 * 32, 256), (unsigned)ctx->n_sm * 8u), 256, (size_t)ctx->Wcv * 4 * 10, st>>>(cvs, ss, par.cv_allele.as<uint32_t>(), off.cv_allele.as<uint32_t>(), D.father.as<uint32_t>(),
                                                                                                         D.mother.as<uint32_t>(), D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.start_hap.as<uint8_t>());
        else   // unsorted cv.info rows, or more CV words per row than eight warps can hold in shared memory
            cv_propagate_bits_kernel<<<ctx->ctrl_grid(tot, 256), 256, 0, st>>>(cvs, ss, par.cv_allele.as<uint32_t>(), off.cv_allele.as<uint32_t>(), D.father.as<uint32_t>(),
                                                                     D.mother.as<uint32_t>(), D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.start_hap.as<uint8_t>());
        GE_TRY(ctx->check_launch("cv_propagate_bits"));
        if (ctx->use_root) {
            uint64_t tr = cap * 2 * ctx->n_cv_tot;
            cv_root_propagate_kernel<<<ctx->grid_for(tr, 256), 256, 0, st>>>(ctx->cvset(), ss, par.cv_root.as<uint8_t>(), off.cv_root.as<uint8_t>(), D.father.as<uint32_t>(), D.mother.as<uint32_t>(),
                                                                    D.xo_off.as<uint64_t>(), D.xo_bp.as<uint32_t>(), D.start_hap.as<uint8_t>());
            GE_TRY(ctx->check_launch("cv_root_propagate"));
        }
    }
    // ---- founder segments
    if (ctx->segs()) GE_TRY(seg_recombine(ctx, pop, dr ? dr->n_offspring : 0));
    // ---- mutation lists (inherit + this generation's hits)
    if (P.has_mut || par.has_hm) {
        MutArgs a;
        a.n_chr = C; a.ss = ss; a.father = D.father.as<uint32_t>(); a.mother = D.mother.as<uint32_t>();
        a.xo_off = D.xo_off.as<uint64_t>(); a.xo_bp = D.xo_bp.as<uint32_t>(); a.start_hap = D.start_hap.as<uint8_t>();
        a.par_hm_off = par.has_hm ? par.hm_off.as<uint64_t>() : nullptr; a.par_hm_bp = par.hm_bp.as<uint32_t>();
        bool hits = P.has_mut && (dr ? dr->mut_off != nullptr : true);
        a.mut_off = hits ? P.mut_off.as<uint64_t>() : nullptr; a.mut_bp = P.mut_bp.as<uint32_t>(); a.mut_gam = P.mut_gam.as<uint8_t>();
        a.cov_lo = P.d_cov_lo.as<uint32_t>(); a.cov_hi = P.d_cov_hi.as<uint32_t>();
        // room for the offspring lists: what the parents carry can at most double on the way down (both parental haplotypes inherited
        // by every child of a growing population), plus this generation's hits.  par.n_hm is exact: it came with the last read-back.
        const double growth = std::max(1.0, (double)cap / (double)std::max<uint64_t>(par.n, 1));
        const uint64_t need = (uint64_t)(2.5 * growth * (double)par.n_hm) + P.hs.mut_cap + 1024;
        GE_TRY(ctx->ensure(off.hm_off, (slots + 1) * 8));
        if (need * 4 > off.hm_bp.cap) GE_TRY(ctx->ensure(off.hm_bp, need * 4 + need));
        if (P.hs.hm_cap != off.hm_bp.cap / 4) { P.hs.hm_cap = off.hm_bp.cap / 4; GE_TRY(ctx->push_state(P, offsetof(StepState, hm_cap), 8)); }
        const unsigned g = ctx->grid_for(slots, 128);
        mutation_lists_kernel<false><<<g, 128, 0, st>>>(a, ctx->genome(), ctx->cvset(), P.cnt32.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr);
        GE_TRY(ctx->check_launch("mutation_lists<count>"));
        GE_TRY(ctx->scan(st, P.cnt32.as<uint32_t>(), devn(&ss->n_off, (uint64_t)C * 2), slots, off.hm_off.as<uint64_t>(), StoreTotal{off.d_n_hm, &ss->err, &ss->hm_cap, SE_CAP_HM}));
        if (bulk_launched) GE_TRY(ctx->join_bulk());  // the fill pass toggles bits of the freshly propagated rows
        mutation_lists_kernel<true><<<g, 128, 0, st>>>(a, ctx->genome(), ctx->cvset(), nullptr, off.hm_off.as<uint64_t>(), off.hm_bp.as<uint32_t>(),
                                                       ctx->bits() ? off.hap.as<uint32_t>() : nullptr, ctx->n_cv_tot ? off.cv_allele.as<uint32_t>() : nullptr);
        GE_TRY(ctx->check_launch("mutation_lists<fill>"));
        off.has_hm = true;
    } else off.has_hm = false;
    off.rowmap = nullptr;   // a fresh generation is written in identity order
    P.cur ^= 1;
    return GE_OK;
}
// a failed step leaves the population where it was: the parents' buffers are untouched (offspring go to the other set)
static void rollback_reproduce(ge_ctx *ctx, int pop) { PopDev &P = ctx->pop[pop]; P.cur ^= 1; P.dcur ^= 1; P.st[P.cur ^ 1].seg.valid = false; }

int ge_reproduce(ge_ctx *ctx, int pop, int gen, const ge_draws *dr) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    if (!ctx->gen0_done || P.st[P.cur].n == 0) return fail(GE_ERR_INVALID, "ge_reproduce before ge_init_generation0");
    if (dr) GE_TRY(validate_draws(ctx, P, dr, P.st[P.cur].n));
    else {
        if (ctx->cfg.rng_mode != GE_RNG_PHILOX) return fail(GE_ERR_INVALID, "replay mode needs draws");
        if (P.n_couples == 0) return fail(GE_ERR_INVALID, "no couples: call ge_mate or ge_set_couples first");
    }
    if (P.hs.gen != gen) { P.hs.gen = gen; GE_TRY(ctx->push_state(P, offsetof(StepState, gen), 4)); }
    GE_TRY(enqueue_reproduce(ctx, pop, dr));
    int rc = ctx->pull_state("ge_reproduce");   // (also: caller buffers were read asynchronously)
    if (rc != GE_OK) rollback_reproduce(ctx, pop);
    return rc;
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
    GE_TRY(ctx->pull_state("ge_do_migration"));   // the gather lists are built on the host from the population sizes
    GE_TRY(seg_finish_all(ctx));
    return migrate(ctx, gen, row);
}

// everything of one generation except the final read-back (sim_next_generation :1890-2082, the reference's order)
static int enqueue_generation(ge_ctx *ctx, int gen, const ge_gen_params *gp, const double *mig, const ge_draws *dr, int &reproduced, bool &migrated) {
    const int nf = ctx->cfg.n_phen, np = ctx->cfg.n_pop;
    // one phenotype, one population, no --gamma shift between the two: mating and selection values come out of the phenotype pass
    bool fused_mv_sv = nf == 1 && np == 1;
    for (double g : ctx->gamma) if (g != 0) fused_mv_sv = false;
    for (int p = 0; p < np; p++) {
        GE_TRY(enqueue_step_begin(ctx, p, gen, gp[p]));
        if (!dr) { ge_ctx::PhaseTimer timer(ctx, GE_PHASE_MATE); GE_TRY(enqueue_mate(ctx, p, gp[p])); }
        GE_TRY(enqueue_reproduce(ctx, p, dr ? &dr[p] : nullptr));
        reproduced = p + 1;
        GE_TRY(enqueue_AD(ctx, p));
        {
            ge_ctx::PhaseTimer timer(ctx, GE_PHASE_PHENOTYPE);
            const uint64_t n = dr ? dr[p].n_offspring : 0;
            for (int f = 0; f < nf; f++) GE_TRY(enqueue_GEF(ctx, p, f, false, (dr && dr[p].e_raw) ? dr[p].e_raw + (uint64_t)f * n : nullptr, nullptr, fused_mv_sv));
        }
    }
    for (int f = 0; f < nf; f++) GE_TRY(ge_environmental_effects_specific_to_each_population(ctx, f));
    if (!fused_mv_sv) for (int p = 0; p < np; p++) GE_TRY(enqueue_mv_sv(ctx, p));
    if (np > 1 && mig) {
        GE_TRY(ctx->pull_state("generation"));   // migration builds its gather lists on the host
        GE_TRY(seg_finish_all(ctx));
        GE_TRY(migrate(ctx, gen, mig));
        migrated = true;
    }
    for (int p = 0; p < np; p++) if (ctx->needs_prev) GE_TRY(enqueue_save_prev(ctx, p));
    return GE_OK;
}

// ---- the control chain of a generation as a CUDA graph ----
// With every size on the device, no kernel argument of the chain changes from one generation to the next except the generation-table
// row (step_begin_kernel's argument) and which of the two buffer sets is read or written (two graphs, by parity).  The launch-bound
// configurations (1 000 - 10 000 individuals: forty kernels of a few microseconds) then cost one graph launch, one node-parameter
// update and one read-back per generation.  The bulk-stream work (bit-packed copy, segment plan + gather) stays outside the graph —
// it overlaps the NEXT generation's chain, which a graph launched behind it could not — tied in by external event nodes.
static bool graphable(ge_ctx *ctx, const double *mig, const ge_draws *dr) {
    if (!ctx->use_graph || dr || mig || ctx->cfg.rng_mode != GE_RNG_PHILOX || ctx->phase_timing || ctx->allreduce /* a host hook cannot be captured; ge_set_allreduce_nccl can */ || ctx->serial || ctx->cfg.n_pop != 1) return false;
    for (double g : ctx->gamma) if (g != 0) return false;
    if (ctx->segs() && (ctx->cfg.seg_capacity == 0 || ctx->cv_from_segments)) return false;   // the host sizes the segment buffer between the passes
    for (PopDev &P : ctx->pop) if (P.has_mut || P.st[P.cur].has_hm || P.st[P.cur].rowmap) return false;   // the mutation pass waits for the bulk copy mid-chain
    return true;
}
static void replay_host_state(ge_ctx *ctx, int p, int gen, const ge_gen_params &gp) {   // what enqueue_* change on the host
    PopDev &P = ctx->pop[p];
    host_row(P, step_row(gen, gp));
    P.dcur ^= 1;
    P.have_couple_of = true; P.have_e_raw = true;
    GenState &off = P.st[P.cur ^ 1];
    off.has_hm = false; off.rowmap = nullptr;
    if (ctx->segs()) off.seg.valid = true;
    P.cur ^= 1;
}
static int step_with_graph(ge_ctx *ctx, int gen, const ge_gen_params *gp, int &reproduced) {
    const int np = ctx->cfg.n_pop;
    std::string key;
    for (int p = 0; p < np; p++) {
        PopDev &P = ctx->pop[p];
        key += (char)('0' + P.cur + 2 * P.dcur + 4 * (P.RM ? 1 : 0) + 8 * ((gp[p].offspring_dist == 'p' || gp[p].offspring_dist == 'P') ? 1 : 0));
    }
    key += ctx->bulk_busy ? 'b' : 'i';
    key += (char)('A' + ctx->thin_now);
    ge_ctx::StepGraph &G = ctx->graphs[key];
    if (G.epoch != ctx->graph_epoch) {   // a buffer moved since this graph was recorded (or warmed): its nodes hold stale pointers
        if (G.exec) cudaGraphExecDestroy(G.exec);
        if (G.graph) cudaGraphDestroy(G.graph);
        G = ge_ctx::StepGraph();
        G.epoch = ctx->graph_epoch;
    }
    bool dummy = false;
    // The first generation of a context is queued kernel by kernel (it may still allocate), and so is the first pass of a key if no
    // such generation has completed since a buffer last moved; after that a key is recorded the first time it comes up, so that the
    // two buffer parities are graphs from the third generation on (a bench with three warm-up steps times replays only).
    const bool settled = ctx->plain_epoch == ctx->graph_epoch && ctx->plain_gens >= 1;
    if (!G.exec && (G.warm < 0 || (G.warm < 1 && !settled))) {
        int rc = enqueue_generation(ctx, gen, gp, nullptr, nullptr, reproduced, dummy);
        if (rc == GE_OK && ctx->graph_epoch == G.epoch) {
            G.warm++;
            if (ctx->plain_epoch != ctx->graph_epoch) { ctx->plain_epoch = ctx->graph_epoch; ctx->plain_gens = 0; }
            ctx->plain_gens++;
        }
        return rc;
    }
    if (!G.exec) {   // record
        const uint64_t launches0 = ctx->launches;
        struct Saved { int cur, dcur; } saved[16];
        for (int p = 0; p < np; p++) saved[p] = {ctx->pop[p].cur, ctx->pop[p].dcur};
        ctx->deferred.clear();
        ctx->capturing = true;
        cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
        int rc = e == cudaSuccess ? enqueue_generation(ctx, gen, gp, nullptr, nullptr, reproduced, dummy) : GE_ERR_CUDA;
        cudaGraph_t g = nullptr;
        if (e == cudaSuccess) e = cudaStreamEndCapture(ctx->stream, &g);
        ctx->capturing = false;
        if (rc == GE_OK && e == cudaSuccess && g) e = cudaGraphInstantiate(&G.exec, g, 0);
        if (rc != GE_OK || e != cudaSuccess || !G.exec) {   // not capturable (yet): back to plain launches — two more plain passes, then for good
            cudaGetLastError();
            if (g) cudaGraphDestroy(g);
            G.exec = nullptr; G.warm = ++G.fails >= 3 ? -1000000 : -2;
            ctx->deferred.clear();
            ctx->launches = launches0;
            for (int p = 0; p < np; p++) { ctx->pop[p].cur = saved[p].cur; ctx->pop[p].dcur = saved[p].dcur; }
            reproduced = 0;
            return enqueue_generation(ctx, gen, gp, nullptr, nullptr, reproduced, dummy);
        }
        G.graph = g;
        G.bulk = std::move(ctx->deferred);
        ctx->deferred.clear();
        G.launches = ctx->launches - launches0;
        size_t n_nodes = 0;
        CUDA_TRY(cudaGraphGetNodes(g, nullptr, &n_nodes));
        std::vector<cudaGraphNode_t> nodes(n_nodes);
        CUDA_TRY(cudaGraphGetNodes(g, nodes.data(), &n_nodes));
        G.begin_nodes.assign(np, nullptr);
        for (cudaGraphNode_t nd : nodes) {
            cudaGraphNodeType ty;
            CUDA_TRY(cudaGraphNodeGetType(nd, &ty));
            if (ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp{};
            if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) { cudaGetLastError(); continue; }   // another library's kernel (the NCCL collective): not ours to patch
            if (kp.func != (void *)step_begin_kernel) continue;
            StepState *target = *static_cast<StepState **>(kp.kernelParams[0]);
            for (int p = 0; p < np; p++) if (ctx->pop[p].d_ss == target) G.begin_nodes[p] = nd;
        }
        for (int p = 0; p < np; p++) if (!G.begin_nodes[p]) return fail(GE_ERR_CUDA, "captured generation graph has no step_begin node");
        // the host state advanced while recording; the graph itself has not run yet
    } else {         // replay: the one argument that changes, then the host-side bookkeeping enqueue_* would have done
        for (int p = 0; p < np; p++) {
            StepState *d = ctx->pop[p].d_ss;
            StepRow r = step_row(gen, gp[p]);
            void *args[2] = {&d, &r};
            cudaKernelNodeParams kp{};
            kp.func = (void *)step_begin_kernel; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
            CUDA_TRY(cudaGraphExecKernelNodeSetParams(G.exec, G.begin_nodes[p], &kp));
            replay_host_state(ctx, p, gen, gp[p]);
            reproduced = p + 1;
        }
        ctx->launches += G.launches;
        ctx->graph_replays++;
    }
    CUDA_TRY(cudaGraphLaunch(G.exec, ctx->stream));
    for (auto &fn : G.bulk) GE_TRY(fn());
    return GE_OK;
}

int ge_step_generation(ge_ctx *ctx, int gen, const ge_gen_params *gp, const double *mig, const ge_draws *dr) {  // sim_next_generation :1890-2082
    CHECK_CTX(ctx);
    if (!gp) return fail(GE_ERR_INVALID, "null params");
    if (!ctx->gen0_done) return fail(GE_ERR_INVALID, "ge_step_generation before ge_init_generation0");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    const int np = ctx->cfg.n_pop;
    if (dr) for (int p = 0; p < np; p++) GE_TRY(validate_draws(ctx, ctx->pop[p], &dr[p], ctx->pop[p].st[ctx->pop[p].cur].n));
    else if (ctx->cfg.rng_mode != GE_RNG_PHILOX) return fail(GE_ERR_INVALID, "replay mode needs draws");
    int reproduced = 0;
    bool migrated = false;
    int rc = graphable(ctx, mig, dr) ? step_with_graph(ctx, gen, gp, reproduced) : enqueue_generation(ctx, gen, gp, mig, dr, reproduced, migrated);
    if (rc == GE_OK) rc = ctx->pull_state("generation");   // THE host synchronisation of the generation: sizes for the host's bookkeeping, errors
    if (rc != GE_OK && !migrated) for (int p = 0; p < reproduced; p++) rollback_reproduce(ctx, p);
    return rc;
}

// ---------------- results ----------------

int ge_get_population_size(ge_ctx *ctx, int pop, uint64_t *n) { CHECK_POP(ctx, pop); *n = ctx->pop[pop].st[ctx->pop[pop].cur].n; return GE_OK; }

int ge_download_individuals(ge_ctx *ctx, int pop, ge_indiv_soa *o) {
    CHECK_POP(ctx, pop);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    uint64_t n = S.n; int nf = ctx->cfg.n_phen;
    const uint64_t cap = ctx->cfg.capacity;
    cudaStream_t st = ctx->stream;
    auto cp = [&](void *dst, const Buf &src, size_t bytes) -> cudaError_t { return dst ? cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
    // per-phenotype columns: [n_phen][n] for the caller, rows `capacity` apart on the device
    auto cols = [&](double *dst, const Buf &src) -> cudaError_t {
        return (dst && n) ? cudaMemcpy2DAsync(dst, n * 8, src.p, cap * 8, n * 8, (size_t)nf, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    };
    CUDA_TRY(cp(o->ids, S.ids, n * 56)); CUDA_TRY(cp(o->sex, S.sex, n));
    CUDA_TRY(cols(o->A, S.A)); CUDA_TRY(cols(o->D, S.D)); CUDA_TRY(cols(o->G, S.G)); CUDA_TRY(cols(o->C, S.C));
    CUDA_TRY(cols(o->E, S.E)); CUDA_TRY(cols(o->F, S.F)); CUDA_TRY(cols(o->P, S.P));
    CUDA_TRY(cp(o->mv, S.mv, n * 8)); CUDA_TRY(cp(o->sv, S.sv, n * 8)); CUDA_TRY(cp(o->svf, S.svf, n * 8));
    CUDA_TRY(cudaStreamSynchronize(st));
    return GE_OK;
}

int ge_get_moments(ge_ctx *ctx, int pop, int f, ge_moments *m) {
    CHECK_POP(ctx, pop); CHECK_PHEN(ctx, f);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    uint64_t o = (uint64_t)f * ctx->cfg.capacity;
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
    ge_ctx::TmpScope scratch_only(ctx);
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
int ge_download_haplotypes_from_segments(ge_ctx *ctx, int pop, int c, uint8_t *al) {   // ras_convert_interval_to_hap_matrix :1186-1230, whatever else the context carries
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    ge_ctx::TmpScope scratch_only(ctx);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_download_haplotypes_from_segments needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(seg_finish_all(ctx));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    uint64_t tot = (uint64_t)2 * S.n * ctx->chr_nloci[c];
    if (tot == 0) return GE_OK;
    Buf tmp;
    GE_TRY(ctx->ensure_exact(tmp, tot));
    GE_TRY(seg_materialise(ctx, pop, c, tmp.as<uint8_t>()));
    CUDA_TRY(cudaMemcpyAsync(al, tmp.p, tot, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tmp);
    return GE_OK;
}
int ge_download_haplotypes_packed(ge_ctx *ctx, int pop, int c, uint32_t *words) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    ge_ctx::TmpScope scratch_only(ctx);
    if (!ctx->bits() && ctx->seg_per_thread) return fail(GE_ERR_UNSUPPORTED, "packed download from segment lists that need not be sorted: use ge_download_haplotypes");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GenState &S = ctx->pop[pop].st[ctx->pop[pop].cur];
    uint32_t nw = (ctx->chr_nloci[c] + 31) / 32;
    uint64_t tot = (uint64_t)2 * S.n * nw;
    if (tot == 0) return GE_OK;
    Buf tmp;
    GE_TRY(ctx->ensure_exact(tmp, tot * 4));
    if (ctx->bits()) {
        GE_TRY(ctx->join_bulk());
        gather_packed_chr_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(S.hap.as<uint32_t>(), S.rowmap, ctx->W, ctx->chr_word_off[c], (uint32_t)(2 * S.n), nw,
                                                                          tmp.as<uint32_t>());
        GE_TRY(ctx->check_launch("gather_packed"));
    } else GE_TRY(seg_materialise_packed(ctx, pop, c, tmp.as<uint32_t>()));
    CUDA_TRY(cudaMemcpyAsync(words, tmp.p, tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tmp);
    return GE_OK;
}

int ge_get_segment_count(ge_ctx *ctx, int pop, int c, uint64_t *ns, uint64_t *nm) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    ge_ctx::TmpScope scratch_only(ctx);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_count(ctx, pop, c, ns, nm);
}
int ge_download_segments(ge_ctx *ctx, int pop, int c, uint64_t *off, uint64_t *seg, uint64_t *moff, uint64_t *mbp) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    ge_ctx::TmpScope scratch_only(ctx);
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_download(ctx, pop, c, off, seg, moff, mbp);
}

int ge_download_cv_alleles(ge_ctx *ctx, int pop, int f, int c, uint8_t *out) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c); CHECK_PHEN(ctx, f);
    ge_ctx::TmpScope scratch_only(ctx);
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

int ge_rebase_founders(ge_ctx *ctx, int keep_history) {
    CHECK_CTX(ctx);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_rebase_founders needs GE_REP_SEGMENTS");
    if (!ctx->gen0_done) return fail(GE_ERR_INVALID, "ge_rebase_founders before ge_init_generation0");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_rebase(ctx, keep_history);
}
int ge_get_segment_count_gen0(ge_ctx *ctx, int pop, int c, uint64_t *ns) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    ge_ctx::TmpScope scratch_only(ctx);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_get_segment_count_gen0 needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    GE_TRY(seg_download_gen0(ctx, pop, c, ctx->gen0_off, ctx->gen0_seg));
    ctx->gen0_pop = pop; ctx->gen0_chr = c;
    *ns = ctx->gen0_seg.size() / 4;
    return GE_OK;
}
int ge_download_segments_gen0(ge_ctx *ctx, int pop, int c, uint64_t *off, uint64_t *seg) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, c);
    ge_ctx::TmpScope scratch_only(ctx);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_download_segments_gen0 needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    if (ctx->gen0_pop != pop || ctx->gen0_chr != c) GE_TRY(seg_download_gen0(ctx, pop, c, ctx->gen0_off, ctx->gen0_seg));   // (ge_get_segment_count_gen0 leaves the composed lists here)
    std::memcpy(off, ctx->gen0_off.data(), ctx->gen0_off.size() * 8);
    if (!ctx->gen0_seg.empty()) std::memcpy(seg, ctx->gen0_seg.data(), ctx->gen0_seg.size() * 8);
    ctx->gen0_pop = ctx->gen0_chr = -1;
    std::vector<uint64_t>().swap(ctx->gen0_off); std::vector<uint64_t>().swap(ctx->gen0_seg);
    return GE_OK;
}

int ge_recompute_cv_from_segments(ge_ctx *ctx, int pop) {  // ras_find_cv :2752-2815 literally: scan the parts of every haplotype
    CHECK_POP(ctx, pop);
    ge_ctx::TmpScope scratch_only(ctx);
    if (!ctx->segs()) return fail(GE_ERR_UNSUPPORTED, "ge_recompute_cv_from_segments needs GE_REP_SEGMENTS");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    return seg_find_cv(ctx, pop);
}

int ge_ibd_sharing(ge_ctx *ctx, int pop, int chr, const uint64_t *ind_a, const uint64_t *ind_b, uint64_t n_pairs, uint64_t min_bp, uint64_t *shared_bp, uint32_t *n_runs) {
    CHECK_POP(ctx, pop); CHECK_CHR(ctx, chr);
    ge_ctx::TmpScope scratch_only(ctx);
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
int ge_set_profiling(ge_ctx *ctx, int level) { CHECK_CTX(ctx); ctx->profiling = level != 0; ctx->phase_timing = level >= 2; return GE_OK; }
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
int ge_get_graph_replays(ge_ctx *ctx, uint64_t *n) { CHECK_CTX(ctx); *n = ctx->graph_replays; return GE_OK; }
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
