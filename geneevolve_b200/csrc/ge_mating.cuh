// ge_mating.cuh — parent-pair selection on the device: random_mate (:2090-2157) and assort_mate (:2167-2360)
// under the Philox streams of DESIGN.md §RNG, and migration (:877-989).
//
// The reference has no fitness-weighted roulette wheel: selection is Bernoulli thinning of who may mate
// (U_i < selection_value_func_i) followed by uniform index draws (random mating) or by pairing the males and
// females, both sorted by mating value, through the ranks of a bivariate-normal template (assortative).
// So the scans here are stream compaction and offspring-offset prefix sums.
//
// Nothing in the chain comes back to the host: list lengths, the couple count, the number of inbred couples and the
// Poisson mean live in StepState (ge_kernels.cuh), every kernel loops up to the device-side count on a grid sized from
// the capacity, and which sex list gets trimmed is decided on the device.  Sorting (mating values, template values,
// trim keys) is a stable tile sort + rank-by-search merges written here; their stability fixes the tie order the oracle
// also uses.
#pragma once
#include "ge_context.cuh"


namespace gek {

__device__ __forceinline__ uint64_t key64(const uint32_t w[4]) { return ((uint64_t)w[0] << 32) | w[1]; }
__device__ __forceinline__ uint64_t sortable(double x) {  // monotone map double -> uint64
    if (x == 0.0) x = 0.0;  // -0.0 and +0.0 compare equal
    uint64_t b = (uint64_t)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// thinning (:2105-2117 / :2186-2216): number of list entries individual i contributes — males in the low, females in
// the high 32 bits, so ONE scan gives the offsets into both lists
struct ThinIn {   // scanned as it is drawn; replayed draws (ge_mate_replay): thin_u / mm_u are the reference's own uniforms instead of the Philox ones
    Stream st; const StepState *ss; int pop; const uint8_t *sex; const double *svf; int with_mm; double mm; const double *thin_u, *mm_u;
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const {
        double r, r2;
        if (thin_u) { r = thin_u[i]; r2 = mm_u[i]; }
        else {
            uint32_t w[4];
            draw(st, P_THIN, pop, ss->gen, i, 0, 0, w);
            r = u01(w[0], w[1]); r2 = u01(w[2], w[3]);
        }
        uint64_t c = 0;
        if (r < svf[i]) c = (with_mm && r2 < mm) ? 2u : 1u;
        return sex[i] == 1 ? c : (sex[i] == 2 ? c << 32 : 0ull);
    }
};
// grand total of the thinning scan: list lengths, the couple count and the errors of :2125-2129 / :2226-2230
struct ThinTotal {
    StepState *ss; int random_mating;
    __device__ __forceinline__ void operator()(uint64_t t) const {
        const uint64_t n_m = t & 0xFFFFFFFFull, n_f = t >> 32;
        ss->n_m = n_m; ss->n_f = n_f;
        ss->n2 = n_m < n_f ? n_m : n_f;
        ss->trim_which = n_f > n_m ? 1u : 0u;
        ss->n_trim = n_m > n_f ? n_m - n_f : n_f - n_m;
        ss->n_trim_list = (random_mating || n_m == n_f) ? 0 : (n_m > n_f ? n_m : n_f);   // what the trim's sort and scan run over
        if (random_mating) {
            ss->n_couples = ss->pop_size;
            if (n_m == 0 || n_f == 0) atomicOr(&ss->err, (uint32_t)SE_NO_MATES_RM);
            if (ss->pop_size > ss->couples_cap) atomicOr(&ss->err, (uint32_t)SE_CAP_COUPLES);
        } else {
            ss->n_couples = ss->n2;
            if (ss->n2 == 0) atomicOr(&ss->err, (uint32_t)SE_NO_COUPLES);
        }
    }
};
__global__ void thin_fill_kernel(const uint64_t *__restrict__ n_ind, const uint64_t *__restrict__ off, uint32_t *__restrict__ list_m, uint32_t *__restrict__ list_f) {
    const uint64_t n = *n_ind;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = off[i], b = off[i + 1];
        for (uint32_t k = (uint32_t)a; k < (uint32_t)b; k++) list_m[k] = (uint32_t)i;
        for (uint32_t k = (uint32_t)(a >> 32); k < (uint32_t)(b >> 32); k++) list_f[k] = (uint32_t)i;
    }
}
__global__ void rm_pair_kernel(Stream st, const StepState *__restrict__ ss, int pop, const uint32_t *__restrict__ list_m, const uint32_t *__restrict__ list_f,
                               const uint32_t *__restrict__ idx_m /* replayed index draws, or null */, const uint32_t *__restrict__ idx_f,
                               uint32_t *__restrict__ male, uint32_t *__restrict__ female, uint8_t *__restrict__ inbreed, int32_t *__restrict__ noff) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n_couples = ss->n_couples, n_m = ss->n_m, n_f = ss->n_f;
    const int gen = ss->gen;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_couples; k += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t im, jf;
        if (idx_m) { im = min((uint64_t)idx_m[k], n_m - 1); jf = min((uint64_t)idx_f[k], n_f - 1); }
        else {
            uint32_t w[4];
            draw(st, P_RM_PAIR, pop, gen, k, 0, 0, w);
            im = ((uint64_t)w[0] * n_m) >> 32; jf = ((uint64_t)w[1] * n_f) >> 32;
        }
        male[k] = list_m[im];
        female[k] = list_f[jf];
        inbreed[k] = 0; noff[k] = 1;
    }
}
// trim of the longer sex list (:2233-2246): the n_trim entries with the smallest (Philox key, position) leave, survivors keep
// their order.  Keys of the longer list (nothing when the lists are equal) ...
__global__ void trim_keys_kernel(Stream st, const StepState *__restrict__ ss, int pop, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    const uint64_t n = ss->n_trim_list;
    const uint32_t sub = ss->trim_which;
    const int gen = ss->gen;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[4];
        draw(st, P_TRIM, pop, gen, k, sub, 0, w);
        keys[k] = key64(w); idx[k] = (uint32_t)k;
    }
}
// ... and, after the stable sort, who stays: flag[position] = 0 for the first n_trim of the sorted order
__global__ void trim_flag_kernel(const StepState *__restrict__ ss, const uint32_t *__restrict__ sorted_idx, uint32_t *__restrict__ flag) {
    const uint64_t n = ss->n_trim_list, n_trim = ss->n_trim;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) flag[sorted_idx[k]] = k >= n_trim;
}

// sort keys of one sex list by mating value (:2251-2252); the trimmed list is compacted on the way (survivors keep their order)
__global__ void mv_keys_kernel(const StepState *__restrict__ ss, uint32_t which, const uint32_t *__restrict__ list, const uint32_t *__restrict__ flag,
                               const uint64_t *__restrict__ keep_off, const double *__restrict__ mv, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = which ? ss->n_f : ss->n_m;
    const bool trimmed = ss->n_trim != 0 && ss->trim_which == which;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t o = k;
        if (trimmed) { if (!flag[k]) continue; o = keep_off[k]; }
        const uint32_t ind = list[k];
        keys[o] = sortable(mv[ind]); vals[o] = ind;
    }
}
// bivariate-normal template (ras_mvnorm, src/RasRandomNumber.cpp:15-53, with U = [[1,rho],[0,sqrt(1-rho^2)]])
__global__ void template_kernel(Stream st, const StepState *__restrict__ ss, int pop, uint64_t *__restrict__ k1, uint64_t *__restrict__ k2, uint32_t *__restrict__ idx) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = ss->n2;
    const int gen = ss->gen;
    const double rho = ss->mat_cor, u11 = ss->u11;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double z0, z1;
        normal2(st, P_TEMPLATE, pop, gen, i, 0, z0, z1);
        k1[i] = sortable(z0);
        k2[i] = sortable(z0 * rho + z1 * u11);
        idx[i] = (uint32_t)i;
    }
}
// the same from replayed template values (ras_mvnorm's two columns as the reference drew them)
__global__ void template_from_kernel(const StepState *__restrict__ ss, const double *__restrict__ t1, const double *__restrict__ t2, uint64_t *__restrict__ k1,
                                     uint64_t *__restrict__ k2, uint32_t *__restrict__ idx) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = ss->n2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        k1[i] = sortable(t1[i]); k2[i] = sortable(t2[i]); idx[i] = (uint32_t)i;
    }
}
__global__ void rank_scatter_kernel(const StepState *__restrict__ ss, const uint32_t *__restrict__ sorted_idx, uint32_t *__restrict__ rank) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = ss->n2;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (uint64_t)gridDim.x * blockDim.x) rank[sorted_idx[k]] = (uint32_t)k;
}
// couple i = (male at rank(t1_i), female at rank(t2_i)) and the sib/cousin exclusion (:2296-2320)
__global__ void pair_kernel(StepState *__restrict__ ss, const uint32_t *__restrict__ males, const uint32_t *__restrict__ females, const uint32_t *__restrict__ r1,
                            const uint32_t *__restrict__ r2, const uint64_t *__restrict__ ids, int avoid_inbreeding, uint32_t *__restrict__ male,
                            uint32_t *__restrict__ female, uint8_t *__restrict__ inbreed) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = ss->n2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t pm = males[r1[i]], pf = females[r2[i]];
        male[i] = pm; female[i] = pf;
        uint8_t ib = 0;
        if (avoid_inbreeding) {
            const uint64_t *a = ids + (uint64_t)pm * 7, *b = ids + (uint64_t)pf * 7;
            bool sib = a[1] == b[1];
            bool cousin = (a[3] == b[3] || a[3] == b[5] || a[5] == b[3] || a[5] == b[5] || a[4] == b[4] || a[4] == b[6] || a[6] == b[4] || a[6] == b[6]);
            ib = sib || cousin;
            if (ib) atomicAdd((unsigned long long *)&ss->n_inbreed, 1ull);   // rare
        }
        inbreed[i] = ib;
    }
}
// family sizes.  Poisson (:2329-2337): exact Poisson(lam), lam = pop_size / marriageable couples, as a sum of independent
// Poisson(<=32) chunks, each by sequential-search inversion (ras_rpois, src/RasRandomNumber.cpp:57-67).  Fixed (:2338-2355): floor.
// poisson: 1 = draw, 0 = fixed size, 2 = the sizes are already in noff (replayed draws): only the all-inbred check
__global__ void family_kernel(Stream st, StepState *__restrict__ ss, int pop, int poisson, int32_t *__restrict__ noff) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = ss->n2, n_ok = n - ss->n_inbreed;
    if (n_ok == 0) {   // every couple is inbred
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ss->err, (uint32_t)SE_ALL_INBRED);
        return;
    }
    if (poisson == 2) return;
    const int gen = ss->gen;
    const double lam = (double)ss->pop_size / (double)n_ok;
    const int32_t nfix = (int32_t)floor((double)ss->pop_size / (double)n_ok);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (!poisson) { noff[i] = nfix; continue; }
        int total = 0; uint32_t blk = 0; double rem = lam;
        while (rem > 0) {
            double l = rem > 32.0 ? 32.0 : rem;
            rem -= l;
            uint32_t w[4];
            draw(st, P_POISSON, pop, gen, i, 0, blk++, w);
            double u = u01(w[0], w[1]);
            double pk = exp(-l), F = pk; int k = 0;
            while (u >= F && k < 400) { k++; pk *= l / (double)k; F += pk; }
            total += k;
        }
        noff[i] = total;
    }
}
// fixed family size: the remainder goes to distinct random couples that may marry (the reference indexes an empty list here when
// --avoid_inbreeding is on, :2314-2353; this is its evident intent): couples by Philox key, inbred ones last ...
__global__ void remainder_keys_kernel(Stream st, const StepState *__restrict__ ss, int pop, const uint8_t *__restrict__ inbreed, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n = ss->n2;
    const int gen = ss->gen;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w[4];
        draw(st, P_REMAINDER, pop, gen, i, 0, 0, w);
        keys[i] = inbreed[i] ? 0xFFFFFFFFFFFFFFFFull : key64(w);
        idx[i] = (uint32_t)i;
    }
}
// ... and the first `remainder` of the sorted order get one more child
__global__ void remainder_add_kernel(const StepState *__restrict__ ss, const uint32_t *__restrict__ sorted_idx, const uint8_t *__restrict__ inbreed, int32_t *__restrict__ noff) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n_ok = ss->n2 - ss->n_inbreed;
    if (n_ok == 0) return;
    const uint64_t nfix = (uint64_t)floor((double)ss->pop_size / (double)n_ok);
    const uint64_t remain = ss->pop_size - nfix * n_ok;
    const uint64_t n_add = remain < n_ok ? remain : n_ok;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_add; k += (uint64_t)gridDim.x * blockDim.x)
        if (!inbreed[sorted_idx[k]]) noff[sorted_idx[k]]++;
}

}  // namespace gek

// ------------------------------------------------------------------------------------------------
// Stable sorts of (uint64 key, uint32 value) pairs with the element count on the device.  The mating chain sorts <= N/2 pairs
// five times per generation and sits on the dependency cycle mating -> draws -> CV planes -> genetic values -> phenotypes ->
// selection -> mating: a library radix sort is ten dependent launches and wants its count on the host.
// Here: tiles of 4096 pairs are sorted by one CTA each, then merged by rank — every element finds its final position by binary
// searches of the other runs of its group (up to 32 runs per level): rank = position in its run + #(<= key) in earlier runs +
// #(< key) in later runs, which keeps the sort stable.  One launch up to 4096 pairs, two up to 131 072, three up to 4 194 304, ...
// The tile sort is a bitonic network on (key, position in the tile) in shared memory — the position makes the order total, so
// the network's instability cannot show; values are gathered by position at the end.  It replaces cub::BlockRadixSort on tiles
// of the same size (sixteen 4-bit passes over 64-bit keys): 78 compare-exchange steps, most of them warp-local.
// ------------------------------------------------------------------------------------------------
namespace gek {
constexpr uint32_t SS_GROUP = 32;

// (SS_THREADS, SS_TILE) = (1024, 4096), or (512, 2048) for lists of at most four such tiles: a 2048-network is done in 7 us instead of
// 16 (config 2, 5 700 pairs per list: 0.408 against 0.443 ms per generation), but twice the runs cost the merge of a long list more
// than that — its dependent probes go to L2 beside the bulk copy (config 3, 52 000 pairs: mating phase 0.48 against 0.44 ms)
template <int SS_THREADS, int SS_TILE>
__global__ void __launch_bounds__(SS_THREADS) small_sort_tile_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                                     uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, DevN dn) {
    __shared__ uint64_t sk[SS_TILE];
    __shared__ uint16_t sp[SS_TILE];
    const uint32_t n = (uint32_t)dn.get();
    const uint32_t base = blockIdx.x * SS_TILE;
    if (base >= n) return;
    const uint32_t cnt = min((uint32_t)SS_TILE, n - base);
    uint32_t m = 64;   // the network's size: the power of two that holds the tile's pairs (padding sorts last: all-ones key, later position)
    while (m < cnt) m <<= 1;
    for (uint32_t t = threadIdx.x; t < m; t += SS_THREADS) { sk[t] = t < cnt ? kin[base + t] : ~0ull; sp[t] = (uint16_t)t; }
    __syncthreads();
    for (uint32_t k = 2; k <= m; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t p = threadIdx.x; p < (m >> 1); p += SS_THREADS) {
                const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), l = i | j;
                const uint64_t a = sk[i], b = sk[l];
                const uint16_t pa = sp[i], pb = sp[l];
                const bool gt = a > b || (a == b && pa > pb);
                if (gt == ((i & k) == 0)) { sk[i] = b; sk[l] = a; sp[i] = pb; sp[l] = pa; }
            }
            // pairs at distance j <= 32 stay inside the 64 elements one warp's 32 pair indices cover: a warp barrier orders them; the
            // step after this one decides (the first step of the next phase is at distance k)
            const uint32_t next_j = j > 1 ? (j >> 1) : k;
            if (j > 32 || next_j > 32) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < cnt; t += SS_THREADS) { kout[base + t] = sk[t]; vout[base + t] = vin[base + sp[t]]; }
}

// sorted runs of run_len pairs -> sorted runs of run_len * SS_GROUP pairs
__global__ void small_sort_merge_kernel(const uint64_t *__restrict__ tk, const uint32_t *__restrict__ tv, uint64_t *__restrict__ kout,
                                        uint32_t *__restrict__ vout, DevN dn, uint32_t run_len) {
    const uint32_t n = (uint32_t)dn.get();
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint64_t key = tk[e];
        const uint32_t mine = e / run_len, n_runs = (n + run_len - 1) / run_len;
        const uint32_t g0 = mine / SS_GROUP * SS_GROUP, g1 = min(g0 + SS_GROUP, n_runs);
        uint32_t rank = e - mine * run_len;
        for (uint32_t b = g0; b < g1; b++) {
            if (b == mine) continue;
            const uint64_t *t = tk + (uint64_t)b * run_len;
            uint32_t lo = 0, hi = min(run_len, n - b * run_len);
            if (b < mine) { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(t + mid) <= key) lo = mid + 1; else hi = mid; } }
            else { while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(t + mid) < key) lo = mid + 1; else hi = mid; } }
            rank += lo;
        }
        kout[(uint64_t)g0 * run_len + rank] = key;
        vout[(uint64_t)g0 * run_len + rank] = tv[e];
    }
}
}  // namespace gek

constexpr uint64_t SORT_MAX = 1ull << 31;   // element indices are 32 bits in the kernels above

// keys_in/vals_in -> keys_out/vals_out; n on the device, n_bound (host) sizes the grids and the scratch
static int sort_pairs_on(ge_ctx *ctx, cudaStream_t st, Buf &tmp, const uint64_t *kin, uint64_t *kout, const uint32_t *vin, uint32_t *vout, DevN n, uint64_t n_bound) {
    if (n_bound == 0) return GE_OK;
    if (n_bound > SORT_MAX) return fail(GE_ERR_UNSUPPORTED, "more than 2^31 list entries to sort");
    n = limited(n, n_bound);
    const bool small = n_bound <= 4 * 2048;
    const uint64_t SS_TILE = small ? 2048 : 4096;
    const unsigned tiles = nblk(n_bound, SS_TILE);
    auto tile_sort = [&](uint64_t *ko, uint32_t *vo) {
        if (small) small_sort_tile_kernel<512, 2048><<<tiles, 512, 0, st>>>(kin, vin, ko, vo, n);
        else small_sort_tile_kernel<1024, 4096><<<tiles, 1024, 0, st>>>(kin, vin, ko, vo, n);
        return ctx->check_launch("small_sort_tile");
    };
    if (tiles == 1) return tile_sort(kout, vout);
    int levels = 0;   // merge levels: runs of SS_TILE * 32^levels pairs cover the bound
    for (uint64_t run = SS_TILE; run < n_bound; run *= SS_GROUP) levels++;
    GE_TRY(ctx->ensure(tmp, 2 * (n_bound * 12 + 16)));
    uint64_t *ak = tmp.as<uint64_t>(), *bk = ak + n_bound;
    uint32_t *av = reinterpret_cast<uint32_t *>(bk + n_bound), *bv = av + n_bound;
    GE_TRY(tile_sort(ak, av));
    const unsigned mgrid = (unsigned)std::min<uint64_t>(nblk(n_bound, 256), 65535);
    uint64_t run = SS_TILE;
    for (int lv = 0; lv < levels; lv++, run *= SS_GROUP) {   // a -> b -> a ...; the last level writes the caller's arrays
        const bool last = lv == levels - 1;
        small_sort_merge_kernel<<<mgrid, 256, 0, st>>>(ak, av, last ? kout : bk, last ? vout : bv, n, (uint32_t)std::min<uint64_t>(run, 1u << 31));
        GE_TRY(ctx->check_launch("small_sort_merge"));
        std::swap(ak, bk); std::swap(av, bv);
    }
    return GE_OK;
}

static void mate_release(ge_ctx *ctx, MateScratch &m) {
    for (Buf *b : {&m.fam_off, &m.keep, &m.keep_off, &m.keys_a, &m.keys_b, &m.idx_a, &m.idx_b, &m.list_m, &m.list_f, &m.t1, &m.t2, &m.rank1, &m.rank2, &m.tmp_sort})
        ctx->release(*b);
}

// ------------------------------------------------------------------------------------------------
// the mating chain of one population, queued on the control stream (and the four sort lanes); see StepState for what it leaves
// ------------------------------------------------------------------------------------------------
static int ensure_couples(ge_ctx *ctx, PopDev &P, uint64_t n) {
    if (n <= P.couples_cap) return GE_OK;
    GE_TRY(ctx->ensure(P.c_male, n * 4)); GE_TRY(ctx->ensure(P.c_female, n * 4));
    GE_TRY(ctx->ensure(P.c_inbreed, n)); GE_TRY(ctx->ensure(P.c_noff, n * 4));
    GE_TRY(ctx->ensure(P.cnt32, (std::max<uint64_t>(n, ctx->cfg.capacity * ctx->cfg.n_chr * 2) + 1) * 4));
    GE_TRY(ctx->ensure(P.mate.fam_off, (n + 1) * 8));
    P.couples_cap = n;
    P.hs.couples_cap = n;
    ctx->graph_epoch++;
    return ctx->push_state(P, offsetof(StepState, couples_cap), 8);
}

// replayed mating draws -> device buffers (zero-padded to the buffers' bounds, values checked against them)
static int upload_f64(ge_ctx *ctx, Buf &b, const double *src, uint64_t n, uint64_t bound) {
    GE_TRY(ctx->ensure(b, std::max<uint64_t>(bound, 1) * 8));
    CUDA_TRY(cudaMemsetAsync(b.p, 0, bound * 8, ctx->stream));
    if (src && n) CUDA_TRY(cudaMemcpyAsync(b.p, src, std::min(n, bound) * 8, cudaMemcpyHostToDevice, ctx->stream));
    return GE_OK;
}
static int upload_idx(ge_ctx *ctx, Buf &b, const uint64_t *src, uint64_t n, uint64_t bound, uint64_t limit, std::vector<uint32_t> &tmp, const char *what) {
    GE_TRY(ctx->ensure(b, std::max<uint64_t>(bound, 1) * 4));
    CUDA_TRY(cudaMemsetAsync(b.p, 0, bound * 4, ctx->stream));
    if (n > bound) return fail(GE_ERR_INVALID, std::string("ge_mate_replay: ") + what + " is longer than the population allows");
    tmp.resize(n);
    for (uint64_t k = 0; k < n; k++) {
        if (src[k] >= limit) return fail(GE_ERR_INVALID, std::string("ge_mate_replay: ") + what + " holds an index out of range");
        tmp[k] = (uint32_t)src[k];
    }
    if (n) CUDA_TRY(cudaMemcpyAsync(b.p, tmp.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));   // tmp is reused by the caller
    return GE_OK;
}

static int enqueue_mate(ge_ctx *ctx, int pop, const ge_gen_params &gp, const ge_mate_draws *md = nullptr) {
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    MateScratch &M = P.mate;
    cudaStream_t st = ctx->stream;
    StepState *ss = P.d_ss;
    const uint64_t cap = ctx->cfg.capacity;
    const bool with_mm = !P.RM;
    const uint64_t lb = (with_mm && P.MM > 0 ? 2 : 1) * cap;        // list entries at most (a --MM duplicate is a second entry)
    const unsigned g_ind = ctx->grid_for(cap, 256);
    const uint64_t cb = lb / 2 + 1;                                  // couples of assortative mating at most: min(n_m, n_f)
    GE_TRY(ensure_couples(ctx, P, P.RM ? std::max<uint64_t>(gp.pop_size, 1) : cb));   // (before the kernels below read couples_cap)
    GE_TRY(ctx->ensure(M.keys_a, (cap + 1) * 8)); GE_TRY(ctx->ensure(M.keys_b, (lb + 1) * 8));
    GE_TRY(ctx->ensure(M.list_m, lb * 4)); GE_TRY(ctx->ensure(M.list_f, lb * 4));
    // who may mate: thinning counts, one scan for both sexes, the two lists
    std::vector<uint32_t> tmp;
    const double *d_thin = nullptr, *d_mm = nullptr;
    if (md) {   // the reference's own uniforms
        if (!md->thin_u || (with_mm && !md->mm_u)) return fail(GE_ERR_INVALID, "ge_mate_replay: thin_u / mm_u missing");
        GE_TRY(upload_f64(ctx, M.t1, md->thin_u, S.n, cap));
        GE_TRY(upload_f64(ctx, M.t2, with_mm ? md->mm_u : nullptr, S.n, cap));
        d_thin = M.t1.as<double>(); d_mm = M.t2.as<double>();
    }
    GE_TRY(ctx->scan_in(st, ThinIn{ctx->rng, ss, pop, S.sex.as<uint8_t>(), S.svf.as<double>(), with_mm ? 1 : 0, P.MM, d_thin, d_mm}, devn(S.d_n), cap, M.keys_b.as<uint64_t>(),
                        ThinTotal{ss, P.RM ? 1 : 0}));
    thin_fill_kernel<<<g_ind, 256, 0, st>>>(S.d_n, M.keys_b.as<uint64_t>(), M.list_m.as<uint32_t>(), M.list_f.as<uint32_t>());
    GE_TRY(ctx->check_launch("thin_fill"));
    if (P.RM) {  // random_mate :2090-2157
        const uint32_t *im = nullptr, *jf = nullptr;
        if (md) {
            if (!md->rm_father_idx || !md->rm_mother_idx || md->n_rm != gp.pop_size) return fail(GE_ERR_INVALID, "ge_mate_replay: random mating needs pop_size index draws for either parent");
            GE_TRY(upload_idx(ctx, M.idx_a, md->rm_father_idx, md->n_rm, std::max<uint64_t>(gp.pop_size, 1), lb, tmp, "rm_father_idx"));
            GE_TRY(upload_idx(ctx, M.idx_b, md->rm_mother_idx, md->n_rm, std::max<uint64_t>(gp.pop_size, 1), lb, tmp, "rm_mother_idx"));
            im = M.idx_a.as<uint32_t>(); jf = M.idx_b.as<uint32_t>();
        }
        rm_pair_kernel<<<ctx->grid_for(gp.pop_size, 256), 256, 0, st>>>(ctx->rng, ss, pop, M.list_m.as<uint32_t>(), M.list_f.as<uint32_t>(), im, jf, P.c_male.as<uint32_t>(),
                                                                        P.c_female.as<uint32_t>(), P.c_inbreed.as<uint8_t>(), P.c_noff.as<int32_t>());
        return ctx->check_launch("rm_pair");
    }
    // assort_mate :2167-2360
    const unsigned g_list = ctx->grid_for(lb, 256), g_c = ctx->grid_for(cb, 256);
    GE_TRY(ctx->ensure(M.t1, lb * 8)); GE_TRY(ctx->ensure(M.t2, lb * 8)); GE_TRY(ctx->ensure(M.idx_a, lb * 4)); GE_TRY(ctx->ensure(M.idx_b, lb * 4));
    GE_TRY(ctx->ensure(M.keep, (lb + 1) * 4)); GE_TRY(ctx->ensure(M.keep_off, (lb + 1) * 8));
    GE_TRY(ctx->ensure(M.rank1, cb * 4)); GE_TRY(ctx->ensure(M.rank2, cb * 4));
    const uint64_t *d_trim_len = &ss->n_trim_list;
    // trim the longer list: Philox keys, stable sort, flags, offsets (all of it idles when the lists are equal)
    if (md) {   // std::random_shuffle's order of the longer list (:2235 / :2242): its first n_trim entries leave
        GE_TRY(upload_idx(ctx, M.idx_b, md->trim_order, md->trim_order ? md->n_trim_order : 0, lb, std::max<uint64_t>(md->n_trim_order, 1), tmp, "trim_order"));
    } else {
        trim_keys_kernel<<<g_list, 256, 0, st>>>(ctx->rng, ss, pop, M.t1.as<uint64_t>(), M.idx_a.as<uint32_t>());
        GE_TRY(ctx->check_launch("trim_keys"));
        GE_TRY(sort_pairs_on(ctx, st, M.tmp_sort, M.t1.as<uint64_t>(), M.t2.as<uint64_t>(), M.idx_a.as<uint32_t>(), M.idx_b.as<uint32_t>(), devn(d_trim_len), lb));
    }
    trim_flag_kernel<<<g_list, 256, 0, st>>>(ss, M.idx_b.as<uint32_t>(), M.keep.as<uint32_t>());
    GE_TRY(ctx->check_launch("trim_flag"));
    GE_TRY(ctx->scan(st, M.keep.as<uint32_t>(), devn(d_trim_len), lb, M.keep_off.as<uint64_t>(), NoTotal{}));
    // template values, then the four independent sorts side by side on the sort lanes
    if (md) {
        if (!md->t1 || !md->t2) return fail(GE_ERR_INVALID, "ge_mate_replay: template values missing");
        GE_TRY(upload_f64(ctx, M.keys_a, md->t1, md->n_couples, cb));   // (keys_a / keys_b: the thinning is done with them)
        GE_TRY(upload_f64(ctx, M.keys_b, md->t2, md->n_couples, cb));
        template_from_kernel<<<g_c, 256, 0, st>>>(ss, M.keys_a.as<double>(), M.keys_b.as<double>(), M.t1.as<uint64_t>(), M.t2.as<uint64_t>(), M.idx_a.as<uint32_t>());
        GE_TRY(ctx->check_launch("template_from"));
    } else {
        template_kernel<<<g_c, 256, 0, st>>>(ctx->rng, ss, pop, M.t1.as<uint64_t>(), M.t2.as<uint64_t>(), M.idx_a.as<uint32_t>());
        GE_TRY(ctx->check_launch("template"));
    }
    GE_TRY(ctx->fork_lanes());
    const DevN n2 = devn(&ss->n2);
    for (int w = 0; w < 2; w++) {   // males, females by mating value
        SortLane &L = ctx->lane[w];
        GE_TRY(ctx->ensure(L.keys_in, cb * 8)); GE_TRY(ctx->ensure(L.vals_in, cb * 4)); GE_TRY(ctx->ensure(L.keys_out, cb * 8)); GE_TRY(ctx->ensure(L.vals_out, cb * 4));
        mv_keys_kernel<<<g_list, 256, 0, L.s>>>(ss, (uint32_t)w, (w ? M.list_f : M.list_m).as<uint32_t>(), M.keep.as<uint32_t>(), M.keep_off.as<uint64_t>(), S.mv.as<double>(),
                                                L.keys_in.as<uint64_t>(), L.vals_in.as<uint32_t>());
        GE_TRY(ctx->check_launch("mv_keys"));
        GE_TRY(sort_pairs_on(ctx, L.s, L.tmp, L.keys_in.as<uint64_t>(), L.keys_out.as<uint64_t>(), L.vals_in.as<uint32_t>(), L.vals_out.as<uint32_t>(), n2, cb));
    }
    for (int w = 0; w < 2; w++) {   // ranks of the two template columns
        SortLane &L = ctx->lane[2 + w];
        GE_TRY(ctx->ensure(L.keys_out, cb * 8)); GE_TRY(ctx->ensure(L.vals_out, cb * 4));
        GE_TRY(sort_pairs_on(ctx, L.s, L.tmp, (w ? M.t2 : M.t1).as<uint64_t>(), L.keys_out.as<uint64_t>(), M.idx_a.as<uint32_t>(), L.vals_out.as<uint32_t>(), n2, cb));
        rank_scatter_kernel<<<g_c, 256, 0, L.s>>>(ss, L.vals_out.as<uint32_t>(), (w ? M.rank2 : M.rank1).as<uint32_t>());
        GE_TRY(ctx->check_launch("rank_scatter"));
    }
    GE_TRY(ctx->join_lanes());
    pair_kernel<<<g_c, 256, 0, st>>>(ss, ctx->lane[0].vals_out.as<uint32_t>(), ctx->lane[1].vals_out.as<uint32_t>(), M.rank1.as<uint32_t>(), M.rank2.as<uint32_t>(), S.ids.as<uint64_t>(),
                                     P.avoid_inbreeding, P.c_male.as<uint32_t>(), P.c_female.as<uint32_t>(), P.c_inbreed.as<uint8_t>());
    GE_TRY(ctx->check_launch("pair"));
    const bool poisson = gp.offspring_dist == 'p' || gp.offspring_dist == 'P';
    if (md && poisson) {   // ras_rpois's family sizes as the reference drew them
        if (!md->family) return fail(GE_ERR_INVALID, "ge_mate_replay: family sizes missing");
        CUDA_TRY(cudaMemsetAsync(P.c_noff.p, 0, cb * 4, st));
        CUDA_TRY(cudaMemcpyAsync(P.c_noff.p, md->family, std::min<uint64_t>(md->n_couples, cb) * 4, cudaMemcpyHostToDevice, st));
    }
    family_kernel<<<g_c, 256, 0, st>>>(ctx->rng, ss, pop, md && poisson ? 2 : (poisson ? 1 : 0), P.c_noff.as<int32_t>());
    GE_TRY(ctx->check_launch("family"));
    if (!poisson) {
        if (md) {   // pos_couple_can_marry after std::random_shuffle (:2350): its first `remainder` couples get one more child
            GE_TRY(upload_idx(ctx, M.idx_b, md->remainder_order, md->remainder_order ? md->n_remainder_order : 0, cb, std::max<uint64_t>(md->n_couples, 1), tmp, "remainder_order"));
        } else {
            remainder_keys_kernel<<<g_c, 256, 0, st>>>(ctx->rng, ss, pop, P.c_inbreed.as<uint8_t>(), M.t1.as<uint64_t>(), M.idx_a.as<uint32_t>());
            GE_TRY(ctx->check_launch("remainder_keys"));
            GE_TRY(sort_pairs_on(ctx, st, M.tmp_sort, M.t1.as<uint64_t>(), M.t2.as<uint64_t>(), M.idx_a.as<uint32_t>(), M.idx_b.as<uint32_t>(), n2, cb));
        }
        remainder_add_kernel<<<g_c, 256, 0, st>>>(ss, M.idx_b.as<uint32_t>(), P.c_inbreed.as<uint8_t>(), P.c_noff.as<int32_t>());
        GE_TRY(ctx->check_launch("remainder_add"));
    }
    return GE_OK;
}


// ------------------------------------------------------------------------------------------------
// migration (ras_do_migration :877-989): whole individuals move between populations.  Each destination
// population is rebuilt by gathering (source population, source position) pairs: its stayers in order, then
// the camps of the other populations in (source, destination) order, each camp in descending source position
// (the reference sorts the sample descending, :922).  Every per-individual column, the bit-packed rows, the CV
// planes and the CSR lists (mutations, segments) are gathered on the device; only the small index lists are
// built on the host.
// ------------------------------------------------------------------------------------------------
namespace gek {

constexpr int MAX_POP = 16;
struct PopPtrs { const void *p[MAX_POP]; uint64_t n[MAX_POP]; };

// Moves the two bit-packed rows of migrant m from (spop[m], row pair sidx[m]) to (dpop[m], row pair didx[m]); one CTA per
// row (rows are multiples of 128 bytes): streaming 16-byte copies, four in flight per thread.
__global__ void move_rows_kernel(PopPtrs src, const uint8_t *__restrict__ spop, const uint32_t *__restrict__ sidx, PopPtrs dst,
                                 const uint8_t *__restrict__ dpop, const uint32_t *__restrict__ didx, uint64_t n_moves, uint32_t row_words) {
    for (uint64_t r = blockIdx.x; r < n_moves * 2; r += gridDim.x) {
        const uint64_t m = r >> 1;
        const uint32_t h = (uint32_t)(r & 1);
        const uint4 *s = reinterpret_cast<const uint4 *>(static_cast<const uint32_t *>(src.p[spop ? spop[m] : 0]) + ((uint64_t)sidx[m] * 2 + h) * row_words);
        uint4 *d = reinterpret_cast<uint4 *>(static_cast<uint32_t *>(const_cast<void *>(dst.p[dpop ? dpop[m] : 0])) + ((uint64_t)didx[m] * 2 + h) * row_words);
        const uint32_t nq = row_words / 4;
        uint32_t c = threadIdx.x;
        for (; c + 3 * blockDim.x < nq; c += 4 * blockDim.x) {
            uint4 a = ld_stream(s + c), b = ld_stream(s + c + blockDim.x), e = ld_stream(s + c + 2 * blockDim.x), f = ld_stream(s + c + 3 * blockDim.x);
            st_stream(d + c, a); st_stream(d + c + blockDim.x, b); st_stream(d + c + 2 * blockDim.x, e); st_stream(d + c + 3 * blockDim.x, f);
        }
        for (; c < nq; c += blockDim.x) st_stream(d + c, ld_stream(s + c));
    }
}
__global__ void gather_bytes_kernel(PopPtrs src, const uint8_t *__restrict__ gpop, const uint32_t *__restrict__ gidx, uint64_t n_dst,
                                    uint32_t bytes_per_ind, uint8_t *__restrict__ dst) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_dst * bytes_per_ind) return;
    uint64_t k = t / bytes_per_ind; uint32_t b = (uint32_t)(t % bytes_per_ind);
    dst[t] = static_cast<const uint8_t *>(src.p[gpop[k]])[(uint64_t)gidx[k] * bytes_per_ind + b];
}
// fp64 columns are stored [f*stride + i], stride = the capacity
__global__ void gather_f64_kernel(PopPtrs src, const uint8_t *__restrict__ gpop, const uint32_t *__restrict__ gidx, uint64_t n_dst, int n_col, uint64_t stride,
                                  double *__restrict__ dst) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_dst * n_col) return;
    uint64_t k = t % n_dst; int f = (int)(t / n_dst);
    int p = gpop[k];
    dst[(uint64_t)f * stride + k] = static_cast<const double *>(src.p[p])[(uint64_t)f * stride + gidx[k]];
}
// CSR gather: slots_per_ind lists per individual; element = ELEM bytes
__global__ void gather_csr_count_kernel(PopPtrs src_off, const uint8_t *__restrict__ gpop, const uint32_t *__restrict__ gidx, uint64_t n_dst,
                                        uint32_t slots_per_ind, uint32_t *__restrict__ cnt) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_dst * slots_per_ind) return;
    uint64_t k = t / slots_per_ind; uint32_t sub = (uint32_t)(t % slots_per_ind);
    const uint64_t *o = static_cast<const uint64_t *>(src_off.p[gpop[k]]);
    uint64_t s = (uint64_t)gidx[k] * slots_per_ind + sub;
    cnt[t] = o ? (uint32_t)(o[s + 1] - o[s]) : 0u;
}
template <class T>
__global__ void gather_csr_fill_kernel(PopPtrs src_off, PopPtrs src_val, const uint8_t *__restrict__ gpop, const uint32_t *__restrict__ gidx,
                                       uint64_t n_dst, uint32_t slots_per_ind, const uint64_t *__restrict__ dst_off, T *__restrict__ dst) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_dst * slots_per_ind) return;
    uint64_t k = t / slots_per_ind; uint32_t sub = (uint32_t)(t % slots_per_ind);
    const uint64_t *o = static_cast<const uint64_t *>(src_off.p[gpop[k]]);
    if (!o) return;
    const T *v = static_cast<const T *>(src_val.p[gpop[k]]);
    uint64_t s = (uint64_t)gidx[k] * slots_per_ind + sub;
    uint64_t d = dst_off[t];
    for (uint64_t e = o[s]; e < o[s + 1]; e++) dst[d++] = v[e];
}
__global__ void migrate_keys_kernel(Stream st, int pop, int gen, uint64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t w[4];
    draw(st, P_MIGRATE, pop, gen, k, 0, 0, w);
    keys[k] = key64(w); idx[k] = (uint32_t)k;
}

}  // namespace gek

static int migrate(ge_ctx *ctx, int gen, const double *row) {
    int np = ctx->cfg.n_pop, C = ctx->cfg.n_chr, nf = ctx->cfg.n_phen;
    if (np > MAX_POP) return fail(GE_ERR_UNSUPPORTED, "too many populations");
    if (!row) return fail(GE_ERR_INVALID, "null migration row");
    cudaStream_t st = ctx->stream;
    // Whole individuals move.  Everything the control chain needs (per-individual columns, causal-variant planes, lists)
    // is gathered into the other generation buffer on the control stream.  The bit-packed rows are NOT re-packed (that
    // would be a second full pass over the generation): stayers keep their physical rows, each incoming migrant's two
    // rows are copied — through a staging buffer, on the bulk stream, behind the copy that still produces them — into a
    // row pair vacated by an emigrant (or appended), and a per-individual row map (logical position -> physical row
    // pair) is handed to the next propagation, whose offspring are written in identity order again.
    const int par = gen & 1;
    std::vector<std::vector<uint64_t>> num_move(np, std::vector<uint64_t>(np, 0));
    for (int i = 0; i < np; i++) {
        double s = 0;
        for (int j = 0; j < np; j++) s += row[i * np + j];
        if (s < 0.99999 || s > 1.00001) return fail(GE_ERR_MIGRATION, "Error: The sum of columns in transition matrix in [--file_migration] must be 1.");
    }
    std::vector<uint64_t> n_old(np);
    for (int i = 0; i < np; i++) n_old[i] = ctx->pop[i].st[ctx->pop[i].cur].n;
    for (int i = 0; i < np; i++) for (int j = 0; j < np; j++) if (i != j) num_move[i][j] = (uint64_t)std::llround(row[i * np + j] * (double)n_old[i]);
    // the migrants of every source population, sorted descending (:921-922)
    std::vector<std::vector<uint64_t>> sample(np);
    for (int i = 0; i < np; i++) {
        uint64_t s = 0;
        for (uint64_t v : num_move[i]) s += v;
        if (s > n_old[i]) return fail(GE_ERR_MIGRATION, "more migrants than individuals");
        if (ctx->cfg.rng_mode == GE_RNG_REPLAY) {
            if ((int)ctx->mig_sample.size() <= i || ctx->mig_sample[i].size() != s) return fail(GE_ERR_INVALID, "replay mode: ge_set_migration_sample must supply the migrants");
            sample[i] = ctx->mig_sample[i];
        } else if (s) {
            // uniform sample without replacement = the s smallest (Philox key, position) pairs
            PopDev &P = ctx->pop[i];
            MateScratch &M = P.mate;
            uint64_t n = n_old[i];
            GE_TRY(ctx->ensure(M.keys_a, (n + 1) * 8)); GE_TRY(ctx->ensure(M.keys_b, (n + 1) * 8)); GE_TRY(ctx->ensure(M.idx_a, n * 4)); GE_TRY(ctx->ensure(M.idx_b, n * 4));
            migrate_keys_kernel<<<nblk(n, 256), 256, 0, st>>>(ctx->rng, i, gen, n, M.keys_a.as<uint64_t>(), M.idx_a.as<uint32_t>());
            GE_TRY(ctx->check_launch("migrate_keys"));
            GE_TRY(sort_pairs_on(ctx, st, M.tmp_sort, M.keys_a.as<uint64_t>(), M.keys_b.as<uint64_t>(), M.idx_a.as<uint32_t>(), M.idx_b.as<uint32_t>(), hostn(n), n));
            std::vector<uint32_t> h(s);
            CUDA_TRY(cudaMemcpyAsync(h.data(), M.idx_b.p, s * 4, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            sample[i].assign(h.begin(), h.end());
        }
        std::sort(sample[i].begin(), sample[i].end(), std::greater<uint64_t>());
        for (uint64_t v : sample[i]) if (v >= n_old[i]) return fail(GE_ERR_INVALID, "migrant position out of range");
    }
    // gather lists per destination
    std::vector<std::vector<uint8_t>> gp(np);
    std::vector<std::vector<uint32_t>> gi(np);
    for (int j = 0; j < np; j++) {
        std::vector<uint8_t> gone(n_old[j], 0);
        for (uint64_t v : sample[j]) gone[v] = 1;
        for (uint64_t k = 0; k < n_old[j]; k++) if (!gone[k]) { gp[j].push_back((uint8_t)j); gi[j].push_back((uint32_t)k); }
    }
    for (int i = 0; i < np; i++) {
        uint64_t k = 0;  // consecutive slices of the sample go to consecutive destinations (the reference does not
        for (int j = 0; j < np; j++) {  // reset k, :924-936; identical whenever the reference itself survives)
            if (i == j) continue;
            for (uint64_t it = 0; it < num_move[i][j]; it++, k++) { gp[j].push_back((uint8_t)i); gi[j].push_back((uint32_t)sample[i][k]); }
        }
    }
    for (int j = 0; j < np; j++) if (gp[j].size() > ctx->cfg.capacity) return fail(GE_ERR_CAPACITY, "population outgrew capacity through migration");
    // source tables
    auto table = [&](auto getter) { PopPtrs t{}; for (int p = 0; p < np; p++) { GenState &S = ctx->pop[p].st[ctx->pop[p].cur]; t.p[p] = getter(S); t.n[p] = S.n; } return t; };
    bool any_hm = false, any_seg = ctx->segs();
    for (int p = 0; p < np; p++) any_hm |= ctx->pop[p].st[ctx->pop[p].cur].has_hm;
    for (int j = 0; j < np; j++) {
        PopDev &P = ctx->pop[j];
        GenState &D = P.st[P.cur ^ 1];
        uint64_t n = gp[j].size();
        Buf &d_gp = P.mig_pop[0], &d_gi = P.mig_idx[0];   // read on the control stream only
        GE_TRY(ctx->upload(d_gp, gp[j])); GE_TRY(ctx->upload(d_gi, gi[j]));
        const uint8_t *g8 = d_gp.as<uint8_t>(); const uint32_t *g32 = d_gi.as<uint32_t>();
        if (n) {
            if (ctx->n_cv_tot) {
                gather_bytes_kernel<<<nblk(n * 8 * ctx->Wcv, 256), 256, 0, st>>>(table([](GenState &S) { return S.cv_allele.p; }), g8, g32, n, 8 * ctx->Wcv, D.cv_allele.as<uint8_t>());
                GE_TRY(ctx->check_launch("gather_cv"));
                if (ctx->use_root) {
                    gather_bytes_kernel<<<nblk(n * 2 * ctx->n_cv_tot, 256), 256, 0, st>>>(table([](GenState &S) { return S.cv_root.p; }), g8, g32, n, 2 * ctx->n_cv_tot, D.cv_root.as<uint8_t>());
                    GE_TRY(ctx->check_launch("gather_cv_root"));
                }
            }
            gather_bytes_kernel<<<nblk(n * 56, 256), 256, 0, st>>>(table([](GenState &S) { return S.ids.p; }), g8, g32, n, 56, D.ids.as<uint8_t>());
            GE_TRY(ctx->check_launch("gather_ids"));
            gather_bytes_kernel<<<nblk(n, 256), 256, 0, st>>>(table([](GenState &S) { return S.sex.p; }), g8, g32, n, 1, D.sex.as<uint8_t>());
            GE_TRY(ctx->check_launch("gather_sex"));
            struct Col { Buf GenState::*m; int cols; };
            const Col cols[] = {{&GenState::A, nf}, {&GenState::D, nf}, {&GenState::G, nf}, {&GenState::C, nf}, {&GenState::E, nf}, {&GenState::F, nf},
                                {&GenState::P, nf}, {&GenState::mv, 1}, {&GenState::sv, 1}, {&GenState::svf, 1}};
            for (const Col &cc : cols) {
                PopPtrs t{};
                for (int p = 0; p < np; p++) { GenState &S = ctx->pop[p].st[ctx->pop[p].cur]; t.p[p] = (S.*(cc.m)).p; t.n[p] = S.n; }
                gather_f64_kernel<<<nblk(n * cc.cols, 256), 256, 0, st>>>(t, g8, g32, n, cc.cols, ctx->cfg.capacity, (D.*(cc.m)).as<double>());
                GE_TRY(ctx->check_launch("gather_f64"));
            }
        }
        uint32_t spi = (uint32_t)(C * 2);
        if (any_hm) {
            PopPtrs to = table([](GenState &S) { return S.has_hm ? S.hm_off.p : nullptr; }), tv = table([](GenState &S) { return S.hm_bp.p; });
            GE_TRY(ctx->ensure(P.cnt32, (n * spi + 1) * 4)); GE_TRY(ctx->ensure(D.hm_off, (n * spi + 1) * 8));
            if (n) { gather_csr_count_kernel<<<nblk(n * spi, 256), 256, 0, st>>>(to, g8, g32, n, spi, P.cnt32.as<uint32_t>()); GE_TRY(ctx->check_launch("gather_csr_count")); }
            GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), n * spi, D.hm_off.as<uint64_t>(), &D.n_hm));
            GE_TRY(ctx->ensure(D.hm_bp, std::max<uint64_t>(D.n_hm, 1) * 4));
            if (n) { gather_csr_fill_kernel<uint32_t><<<nblk(n * spi, 256), 256, 0, st>>>(to, tv, g8, g32, n, spi, D.hm_off.as<uint64_t>(), D.hm_bp.as<uint32_t>()); GE_TRY(ctx->check_launch("gather_csr_fill")); }
            D.has_hm = true;
        } else D.has_hm = false;
        if (any_seg) {
            PopPtrs to = table([](GenState &S) { return S.seg.valid ? S.seg.off.p : nullptr; }), tv = table([](GenState &S) { return S.seg.seg.p; });
            GE_TRY(ctx->ensure(P.cnt32, (n * spi + 1) * 4)); GE_TRY(ctx->ensure(D.seg.off, (n * spi + 1) * 8));
            if (n) { gather_csr_count_kernel<<<nblk(n * spi, 256), 256, 0, st>>>(to, g8, g32, n, spi, P.cnt32.as<uint32_t>()); GE_TRY(ctx->check_launch("gather_csr_count")); }
            GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), n * spi, D.seg.off.as<uint64_t>(), &D.seg.n_seg));
            GE_TRY(ctx->ensure(D.seg.seg, std::max<uint64_t>(std::max<uint64_t>(D.seg.n_seg, ctx->cfg.seg_capacity), 1) * ctx->seg_esz()));
            if (n && ctx->seg_packed) { gather_csr_fill_kernel<uint2><<<nblk(n * spi, 256), 256, 0, st>>>(to, tv, g8, g32, n, spi, D.seg.off.as<uint64_t>(), D.seg.seg.as<uint2>()); GE_TRY(ctx->check_launch("gather_csr_fill")); }
            else if (n) { gather_csr_fill_kernel<uint4><<<nblk(n * spi, 256), 256, 0, st>>>(to, tv, g8, g32, n, spi, D.seg.off.as<uint64_t>(), D.seg.seg.as<uint4>()); GE_TRY(ctx->check_launch("gather_csr_fill")); }
            D.seg.valid = true;
        }
        D.n = n;
    }
    if (ctx->bits()) {
        for (int j = 0; j < np; j++)
            if (ctx->pop[j].st[ctx->pop[j].cur].rowmap) return fail(GE_ERR_INVALID, "two migrations without a generation in between");
        // physical placement of the rows: row map per destination, move list over all populations
        std::vector<uint8_t> m_spop, m_dpop;
        std::vector<uint32_t> m_sidx, m_didx, m_stage;
        for (int j = 0; j < np; j++) {
            PopDev &P = ctx->pop[j];
            GenState &D = P.st[P.cur ^ 1];
            std::vector<uint64_t> vacated(sample[j].rbegin(), sample[j].rend());   // ascending positions of the emigrants
            size_t next_free = 0;
            uint64_t next_new = n_old[j];
            std::vector<uint32_t> rowmap(gp[j].size());
            for (size_t k = 0; k < gp[j].size(); k++) {
                if (gp[j][k] == j) { rowmap[k] = gi[j][k]; continue; }
                uint32_t slot = next_free < vacated.size() ? (uint32_t)vacated[next_free++] : (uint32_t)next_new++;
                rowmap[k] = slot;
                m_stage.push_back((uint32_t)m_spop.size());
                m_spop.push_back(gp[j][k]); m_sidx.push_back(gi[j][k]); m_dpop.push_back((uint8_t)j); m_didx.push_back(slot);
            }
            if (next_new > ctx->cfg.capacity) return fail(GE_ERR_CAPACITY, "population outgrew capacity through migration");
            // the map of generation g-2 in this slot was read by the propagation of generation g-1 (the previous draw set)
            DrawSet &prev = P.ds[P.dcur ^ 1];
            if (prev.bulk_pending) CUDA_TRY(cudaStreamWaitEvent(st, prev.bulk_done, 0));
            GE_TRY(ctx->upload(P.rowmap_buf[par], rowmap));
            D.rowmap = P.rowmap_buf[par].as<uint32_t>();
        }
        const uint64_t n_moves = m_spop.size();
        if (n_moves) {
            cudaStream_t bulk = ctx->serial ? st : ctx->bulk;
            if (ctx->mig_pending[par]) { CUDA_TRY(cudaStreamWaitEvent(st, ctx->mig_done[par], 0)); ctx->mig_pending[par] = false; }
            if (!ctx->mig_done[par]) CUDA_TRY(cudaEventCreateWithFlags(&ctx->mig_done[par], cudaEventDisableTiming));
            Buf *mb = ctx->mig_lists[par];
            GE_TRY(ctx->upload(mb[0], m_spop)); GE_TRY(ctx->upload(mb[1], m_sidx)); GE_TRY(ctx->upload(mb[2], m_dpop)); GE_TRY(ctx->upload(mb[3], m_didx));
            GE_TRY(ctx->upload(mb[4], m_stage));
            GE_TRY(ctx->ensure(ctx->mig_stage, n_moves * 2 * (size_t)ctx->W * 4));
            CUDA_TRY(cudaEventRecord(ctx->ev_ready, st));           // the lists are on the device
            CUDA_TRY(cudaStreamWaitEvent(bulk, ctx->ev_ready, 0));
            PopPtrs rows = table([](GenState &S) { return S.hap.p; }), stage{};
            stage.p[0] = ctx->mig_stage.p;
            const unsigned grid = (unsigned)std::min<uint64_t>(n_moves * 2, 1u << 20);
            move_rows_kernel<<<grid, 256, 0, bulk>>>(rows, mb[0].as<uint8_t>(), mb[1].as<uint32_t>(), stage, nullptr, mb[4].as<uint32_t>(), n_moves, ctx->W);
            GE_TRY(ctx->check_launch("move_rows<stage>"));
            move_rows_kernel<<<grid, 256, 0, bulk>>>(stage, nullptr, mb[4].as<uint32_t>(), rows, mb[2].as<uint8_t>(), mb[3].as<uint32_t>(), n_moves, ctx->W);
            GE_TRY(ctx->check_launch("move_rows<place>"));
            CUDA_TRY(cudaEventRecord(ctx->mig_done[par], bulk));
            ctx->mig_pending[par] = true;
        }
        for (int j = 0; j < np; j++) { PopDev &P = ctx->pop[j]; std::swap(P.st[P.cur].hap, P.st[P.cur ^ 1].hap); }   // the rows stay where they are
    }
    for (int j = 0; j < np; j++) {   // the host decided the new sizes: tell the device-resident step state
        PopDev &P = ctx->pop[j];
        P.cur ^= 1;
        GenState &S = P.st[P.cur];
        P.hs.n[P.cur] = S.n; P.hs.n_hm[P.cur] = S.n_hm; P.hs.n_seg[P.cur] = S.seg.n_seg;
        GE_TRY(ctx->push_state(P, offsetof(StepState, n), offsetof(StepState, prev_n) - offsetof(StepState, n)));
    }
    ctx->mig_sample.clear();
    return GE_OK;
}
