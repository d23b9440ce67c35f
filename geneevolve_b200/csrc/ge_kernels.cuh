// ge_kernels.cuh — hand-written sm_100a kernels of the GeneEvolve reproduction hot path.
//
// Everything here is HBM-bound integer/bit work or tiny fp64 reductions; there is no dense contraction, so
// no tensor-core (tcgen05) path — see DESIGN.md §Kernels for the roofline that bounds each kernel.
// Reference lines cited as :N are src/Simulation.cpp:N of MMesbahU/GeneEvolve.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gek {

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (replaces RasRandomNumber / std::minstd_rand0 / rand())
// ------------------------------------------------------------------------------------------------
enum Purpose : uint32_t {
    P_THIN = 1, P_RM_PAIR = 2, P_TRIM = 3, P_TEMPLATE = 4, P_POISSON = 5, P_REMAINDER = 6, P_XO = 7, P_MUT = 8,
    P_SEX = 9, P_ENOISE = 10, P_F0 = 11, P_COMMON = 12, P_MIGRATE = 13
};

struct Stream { uint32_t k0, k1; };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                       uint32_t c2, uint32_t c3, uint32_t w[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
#ifdef __CUDA_ARCH__
        uint32_t h0 = __umulhi(M0, c0), l0 = M0 * c0, h1 = __umulhi(M1, c2), l1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += W0; k1 += W1;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

__device__ __forceinline__ void draw(const Stream s, uint32_t purpose, int pop, int gen, uint64_t entity,
                                     uint32_t sub, uint32_t block, uint32_t w[4]) {
    uint32_t c3 = (purpose << 24) | ((uint32_t)pop << 20) | ((uint32_t)gen & 0xFFFFFu);
    philox4x32_10(s.k0, s.k1, block, (uint32_t)entity, sub, c3, w);
}

__device__ __forceinline__ double u01(uint32_t a, uint32_t b) {  // [0,1), 53 bits
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void normal2(const Stream s, uint32_t purpose, int pop, int gen, uint64_t entity,
                                        uint32_t sub, double &z0, double &z1) {
    uint32_t w[4];
    draw(s, purpose, pop, gen, entity, sub, 0, w);
    double u1 = 1.0 - u01(w[0], w[1]);
    double u2 = u01(w[2], w[3]);
    double r = sqrt(-2.0 * log(u1));
    double th = 6.283185307179586476925286766559 * u2;
    double sn, cs;
    sincos(th, &sn, &cs);
    z0 = r * cs;
    z1 = r * sn;
}

// ------------------------------------------------------------------------------------------------
// Device-resident state of one population's generation step.  Every size that is only known once a kernel of the
// step has run (list lengths after thinning, couples, offspring, crossovers, mutation hits) lives HERE and is read
// by the kernels that follow, so nothing between mating and the phenotypes of a generation needs the host: grids
// are sized from host-known bounds (capacity, the generation table's pop_size) and loop grid-stride up to the
// device-side count.  The host reads the struct back once, at the end of the step (sizes for its own
// bookkeeping, err for the reference's "No one can marry" class of errors).  Being plain device memory it is
// also what lets a whole generation be replayed as a CUDA graph: no kernel argument depends on a count.
// ------------------------------------------------------------------------------------------------
enum StepErr : uint32_t {
    SE_NAN = 1u << 0,            // A or D is NaN (:2716-2720)
    SE_PARENT_ID = 1u << 1,      // parent ID outside the previous generation (:3118-3133 reads out of bounds there)
    SE_NO_MATES_RM = 1u << 2,    // random_mate: an empty sex list (:2125-2129)
    SE_NO_COUPLES = 1u << 3,     // assort_mate: couples = 0 (:2226-2230)
    SE_ALL_INBRED = 1u << 4,     // every couple is inbred
    SE_NO_OFFSPRING = 1u << 5,
    SE_CAP_OFFSPRING = 1u << 6,  // offspring exceed ge_config.capacity
    SE_CAP_XO = 1u << 7,         // crossovers exceed the draw buffers (sized from the genetic map at generation 0)
    SE_CAP_MUT = 1u << 8,        // mutation hits exceed the draw buffers
    SE_CAP_HM = 1u << 9,         // per-haplotype mutation lists exceed their buffer
    SE_CAP_COUPLES = 1u << 10,
    SE_CAP_SEG = 1u << 11,       // founder-segment lists exceed ge_config.seg_capacity
    SE_SEG_UNSORTED = 1u << 12,  // packed 8-byte parts cannot hold the pieces of a gamete whose crossover positions do not ascend
};
// Sizes of one draw set (double-buffered like the draws themselves).  The bulk stream reads THESE, not the step's live counters: its
// kernels run a generation behind the control stream, which by then is already overwriting n_off and n_xo for the next generation.
struct DrawCounts { uint64_t n_off, n_xo, n_iv; uint32_t fatal, pad; };
struct StepState {
    // the generation-table row, written by step_begin_kernel
    int32_t gen, sel_func, offspring_dist, pad0;
    uint64_t pop_size;
    double mat_cor, u11, sel_par1, sel_par2;
    // sizes
    uint64_t n[2];                        // individuals in generation buffers st[0], st[1]
    uint64_t n_m, n_f, n2, n_trim;        // thinned sex lists; n2 = min(n_m, n_f); n_trim entries leave the longer one
    uint64_t n_trim_list;                 // length of the list being trimmed (0: nothing to trim)
    uint64_t n_couples, n_inbreed;
    uint64_t n_off, n_xo, n_mut;          // offspring and draws of this generation
    uint64_t n_hm[2];                     // entries of the per-haplotype mutation lists of st[0], st[1]
    uint64_t n_iv;                        // intervals of the segment plan: n_xo + offspring slots
    uint64_t n_seg[2];                    // parts in the founder-segment lists of st[0], st[1]
    DrawCounts dc[2];                     // per draw set, frozen when its crossovers have been placed
    uint64_t prev_n;                      // size of the snapshot read by vertical transmission (ras_save_human_info_to_Pop_info_prev_gen)
    uint32_t trim_which;                  // 0: males are trimmed, 1: females
    uint32_t err;                         // StepErr bits (sticky until the host clears them)
    // capacities the kernels clamp to (host constants, set once)
    uint64_t cap, xo_cap, mut_cap, hm_cap, couples_cap;
};
constexpr uint32_t SE_FATAL = SE_NO_MATES_RM | SE_NO_COUPLES | SE_ALL_INBRED | SE_NO_OFFSPRING | SE_CAP_OFFSPRING | SE_CAP_XO | SE_CAP_MUT | SE_CAP_HM | SE_CAP_COUPLES | SE_CAP_SEG | SE_SEG_UNSORTED;
struct StepRow { int32_t gen, sel_func, offspring_dist, pad; uint64_t pop_size; double mat_cor, sel_par1, sel_par2; };
// first kernel of a step: the host's generation-table row -> device (a kernel argument, so a captured graph replays
// with one node-parameter update) and the per-step counters back to zero
__global__ void step_begin_kernel(StepState *ss, StepRow r) {
    ss->gen = r.gen; ss->sel_func = r.sel_func; ss->offspring_dist = r.offspring_dist; ss->pop_size = r.pop_size;
    ss->mat_cor = r.mat_cor; ss->u11 = sqrt(1.0 - r.mat_cor * r.mat_cor); ss->sel_par1 = r.sel_par1; ss->sel_par2 = r.sel_par2;
    ss->n_inbreed = 0;
}

// ------------------------------------------------------------------------------------------------
// genome layout of one bit-packed haplotype row
// ------------------------------------------------------------------------------------------------
struct Genome {
    int n_chr;
    uint32_t W;                    // u32 words per haplotype row (multiple of 32 -> rows are 128 B aligned)
    const uint32_t *chr_word_off;  // [n_chr] first word of each chromosome (multiple of 4 -> 16 B aligned)
    const uint32_t *chr_nloci;     // [n_chr]
    const uint32_t *locus_off;     // [n_chr+1] offsets into pos
    const uint32_t *pos;           // concatenated locus positions (bp), ascending inside a chromosome
    // coarse position index: bkt[bkt_off[c] + b] = first locus of chromosome c with pos >= b << bkt_shift[c]
    const uint32_t *bkt_off;       // [n_chr+1]
    const uint32_t *bkt_shift;     // [n_chr]
    const uint32_t *bkt;
};

__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *a, uint32_t n, uint32_t key) {
    uint32_t lo = 0, hi = n;  // first index with a[idx] >= key
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// first locus of chromosome c with position >= key: one bucket lookup, then a search over the few loci of the
// bucket (about one on average) instead of ~20 dependent loads over the whole position table
__device__ __forceinline__ uint32_t locus_lower_bound(const Genome &g, int c, uint32_t key) {
    const uint32_t *pos = g.pos + g.locus_off[c];
    const uint32_t nb = g.bkt_off[c + 1] - g.bkt_off[c] - 1;  // buckets 0..nb-1, entry nb = n_loci
    const uint32_t b = key >> g.bkt_shift[c];
    if (b >= nb) return g.chr_nloci[c];
    const uint32_t *t = g.bkt + g.bkt_off[c] + b;
    uint32_t lo = __ldg(t), hi = __ldg(t + 1);
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(pos + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Crossover positions (bp) -> locus indices.  A locus s of chromosome c takes parental haplotype
// start ^ (#{crossovers b <= pos[s]} & 1) (recombine :2903-2958 followed by materialisation :1186-1230), i.e.
// the haplotype flips at locus index lower_bound(pos, b).  One thread per (offspring, chromosome, gamete) slot.
__global__ void xo_to_flips_kernel(Genome g, uint64_t n_slots, const uint64_t *__restrict__ xo_off,
                                   const uint32_t *__restrict__ xo_bp, uint32_t *__restrict__ flips) {
    uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    int c = (int)((slot >> 1) % (uint64_t)g.n_chr);
    for (uint64_t e = xo_off[slot]; e < xo_off[slot + 1]; e++) flips[e] = locus_lower_bound(g, c, xo_bp[e]);
}

// ------------------------------------------------------------------------------------------------
// THE hot kernel: bit-packed haplotype propagation (G5 in SURVEY.md §8a)
//   offspring row (i, g)  <-  alternating runs of the two haplotype rows of parent (father|mother)[i]
// One CTA per offspring; its warps pull work items (gamete, chromosome tile) from a shared counter, so long
// and short chromosomes balance.  Inside an item the control flow is warp-uniform: whole 16-byte chunks
// between two crossovers are a pure LDG.128/STG.128 copy with 4 independent loads in flight per lane; the
// (at most a few) chunks that contain a crossover are mask-merged from both parental rows.
// Algorithmic HBM traffic: 2 bits read + 2 bits written per individual-locus = 0.5 byte.
// ------------------------------------------------------------------------------------------------
struct TileTable {
    uint32_t n_items;          // items of ONE gamete; the CTA processes 2*n_items
    const uint32_t *chr;       // [n_items]
    const uint32_t *chunk0;    // [n_items] first 16 B chunk inside the chromosome
    const uint32_t *nchunk;    // [n_items]
};

// a / b and a % b for 64-bit a that almost always fits 32 bits (thread and slot indices): the 64-bit division is ~100 instructions
__device__ __forceinline__ void divmod_idx(uint64_t a, uint32_t b, uint64_t &q, uint32_t &r) {
    if (a <= 0xFFFFFFFFull) { const uint32_t a32 = (uint32_t)a, q32 = a32 / b; q = q32; r = a32 - q32 * b; }
    else { q = a / b; r = (uint32_t)(a - q * b); }
}

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    // (L2::256B: the sector pair comes with one request — measured 0.7 % on the copy alone; an explicit prefetch.global.L2 pass over an
    //  item's runs before its first copy was measured too and LOSES 8 %: profiles/r2i_copy_variants.md)
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// whole chunks [q, qe) of one run: four independent 16-byte loads in flight per lane
__device__ __forceinline__ void warp_copy_chunks(uint4 *__restrict__ dst, const uint4 *__restrict__ src, uint32_t q, uint32_t qe, int lane) {
    uint32_t c = q + lane;
    for (; c + 96 < qe; c += 128) {
        uint4 a = ld_stream(src + c), b = ld_stream(src + c + 32), d = ld_stream(src + c + 64), e = ld_stream(src + c + 96);
        st_stream(dst + c, a); st_stream(dst + c + 32, b); st_stream(dst + c + 64, d); st_stream(dst + c + 96, e);
    }
    // the remainder (the mean run between crossovers is about one such iteration long): all of a lane's loads are issued
    // before its first store, so the tail costs one memory latency instead of one per 512 bytes
    if (c < qe) {
        const bool p1 = c + 32 < qe, p2 = c + 64 < qe;
        uint4 a = ld_stream(src + c), b = a, d = a;
        if (p1) b = ld_stream(src + c + 32);
        if (p2) d = ld_stream(src + c + 64);
        st_stream(dst + c, a);
        if (p1) st_stream(dst + c + 32, b);
        if (p2) st_stream(dst + c + 64, d);
    }
}

constexpr int PROP_THREADS = 256;
constexpr int PROP_SMEM_FLIPS = 384;   // flips of one offspring staged in shared memory (mean 2 x 36 on the 22 autosomes)

// dynamic shared memory: uint64 xo_off[2*n_chr+1] | uint32 flips[PROP_SMEM_FLIPS] | uint8 start[2*n_chr]
static inline size_t prop_smem_bytes(int n_chr) { return (size_t)(2 * n_chr + 1) * 8 + PROP_SMEM_FLIPS * 4 + (size_t)((2 * n_chr + 15) & ~15); }

// The two loads of a crossover chunk are issued before the run copy that precedes it.  Same-box measurements on config 3
// (kernel alone / pipelined step): 40 registers, 6 CTAs/SM with this prefetch 6.55 / 7.67 ms; 32 registers, 8 CTAs/SM
// without it 6.85 / 7.74 ms; 32 registers with it spills (6.95 / 8.02 ms).  Rejected variants (eight loads per lane, a
// persistent grid, TMA-staged bulk copies: profiles/r2a_tma_variant.md) are in DESIGN.md §3.
// The grid is sized from the capacity; CTAs beyond the device-side offspring count exit at once.
__global__ void __launch_bounds__(PROP_THREADS, 6)
propagate_bits_kernel(Genome g, TileTable tt, const DrawCounts *__restrict__ dc, const uint32_t *__restrict__ par_rows, const uint32_t *__restrict__ par_rowmap,
                      uint32_t *__restrict__ off_rows, const uint32_t *__restrict__ father, const uint32_t *__restrict__ mother,
                      const uint64_t *__restrict__ xo_off, const uint32_t *__restrict__ flips, const uint8_t *__restrict__ start_hap) {
    if (dc->fatal) return;
    const uint32_t n_off = (uint32_t)dc->n_off;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_next;
    const int n_ls = 2 * g.n_chr;  // slots of one offspring: chromosome-major, gamete-minor — contiguous in every draw array
    uint64_t *s_off = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_fl = reinterpret_cast<uint32_t *>(smem_raw + (size_t)(n_ls + 1) * 8);
    uint8_t *s_start = smem_raw + (size_t)(n_ls + 1) * 8 + PROP_SMEM_FLIPS * 4;
    const int lane = threadIdx.x & 31;
    for (uint32_t oi = blockIdx.x; oi < n_off; oi += gridDim.x) {
        const uint64_t i = oi;
        const uint64_t slot0 = i * (uint64_t)n_ls;
        __syncthreads();
        // stage this offspring's crossover metadata once, coalesced: the work items below never wait on a
        // dependent global load before their first copy
        for (int t = threadIdx.x; t <= n_ls; t += blockDim.x) s_off[t] = xo_off[slot0 + t];
        for (int t = threadIdx.x; t < n_ls; t += blockDim.x) s_start[t] = start_hap[slot0 + t];
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        const uint64_t e_base = s_off[0];
        const uint32_t n_fl = (uint32_t)(s_off[n_ls] - e_base);
        const bool staged = n_fl <= PROP_SMEM_FLIPS;
        if (staged) for (uint32_t t = threadIdx.x; t < n_fl; t += blockDim.x) s_fl[t] = flips[e_base + t];
        // after a migration the parents' rows are not in logical order: par_rowmap gives the physical row pair
        const uint32_t pf = par_rowmap ? par_rowmap[father[i]] : father[i], pm = par_rowmap ? par_rowmap[mother[i]] : mother[i];
        __syncthreads();
        for (;;) {
            uint32_t item = 0;
            if (lane == 0) item = atomicAdd(&s_next, 1u);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= 2 * tt.n_items) break;
            const int gam = item >= tt.n_items;
            const uint32_t it = gam ? item - tt.n_items : item;
            const uint32_t c = tt.chr[it], q0 = tt.chunk0[it], q1 = q0 + tt.nchunk[it];
            const int ls = (int)c * 2 + gam;
            const uint64_t e0 = s_off[ls];
            const uint32_t k = (uint32_t)(s_off[ls + 1] - e0);
            const uint32_t *fl = staged ? s_fl + (e0 - e_base) : flips + e0;
            const uint32_t woff = g.chr_word_off[c];
            const uint32_t prow = gam ? pm : pf;
            const uint4 *h0 = reinterpret_cast<const uint4 *>(par_rows + (uint64_t)(2 * prow) * g.W + woff);
            const uint4 *h1 = reinterpret_cast<const uint4 *>(par_rows + (uint64_t)(2 * prow + 1) * g.W + woff);
            uint4 *dst = reinterpret_cast<uint4 *>(off_rows + (uint64_t)(2 * i + gam) * g.W + woff);
            // flips at or before the first locus of the tile only set the parity
            uint32_t j = 0;
            const uint32_t x0 = q0 << 7;
            while (j < k && fl[j] <= x0) j++;
            uint32_t cur = (s_start[ls] ^ j) & 1u;
            uint32_t q = q0;
            while (q < q1) {
                const uint32_t f = j < k ? fl[j] : 0xFFFFFFFFu;
                const uint32_t qb = f >> 7;
                if (j >= k || qb >= q1) { warp_copy_chunks(dst, cur ? h1 : h0, q, q1, lane); break; }
                // chunk qb holds one or more flips: merge both parental chunks under a 128-bit mask
                // (mask bit = 1 -> haplotype 1).  A flip on the chunk's first locus gives bit offset 0.
                // Its two loads are issued before the run that precedes it, so their latency hides behind that copy.
                uint4 a = make_uint4(0, 0, 0, 0), b = a;
                if (lane == 0) { a = ld_stream(h0 + qb); b = ld_stream(h1 + qb); }
                warp_copy_chunks(dst, cur ? h1 : h0, q, qb, lane);
                if (lane == 0) {
                    const uint32_t fill = cur ? 0xFFFFFFFFu : 0u;
                    uint32_t m0 = fill, m1 = fill, m2 = fill, m3 = fill;
                    const uint32_t base = qb << 7;
                    for (uint32_t jj = j; jj < k && fl[jj] < base + 128u; jj++) {
                        const uint32_t r = fl[jj] - base;
                        m0 ^= r < 32u ? 0xFFFFFFFFu << r : 0u;
                        m1 ^= r <= 32u ? 0xFFFFFFFFu : (r < 64u ? 0xFFFFFFFFu << (r - 32u) : 0u);
                        m2 ^= r <= 64u ? 0xFFFFFFFFu : (r < 96u ? 0xFFFFFFFFu << (r - 64u) : 0u);
                        m3 ^= r <= 96u ? 0xFFFFFFFFu : 0xFFFFFFFFu << (r - 96u);
                    }
                    uint4 o;
                    o.x = (a.x & ~m0) | (b.x & m0); o.y = (a.y & ~m1) | (b.y & m1);
                    o.z = (a.z & ~m2) | (b.z & m2); o.w = (a.w & ~m3) | (b.w & m3);
                    st_stream(dst + qb, o);
                }
                while (j < k && fl[j] < ((qb + 1) << 7)) { j++; cur ^= 1u; }
                q = qb + 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Founder panel (Hap_SNP bytes, hap-major) -> bit-packed generation-0 rows, masked to the covered range
// [rmap.bp[0], rmap.bp[last]) (a locus outside it lies in no `part`, :3029-3034, and reads 0 at :1186-1230).
// ------------------------------------------------------------------------------------------------
__global__ void pack_panel_kernel(const uint8_t *__restrict__ alleles, uint32_t n_rows, uint32_t n_loci,
                                  const uint32_t *__restrict__ pos, uint32_t cov_lo, uint32_t cov_hi,
                                  uint32_t *__restrict__ rows, uint32_t W, uint32_t woff) {
    uint32_t nw = (n_loci + 31) >> 5;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_rows * nw) return;
    uint32_t r = (uint32_t)(t / nw), w = (uint32_t)(t % nw);
    uint32_t v = 0;
    for (uint32_t b = 0; b < 32; b++) {
        uint32_t s = w * 32 + b;
        if (s < n_loci) {
            uint32_t p = pos[s];
            if (p >= cov_lo && p < cov_hi && alleles[(uint64_t)r * n_loci + s]) v |= 1u << b;
        }
    }
    rows[(uint64_t)r * W + woff + w] = v;
}

// host-packed panel words -> rows, masking loci outside the covered range
__global__ void mask_packed_panel_kernel(const uint32_t *__restrict__ words, uint32_t n_rows, uint32_t n_loci, const uint32_t *__restrict__ pos,
                                         uint32_t cov_lo, uint32_t cov_hi, uint32_t *__restrict__ rows, uint32_t W, uint32_t woff) {
    uint32_t nw = (n_loci + 31) >> 5;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_rows * nw) return;
    uint32_t r = (uint32_t)(t / nw), w = (uint32_t)(t % nw);
    uint32_t m = 0;
    for (uint32_t b = 0; b < 32; b++) {
        uint32_t s = w * 32 + b;
        if (s < n_loci) { uint32_t p = pos[s]; if (p >= cov_lo && p < cov_hi) m |= 1u << b; }
    }
    rows[(uint64_t)r * W + woff + w] = words[t] & m;
}

__global__ void unpack_rows_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ rowmap, uint32_t W, uint32_t woff, uint32_t n_rows,
                                   uint32_t n_loci, uint8_t *__restrict__ alleles) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_rows * n_loci) return;
    uint32_t r = (uint32_t)(t / n_loci), s = (uint32_t)(t % n_loci);
    if (rowmap) r = rowmap[r >> 1] * 2 + (r & 1);   // logical haplotype row -> physical row (after a migration)
    alleles[t] = (rows[(uint64_t)r * W + woff + (s >> 5)] >> (s & 31)) & 1u;
}

__global__ void gather_packed_chr_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ rowmap, uint32_t W, uint32_t woff, uint32_t n_rows,
                                         uint32_t nw, uint32_t *__restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_rows * nw) return;
    uint32_t r = (uint32_t)(t / nw), w = (uint32_t)(t % nw);
    if (rowmap) r = rowmap[r >> 1] * 2 + (r & 1);
    out[t] = rows[(uint64_t)r * W + woff + w];
}

// ------------------------------------------------------------------------------------------------
// Causal variants: a second, tiny locus set carried as BIT planes, one row of Wcv u32 words per haplotype
// (row 2*i+h).  Block (phenotype f, chromosome c) = CVs [block_off[b], block_off[b+1]) and starts on a word
// boundary, word_off[b]; CV k of block b is bit (k - block_off[b]) of that block's words.  With 1 000 CVs on 22
// chromosomes a row is 176 B, so the whole plane of a 100k population (35 MB) lives in L2.  With more than one
// population the root population of every (haplotype, CV) is carried next to it as a byte plane
// [row][n_cv_tot] because ras_find_cv takes the effect sizes a, d from the root population (:2776-2786).
// ------------------------------------------------------------------------------------------------
struct CvSet {
    int n_chr, n_phen;
    uint32_t n_cv_tot;
    uint32_t Wcv;               // words per row (multiple of 4)
    int sorted;                 // every block lists its positions in ascending order (binary search allowed)
    const uint32_t *block_off;  // [n_phen*n_chr + 1] first CV of each block
    const uint32_t *word_off;   // [n_phen*n_chr + 1] first word of each block
    const uint32_t *word_blk;   // [Wcv] block of each word, 0xFFFFFFFF for padding words
    const uint32_t *bp;         // [n_cv_tot] CV positions
    const uint32_t *chr_of;     // [n_cv_tot]
};

// generation 0: allele = founder CV allele when covered, root = population (ras_find_cv :2752-2815).
// One thread per (row, word).
__global__ void cv_init_kernel(CvSet cs, const uint8_t *__restrict__ founder_cv /* [n_rows][n_cv_tot] */, uint32_t n_rows,
                               const uint32_t *__restrict__ cov_lo, const uint32_t *__restrict__ cov_hi, uint8_t root,
                               uint32_t *__restrict__ bits, uint8_t *__restrict__ rootp) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (uint64_t)n_rows * cs.Wcv) return;
    uint32_t w = (uint32_t)(t % cs.Wcv);
    uint64_t row = t / cs.Wcv;
    uint32_t b = cs.word_blk[w], v = 0;
    if (b != 0xFFFFFFFFu) {
        uint32_t c = b % (uint32_t)cs.n_chr;
        uint32_t k0 = cs.block_off[b] + (w - cs.word_off[b]) * 32u, k1 = min(k0 + 32u, cs.block_off[b + 1]);
        uint32_t lo = cov_lo[c], hi = cov_hi[c];
        for (uint32_t k = k0; k < k1; k++) {
            uint32_t p = cs.bp[k];
            if (p >= lo && p < hi && founder_cv[row * cs.n_cv_tot + k]) v |= 1u << (k - k0);
            if (rootp) rootp[row * cs.n_cv_tot + k] = root;
        }
    }
    bits[t] = v;
}

// One thread per (offspring gamete row, word): the haplotype of CV k is start ^ parity of the crossovers at or
// below its position, so a crossover flips every CV of the word from index lower_bound(bp, crossover) upwards;
// the word is then a mask-merge of the two parental words — the same operation as the boundary chunk of
// propagate_bits_kernel.
__global__ void cv_propagate_bits_kernel(CvSet cs, const StepState *__restrict__ ss, const uint32_t *__restrict__ par_bits, uint32_t *__restrict__ off_bits,
                                         const uint32_t *__restrict__ father, const uint32_t *__restrict__ mother,
                                         const uint64_t *__restrict__ xo_off, const uint32_t *__restrict__ xo_bp,
                                         const uint8_t *__restrict__ start_hap) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n_off = ss->n_off;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_off * 2 * cs.Wcv; t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t w;
    uint64_t row;
    divmod_idx(t, cs.Wcv, row, w);
    uint64_t i = row >> 1;
    int gam = (int)(row & 1);
    uint32_t b = cs.word_blk[w], v = 0;
    if (b != 0xFFFFFFFFu) {
        uint32_t c = b % (uint32_t)cs.n_chr;
        uint32_t k0 = cs.block_off[b] + (w - cs.word_off[b]) * 32u, nk = min(32u, cs.block_off[b + 1] - k0);
        uint64_t slot = (i * (uint64_t)cs.n_chr + c) * 2 + gam;
        uint32_t mask = start_hap[slot] ? 0xFFFFFFFFu : 0u;
        const uint32_t *bp = cs.bp + k0;
        for (uint64_t e = xo_off[slot]; e < xo_off[slot + 1]; e++) {
            const uint32_t x = xo_bp[e];  // CVs at or above the crossover flip
            if (cs.sorted) {
                const uint32_t idx = lower_bound_u32(bp, nk, x);
                if (idx < 32u) mask ^= 0xFFFFFFFFu << idx;
            } else {  // cv.info rows need not be sorted
                uint32_t m = 0;
                for (uint32_t j = 0; j < nk; j++) m |= (uint32_t)(x <= __ldg(bp + j)) << j;
                mask ^= m;
            }
        }
        uint32_t parent = gam ? mother[i] : father[i];
        const uint32_t *pr = par_bits + (uint64_t)parent * 2 * cs.Wcv + w;
        v = (pr[0] & ~mask) | (pr[cs.Wcv] & mask);
    }
    off_bits[(i * 2 + gam) * (uint64_t)cs.Wcv + w] = v;
    }
}

// The same planes, one WARP per offspring gamete row (sorted CV blocks, toggle words in shared memory).  The thread-per-word kernel
// above re-walks the slot's crossovers for every word of the chromosome and ran with 15 of 32 lanes active (ncu, round 2); here the
// ~35 crossovers of the whole gamete are one flat list dealt out to the lanes, each crossover marks "the haplotype flips from CV idx
// on" as ONE toggle bit per phenotype block (idx = lower_bound over the block's positions), and the copy mask of a word is the
// running XOR of the toggle bits before it in its block: a prefix-XOR inside the word (five shifts) and the parity of the block's
// earlier words.
__global__ void __launch_bounds__(256, 6) cv_propagate_rows_kernel(CvSet cs, const StepState *__restrict__ ss, const uint32_t *__restrict__ par_bits, uint32_t *__restrict__ off_bits,
                                                                const uint32_t *__restrict__ father, const uint32_t *__restrict__ mother,
                                                                const uint64_t *__restrict__ xo_off, const uint32_t *__restrict__ xo_bp,
                                                                const uint8_t *__restrict__ start_hap) {
    extern __shared__ uint32_t s_toggle[];   // [warps per CTA][Wcv] toggle words, then [Wcv] chromosome and [Wcv] first block word of every word
    if (ss->err & SE_FATAL) return;
    const uint64_t n_rows = ss->n_off * 2;
    const int lane = threadIdx.x & 31;
    const uint32_t C = (uint32_t)cs.n_chr, Wcv = cs.Wcv;
    uint32_t *T = s_toggle + (threadIdx.x >> 5) * Wcv;
    uint32_t *s_chr = s_toggle + (blockDim.x >> 5) * Wcv, *s_first = s_chr + Wcv;
    for (uint32_t w = threadIdx.x; w < Wcv; w += blockDim.x) {   // (the same for every row this CTA will take)
        const uint32_t b = cs.word_blk[w];
        s_chr[w] = b == 0xFFFFFFFFu ? b : b % C;
        s_first[w] = b == 0xFFFFFFFFu ? 0u : cs.word_off[b];
    }
    for (uint32_t w = lane; w < Wcv; w += 32) T[w] = 0u;
    __syncthreads();
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
        const uint64_t i = row >> 1;
        const uint32_t gam = (uint32_t)(row & 1);
        const uint64_t slot0 = i * C * 2 + gam;     // slot of chromosome c: slot0 + 2c
        for (uint32_t c0 = 0; c0 < C; c0 += 32) {   // the crossovers of up to 32 chromosomes of this gamete, as one flat list
            const uint32_t c = c0 + lane;
            uint64_t e0 = 0;
            uint32_t cnt = 0;
            if (c < C) { const uint64_t slot = slot0 + 2 * c; e0 = xo_off[slot]; cnt = (uint32_t)(xo_off[slot + 1] - e0); }
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            for (uint32_t r0 = 0; r0 < total; r0 += 32) {
                const bool act = r0 + lane < total;
                const uint32_t r = min(r0 + lane, total - 1);
                int lo = 0, hi = 31;   // owner: the first lane whose inclusive prefix exceeds r
#pragma unroll
                for (int step = 0; step < 5; step++) {
                    const int mid = (lo + hi) >> 1;
                    const uint32_t v = __shfl_sync(0xffffffffu, incl, mid);
                    if (v > r) hi = mid; else lo = mid + 1;
                }
                const uint32_t q = r - (__shfl_sync(0xffffffffu, incl, lo) - __shfl_sync(0xffffffffu, cnt, lo));
                const uint64_t e = __shfl_sync(0xffffffffu, e0, lo) + q;
                if (act) {
                    const uint32_t x = xo_bp[e];
                    for (int f = 0; f < cs.n_phen; f++) {
                        const uint32_t b = (uint32_t)f * C + c0 + (uint32_t)lo;
                        const uint32_t k0 = cs.block_off[b], nk = cs.block_off[b + 1] - k0;
                        const uint32_t idx = lower_bound_u32(cs.bp + k0, nk, x);   // CVs at or above the crossover flip
                        if (idx < nk) atomicXor(&T[cs.word_off[b] + (idx >> 5)], 1u << (idx & 31u));
                    }
                }
            }
        }
        __syncwarp();
        const uint32_t parent = gam ? mother[i] : father[i];
        const uint32_t *pr = par_bits + (uint64_t)parent * 2 * Wcv;
        for (uint32_t w = lane; w < Wcv; w += 32) {
            const uint32_t c = s_chr[w];
            uint32_t v = 0;
            if (c != 0xFFFFFFFFu) {
                uint32_t m = T[w];
                m ^= m << 1; m ^= m << 2; m ^= m << 4; m ^= m << 8; m ^= m << 16;
                uint32_t par = start_hap[slot0 + 2 * c] & 1u;
                for (uint32_t w2 = s_first[w]; w2 < w; w2++) par ^= (uint32_t)__popc(T[w2]) & 1u;
                if (par) m = ~m;
                v = (pr[w] & ~m) | (pr[Wcv + w] & m);
            }
            off_bits[row * Wcv + w] = v;
        }
        __syncwarp();   // every lane has read what it needs of the toggle words: cleared for the next row
        for (uint32_t w = lane; w < Wcv; w += 32) T[w] = 0u;
        __syncwarp();
    }
}

// root-population plane (only with more than one population): one thread per (offspring gamete row, CV)
__global__ void cv_root_propagate_kernel(CvSet cs, const StepState *__restrict__ ss, const uint8_t *__restrict__ par_root, uint8_t *__restrict__ off_root,
                                         const uint32_t *__restrict__ father, const uint32_t *__restrict__ mother,
                                         const uint64_t *__restrict__ xo_off, const uint32_t *__restrict__ xo_bp,
                                         const uint8_t *__restrict__ start_hap) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n_off = ss->n_off;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_off * 2 * cs.n_cv_tot; t += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t k = (uint32_t)(t % cs.n_cv_tot);
    uint64_t row = t / cs.n_cv_tot;
    uint64_t i = row >> 1;
    int gam = (int)(row & 1);
    uint32_t c = cs.chr_of[k], p = cs.bp[k];
    uint64_t slot = (i * (uint64_t)cs.n_chr + c) * 2 + gam;
    int h = start_hap[slot];
    for (uint64_t e = xo_off[slot]; e < xo_off[slot + 1]; e++) h ^= (xo_bp[e] <= p);
    uint32_t parent = gam ? mother[i] : father[i];
    off_root[(i * 2 + gam) * (uint64_t)cs.n_cv_tot + k] = par_root[((uint64_t)parent * 2 + (h & 1)) * cs.n_cv_tot + k];
    }
}

// allele count per CV over the population (frq numerator, :2647-2663).  A CTA takes 32 word columns x rows_per_cta rows:
// a warp reads 32 consecutive words of one row (128 B), 8 warps stride the rows.  Every thread counts the 32 bit columns of its
// word in eight bit-sliced planes (plane l holds bit l of all 32 counters: adding a word is a ripple of AND/XOR pairs, 16
// logic operations instead of 96 shift-mask-adds), unpacks them once at the end, shared-memory atomics fold the 8 warps and
// one 64-bit atomic per CV and CTA goes to memory.  rows_per_cta <= 8 * 255.
__global__ void cv_count_bits_kernel(CvSet cs, const uint32_t *__restrict__ bits, const uint64_t *__restrict__ n_ind, uint32_t rows_per_cta, unsigned long long *__restrict__ count) {
    __shared__ unsigned int sh[32][33];
    const uint32_t w = blockIdx.x * 32 + threadIdx.x;
    const uint64_t n_rows = 2 * *n_ind;
    const uint64_t r0 = (uint64_t)blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, n_rows);
    if (r0 >= n_rows) return;
    for (int q = threadIdx.y; q < 32; q += 8) sh[q][threadIdx.x] = 0;
    __syncthreads();
    uint32_t plane[8];
#pragma unroll
    for (int l = 0; l < 8; l++) plane[l] = 0;
    if (w < cs.Wcv)
        for (uint64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            uint32_t x = bits[r * cs.Wcv + w];
#pragma unroll
            for (int l = 0; l < 8; l++) { const uint32_t carry = plane[l] & x; plane[l] ^= x; x = carry; }
        }
#pragma unroll
    for (int b = 0; b < 32; b++) {
        unsigned int c = 0;
#pragma unroll
        for (int l = 0; l < 8; l++) c |= ((plane[l] >> b) & 1u) << l;
        if (c) atomicAdd(&sh[threadIdx.x][b], c);
    }
    __syncthreads();
    if (w < cs.Wcv) {
        uint32_t blk = cs.word_blk[w];
        if (blk != 0xFFFFFFFFu) {
            uint32_t k0 = cs.block_off[blk] + (w - cs.word_off[blk]) * 32u, k1 = cs.block_off[blk + 1];
            for (uint32_t b = threadIdx.y; b < 32u; b += 8) {
                unsigned int v = sh[threadIdx.x][b];
                if (v && k0 + b < k1) atomicAdd(&count[k0 + b], (unsigned long long)v);
            }
        }
    }
}

// A and D per individual (ras_compute_AD :2686-2746): one warp per individual and phenotype; lanes stride
// the CVs of each chromosome block and the warp reduces once at the end (fp64, fixed shuffle tree).
__global__ void genetic_value_kernel(CvSet cs, const uint32_t *__restrict__ bits, const uint8_t *__restrict__ rootp,
                                     const unsigned long long *__restrict__ count, const uint64_t *__restrict__ n_ind /* individuals (also in frq) */,
                                     const double *__restrict__ a_eff /* [n_pop][n_cv_tot] */, const double *__restrict__ d_eff,
                                     const uint8_t *__restrict__ vd_zero /* [n_phen] */, uint64_t stride /* column stride = capacity */, double *__restrict__ A,
                                     double *__restrict__ D, double *__restrict__ Gv, uint32_t *__restrict__ err) {
    const uint64_t n = *n_ind, n_count = n;
    int lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < n * cs.n_phen; wid += n_warps) {
    uint64_t i = wid % n;
    int f = (int)(wid / n);
    const uint32_t *al0 = bits + (i * 2) * (uint64_t)cs.Wcv, *al1 = al0 + cs.Wcv;
    const uint8_t *r0 = rootp ? rootp + (i * 2) * (uint64_t)cs.n_cv_tot : nullptr, *r1 = rootp ? r0 + cs.n_cv_tot : nullptr;
    const bool no_d = vd_zero[f] != 0;
    const double two_n = (double)(2 * n_count);
    double Ac = 0, Dc = 0;
    for (int c = 0; c < cs.n_chr; c++) {
        const int blk = f * cs.n_chr + c;
        const uint32_t b0 = cs.block_off[blk], b1 = cs.block_off[blk + 1], wo = cs.word_off[blk];
        for (uint32_t k = b0 + lane; k < b1; k += 32) {
            const uint32_t j = k - b0;
            const unsigned t = ((al0[wo + (j >> 5)] >> (j & 31)) & 1u) + ((al1[wo + (j >> 5)] >> (j & 31)) & 1u);
            uint64_t o0 = r0 ? (uint64_t)r0[k] * cs.n_cv_tot + k : k, o1 = r1 ? (uint64_t)r1[k] * cs.n_cv_tot + k : k;
            double a = (a_eff[o0] + a_eff[o1]) / 2;
            double d = no_d ? 0.0 : (d_eff[o0] + d_eff[o1]) / 2;
            double p = (double)count[k] / two_n, q = 1 - p;
            double alpha = a + d * (q - p);
            Ac += ((double)t - 2 * p) * alpha;
            double ct = t == 0 ? -2 * p * p : (t == 1 ? 2 * p * q : -2 * q * q);
            Dc += ct * d;
        }
    }
    for (int o = 16; o; o >>= 1) { Ac += __shfl_xor_sync(0xffffffffu, Ac, o); Dc += __shfl_xor_sync(0xffffffffu, Dc, o); }
    if (lane == 0) {
        A[(uint64_t)f * stride + i] = Ac; D[(uint64_t)f * stride + i] = Dc; Gv[(uint64_t)f * stride + i] = Ac + Dc;
        if (isnan(Ac) || isnan(Dc)) atomicOr(err, (uint32_t)SE_NAN);
    }
    }
}

// One population: the per-CV terms do not depend on the individual, so they are tabulated once per generation —
// LAD[k][t] = {(t - 2p) * alpha, c_t * d} for genotype t in {0,1,2} (the very products of :2691-2712) —
// and the per-individual kernel only gathers and adds them.
__global__ void cv_tables_kernel(CvSet cs, const unsigned long long *__restrict__ count, const uint64_t *__restrict__ n_ind, const double *__restrict__ a_eff,
                                 const double *__restrict__ d_eff, const uint8_t *__restrict__ vd_zero, double2 *__restrict__ LAD) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cs.n_cv_tot) return;
    int blk = 0;
    while (cs.block_off[blk + 1] <= k) blk++;
    const int f = blk / cs.n_chr;
    const double two_n = (double)(2 * *n_ind);
    double a = (a_eff[k] + a_eff[k]) / 2;
    double d = vd_zero[f] ? 0.0 : (d_eff[k] + d_eff[k]) / 2;
    double p = (double)count[k] / two_n, q = 1 - p;
    double alpha = a + d * (q - p);
    for (int t = 0; t < 3; t++) {
        double ct = t == 0 ? -2 * p * p : (t == 1 ? 2 * p * q : -2 * q * q);
        LAD[k * 3 + t] = make_double2(((double)t - 2 * p) * alpha, ct * d);
    }
}
// The per-individual sums take four CVs at a time: a nibble of the two allele words of an individual indexes LG[group][256], the sum of
// the four {LA, LD} terms for that combination of genotypes (CVs a block does not have contribute nothing, their bits are 0 in every
// row).  One 16-byte load and two additions per four CVs; the first version — one table load per CV, located through a bit-position
// list — ran 298 M warp instructions per 250k individuals (ncu launch lists of round 2).
__global__ void cv_group_tables_kernel(CvSet cs, const double2 *__restrict__ LAD, double2 *__restrict__ LG) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cs.Wcv * 8u * 256u) return;
    const uint32_t idx = t & 255u, grp = t >> 8, w = grp >> 3, q = grp & 7u, b = cs.word_blk[w];
    double2 s = make_double2(0.0, 0.0);
    if (b != 0xFFFFFFFFu) {
        const uint32_t k0 = cs.block_off[b] + (w - cs.word_off[b]) * 32u + q * 4u, k1 = cs.block_off[b + 1];
        for (uint32_t j = 0; j < 4u && k0 + j < k1; j++) {
            const uint32_t g = ((idx >> j) & 1u) + ((idx >> (4 + j)) & 1u);
            const double2 v = LAD[(uint64_t)(k0 + j) * 3 + g];
            s.x += v.x; s.y += v.y;
        }
    }
    LG[t] = s;
}
// one warp per individual; lanes stride the half words of the phenotype (its blocks are consecutive words of the row): two allele loads
// and four table loads per lane and round
__global__ void __launch_bounds__(256) genetic_value_groups_kernel(CvSet cs, const uint32_t *__restrict__ bits, const double2 *__restrict__ LG, const uint64_t *__restrict__ n_ind,
                                                                   uint64_t stride /* column stride = capacity */, double *__restrict__ A, double *__restrict__ D, double *__restrict__ Gv,
                                                                   uint32_t *__restrict__ err) {
    const uint64_t n = *n_ind;
    if (n == 0) return;
    const int lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
        const uint32_t *al0 = bits + (i * 2) * (uint64_t)cs.Wcv, *al1 = al0 + cs.Wcv;
        for (int f = 0; f < cs.n_phen; f++) {
            const uint32_t h1 = 2 * cs.word_off[(f + 1) * cs.n_chr];
            double Ac = 0, Dc = 0;
            for (uint32_t h = 2 * cs.word_off[f * cs.n_chr] + lane; h < h1; h += 32) {
                const uint32_t w = h >> 1, sh = (h & 1u) * 16u;
                const uint32_t a0 = al0[w] >> sh, a1 = (al1[w] >> sh) << 4;
                const double2 *G = LG + (uint64_t)h * 1024;   // four groups of 256 entries
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 v = __ldg(G + q * 256 + (((a0 >> (4 * q)) & 15u) | ((a1 >> (4 * q)) & 240u)));
                    Ac += v.x;
                    Dc += v.y;
                }
            }
            for (int o = 16; o; o >>= 1) { Ac += __shfl_xor_sync(0xffffffffu, Ac, o); Dc += __shfl_xor_sync(0xffffffffu, Dc, o); }
            if (lane == 0) {
                A[(uint64_t)f * stride + i] = Ac; D[(uint64_t)f * stride + i] = Dc; Gv[(uint64_t)f * stride + i] = Ac + Dc;
                if (isnan(Ac) || isnan(Dc)) atomicOr(err, (uint32_t)SE_NAN);
            }
        }
    }
}

// bit plane -> bytes of one block (the `--debug` .cvval dump): out[row*ncv + j]
__global__ void cv_unpack_block_kernel(const uint32_t *__restrict__ bits, uint32_t Wcv, uint32_t wo, uint32_t ncv, uint64_t n_rows, uint8_t *__restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * ncv) return;
    uint32_t j = (uint32_t)(t % ncv);
    uint64_t row = t / ncv;
    out[t] = (bits[row * Wcv + wo + (j >> 5)] >> (j & 31)) & 1u;
}

// a count that lives on the device: (*p) * mul + add (p == nullptr: add alone, a host-known count)
// lim: the host-known bound the buffers were sized for — a count a failed step left behind never indexes beyond them
struct DevN {
    const uint64_t *p; uint64_t mul, add, lim;
    __device__ __forceinline__ uint64_t get() const { const uint64_t v = (p ? *p : 0ull) * mul + add; return v < lim ? v : lim; }
};
static inline DevN devn(const uint64_t *p, uint64_t mul = 1, uint64_t add = 0) { return DevN{p, mul, add, ~0ull}; }
static inline DevN hostn(uint64_t n) { return DevN{nullptr, 0, n, ~0ull}; }
static inline DevN limited(DevN n, uint64_t lim) { n.lim = lim < n.lim ? lim : n.lim; return n; }

// ------------------------------------------------------------------------------------------------
// fp64 population moments: sum, then centred sum of squares (two-pass like CommFunc::var, src/CommFunc.cpp:57-68).
// One launch per pass: every CTA leaves its partial sum in a fixed slot, the last CTA to arrive (ticket) adds the
// slots in a fixed order — so the result does not depend on the arrival order — and divides.
// ------------------------------------------------------------------------------------------------
constexpr int MOMENT_MAX_BLOCKS = 1024;
__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double sh[32];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;  // valid in thread 0
}
// scratch: double partial[MOMENT_MAX_BLOCKS] followed by one unsigned ticket (zero between launches)
__device__ __forceinline__ void moment_finish(double s, double *__restrict__ partial, uint64_t n, int minus, double *__restrict__ out) {
    __shared__ bool last;
    unsigned int *ticket = reinterpret_cast<unsigned int *>(partial + MOMENT_MAX_BLOCKS);
    s = block_reduce_sum(s);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {   // lane l adds slots l, l+32, ..., then a shuffle tree: a fixed order
        double t = 0;
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) t += *((volatile double *)(partial + i));
        for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            const double denom = (double)n - (double)minus;
            *out = (denom > 0 && (minus == 0 || n > 1)) ? t / denom : 0.0;
            *ticket = 0;
        }
    }
}
// *out = sum over i < n of x[i] (pow2 = 0, minus = 0: the mean) or of (x[i] - *shift)^2 (pow2 = 1, minus = 1: the variance), / (n - minus)
__global__ void __launch_bounds__(256) moment_kernel(const double *__restrict__ x, DevN dn, const double *__restrict__ shift, int pow2, int minus, double *__restrict__ partial,
                                                      double *__restrict__ out) {
    const uint64_t n = dn.get();
    double s = 0, mu = shift ? *shift : 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double v = x[i] - mu;
        s += pow2 ? v * v : v;
    }
    moment_finish(s, partial, n, minus, out);
}

// ------------------------------------------------------------------------------------------------
// phenotypes: ras_scale_AD_compute_GEF :3075-3206 (elementwise part) and MV/SV :3300-3342
// ------------------------------------------------------------------------------------------------
// environment draws e_i ~ N(0,1) (:3102) and their mean in the same pass
__global__ void __launch_bounds__(256) enoise_mean_kernel(Stream st, const StepState *__restrict__ ss, int pop, int f, const uint64_t *__restrict__ n_ind, double *__restrict__ e,
                                                           double *__restrict__ partial, double *__restrict__ out_mean) {
    const uint64_t n = *n_ind;
    const int gen = ss->gen;
    double s = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double z0, z1;
        normal2(st, P_ENOISE, pop, gen, i, (uint32_t)f, z0, z1);
        e[i] = z0;
        s += z0;
    }
    moment_finish(s, partial, n, 0, out_mean);
}
__global__ void normal_scaled_kernel(Stream st, uint32_t purpose, int pop, int gen, int f, uint64_t n, double sd, double *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double z0, z1;
    normal2(st, purpose, pop, gen, i, (uint32_t)f, z0, z1);
    out[i] = z0 * sd;
}

struct PhenoArgs {
    double s_a, s_d, ve, vf, beta;
    int vt_type;
    // one phenotype, one population: mating and selection values (mv_sv_selection_kernel) in the same pass; mv == nullptr otherwise
    const double *omega, *lambda, *sv0;
    double *mv, *sv, *svf;
};
// ras_selection_func :3386-3428
__device__ __forceinline__ double selection_func(int gen, int func, double par1, double par2, double z) {
    double r = 1.0;
    if (gen != 0) {
        if (func == 0) { double y = exp(0.0 + 1.0 * z); r = y / (1 + y); }
        else if (func == 1) { double y = exp(par1 + par2 * z); r = y / (1 + y); }
        else if (func == 2) r = .5 * (1 + erf((z - par1) / (sqrt(2.0) * par2)));
        else if (func == 3) { const double pi = 3.1415926; double u = (z - par1) / par2; r = 1 / (sqrt(2.0 * pi) * par2) * exp(-0.5 * (u * u)); }
        else if (func == 4) r = z <= par2 ? par1 : 1.0;
    }
    return r;
}
__global__ void phenotype_kernel(PhenoArgs a, const StepState *__restrict__ ss, const uint64_t *__restrict__ n_ind, const uint64_t *__restrict__ prev_n_ptr,
                                 const double *__restrict__ e_raw, const double *__restrict__ var_e,
                                 double *__restrict__ A, double *__restrict__ D, double *__restrict__ G, const double *__restrict__ Cc,
                                 double *__restrict__ E, double *__restrict__ F, double *__restrict__ P,
                                 const uint64_t *__restrict__ ids, const double *__restrict__ prev_P, const double *__restrict__ prev_F,
                                 const double *__restrict__ f0, uint32_t *__restrict__ err) {
    const uint64_t n = *n_ind;
    const int gen = ss->gen;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double s_ev = a.ve > 0 ? sqrt(*var_e / a.ve) : 0.0;
        double e = s_ev > 0 ? e_raw[i] / s_ev : 0.0;
        double av = A[i] / a.s_a;
        double dv = a.s_d > 0 ? D[i] / a.s_d : 0.0;
        double fv = 0.0;
        if (a.vf > 0) {
            if (gen == 0) fv = f0 ? f0[i] : 0.0;
            else {
                const uint64_t prev_n = *prev_n_ptr;
                uint64_t idf = ids[i * 7 + 1], idm = ids[i * 7 + 2];
                if (idf >= prev_n || idm >= prev_n) atomicOr(err, (uint32_t)SE_PARENT_ID);
                else {
                    const double *src = a.vt_type == 1 ? prev_P : prev_F;
                    fv = a.beta * (src[idf] + src[idm]);
                }
            }
        }
        E[i] = e; A[i] = av; D[i] = dv; G[i] = av + dv; F[i] = fv;
        const double p = av + dv + Cc[i] + e + fv;
        P[i] = p;
        if (a.mv) {   // :3300-3342 for the one phenotype
            a.mv[i] = a.omega[0] * p;
            const double s = a.lambda[0] * p, mean0 = a.sv0[0], var0 = a.sv0[1];
            double z = s - mean0;
            if (var0 > 0) z = (s - mean0) / sqrt(var0);
            a.sv[i] = z;
            a.svf[i] = selection_func(gen, ss->sel_func, ss->sel_par1, ss->sel_par2, z);
        }
    }
}

// mating value and raw selection value (:3300-3322); with sv0 = {mean, var} of generation 0 also the standardised selection value
// and its map to a mating probability (ras_selection_func :3386-3428).  Generation 0 runs it twice: raw values, moments, then all.
__global__ void mv_sv_selection_kernel(const StepState *__restrict__ ss, const uint64_t *__restrict__ n_ind, int n_phen, uint64_t stride, const double *__restrict__ P,
                                       const double *__restrict__ omega, const double *__restrict__ lambda, const double *__restrict__ sv0 /* null: raw only */,
                                       double *__restrict__ mv, double *__restrict__ sv, double *__restrict__ svf) {
    const uint64_t n = *n_ind;
    const int gen = ss->gen, func = ss->sel_func;
    const double par1 = ss->sel_par1, par2 = ss->sel_par2;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        double m = 0, s = 0;
        for (int f = 0; f < n_phen; f++) { double p = P[(uint64_t)f * stride + i]; m += omega[f] * p; s += lambda[f] * p; }
        mv[i] = m;
        if (!sv0) { sv[i] = s; continue; }
        const double mean0 = sv0[0], var0 = sv0[1];
        double z = s - mean0;
        if (var0 > 0) z = (s - mean0) / sqrt(var0);
        sv[i] = z;
        svf[i] = selection_func(gen, func, par1, par2, z);
    }
}
__global__ void add_scalar_kernel(double *__restrict__ x, uint64_t n, double v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += v;
}
// ras_save_human_info_to_Pop_info_prev_gen (:3211-3236): phenotype and parental effect of the generation, by position
__global__ void save_prev_kernel(const uint64_t *__restrict__ n_ind, int n_phen, uint64_t stride, const double *__restrict__ P, const double *__restrict__ F,
                                 double *__restrict__ prev_P, double *__restrict__ prev_F, uint64_t *__restrict__ prev_n) {
    const uint64_t n = *n_ind;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * n_phen; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t f = t / n, i = t - f * n;
        prev_P[f * stride + i] = P[f * stride + i];
        prev_F[f * stride + i] = F[f * stride + i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *prev_n = n;
}

// pedigree of the offspring (:2471-2479) — replayed draws; the Philox path writes it in offspring_kernel
__global__ void pedigree_kernel(uint64_t n_off, const uint32_t *__restrict__ father, const uint32_t *__restrict__ mother,
                                const uint64_t *__restrict__ par_ids, uint64_t *__restrict__ ids) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_off) return;
    const uint64_t *fa = par_ids + (uint64_t)father[i] * 7, *mo = par_ids + (uint64_t)mother[i] * 7;
    uint64_t *d = ids + i * 7;
    d[0] = i; d[1] = fa[0]; d[2] = mo[0]; d[3] = fa[1]; d[4] = fa[2]; d[5] = mo[1]; d[6] = mo[2];
}

// ------------------------------------------------------------------------------------------------
// generic exclusive scan of counts into uint64 offsets, with the element count read from device memory
// ------------------------------------------------------------------------------------------------
// what the thread that owns out[n] does with the grand total (besides storing it)
struct NoTotal { __device__ __forceinline__ void operator()(uint64_t) const {} };
struct StoreTotal {   // *dst = total; err |= bit when it exceeds cap
    uint64_t *dst; uint32_t *err; const uint64_t *cap; uint32_t bit;
    __device__ __forceinline__ void operator()(uint64_t t) const { *dst = t; if (cap && t > *cap) atomicOr(err, bit); }
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread -> 2048 per block
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
template <int THREADS>
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t &total) {
    __shared__ uint64_t wsum[THREADS / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    __syncthreads();
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < THREADS / 32 ? wsum[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { uint64_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        if (lane < THREADS / 32) wsum[lane] = w;
    }
    __syncthreads();
    uint64_t base = warp ? wsum[warp - 1] : 0;
    total = wsum[THREADS / 32 - 1];
    return base + x - v;
}
// What is scanned: an array (PtrIn), or values computed on the fly by a functor `uint64_t operator()(uint64_t i) const` — the thinning
// counts and the family sizes are scanned as they are produced, without a kernel and an array in between.
template <class T> struct PtrIn { const T *p; __device__ __forceinline__ uint64_t operator()(uint64_t i) const { return p[i]; } };
// three launches, any size: block sums, scan of the block sums (one CTA), final pass.  Blocks beyond the device-side n idle.
template <class In>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(In in, DevN dn, uint64_t *__restrict__ block_sums) {
    const uint64_t n = dn.get();
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    if (base > n) return;
    uint64_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + (uint64_t)threadIdx.x * SCAN_ITEMS + k; if (i < n) s += in(i); }
    uint64_t total; block_exclusive_scan<SCAN_THREADS>(s, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_single_block_kernel(uint64_t *__restrict__ block_sums, DevN dn) {
    // serial over chunks of SCAN_THREADS entries; nb is small (n / 2048 + 1)
    const uint32_t nb = (uint32_t)(dn.get() / SCAN_TILE) + 1;
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < nb; b += SCAN_THREADS) {
        uint32_t i = b + threadIdx.x;
        uint64_t v = i < nb ? block_sums[i] : 0, total;
        uint64_t ex = block_exclusive_scan<SCAN_THREADS>(v, total);
        uint64_t c = carry;
        if (i < nb) block_sums[i] = ex + c;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}
template <class In, class OnTotal>
__global__ void __launch_bounds__(SCAN_THREADS) scan_final_kernel(In in, DevN dn, const uint64_t *__restrict__ block_sums, uint64_t *__restrict__ out, OnTotal on_total) {
    const uint64_t n = dn.get();
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    if (base > n) return;
    uint64_t v[SCAN_ITEMS], s = 0;
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + (uint64_t)threadIdx.x * SCAN_ITEMS + k; v[k] = i < n ? in(i) : 0; s += v[k]; }
    uint64_t total, ex = block_exclusive_scan<SCAN_THREADS>(s, total);
    uint64_t run = block_sums[blockIdx.x] + ex;
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint64_t i = base + (uint64_t)threadIdx.x * SCAN_ITEMS + k;
        if (i <= n) out[i] = run;                    // out[n] = grand total (everything at or beyond n counts zero)
        if (i == n) on_total(run);
        run += v[k];
    }
}
// one launch, small arrays (the launch-bound configurations): ONE CTA walks the array in chunks with a running carry
constexpr int SCAN1_THREADS = 1024;
constexpr int SCAN1_CHUNK = SCAN1_THREADS * SCAN_ITEMS;
template <class In, class OnTotal>
__global__ void __launch_bounds__(SCAN1_THREADS) scan_one_cta_kernel(In in, DevN dn, uint64_t *__restrict__ out, OnTotal on_total) {
    const uint64_t n = dn.get();
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base <= n; base += SCAN1_CHUNK) {
        uint64_t v[SCAN_ITEMS], s = 0;
        for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + (uint64_t)threadIdx.x * SCAN_ITEMS + k; v[k] = i < n ? in(i) : 0; s += v[k]; }
        uint64_t total, ex = block_exclusive_scan<SCAN1_THREADS>(s, total);
        const uint64_t c = carry;
        uint64_t run = c + ex;
        for (int k = 0; k < SCAN_ITEMS; k++) {
            uint64_t i = base + (uint64_t)threadIdx.x * SCAN_ITEMS + k;
            if (i <= n) out[i] = run;
            if (i == n) on_total(run);
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Philox crossover / mutation sampling: exact skip-sampler over the survival table T[k] = prod_{i<k}(1-p_i)
// (same law as one Bernoulli(p_j) per map row, ras_sim_loc_rec :2973-2995, one uniform per crossover)
// ------------------------------------------------------------------------------------------------
struct MapDev {
    const uint32_t *row_off;  // [n_chr+1] offsets into bp (rows per chromosome)
    const uint32_t *bp;       // concatenated map rows
    const double *T;          // concatenated survival tables, chromosome c starts at row_off[c] + c (R_c + 1 entries)
    // value index of T (exact acceleration of the search below): vb[vb_off[c] + b] = first row k with
    // T[k+1] < 1 - b / vb_scale[c], b = 0 .. B_c where B_c = vb_off[c+1] - vb_off[c] - 1
    const uint32_t *vb_off;   // [n_chr+1]
    const uint32_t *vb;
    const double *vb_scale;   // [n_chr]
    const uint32_t *bp_dist;  // [n_chr]
    const uint32_t *chr_id;   // [n_chr] index in the full genome (Philox counter; differs from c on a sharded context)
};
// first row k >= j with T[k+1] < v, or -1.  T is non-increasing, so {k : T[k+1] < v} is an up-set with minimum k_g(v) and the
// answer is max(j, k_g).  k_g lies between the index entries of v's bucket; one bucket of slack on either side absorbs
// the rounding of the bucket number, and the binary search over that range returns exactly what a search over [0, R) would.
// (the constants of the chromosome come in registers: the sampler is bound by its load instructions, profiles/README.md r2r)
struct ChrTable { const double *T; const uint32_t *vb; double T_last, vb_scale; uint32_t R, B; };
__device__ __forceinline__ ChrTable chr_table(const MapDev &m, uint32_t c) {
    ChrTable t;
    const uint32_t r0 = m.row_off[c], v0 = m.vb_off[c];
    t.R = m.row_off[c + 1] - r0;
    t.T = m.T + r0 + c;
    t.T_last = t.T[t.R];
    t.vb = m.vb + v0;
    t.B = m.vb_off[c + 1] - v0 - 1;
    t.vb_scale = m.vb_scale[c];
    return t;
}
__device__ __forceinline__ long long next_success(const ChrTable &t, uint32_t j, double v) {
    if (j >= t.R || !(t.T_last < v)) return -1;
    const uint32_t b = (uint32_t)fmin((1.0 - v) * t.vb_scale, (double)(t.B - 1));
    uint32_t lo = __ldg(t.vb + (b > 0 ? b - 1 : 0)), hi = __ldg(t.vb + min(b + 2, t.B));
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(t.T + mid + 1) < v) hi = mid; else lo = mid + 1; }
    if (lo < j) lo = j;
    return (long long)lo;
}
__device__ __forceinline__ long long next_success(const MapDev &m, int c, const double *, uint32_t, uint32_t j, double v) { return next_success(chr_table(m, (uint32_t)c), j, v); }
// Counts the crossovers of every slot, writes start_hap and stashes the first XO_STASH positions at a fixed stride; after
// the scan, xo_place_kernel moves the stash into the CSR (re-drawing only the rare longer lists) and converts positions to
// locus indices in the same sweep.
// A slot takes 1 + (its crossovers) draws — 1 for most of chromosome 22, up to a dozen for chromosome 1 — so one thread per slot
// kept 12 of 32 lanes busy (ncu, round 2).  Here a warp owns a chunk of consecutive slots and its lanes take them one at a time: a
// lane whose slot is finished picks the next free slot of the chunk at the end of the round (ballot + rank), so every lane draws in
// every round until the chunk runs dry.  The draws are keyed by (individual, chromosome, gamete, block): the schedule cannot change them.
constexpr int XO_STASH = 8;
__global__ void sample_xo_kernel(Stream st, MapDev m, const StepState *__restrict__ ss, int n_chr, int pop, uint32_t *__restrict__ count,
                                 uint8_t *__restrict__ start_hap, uint32_t *__restrict__ stash) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n_slots = ss->n_off * (uint64_t)n_chr * 2;
    const int gen = ss->gen;
    const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5, warp_id = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    // chunk: long enough to keep the lanes fed (the tail of a chunk idles), short enough that a small population still covers the grid,
    // and the same number of chunks for every warp (1.5 chunks per warp would leave half the warps waiting for the other half)
    const uint64_t per_warp = (n_slots + n_warps - 1) / n_warps, passes = (per_warp + 1023) / 1024;
    const uint64_t even = passes ? ((per_warp + passes - 1) / passes + 31) & ~31ull : 32;
    const uint64_t chunk = even < 32 ? 32 : even;
    for (uint64_t c0 = warp_id * chunk; c0 < n_slots; c0 += n_warps * chunk) {
        const uint64_t end = min(c0 + chunk, n_slots);
        uint64_t next = c0 + 32, slot = c0 + lane, i = 0;
        bool have = slot < end, fresh = have;
        uint32_t c = 0, j = 0, blk = 0, n = 0, key = 0, bp_dist = 0;
        const uint32_t *bp = nullptr;
        ChrTable tab{};
        while (__any_sync(0xffffffffu, have)) {
            if (have) {
                if (fresh) {
                    divmod_idx(slot >> 1, (uint32_t)n_chr, i, c);
                    tab = chr_table(m, c);
                    bp = m.bp + m.row_off[c];
                    bp_dist = m.bp_dist[c];
                    key = m.chr_id[c] * 2u + (uint32_t)(slot & 1);
                    j = 0; blk = 0; n = 0; fresh = false;
                }
                uint32_t w[4];
                draw(st, P_XO, pop, gen, i, key, blk, w);
                if (blk == 0) start_hap[slot] = (uint8_t)(w[3] & 1u);
                blk++;
                bool done = j >= tab.R;
                if (!done) {
                    const double v = (1.0 - u01(w[0], w[1])) * __ldg(tab.T + j);
                    const long long k = next_success(tab, j, v);
                    if (k < 0) done = true;
                    else {
                        if (n < XO_STASH) stash[slot * XO_STASH + n] = __ldg(bp + (uint32_t)k) + (uint32_t)(((uint64_t)w[2] * bp_dist) >> 32);
                        n++;
                        j = (uint32_t)k + 1;
                    }
                }
                if (done) { count[slot] = n; have = false; }
            }
            const unsigned idle = __ballot_sync(0xffffffffu, !have);
            if (idle && next < end) {   // (warp-uniform) idle lanes take the next slots of the chunk, in lane order
                if (!have) { slot = next + __popc(idle & lt_mask); have = slot < end; fresh = have; }
                next += __popc(idle);
            }
        }
    }
}
// stash -> CSR (+ locus indices of the flips when the bit-packed rows are kept); slots longer than the stash re-draw.
// One thread per slot ran at 4.6 of 32 lanes active (ncu, round 2: most slots hold no crossover, a few hold many).  A WARP now takes 32
// consecutive slots and deals their crossovers — one flat list, prefix-summed over the lanes — out one per lane and round: every lane
// moves one position (and looks one locus index up) per round, and the stores of a round are consecutive CSR entries.
__global__ void xo_place_kernel(Stream st, MapDev m, Genome g, const StepState *__restrict__ ss, int n_chr, int pop, const uint64_t *__restrict__ xo_off,
                                const uint32_t *__restrict__ stash, uint32_t *__restrict__ xo_bp, uint32_t *__restrict__ flips) {
    if (ss->err & SE_FATAL) return;   // (also: more crossovers than the draw buffers hold — nothing is written)
    const uint64_t n_slots = ss->n_off * (uint64_t)n_chr * 2;
    const int gen = ss->gen;
    const int lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; w0 < n_slots; w0 += n_warps * 32) {
        const uint64_t slot = w0 + lane;
        uint64_t o = 0;
        uint32_t cnt = 0, cc = 0;
        uint64_t i = 0;
        if (slot < n_slots) {
            o = xo_off[slot];
            cnt = (uint32_t)(xo_off[slot + 1] - o);
            divmod_idx(slot >> 1, (uint32_t)n_chr, i, cc);
        }
        const uint32_t coop = cnt <= XO_STASH ? cnt : 0u;   // what this lane's slot hands to the cooperative rounds
        uint32_t incl = coop;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const uint32_t r = min(r0 + lane, total - 1);   // (lanes beyond the list redo its last entry)
            int lo = 0, hi = 31;                            // owner: the first lane whose inclusive prefix exceeds r
#pragma unroll
            for (int step = 0; step < 5; step++) {
                const int mid = (lo + hi) >> 1;
                const uint32_t v = __shfl_sync(0xffffffffu, incl, mid);
                if (v > r) hi = mid; else lo = mid + 1;
            }
            const uint32_t q = r - (__shfl_sync(0xffffffffu, incl, lo) - __shfl_sync(0xffffffffu, coop, lo));
            const uint64_t oo = __shfl_sync(0xffffffffu, o, lo);
            const uint32_t oc = __shfl_sync(0xffffffffu, cc, lo);
            const uint32_t x = stash[(w0 + lo) * XO_STASH + q];
            xo_bp[oo + q] = x;
            if (flips) flips[oo + q] = locus_lower_bound(g, (int)oc, x);
        }
        if (cnt > XO_STASH) {   // the rare long list: drawn again by its own lane
            const int c = (int)cc, gam = (int)(slot & 1);
            const uint32_t r0 = m.row_off[c], R = m.row_off[c + 1] - r0;
            const double *T = m.T + r0 + c;
            uint32_t j = 0, blk = 0, n = 0;
            for (;;) {
                uint32_t w[4];
                draw(st, P_XO, pop, gen, i, m.chr_id[c] * 2u + (uint32_t)gam, blk++, w);
                if (j >= R) break;
                double v = (1.0 - u01(w[0], w[1])) * T[j];
                long long k = next_success(m, c, T, R, j, v);
                if (k < 0) break;
                uint32_t x = m.bp[r0 + (uint32_t)k] + (uint32_t)(((uint64_t)w[2] * m.bp_dist[c]) >> 32);
                xo_bp[o + n] = x;
                if (flips) flips[o + n] = locus_lower_bound(g, c, x);
                n++;
                j = (uint32_t)k + 1;
            }
        }
    }
}
// the grand total of the crossover scan: n_xo, the interval count of the segment plan, and the capacity check
struct XoTotal {
    StepState *ss; uint64_t slots_per_ind; int dset;
    __device__ __forceinline__ void operator()(uint64_t t) const {
        ss->n_xo = t; ss->n_iv = t + ss->n_off * slots_per_ind;
        if (t > ss->xo_cap) atomicOr(&ss->err, (uint32_t)SE_CAP_XO);
        DrawCounts &d = ss->dc[dset];   // what the bulk stream will read, a generation behind
        d.n_off = ss->n_off; d.n_xo = t; d.n_iv = ss->n_iv; d.fatal = ((ss->err & SE_FATAL) != 0) || t > ss->xo_cap;
    }
};

// mutations: one thread per (offspring, chromosome)
template <bool FILL>
__global__ void sample_mut_kernel(Stream st, MapDev m, const StepState *__restrict__ ss, int n_chr, int pop, uint32_t *__restrict__ count,
                                  const uint64_t *__restrict__ mut_off, uint32_t *__restrict__ mut_bp, uint8_t *__restrict__ mut_gam) {
    if (ss->err & SE_FATAL) return;
    const uint64_t n_items = ss->n_off * (uint64_t)n_chr;
    const int gen = ss->gen;
    for (uint64_t item = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = item / (uint64_t)n_chr;
    int c = (int)(item % (uint64_t)n_chr);
    uint32_t r0 = m.row_off[c], R = m.row_off[c + 1] - r0;
    const double *T = m.T + r0 + c;
    uint32_t j = 1, blk = 0, n = 0;
    uint64_t o = FILL ? mut_off[item] : 0;
    for (;;) {
        if (j >= R) break;
        uint32_t w[4];
        draw(st, P_MUT, pop, gen, i, m.chr_id[c], blk++, w);
        double v = (1.0 - u01(w[0], w[1])) * T[j];
        long long k = next_success(m, c, T, R, j, v);
        if (k < 0) break;
        if (FILL) {
            uint32_t s0 = m.bp[r0 + (uint32_t)k - 1], s1 = m.bp[r0 + (uint32_t)k];
            mut_bp[o + n] = s0 + (uint32_t)(((uint64_t)w[2] * (uint64_t)(s1 - s0 + 1)) >> 32);
            mut_gam[o + n] = (uint8_t)(w[3] & 1u);
        }
        n++;
        j = (uint32_t)k + 1;
    }
    if (!FILL) count[item] = n;
    }
}
__global__ void sex_kernel(Stream st, int pop, int gen, uint64_t n, uint8_t *__restrict__ sex) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[4];
    draw(st, P_SEX, pop, gen, i, 0, 0, w);
    sex[i] = (uint8_t)((w[0] & 1u) + 1);
}

// ------------------------------------------------------------------------------------------------
// per-haplotype mutation lists (the bit path's analogue of part::mutation_pos): inherit by crossover
// parity, append this generation's hits, toggle a locus only on its first hit in the lineage (the reference
// looks positions up with std::find, :1218-1222/:2770-2774, so repeated hits do not toggle back).
// One thread per offspring slot; pass 0 counts, pass 1 fills + applies.
// An entry is a position in bp; bit 31 (HM_BAKED) marks a toggle that a re-based founder panel already holds
// (ge_rebase_founders): it still counts as "seen" for the first-hit rule but is not applied again at materialisation.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t HM_BAKED = 0x80000000u;
struct MutArgs {
    int n_chr;
    const StepState *ss;
    const uint32_t *father, *mother;
    const uint64_t *xo_off; const uint32_t *xo_bp; const uint8_t *start_hap;
    const uint64_t *par_hm_off; const uint32_t *par_hm_bp;   // parent lists, slot (par*n_chr + c)*2 + h; may be null (empty)
    const uint64_t *mut_off; const uint32_t *mut_bp; const uint8_t *mut_gam;  // this generation's hits per (i,c); may be null
    const uint32_t *cov_lo, *cov_hi;                          // [n_chr]
};
__device__ __forceinline__ int parity_at(const MutArgs &a, uint64_t slot, uint32_t pos) {
    int h = a.start_hap[slot];
    for (uint64_t e = a.xo_off[slot]; e < a.xo_off[slot + 1]; e++) h ^= (a.xo_bp[e] <= pos);
    return h & 1;
}
template <bool FILL>
__global__ void mutation_lists_kernel(MutArgs a, Genome g, CvSet cs, uint32_t *__restrict__ count, const uint64_t *__restrict__ hm_off,
                                      uint32_t *__restrict__ hm_bp, uint32_t *__restrict__ off_rows, uint32_t *__restrict__ cv_bits) {
    if (a.ss->err & SE_FATAL) return;
    const uint64_t n_slots = a.ss->n_off * (uint64_t)a.n_chr * 2;
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = (slot >> 1) / (uint64_t)a.n_chr;
    int c = (int)((slot >> 1) % (uint64_t)a.n_chr), gam = (int)(slot & 1);
    uint32_t parent = gam ? a.mother[i] : a.father[i];
    uint32_t n = 0;
    uint64_t o = FILL ? hm_off[slot] : 0;
    if (a.par_hm_off) {
        for (int h = 0; h < 2; h++) {
            uint64_t ps = ((uint64_t)parent * a.n_chr + c) * 2 + h;
            for (uint64_t e = a.par_hm_off[ps]; e < a.par_hm_off[ps + 1]; e++) {
                uint32_t m = a.par_hm_bp[e];
                if (parity_at(a, slot, m & ~HM_BAKED) == h) { if (FILL) hm_bp[o + n] = m; n++; }
            }
        }
    }
    if (a.mut_off) {
        uint64_t item = i * a.n_chr + c;
        for (uint64_t e = a.mut_off[item]; e < a.mut_off[item + 1]; e++) {
            if (a.mut_gam[e] != gam) continue;
            uint32_t m = a.mut_bp[e];
            if (m < a.cov_lo[c] || m >= a.cov_hi[c]) continue;
            if (FILL) {
                bool seen = false;
                for (uint32_t q = 0; q < n; q++) if ((hm_bp[o + q] & ~HM_BAKED) == m) { seen = true; break; }
                hm_bp[o + n] = m;
                if (!seen) {
                    if (off_rows) {
                        const uint32_t *pos = g.pos + g.locus_off[c];
                        uint32_t nl = g.chr_nloci[c];
                        for (uint32_t s = lower_bound_u32(pos, nl, m); s < nl && pos[s] == m; s++)
                            off_rows[(uint64_t)(2 * i + gam) * g.W + g.chr_word_off[c] + (s >> 5)] ^= 1u << (s & 31);
                    }
                    if (cv_bits) {  // the words of block (f, c) of this row belong to this thread alone
                        for (int f = 0; f < cs.n_phen; f++) {
                            const int blk = f * cs.n_chr + c;
                            const uint32_t b0 = cs.block_off[blk], b1 = cs.block_off[blk + 1];
                            for (uint32_t k = b0; k < b1; k++)
                                if (cs.bp[k] == m) cv_bits[(i * 2 + gam) * (uint64_t)cs.Wcv + cs.word_off[blk] + ((k - b0) >> 5)] ^= 1u << ((k - b0) & 31);
                        }
                    }
                }
            }
            n++;
        }
    }
    if (!FILL) count[slot] = n;
    }
}

// ------------------------------------------------------------------------------------------------
// couples -> offspring (reproduce :2394-2493).  Offsets = exclusive scan of the family sizes of the couples that may
// marry (:2402-2406); one thread per couple then writes, for each of its children in birth order, the parents'
// positions, the sex draw (:2472), the pedigree (:2471-2479) and the sibling-common effect (one draw per couple and
// phenotype, :2417-2429, :2481-2484).
// ------------------------------------------------------------------------------------------------
struct FamilyIn {   // family size of couple k as the offspring scan sees it: couples that may not marry have none (:2402-2406)
    const StepState *ss; const uint8_t *inbreed; const int32_t *noff;
    __device__ __forceinline__ uint64_t operator()(uint64_t k) const { return ((ss->err & SE_FATAL) || inbreed[k]) ? 0u : (uint64_t)max(noff[k], 0); }
};
struct OffspringTotal {   // grand total of the family-size scan
    StepState *ss;
    __device__ __forceinline__ void operator()(uint64_t t) const {
        ss->n_off = t;
        if (t == 0) atomicOr(&ss->err, (uint32_t)SE_NO_OFFSPRING);
        if (t > ss->cap) atomicOr(&ss->err, (uint32_t)SE_CAP_OFFSPRING);
    }
};
struct CommonArgs { int n_phen; uint64_t stride; double sd[8]; };   // sd[f] = sqrt(vc) or 0
__global__ void offspring_kernel(Stream st, const StepState *__restrict__ ss, int pop, const uint64_t *__restrict__ fam_off, const uint32_t *__restrict__ male,
                                 const uint32_t *__restrict__ female, const uint64_t *__restrict__ par_ids, CommonArgs ca, uint32_t *__restrict__ father,
                                 uint32_t *__restrict__ mother, uint32_t *__restrict__ couple_of, uint8_t *__restrict__ sex, uint64_t *__restrict__ ids,
                                 double *__restrict__ Cc, uint64_t *__restrict__ n_new /* size of the offspring generation */) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_new = (ss->err & SE_FATAL) ? 0 : ss->n_off;
    if (ss->err & SE_FATAL) return;
    const uint64_t n_couples = ss->n_couples;
    const int gen = ss->gen;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_couples; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i0 = fam_off[k], i1 = fam_off[k + 1];
        if (i0 == i1) continue;
        const uint32_t pm = male[k], pf = female[k];
        const uint64_t *fa = par_ids + (uint64_t)pm * 7, *mo = par_ids + (uint64_t)pf * 7;
        const uint64_t a0 = fa[0], a1 = fa[1], a2 = fa[2], b0 = mo[0], b1 = mo[1], b2 = mo[2];
        double cm[8];
        for (int f = 0; f < ca.n_phen; f++) {
            cm[f] = 0.0;
            if (ca.sd[f] > 0) { double z0, z1; normal2(st, P_COMMON, pop, gen, k, (uint32_t)f, z0, z1); cm[f] = z0 * ca.sd[f]; }
        }
        for (uint64_t i = i0; i < i1; i++) {
            father[i] = pm; mother[i] = pf; couple_of[i] = (uint32_t)k;
            uint32_t w[4];
            draw(st, P_SEX, pop, gen, i, 0, 0, w);
            sex[i] = (uint8_t)((w[0] & 1u) + 1);
            uint64_t *d = ids + i * 7;
            d[0] = i; d[1] = a0; d[2] = b0; d[3] = a1; d[4] = a2; d[5] = b1; d[6] = b2;
            for (int f = 0; f < ca.n_phen; f++) Cc[(uint64_t)f * ca.stride + i] = cm[f];
        }
    }
}

}  // namespace gek
