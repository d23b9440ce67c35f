// ge_segments.cuh — founder-segment representation (GE_REP_SEGMENTS): the reference's own chromosome model,
// a sorted list of `class part {st,en,hap_index,root_population}` per haplotype (src/Population.h:20-51).
// Needed where bit-packed rows cannot fit (BASELINE config 5: 1M individuals x 10M loci = 2.5 TB per
// generation) and for the `.int` output; its cost is proportional to crossovers, not to loci.
//
// Device layout: one CSR over haplotype slots (i*n_chr + c)*2 + h.  A part is either the reference's four fields
// (uint4, 16 B) or, wherever lists are sorted tilings, the packed form {st, hap_index | root_population << 27} (uint2, 8 B)
// whose end is the next part's start (see part_get below; the format is fixed in seg_init_gen0).
// part::mutation_pos lives in the per-haplotype mutation lists shared with the bit path
// (ge_kernels.cuh, mutation_lists_kernel): inside a haplotype the parts are disjoint, so "position is in the
// part that covers it" and "position is in the haplotype's list" are the same predicate.
//
// Kernels: seg_recombine_kernel (one thread per slot, the reference's loop verbatim: maps that can give unsorted lists,
// and the cross-check of the GPU tests) and the default seg_plan_kernel + seg_gather_kernel.  Sizes (offspring, crossovers,
// intervals, parts) are read from StepState on the device; grids are sized from the capacity.
#pragma once
#include "ge_context.cuh"

namespace gek {

// part e of a list that ends at e_end, in either format, as {st, en, hap_index, root_population}
constexpr uint32_t SEG_ID_BITS = 27;   // packed id = hap_index | root_population << 27
__device__ __forceinline__ uint4 part_get(const uint4 *__restrict__ seg, uint64_t e, uint64_t, uint32_t) { return seg[e]; }
__device__ __forceinline__ uint4 part_get(const uint2 *__restrict__ seg, uint64_t e, uint64_t e_end, uint32_t hi_c) {
    const uint2 q = seg[e];
    return make_uint4(q.x, e + 1 < e_end ? seg[e + 1].x : hi_c, q.y & ((1u << SEG_ID_BITS) - 1u), q.y >> SEG_ID_BITS);
}
__device__ __forceinline__ void part_put(uint4 *__restrict__ seg, uint64_t e, const uint4 v) { seg[e] = v; }
__device__ __forceinline__ void part_put(uint2 *__restrict__ seg, uint64_t e, const uint4 v) { seg[e] = make_uint2(v.x, v.z | (v.w << SEG_ID_BITS)); }

template <class T>
__global__ void seg_init_kernel(uint64_t n, int n_chr, int pop, const uint32_t *__restrict__ cov_lo, const uint32_t *__restrict__ cov_hi,
                                uint64_t *__restrict__ off, T *__restrict__ seg) {
    uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t n_slots = n * n_chr * 2;
    if (slot > n_slots) return;
    off[slot] = slot;
    if (slot == n_slots) return;
    uint64_t i = (slot >> 1) / (uint64_t)n_chr;
    int c = (int)((slot >> 1) % (uint64_t)n_chr), h = (int)(slot & 1);
    part_put(seg, slot, make_uint4(cov_lo[c], cov_hi[c], (uint32_t)(2 * i + h), (uint32_t)pop));  // :3029-3034
}

struct SegArgs {
    int n_chr;
    const StepState *ss;      // live error bits of the segment lists (set by the bulk stream itself, a generation earlier)
    const DrawCounts *dc;     // sizes of the draw set being recombined (the bulk stream runs behind the control stream's counters)
    __device__ __forceinline__ bool dead() const { return dc->fatal || (ss->err & (SE_CAP_SEG | SE_SEG_UNSORTED)); }
    const uint32_t *father, *mother;
    const uint64_t *xo_off; const uint32_t *xo_bp; const uint8_t *start_hap;
    const uint64_t *par_off; const void *par_seg;   // uint4 parts, or uint2 packed parts (plan + gather only)
    const uint32_t *cov_lo, *cov_hi;
};

// Simulation::recombine (:2903-2958), branch for branch.  One thread per offspring haplotype slot; pass 0 counts
// the pieces, pass 1 writes them at the scanned offsets.  No merging of adjacent same-founder pieces, zero-length
// and clipped pieces exactly as the reference emits them.
// The reference rescans the parental list from part 0 for every interval (O(parts x crossovers)); the interval starts
// L ascend, so one cursor per parental haplotype gives the same first part in O(parts + crossovers) — at generation
// 100 a list holds ~160 parts.  A descending L (only possible with maps whose rows are closer than bp_dist_in_rmap)
// resets the cursors, which restores the reference's rescan.
template <bool FILL>
__global__ void seg_recombine_kernel(SegArgs a, uint32_t *__restrict__ count, const uint64_t *__restrict__ off_off, uint4 *__restrict__ off_seg, uint64_t cap) {
    if (a.dead()) return;
    const uint64_t n_total = a.dc->n_off * a.n_chr * 2;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_total; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t slot = t;
        if (FILL && off_off[slot + 1] > cap) continue;   // (a generation that outgrew the buffer: reported through StepState::err)
        uint64_t i = (slot >> 1) / (uint64_t)a.n_chr;
        int c = (int)((slot >> 1) % (uint64_t)a.n_chr), gam = (int)(slot & 1);
        uint32_t parent = gam ? a.mother[i] : a.father[i];
        uint64_t ps = ((uint64_t)parent * a.n_chr + c) * 2;
        const uint64_t o0 = a.par_off[ps], o1 = a.par_off[ps + 1], o2 = a.par_off[ps + 2];
        const uint4 *H0 = static_cast<const uint4 *>(a.par_seg) + o0, *H1 = static_cast<const uint4 *>(a.par_seg) + o1;
        const uint32_t n0 = (uint32_t)(o1 - o0), n1 = (uint32_t)(o2 - o1);
        uint64_t e0 = a.xo_off[slot];
        uint32_t k = (uint32_t)(a.xo_off[slot + 1] - e0);
        int hi = a.start_hap[slot] & 1;
        uint4 *out = FILL ? off_seg + off_off[slot] : nullptr;
        uint32_t n = 0;
        if (k == 0) {  // recombination_locs.size() < 3: the chosen parental haplotype unchanged (:2910)
            const uint4 *H = hi ? H1 : H0;
            n = hi ? n1 : n0;
            if (FILL) for (uint32_t q = 0; q < n; q++) out[q] = H[q];
        } else {
            uint32_t cur0 = 0, cur1 = 0;   // first part of each parental haplotype that can still end after L
            uint32_t prevL = 0;
            for (uint32_t i1 = 1; i1 <= k + 1; i1++) {
                uint32_t L = i1 == 1 ? a.cov_lo[c] : a.xo_bp[e0 + i1 - 2];
                uint32_t R = i1 == k + 1 ? a.cov_hi[c] : a.xo_bp[e0 + i1 - 1];
                if (L < prevL) { cur0 = 0; cur1 = 0; }
                prevL = L;
                const uint4 *H = hi ? H1 : H0;
                uint32_t nH = hi ? n1 : n0, i2 = hi ? cur1 : cur0;
                while (i2 < nH && H[i2].y <= L) i2++;
                if (hi) cur1 = i2; else cur0 = i2;
                if (i2 < nH) {
                    uint4 q = H[i2];
                    if (q.x < L && L < q.y && R < q.y) { if (FILL) out[n] = make_uint4(L, R, q.z, q.w); n++; i2++; }
                }
                if (i2 < nH) {
                    uint4 q = H[i2];
                    if (q.x < L && L < q.y && R >= q.y) { if (FILL) out[n] = make_uint4(L, q.y, q.z, q.w); n++; i2++; }
                }
                while (i2 < nH) {
                    uint4 q = H[i2];
                    if (!(q.y <= R && L <= q.x)) break;
                    if (FILL) out[n] = q;
                    n++; i2++;
                }
                if (i2 < nH) {
                    uint4 q = H[i2];
                    if (q.x < R && R < q.y) { if (FILL) out[n] = make_uint4(q.x, R, q.z, q.w); n++; }
                }
                hi ^= 1;
            }
        }
        if (!FILL) count[slot] = n;
    }
}

// ------------------------------------------------------------------------------------------------
// Plan + gather: the form the segment path runs by default (every list length).  ncu on the walk kernels above
// (1M individuals, generation 41) showed both passes ISSUE-bound (81 % / 65 % issue-active at 1.6 / 2.9 TB/s of DRAM
// traffic): the walk evaluates the clip predicate twice per part (count, fill) and reads the other haplotype's parts
// just to skip them.  Here the first pass only COUNTS prefixes and the second is a clipped copy:
//
//  positions   P_0 = cov_lo, P_j = j-th crossover (j = 1..k), P_{k+1} = cov_hi; interval j = [P_j, P_{j+1}) on
//              haplotype start_hap ^ (j & 1).
//  predicate   part (x, y) is emitted for [L, R) iff  y > L  and  (y <= R  or  x < R)  — the union of the four branches of
//              :2922-2954 for L <= R — and what is emitted is (max(x, L), min(y, R)).
//  sorted      parental lists are sorted tilings (x and y non-decreasing), so the emitted parts of an interval are the index
//              range [i0, i1) with i0 = #(y <= L), i1 = max(#(y <= R), #(x < R)).
//
//  seg_plan_kernel    one thread per slot: for every interval two binary searches over the .y column of its haplotype (from
//                     that haplotype's previous end); writes a 16-byte copy descriptor and the part count of every interval.
//  (scan of the interval counts -> absolute output offset of every interval; seg_slot_offsets_kernel -> the new CSR offsets)
//  seg_gather_kernel  flat clipped copy over all intervals, load-balanced by output part, fully coalesced stores.
// Warp-per-slot forms were built and measured first in round 1 (a walk with ballots in both passes, balloted prefix counts, a
// per-slot gather; DESIGN.md §3 has the numbers): the per-slot prologue — a division, six dependent loads — is what all of them
// paid in both passes; here only the thread-per-slot plan pays it.
// Slots whose positions do not ascend are done by one thread with the reference's loop verbatim (descriptor y = 0xFFFFFFFF).
// ------------------------------------------------------------------------------------------------
template <bool FILL>
__device__ __forceinline__ uint32_t seg_recombine_verbatim(const uint4 *H0, uint32_t n0, const uint4 *H1, uint32_t n1, const uint32_t *__restrict__ xo, uint32_t k,
                                                           uint32_t lo_c, uint32_t hi_c, int hi, uint4 *out) {
    uint32_t n = 0, cur0 = 0, cur1 = 0, prevL = 0;
    for (uint32_t i1 = 1; i1 <= k + 1; i1++) {
        uint32_t L = i1 == 1 ? lo_c : xo[i1 - 2];
        uint32_t R = i1 == k + 1 ? hi_c : xo[i1 - 1];
        if (L < prevL) { cur0 = 0; cur1 = 0; }
        prevL = L;
        const uint4 *H = hi ? H1 : H0;
        uint32_t nH = hi ? n1 : n0, i2 = hi ? cur1 : cur0;
        while (i2 < nH && H[i2].y <= L) i2++;
        if (hi) cur1 = i2; else cur0 = i2;
        if (i2 < nH) { uint4 q = H[i2]; if (q.x < L && L < q.y && R < q.y) { if (FILL) out[n] = make_uint4(L, R, q.z, q.w); n++; i2++; } }
        if (i2 < nH) { uint4 q = H[i2]; if (q.x < L && L < q.y && R >= q.y) { if (FILL) out[n] = make_uint4(L, q.y, q.z, q.w); n++; i2++; } }
        while (i2 < nH) { uint4 q = H[i2]; if (!(q.y <= R && L <= q.x)) break; if (FILL) out[n] = q; n++; i2++; }
        if (i2 < nH) { uint4 q = H[i2]; if (q.x < R && R < q.y) { if (FILL) out[n] = make_uint4(q.x, R, q.z, q.w); n++; } }
        hi ^= 1;
    }
    return n;
}

struct SegSlot {   // what both passes need to know about one offspring haplotype slot
    uint64_t b0, b1;   // absolute index of the first part of the two parental lists
    uint32_t n0, n1, k, c;
    uint64_t e0, slot;
    int hi;
};
__device__ __forceinline__ SegSlot seg_slot(const SegArgs &a, uint64_t t) {
    SegSlot s;
    const uint32_t per = (uint32_t)a.n_chr * 2u;
    uint64_t i; uint32_t r;
    if (t <= 0xFFFFFFFFull) { const uint32_t t32 = (uint32_t)t, q = t32 / per; i = q; r = t32 - q * per; }   // 32-bit division: the 64-bit one is ~100 instructions
    else { i = t / per; r = (uint32_t)(t - i * per); }
    s.slot = t;
    s.c = r >> 1;
    const uint32_t parent = (r & 1u) ? a.mother[i] : a.father[i];
    const uint64_t ps = ((uint64_t)parent * a.n_chr + s.c) * 2;
    const uint64_t o0 = a.par_off[ps], o1 = a.par_off[ps + 1], o2 = a.par_off[ps + 2];
    s.b0 = o0; s.b1 = o1;
    s.n0 = (uint32_t)(o1 - o0); s.n1 = (uint32_t)(o2 - o1);
    s.e0 = a.xo_off[s.slot];
    s.k = (uint32_t)(a.xo_off[s.slot + 1] - s.e0);
    s.hi = a.start_hap[s.slot] & 1;
    return s;
}

constexpr uint32_t SEG_PLAN_VERBATIM = 0xFFFFFFFFu;

// Two part formats.  uint4 {st, en, hap_index, root_population} is the reference's `class part`.  uint2 {st, id} is the packed
// form the segment path uses whenever lists are sorted tilings (every context except maps with rows closer than bp_dist_in_rmap
// and the forced walk kernels): consecutive parts of a haplotype are contiguous (en_i = st_{i+1}, the last one ends at cov_hi
// — recombine emits them that way, clip by clip), so en is implied and id = hap_index | root_population << 27.  Half the bytes
// of the one HBM-bound kernel of the path, and half the footprint.
__device__ __forceinline__ uint32_t part_y(const uint4 *__restrict__ H, uint32_t i, uint32_t, uint32_t) { return __ldg(&H[i].y); }
__device__ __forceinline__ uint32_t part_y(const uint2 *__restrict__ H, uint32_t i, uint32_t n, uint32_t hi_c) { return i + 1 < n ? __ldg(&H[i + 1].x) : hi_c; }

// #(y <= X) among parts [lo, n) of a list with non-decreasing y, plus lo — two searches with independent loads in flight
// (XA <= XB in every caller that matters, not required)
template <class T>
__device__ __forceinline__ void seg_count_y_le2(const T *__restrict__ H, uint32_t n, uint32_t hi_c, uint32_t XA, uint32_t XB, uint32_t loA, uint32_t hiA, uint32_t loB, uint32_t hiB,
                                                uint32_t &rA, uint32_t &rB) {   // [loA, hiA), [loB, hiB): windows known to hold the two answers
    while (loA < hiA || loB < hiB) {
        const bool actA = loA < hiA, actB = loB < hiB;
        const uint32_t mA = (loA + hiA) >> 1, mB = (loB + hiB) >> 1;
        uint32_t yA = 0, yB = 0;
        if (actA) yA = part_y(H, mA, n, hi_c);
        if (actB) yB = part_y(H, mB, n, hi_c);
        if (actA) { if (yA <= XA) loA = mA + 1; else hiA = mA; }
        if (actB) { if (yB <= XB) loB = mB + 1; else hiB = mB; }
    }
    rA = loA; rB = loB;
}

// The plan.  Interval g = xo_off[slot] + slot + j gets a copy descriptor {absolute index of its first parental part (64 bit), L, R} and
// its part count; the scan of the counts gives every interval its absolute output offset, so the copy itself knows nothing of slots.
// Round 1 ran one THREAD per slot through its k + 1 intervals: with k = 0 for half the slots and up to a dozen for chromosome 1, ncu
// (profiles/r2k_seg_plan_gather_ncu_full_raw.csv) found 11 of 32 lanes active in the search loop and the kernel issue-bound on that
// waste (1.13 ms for 250k individuals at generation 41).  Now a WARP takes 32 consecutive slots: every lane reads its slot's header
// and does the per-slot checks, the headers go to shared memory, and the intervals of all 32 slots — a flat list, prefix-summed —
// are dealt out to the lanes one each per round, so every lane of every round does exactly one interval: two binary searches, side
// by side, over the .y column of its haplotype.  Slots without a crossover (one unclipped interval) and slots whose positions do not
// ascend (the reference's loop verbatim, or refused with packed parts) are finished by their own lane.
template <class T>
__global__ void __launch_bounds__(128) seg_plan_kernel(SegArgs a, uint32_t *__restrict__ iv_count, uint4 *__restrict__ desc, unsigned int *__restrict__ n_verbatim,
                                                       uint64_t *__restrict__ verb_list, uint32_t verb_cap, uint32_t *__restrict__ err) {
    constexpr bool PACKED = sizeof(T) == 8;
    if (a.dead()) return;   // also: the parental lists outgrew their buffer
    __shared__ SegSlot s_slot[4][32];
    const T *par = static_cast<const T *>(a.par_seg);
    const uint64_t n_total = a.dc->n_off * a.n_chr * 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SegSlot *mine = s_slot[warp];
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w0 = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; w0 < n_total; w0 += n_warps * 32) {
        const uint64_t t = w0 + lane;
        uint32_t cnt = 0;   // intervals this lane's slot hands to the cooperative rounds
        if (t < n_total) {
            const SegSlot s = seg_slot(a, t);
            const uint64_t g = s.e0 + s.slot;
            if (s.k == 0) {   // the chosen parental haplotype unchanged (:2910): one unclipped interval
                const uint64_t src = s.hi ? s.b1 : s.b0;
                desc[g] = make_uint4((uint32_t)src, (uint32_t)(src >> 32), 0u, 0xFFFFFFFFu);
                iv_count[g] = s.hi ? s.n1 : s.n0;
            } else {
                const T *H0 = par + s.b0, *H1 = par + s.b1;
                const uint32_t lo_c = a.cov_lo[s.c], hi_c = a.cov_hi[s.c];
                const uint32_t *xo = a.xo_bp + s.e0;
                bool fast = true;   // positions must ascend: X_1 <= ... <= X_k (positions below cov_lo only make the first interval empty)
                uint32_t prev = 0;
                for (uint32_t j = 0; j < s.k; j++) { const uint32_t x = __ldg(xo + j); fast &= x >= prev; prev = x; }
                if (fast && prev > hi_c) {
                    // a crossover in the last map row lies beyond cov_hi, so the last interval has L > R.  The reference's loop first skips
                    // every part with y <= L: if that is the whole list (the normal case, lists end at cov_hi) the interval emits nothing,
                    // which is what the index range gives (i0 = n).  Anything else goes to the verbatim loop.
                    const int hl = s.hi ^ (int)(s.k & 1u);
                    const uint32_t nl = hl ? s.n1 : s.n0;
                    fast = nl == 0 || part_y(hl ? H1 : H0, nl - 1, nl, hi_c) <= prev;
                }
                if (fast) { cnt = s.k + 1; mine[lane] = s; }
                else if constexpr (PACKED) {   // pieces that do not tile cannot be stored with an implied end: refuse (ge_last_error names the 16-byte format)
                    atomicOr(err, (uint32_t)SE_SEG_UNSORTED);
                    for (uint32_t j = 0; j <= s.k; j++) { desc[g + j] = make_uint4(0u, 0u, 0u, 0u); iv_count[g + j] = 0u; }
                } else {   // the reference's loop verbatim (seg_verbatim_fill_kernel writes the parts)
                    const uint32_t n = seg_recombine_verbatim<false>(H0, s.n0, H1, s.n1, xo, s.k, lo_c, hi_c, s.hi, nullptr);
                    desc[g] = make_uint4(0u, SEG_PLAN_VERBATIM, 0u, 0u);
                    iv_count[g] = n;
                    for (uint32_t j = 1; j <= s.k; j++) { desc[g + j] = make_uint4(0u, 0u, 0u, 0u); iv_count[g + j] = 0u; }
                    const unsigned int w = atomicAdd(n_verbatim, 1u);
                    if (w < verb_cap) verb_list[w] = t;
                }
            }
        }
        // the warp's intervals as one flat list: inclusive prefix of the per-lane counts
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        for (uint32_t r0 = 0; r0 < total; r0 += 32) {
            const uint32_t r = min(r0 + lane, total - 1);   // (lanes beyond the list redo its last interval: no divergence, the same stores)
            int lo = 0, hi = 31;                            // owner: the first lane whose inclusive prefix exceeds r
#pragma unroll
            for (int step = 0; step < 5; step++) {
                const int mid = (lo + hi) >> 1;
                const uint32_t v = __shfl_sync(0xffffffffu, incl, mid);
                if (v > r) hi = mid; else lo = mid + 1;
            }
            const uint32_t owner_incl = __shfl_sync(0xffffffffu, incl, lo), owner_cnt = __shfl_sync(0xffffffffu, cnt, lo);
            const uint32_t j = r - (owner_incl - owner_cnt);
            const SegSlot &s = mine[lo];
            const uint32_t lo_c = a.cov_lo[s.c], hi_c = a.cov_hi[s.c];
            const uint32_t *xo = a.xo_bp + s.e0;
            const uint32_t L = j == 0 ? lo_c : __ldg(xo + j - 1), R = j == s.k ? hi_c : __ldg(xo + j);
            const int h = s.hi ^ (int)(j & 1u);
            const T *H = par + (h ? s.b1 : s.b0);
            const uint32_t nH = h ? s.n1 : s.n0;
            // first part with y > L; #(y <= R) — searched side by side.  The two ends of a chromosome need no search (one probe says so): the
            // first interval starts at the first part, the last one ends with the last — which leaves two searches per CROSSOVER, not per
            // interval, and the divergent probes of the searches are what the kernel waits for (552 M warp instructions in 1.14 ms).
            uint32_t aHi = nH, bLo = 0;
            if (nH) {
                if (j == 0 && part_y(H, 0u, nH, hi_c) > L) aHi = 0;
                if (j == s.k && (PACKED || part_y(H, nH - 1, nH, hi_c) <= R)) bLo = nH;
            }
            uint32_t i0, c;
            seg_count_y_le2(H, nH, hi_c, L, R, 0u, aHi, bLo, nH, i0, c);
            uint32_t i1 = max(c, i0);                                         // max(#(y <= R), #(x < R)), x non-decreasing; L > R (first interval with a
            while (i1 < nH && __ldg(&H[i1].x) < R) i1++;                      // position below cov_lo, last one beyond cov_hi) gives an empty range
            const uint64_t src = (h ? s.b1 : s.b0) + i0;
            const uint64_t g = s.e0 + s.slot + j;
            desc[g] = make_uint4((uint32_t)src, (uint32_t)(src >> 32), L, R);
            iv_count[g] = i1 - i0;
        }
        __syncwarp();   // the headers are overwritten in the next round of 32 slots
    }
}

// CSR offsets of the new generation: slot -> output offset of its first interval
__global__ void seg_slot_offsets_kernel(SegArgs a, const uint64_t *__restrict__ xo_off, const uint64_t *__restrict__ iv_off, uint64_t *__restrict__ off) {
    if (a.dead()) return;
    const uint64_t n_slots = a.dc->n_off * (uint64_t)a.n_chr * 2;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= n_slots; t += (uint64_t)gridDim.x * blockDim.x) off[t] = iv_off[xo_off[t] + t];
}
// grand total of a part-count scan: the size of the new generation's lists, and whether they fit
struct SegTotal {
    StepState *ss; int k; uint64_t cap;   // cap == 0: no limit (the host sizes the buffer from the total)
    __device__ __forceinline__ void operator()(uint64_t t) const {
        ss->n_seg[k] = t;
        if (cap && t > cap) atomicOr(&ss->err, (uint32_t)SE_CAP_SEG);
    }
};

// The copy: a CTA takes SEG_GATHER_IV consecutive intervals (descriptors and offsets staged in shared memory with coalesced
// loads) and its threads walk the flat run of output parts those intervals own — every thread finds the interval of its output
// index by a branch-free binary search of the staged offsets, four independent 16-byte loads per thread are in flight before the
// first clip, and the stores of a warp are 512 contiguous bytes.  No per-slot prologue, no idle lanes on short intervals.
constexpr int SEG_GATHER_IV = 128;
__device__ __forceinline__ uint2 ld_stream(const uint2 *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(uint2 *p, const uint2 v) { asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory"); }
__device__ __forceinline__ void part_clip(uint4 &q, uint32_t L, uint32_t R) { q.x = max(q.x, L); q.y = min(q.y, R); }
__device__ __forceinline__ void part_clip(uint2 &q, uint32_t L, uint32_t) { q.x = max(q.x, L); }   // the end is the next part's start

template <class T, int DEPTH>
__global__ void __launch_bounds__(256) seg_gather_kernel(SegArgs a, const uint4 *__restrict__ desc, const uint64_t *__restrict__ iv_off,
                                                         const T *__restrict__ par_seg, T *__restrict__ off_seg, uint64_t cap) {
    if (a.dead()) return;
    const uint64_t n_iv = a.dc->n_iv;
    constexpr int TILE = 256 * 16;                 // outputs per owner tile: one uint4 of owner bytes per thread
    static_assert(SEG_GATHER_IV <= 256 && (16 % DEPTH) == 0, "interval ids are bytes; a tile is a whole number of load batches");
    __shared__ uint32_t s_rel[SEG_GATHER_IV + 1];
    __shared__ uint4 s_desc[SEG_GATHER_IV];
    __shared__ __align__(16) uint8_t s_owner[TILE];
    __shared__ uint32_t s_wmax[8];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint64_t g0 = (uint64_t)blockIdx.x * SEG_GATHER_IV; g0 < n_iv; g0 += (uint64_t)gridDim.x * SEG_GATHER_IV) {
        const uint32_t m = (uint32_t)min((uint64_t)SEG_GATHER_IV, n_iv - g0);
        const uint64_t o0 = iv_off[g0];
        if (tid < m) s_desc[tid] = desc[g0 + tid];
        if (tid <= SEG_GATHER_IV) s_rel[tid] = (uint32_t)(iv_off[g0 + min(tid, m)] - o0);   // entries beyond m repeat the end
        __syncthreads();
        // parts beyond the buffer are never written (the scan's total set SE_CAP_SEG; the host reports GE_ERR_CAPACITY at its next read-back)
        const uint64_t room = cap > o0 ? cap - o0 : 0;
        const uint32_t n = room < (uint64_t)s_rel[m] ? (uint32_t)room : s_rel[m];
        T *out = off_seg + o0;
        for (uint32_t t0 = 0; t0 < n; t0 += TILE) {
            // ---- owner (interval id) of every output of the tile: heads marked by the intervals, then a running maximum.
            // (A binary search of s_rel per output was 62 % of this kernel's instructions and 55 % of its stall samples.)
            reinterpret_cast<uint4 *>(s_owner)[tid] = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();
            if (tid < m) {
                const uint32_t r0 = s_rel[tid], r1 = s_rel[tid + 1];
                if (r1 > r0 && r0 >= t0 && r0 < t0 + TILE) s_owner[r0 - t0] = (uint8_t)tid;   // empty intervals own nothing
            }
            uint32_t seed = 0;   // owner of output t0: last interval whose first output is <= t0 (uniform search, once per tile)
#pragma unroll
            for (int w = SEG_GATHER_IV / 2; w > 0; w >>= 1) if (s_rel[seed + w] <= t0) seed += w;
            __syncthreads();
            uint4 ow = reinterpret_cast<uint4 *>(s_owner)[tid];
            uint32_t wd[4] = {ow.x, ow.y, ow.z, ow.w};
            uint32_t run = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
#pragma unroll
                for (int bsh = 0; bsh < 32; bsh += 8) {
                    run = max(run, (wd[k] >> bsh) & 0xFFu);
                    wd[k] = (wd[k] & ~(0xFFu << bsh)) | (run << bsh);
                }
            }
            uint32_t inc = run;   // inclusive maximum over the threads of the warp, then over the warps before it
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if ((int)lane >= d) inc = max(inc, v); }
            if (lane == 31) s_wmax[warp] = inc;
            uint32_t before = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) before = 0;
            __syncthreads();
            before = max(before, seed);
            for (uint32_t w = 0; w < warp; w++) before = max(before, s_wmax[w]);
            const uint32_t bb = before * 0x01010101u;
            ow = make_uint4(__vmaxu4(wd[0], bb), __vmaxu4(wd[1], bb), __vmaxu4(wd[2], bb), __vmaxu4(wd[3], bb));
            reinterpret_cast<uint4 *>(s_owner)[tid] = ow;
            __syncthreads();
            // ---- the copy: DEPTH independent loads per thread in flight, stores of a warp contiguous
            const uint32_t t1 = min(n, t0 + TILE);
            for (uint32_t base = t0; base < t1; base += 256 * DEPTH) {
                T q[DEPTH];
                uint32_t L[DEPTH], R[DEPTH];
                bool ok[DEPTH];
#pragma unroll
                for (int u = 0; u < DEPTH; u++) {
                    const uint32_t idx = base + u * 256 + tid;
                    const uint32_t lo = s_owner[min(idx - t0, (uint32_t)TILE - 1)];
                    const uint4 d = s_desc[lo];
                    ok[u] = idx < t1 && d.y != SEG_PLAN_VERBATIM;
                    L[u] = d.z; R[u] = d.w;
                    if (ok[u]) q[u] = ld_stream(par_seg + ((((uint64_t)d.y) << 32 | d.x) + (idx - s_rel[lo])));
                }
#pragma unroll
                for (int u = 0; u < DEPTH; u++)
                    if (ok[u]) { part_clip(q[u], L[u], R[u]); st_stream(out + (base + u * 256 + tid), q[u]); }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// slots the plan could not turn into index ranges (rare: positions that do not ascend): the reference's loop writes them.
// Always launched with a small grid: it walks the list the plan appended to (or, if the list overflowed, every slot).
__global__ void seg_verbatim_fill_kernel(SegArgs a, const uint4 *__restrict__ desc, const unsigned int *__restrict__ n_verbatim, const uint64_t *__restrict__ verb_list,
                                         uint32_t verb_cap, const uint64_t *__restrict__ off_off, uint4 *__restrict__ off_seg, uint64_t cap) {
    const unsigned int nv = *n_verbatim;
    if (nv == 0 || a.dead()) return;
    const bool listed = nv <= verb_cap;
    const uint64_t n_total = listed ? nv : a.dc->n_off * a.n_chr * 2;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_total; w += (uint64_t)gridDim.x * blockDim.x) {
        const SegSlot s = seg_slot(a, listed ? verb_list[w] : w);
        if (s.k == 0 || desc[s.e0 + s.slot].y != SEG_PLAN_VERBATIM) continue;
        if (off_off[s.slot + 1] > cap) continue;
        const uint4 *par = static_cast<const uint4 *>(a.par_seg);
        seg_recombine_verbatim<true>(par + s.b0, s.n0, par + s.b1, s.n1, a.xo_bp + s.e0, s.k, a.cov_lo[s.c], a.cov_hi[s.c], s.hi, off_seg + off_off[s.slot]);
    }
}

// ras_find_cv (:2752-2815) on the segment lists: allele bit plane / root byte plane.  One thread per (haplotype
// row, word of the CV bit plane).  The reference scans every part for every CV and lets the LAST covering part win; in a sorted
// tiling (SORTED) that part is the last one that starts at or before the position, found by binary search; lists that need not be
// sorted (maps with rows closer than bp_dist_in_rmap) keep the reference's scan.  founder_root: after a re-base the parts name
// haplotypes of the re-base generation, whose generation-0 root population per CV is kept in this byte plane.
template <class T, bool SORTED>
__global__ void seg_find_cv_kernel(CvSet cs, uint64_t n_rows, const uint64_t *__restrict__ off, const T *__restrict__ seg, const uint32_t *__restrict__ cov_hi,
                                   const uint64_t *__restrict__ hm_off, const uint32_t *__restrict__ hm_bp,
                                   const uint8_t *const *__restrict__ founder_cv /* [n_pop] -> [nh][n_cv_tot] */,
                                   const uint8_t *const *__restrict__ founder_root /* null, or [n_pop] -> [nh][n_cv_tot] */,
                                   uint32_t *__restrict__ bits, uint8_t *__restrict__ rootp) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * cs.Wcv) return;
    uint32_t w = (uint32_t)(t % cs.Wcv);
    uint64_t row = t / cs.Wcv;  // 2*i + h
    uint32_t blk = cs.word_blk[w], out = 0;
    if (blk != 0xFFFFFFFFu) {
        uint32_t c = blk % (uint32_t)cs.n_chr;
        uint32_t k0 = cs.block_off[blk] + (w - cs.word_off[blk]) * 32u, k1 = min(k0 + 32u, cs.block_off[blk + 1]);
        uint64_t slot = ((row >> 1) * cs.n_chr + c) * 2 + (row & 1);
        const uint64_t e_end = off[slot + 1];
        const uint32_t hi_c = cov_hi[c];
        for (uint32_t k = k0; k < k1; k++) {
            uint32_t bp = cs.bp[k];
            uint8_t v = 0, r = 0;
            bool found = false;
            if (SORTED) {
                uint64_t lo = off[slot], hi = e_end;   // first part that starts beyond the position
                while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (seg[mid].x <= bp) lo = mid + 1; else hi = mid; }
                if (lo > off[slot]) {
                    const uint4 q = part_get(seg, lo - 1, e_end, hi_c);
                    if (bp < q.y) { const uint64_t o = (uint64_t)q.z * cs.n_cv_tot + k; v = founder_cv[q.w][o]; r = founder_root ? founder_root[q.w][o] : (uint8_t)q.w; found = true; }
                }
            } else {
                for (uint64_t e = off[slot]; e < e_end; e++) {
                    uint4 q = part_get(seg, e, e_end, hi_c);
                    if (q.x <= bp && bp < q.y) { const uint64_t o = (uint64_t)q.z * cs.n_cv_tot + k; v = founder_cv[q.w][o]; r = founder_root ? founder_root[q.w][o] : (uint8_t)q.w; found = true; }
                }
            }
            if (found && hm_off) {
                for (uint64_t e = hm_off[slot]; e < hm_off[slot + 1]; e++) if (hm_bp[e] == bp) { v ^= 1; break; }
            }
            out |= (uint32_t)(v & 1) << (k - k0);
            if (rootp) rootp[row * cs.n_cv_tot + k] = found ? r : 0;  // covered by no part: the effect tables hold a = d = 0 there (Human_CV ctor)
        }
    }
    bits[t] = out;
}

// ------------------------------------------------------------------------------------------------
// One chromosome out of the slot-major lists (downloads, `.int`, materialisation): a compact CSR over the 2n haplotype rows of
// chromosome c — count, scan, fill — built on the device, so a download moves one chromosome's parts once instead of the whole
// population's lists twice (what seg_count + seg_download used to do per chromosome).
// ------------------------------------------------------------------------------------------------
__global__ void seg_slice_count_kernel(uint64_t n_rows, int n_chr, int c, const uint64_t *__restrict__ off, const uint64_t *__restrict__ hm_off, uint32_t *__restrict__ cnt,
                                       uint32_t *__restrict__ hm_cnt) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t slot = ((row >> 1) * n_chr + c) * 2 + (row & 1);
    cnt[row] = (uint32_t)(off[slot + 1] - off[slot]);
    if (hm_cnt) hm_cnt[row] = hm_off ? (uint32_t)(hm_off[slot + 1] - hm_off[slot]) : 0u;
}
// parts of chromosome c as the four fields of `class part`, uint64 each (the layout ge_download_segments returns) ...
template <class T>
__global__ void seg_slice_fill_u64_kernel(uint64_t n_rows, int n_chr, int c, const uint64_t *__restrict__ off, const T *__restrict__ seg, uint32_t hi_c,
                                          const uint64_t *__restrict__ row_off, uint64_t *__restrict__ out /* [n][4] */) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t slot = ((row >> 1) * n_chr + c) * 2 + (row & 1);
    const uint64_t e0 = off[slot], e1 = off[slot + 1];
    uint64_t o = row_off[row];
    for (uint64_t e = e0; e < e1; e++, o++) {
        const uint4 q = part_get(seg, e, e1, hi_c);
        out[o * 4] = q.x; out[o * 4 + 1] = q.y; out[o * 4 + 2] = q.z; out[o * 4 + 3] = q.w;
    }
}
__global__ void seg_slice_fill_hm_kernel(uint64_t n_rows, int n_chr, int c, const uint64_t *__restrict__ hm_off, const uint32_t *__restrict__ hm_bp,
                                         const uint64_t *__restrict__ row_off, uint64_t *__restrict__ out) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t slot = ((row >> 1) * n_chr + c) * 2 + (row & 1);
    uint64_t o = row_off[row];
    for (uint64_t e = hm_off[slot]; e < hm_off[slot + 1]; e++, o++) out[o] = hm_bp[e] & ~HM_BAKED;
}
// ... or, for the materialisation below, as {first locus index the part covers, founder id}: one thread per part
struct PartLo { uint32_t lo, id; };   // id = hap_index | root_population << 27
template <class T>
__global__ void seg_slice_fill_lo_kernel(Genome g, uint64_t n_rows, int n_chr, int c, const uint64_t *__restrict__ off, const T *__restrict__ seg, uint32_t hi_c,
                                         const uint64_t *__restrict__ row_off, PartLo *__restrict__ out) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t slot = ((row >> 1) * n_chr + c) * 2 + (row & 1);
    const uint64_t e0 = off[slot], e1 = off[slot + 1];
    uint64_t o = row_off[row];
    for (uint64_t e = e0; e < e1; e++, o++) {
        const uint4 q = part_get(seg, e, e1, hi_c);
        out[o] = PartLo{locus_lower_bound(g, c, q.x), q.z | (q.w << SEG_ID_BITS)};
    }
}

// ras_convert_interval_to_hap_matrix (:1186-1230) for sorted tilings: the bit-packed rows of chromosome c from the parts and the
// founder panels.  The reference tests every part against every SNP (O(parts x loci) per haplotype); here ONE WARP walks a row: its
// lanes take 32 consecutive words, each lane steps from the warp's cursor to the part that covers its first locus (parts and words
// both ascend: a few steps per 1024 loci) and ORs together the founder words of the parts that intersect its 32 loci — one founder
// word per output word almost always.  Output words are written coalesced; nothing is scanned twice.
__global__ void __launch_bounds__(256) seg_materialise_rows_kernel(Genome g, int c, uint64_t n_rows, const uint64_t *__restrict__ row_off, const PartLo *__restrict__ parts,
                                                                   uint32_t hi_locus /* first locus at or beyond the covered end */,
                                                                   const uint32_t *const *__restrict__ founder_rows /* [n_pop] packed rows */, uint32_t nw,
                                                                   uint32_t *__restrict__ out /* word w of row r at out[r * out_stride + w] */, uint64_t out_stride) {
    const int lane = threadIdx.x & 31;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t woff = g.chr_word_off[c];
    for (uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += n_warps) {
        const PartLo *P = parts + row_off[row];
        const uint32_t n = (uint32_t)(row_off[row + 1] - row_off[row]);
        uint32_t cur = 0;   // warp-uniform: last part whose first locus is at or before the chunk's first locus (0 if none)
        for (uint32_t w0 = 0; w0 < nw; w0 += 32) {
            const uint32_t w = w0 + lane, s0 = w << 5, s1 = s0 + 32;
            uint32_t k = cur;
            while (k + 1 < n && __ldg(&P[k + 1].lo) <= s0) k++;
            uint32_t word = 0, q = k;
            while (q < n) {
                const uint32_t a = max(__ldg(&P[q].lo), s0);
                if (a >= s1) break;
                const uint32_t b = min(q + 1 < n ? __ldg(&P[q + 1].lo) : hi_locus, s1);
                if (a < b) {
                    const uint32_t id = __ldg(&P[q].id);
                    const uint32_t fw = __ldg(founder_rows[id >> SEG_ID_BITS] + (uint64_t)(id & ((1u << SEG_ID_BITS) - 1u)) * g.W + woff + w);
                    const uint32_t m = (b - s0 >= 32u ? 0xFFFFFFFFu : ((1u << (b - s0)) - 1u)) & ~((1u << (a - s0)) - 1u);
                    word |= fw & m;
                }
                if (b >= s1) break;
                q++;
            }
            if (w < nw) out[row * out_stride + w] = word;
            cur = __shfl_sync(0xffffffffu, k, 31);   // the next chunk's first locus lies beyond lane 31's: its part is a valid place to resume
        }
    }
}
// toggles of the per-haplotype mutation lists that the (re-based) founder panel does not hold yet: one thread per row
__global__ void seg_materialise_toggle_kernel(Genome g, int c, int n_chr, uint64_t n_rows, const uint64_t *__restrict__ hm_off, const uint32_t *__restrict__ hm_bp, uint64_t out_stride,
                                              uint32_t *__restrict__ out) {
    const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint64_t slot = ((row >> 1) * n_chr + c) * 2 + (row & 1);
    const uint32_t *pos = g.pos + g.locus_off[c];
    const uint32_t nl = g.chr_nloci[c];
    for (uint64_t e = hm_off[slot]; e < hm_off[slot + 1]; e++) {
        const uint32_t m = hm_bp[e];
        if (m & HM_BAKED) continue;
        bool first = true;   // a position toggles once however often the lineage was hit there (the reference looks it up with std::find)
        for (uint64_t e2 = hm_off[slot]; e2 < e; e2++) if ((hm_bp[e2] & ~HM_BAKED) == m) { first = false; break; }
        if (!first) continue;
        for (uint32_t s = lower_bound_u32(pos, nl, m); s < nl && pos[s] == m; s++) out[row * out_stride + (s >> 5)] ^= 1u << (s & 31);
    }
}
// the verbatim form for lists that need not be sorted (maps with rows closer than bp_dist_in_rmap): every part against every locus, the
// LAST covering part wins, exactly like the reference's loop
template <class T>
__global__ void seg_materialise_kernel(Genome g, int c, uint64_t n_rows, const uint64_t *__restrict__ off, const T *__restrict__ seg, const uint32_t *__restrict__ cov_hi,
                                       const uint64_t *__restrict__ hm_off, const uint32_t *__restrict__ hm_bp,
                                       const uint32_t *const *__restrict__ founder_rows /* [n_pop] packed rows */, uint8_t *__restrict__ alleles) {
    uint32_t nl = g.chr_nloci[c];
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * nl) return;
    uint32_t s = (uint32_t)(t % nl);
    uint64_t row = t / nl;
    uint32_t pos = g.pos[g.locus_off[c] + s];
    uint64_t slot = ((row >> 1) * g.n_chr + c) * 2 + (row & 1);
    const uint64_t e_end = off[slot + 1];
    const uint32_t hi_c = cov_hi[c];
    uint8_t v = 0;
    bool found = false;
    for (uint64_t e = off[slot]; e < e_end; e++) {
        uint4 q = part_get(seg, e, e_end, hi_c);
        if (q.x <= pos && pos < q.y) { v = (founder_rows[q.w][(uint64_t)q.z * g.W + g.chr_word_off[c] + (s >> 5)] >> (s & 31)) & 1u; found = true; }
    }
    if (found && hm_off) {
        for (uint64_t e = hm_off[slot]; e < hm_off[slot + 1]; e++) if (hm_bp[e] == pos) { v ^= 1; break; }
    }
    alleles[t] = v;
}

}  // namespace gek

namespace gek {
// Segment compaction (the documentation's limitation #2: lists only ever grow because the reference never
// merges): adjacent parts of one haplotype that continue the same founder haplotype (en == next st, same hap_index and
// root population) become one part.  Zero-length parts disappear into their neighbour or are dropped when they carry
// no length of their own.  The materialised haplotype is unchanged; the `.int` listing is no longer the reference's.
// One thread per haplotype slot; pass 0 counts, pass 1 writes.
template <bool FILL, class T>
__global__ void seg_compact_kernel(uint64_t n_slots, int n_chr, const uint32_t *__restrict__ cov_hi, const uint64_t *__restrict__ off, const T *__restrict__ seg,
                                   uint32_t *__restrict__ count, const uint64_t *__restrict__ new_off, T *__restrict__ out) {
    for (uint64_t slot = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e0 = off[slot], e1 = off[slot + 1];
        const uint32_t hi_c = cov_hi[(slot >> 1) % (uint64_t)n_chr];
        const uint64_t o = FILL ? new_off[slot] : 0;
        uint32_t n = 0;
        bool open = false;
        uint4 cur = make_uint4(0, 0, 0, 0);
        for (uint64_t e = e0; e < e1; e++) {
            const uint4 q = part_get(seg, e, e1, hi_c);
            if (q.x == q.y && e1 - e0 > 1) continue;                       // zero-length part: covers no locus
            if (open && q.x == cur.y && q.z == cur.z && q.w == cur.w) { cur.y = q.y; continue; }
            if (open) { if (FILL) part_put(out, o + n, cur); n++; }
            cur = q; open = true;
        }
        if (open) { if (FILL) part_put(out, o + n, cur); n++; }
        if (!FILL) count[slot] = n;
    }
}
}  // namespace gek

namespace gek {
// Identity-by-descent sharing between pairs of individuals on one chromosome (the documentation's Example 10
// derives it from the `.int` files with an external tool): two haplotypes are IBD where their parts name the same founder
// haplotype.  One thread per (pair, haplotype of a, haplotype of b): a two-pointer walk over the two sorted lists; overlapping
// pieces with equal (hap_index, root_population) that touch are one run; runs of at least min_bp count.
template <class T>
__global__ void seg_ibd_kernel(int n_chr, int c, const uint64_t *__restrict__ off, const T *__restrict__ seg, const uint32_t *__restrict__ cov_hi,
                               const uint32_t *__restrict__ ind_a, const uint32_t *__restrict__ ind_b, uint64_t n_pairs, uint32_t min_bp,
                               unsigned long long *__restrict__ shared_bp, uint32_t *__restrict__ n_runs) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs * 4) return;
    const uint64_t pair = t >> 2;
    const uint32_t ha = (uint32_t)(t >> 1) & 1u, hb = (uint32_t)t & 1u;
    const uint32_t hi_c = cov_hi[c];
    const uint64_t sa = ((uint64_t)ind_a[pair] * n_chr + c) * 2 + ha, sb = ((uint64_t)ind_b[pair] * n_chr + c) * 2 + hb;
    uint64_t ea = off[sa], eb = off[sb];
    const uint64_t ea_end = off[sa + 1], eb_end = off[sb + 1];
    if (ea >= ea_end || eb >= eb_end) return;
    uint4 qa = part_get(seg, ea, ea_end, hi_c), qb = part_get(seg, eb, eb_end, hi_c);
    unsigned long long total = 0;
    uint32_t runs = 0, run_lo = 0, run_hi = 0;
    bool open = false;
    auto close = [&]() { if (open && run_hi - run_lo >= min_bp) { total += run_hi - run_lo; runs++; } open = false; };
    for (;;) {
        const uint32_t lo = max(qa.x, qb.x), hi = min(qa.y, qb.y);
        if (lo < hi) {
            if (qa.z == qb.z && qa.w == qb.w) {
                if (open && lo == run_hi) run_hi = hi;
                else { close(); open = true; run_lo = lo; run_hi = hi; }
            } else close();
        }
        const bool adv_a = qa.y <= qb.y, adv_b = qb.y <= qa.y;
        if (adv_a) { if (++ea >= ea_end) break; qa = part_get(seg, ea, ea_end, hi_c); }
        if (adv_b) { if (++eb >= eb_end) break; qb = part_get(seg, eb, eb_end, hi_c); }
    }
    close();
    if (total) atomicAdd(&shared_bp[pair], total);
    if (runs) atomicAdd(&n_runs[pair], runs);
}
}  // namespace gek

static int seg_finish_all(ge_ctx *ctx);
static int seg_ibd(ge_ctx *ctx, int pop, int c, const uint64_t *ind_a, const uint64_t *ind_b, uint64_t n_pairs, uint64_t min_bp, uint64_t *shared_bp, uint32_t *n_runs) {
    GE_TRY(seg_finish_all(ctx));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists (GE_REP_SEGMENTS not enabled)");
    if (n_pairs == 0) return GE_OK;
    if (!ind_a || !ind_b || !shared_bp || !n_runs) return fail(GE_ERR_INVALID, "ge_ibd_sharing: null argument");
    std::vector<uint32_t> a(n_pairs), b(n_pairs);
    for (uint64_t k = 0; k < n_pairs; k++) {
        if (ind_a[k] >= S.n || ind_b[k] >= S.n) return fail(GE_ERR_INVALID, "ge_ibd_sharing: individual index out of range");
        a[k] = (uint32_t)ind_a[k]; b[k] = (uint32_t)ind_b[k];
    }
    Buf da, db, dt, dr;
    GE_TRY(ctx->upload(da, a)); GE_TRY(ctx->upload(db, b));
    GE_TRY(ctx->ensure_exact(dt, n_pairs * 8)); GE_TRY(ctx->ensure_exact(dr, n_pairs * 4));
    CUDA_TRY(cudaMemsetAsync(dt.p, 0, n_pairs * 8, ctx->stream)); CUDA_TRY(cudaMemsetAsync(dr.p, 0, n_pairs * 4, ctx->stream));
    const uint32_t mb = (uint32_t)std::min<uint64_t>(min_bp, 0xFFFFFFFFull);
    if (ctx->seg_packed)
        seg_ibd_kernel<uint2><<<nblk(n_pairs * 4, 128), 128, 0, ctx->stream>>>(ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint2>(), P.d_cov_hi.as<uint32_t>(), da.as<uint32_t>(),
                                                                             db.as<uint32_t>(), n_pairs, mb, dt.as<unsigned long long>(), dr.as<uint32_t>());
    else
        seg_ibd_kernel<uint4><<<nblk(n_pairs * 4, 128), 128, 0, ctx->stream>>>(ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), P.d_cov_hi.as<uint32_t>(), da.as<uint32_t>(),
                                                                             db.as<uint32_t>(), n_pairs, mb, dt.as<unsigned long long>(), dr.as<uint32_t>());
    GE_TRY(ctx->check_launch("seg_ibd"));
    CUDA_TRY(cudaMemcpyAsync(shared_bp, dt.p, n_pairs * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(n_runs, dr.p, n_pairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (Buf *x : {&da, &db, &dt, &dr}) ctx->release(*x);
    return GE_OK;
}

static void seg_release(ge_ctx *ctx, SegState &s) {
    ctx->release(s.off); ctx->release(s.seg);
    s.valid = false;
}

static int seg_init_gen0(ge_ctx *ctx, int p, uint64_t n) {
    PopDev &P = ctx->pop[p];
    SegState &S = P.st[P.cur].seg;
    uint64_t n_slots = n * ctx->cfg.n_chr * 2;
    // the part format is fixed here, once the maps are known: packed 8-byte parts wherever lists are sorted tilings and the
    // reference's verbatim loop (which needs en in memory) is not asked for
    if (p == 0) {
        bool same_range = true;
        for (int q = 1; q < ctx->cfg.n_pop; q++)
            for (int c = 0; c < ctx->cfg.n_chr; c++)
                same_range &= ctx->pop[q].rmap_bp[c].front() == ctx->pop[0].rmap_bp[c].front() && ctx->pop[q].rmap_bp[c].back() == ctx->pop[0].rmap_bp[c].back();
        uint64_t max_haps = 0;
        for (PopDev &Q : ctx->pop) max_haps = std::max<uint64_t>(max_haps, std::max<uint64_t>(Q.cv[0][0].nhap, 2 * ctx->cfg.capacity));   // (re-basing makes every haplotype of a generation a founder)
        ctx->seg_packed = !ctx->seg_per_thread && !ctx->seg_wide && same_range && max_haps < (1ull << SEG_ID_BITS) && ctx->cfg.n_pop <= (1 << (32 - SEG_ID_BITS));
    }
    const size_t esz = ctx->seg_esz();
    const uint64_t slots_cap = ctx->cfg.capacity * ctx->cfg.n_chr * 2;
    for (GenState &G : P.st) GE_TRY(ctx->ensure(G.seg.off, (slots_cap + 1) * 8));
    GE_TRY(ctx->ensure(S.seg, std::max<uint64_t>(std::max<uint64_t>(n_slots, ctx->cfg.seg_capacity), 1) * esz));
    if (ctx->seg_packed) seg_init_kernel<uint2><<<nblk(n_slots + 1, 256), 256, 0, ctx->stream>>>(n, ctx->cfg.n_chr, p, P.d_cov_lo.as<uint32_t>(), P.d_cov_hi.as<uint32_t>(),
                                                                                              S.off.as<uint64_t>(), S.seg.as<uint2>());
    else seg_init_kernel<uint4><<<nblk(n_slots + 1, 256), 256, 0, ctx->stream>>>(n, ctx->cfg.n_chr, p, P.d_cov_lo.as<uint32_t>(), P.d_cov_hi.as<uint32_t>(),
                                                                                S.off.as<uint64_t>(), S.seg.as<uint4>());
    GE_TRY(ctx->check_launch("seg_init"));
    S.n_seg = n_slots; S.valid = true;
    P.hs.n_seg[P.cur] = n_slots;
    return ctx->push_state(P, offsetof(StepState, n_seg), 16);
}

// before anything on the control stream (or the host) reads segment lists: the bulk stream's plan + gather is complete and the host's
// sizes (n_seg among them) are exact; a generation that outgrew seg_capacity or could not be stored in packed parts is reported here
static int seg_finish_all(ge_ctx *ctx) {
    if (!ctx->segs()) return GE_OK;
    GE_TRY(ctx->join_bulk());
    return ctx->pull_state("founder segments");
}

static int seg_recombine(ge_ctx *ctx, int pop, uint64_t) {
    PopDev &P = ctx->pop[pop];
    GenState &par = P.st[P.cur], &off = P.st[P.cur ^ 1];
    if (!par.seg.valid) return fail(GE_ERR_INVALID, "parent generation has no segment lists");
    const int C = ctx->cfg.n_chr, k_off = P.cur ^ 1;
    const uint64_t slots_cap = ctx->cfg.capacity * C * 2;
    StepState *ss = P.d_ss;
    SegArgs a;
    DrawSet &D = P.draws();
    const DrawCounts *dc = &ss->dc[P.dcur];
    a.n_chr = C; a.ss = ss; a.dc = dc; a.father = D.father.as<uint32_t>(); a.mother = D.mother.as<uint32_t>();
    a.xo_off = D.xo_off.as<uint64_t>(); a.xo_bp = D.xo_bp.as<uint32_t>(); a.start_hap = D.start_hap.as<uint8_t>();
    a.par_off = par.seg.off.as<uint64_t>(); a.par_seg = par.seg.seg.p; a.cov_lo = P.d_cov_lo.as<uint32_t>(); a.cov_hi = P.d_cov_hi.as<uint32_t>();
    const size_t esz = ctx->seg_esz();
    const bool plan = !ctx->seg_per_thread;
    // With seg_capacity the output buffer is sized once, so nothing between the passes needs the host: the whole chain is queued on
    // the bulk stream behind the draws (like the bit-packed copy) and the control chain of the next generation overlaps it.  Without
    // it the host reads the total between the passes and grows the buffer.
    const bool fixed = ctx->cfg.seg_capacity != 0;
    const bool async = fixed && !ctx->serial;
    cudaStream_t st = async ? ctx->bulk : ctx->stream;
    uint64_t cap = 0;
    if (fixed) { GE_TRY(ctx->ensure_exact(off.seg.seg, ctx->cfg.seg_capacity * esz)); cap = ctx->cfg.seg_capacity; }
    PopDev *Pp = &P; GenState *parp = &par, *offp = &off; DrawSet *Dp = &D;
    auto chain = [=]() mutable -> int {
    PopDev &P = *Pp; GenState &par = *parp, &off = *offp; DrawSet &D = *Dp;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (ctx->profiling) for (cudaEvent_t &e : ev) e = ctx->get_event();
    auto grow = [&]() -> int {   // seg_capacity == 0: the total comes to the host, the buffer grows geometrically (a reallocation of tens of GB costs more than a generation)
        uint64_t n_seg = 0;
        CUDA_TRY(cudaMemcpyAsync(&n_seg, &ss->n_seg[k_off], 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        uint64_t want = std::max<uint64_t>(n_seg, 1);
        if (want * esz > off.seg.seg.cap) want += want / 2;
        GE_TRY(ctx->ensure_exact(off.seg.seg, want * esz));
        cap = off.seg.seg.cap / esz;
        return GE_OK;
    };
    if (ev[0]) CUDA_TRY(cudaEventRecord(ev[0], st));
    if (!plan) {   // ---- the reference's loop, one thread per slot: count, scan, fill
        const unsigned g = ctx->ctrl_grid(slots_cap, 128);
        GE_TRY(ctx->ensure(ctx->seg_cnt, (slots_cap + 1) * 4));   // (its own scratch: this may run on the bulk stream beside the next generation's control chain)
        seg_recombine_kernel<false><<<g, 128, 0, st>>>(a, ctx->seg_cnt.as<uint32_t>(), nullptr, nullptr, 0);
        GE_TRY(ctx->check_launch("seg_recombine<count>"));
        if (ev[1]) CUDA_TRY(cudaEventRecord(ev[1], st));
        GE_TRY(ctx->scan(st, ctx->seg_cnt.as<uint32_t>(), devn(&dc->n_off, (uint64_t)C * 2), slots_cap, off.seg.off.as<uint64_t>(), SegTotal{ss, k_off, cap}));
        if (!fixed) GE_TRY(grow());
        if (ev[2]) CUDA_TRY(cudaEventRecord(ev[2], st));
        seg_recombine_kernel<true><<<g, 128, 0, st>>>(a, nullptr, off.seg.off.as<uint64_t>(), off.seg.seg.as<uint4>(), cap);
        GE_TRY(ctx->check_launch("seg_recombine<fill>"));
    } else {       // ---- plan + gather
        const uint64_t iv_cap = P.hs.xo_cap + slots_cap;          // intervals: one more than crossovers in every slot
        constexpr uint32_t VERB_CAP = 1u << 18;
        GE_TRY(ctx->ensure(ctx->seg_cnt, (iv_cap + 1) * 4));
        GE_TRY(ctx->ensure(ctx->seg_desc, (size_t)(iv_cap + 1) * 16));
        GE_TRY(ctx->ensure(ctx->seg_iv_off, (size_t)(iv_cap + 1) * 8));
        GE_TRY(ctx->ensure(ctx->seg_flags, 16));
        GE_TRY(ctx->ensure(ctx->seg_verb, (size_t)VERB_CAP * 8));
        CUDA_TRY(cudaMemsetAsync(ctx->seg_flags.p, 0, 16, st));
        const unsigned pgrid = ctx->grid_for(slots_cap, 128);
        if (ctx->seg_packed) seg_plan_kernel<uint2><<<pgrid, 128, 0, st>>>(a, ctx->seg_cnt.as<uint32_t>(), ctx->seg_desc.as<uint4>(), ctx->seg_flags.as<unsigned int>(), ctx->seg_verb.as<uint64_t>(), VERB_CAP, &ss->err);
        else seg_plan_kernel<uint4><<<pgrid, 128, 0, st>>>(a, ctx->seg_cnt.as<uint32_t>(), ctx->seg_desc.as<uint4>(), ctx->seg_flags.as<unsigned int>(), ctx->seg_verb.as<uint64_t>(), VERB_CAP, &ss->err);
        GE_TRY(ctx->check_launch("seg_plan"));
        if (ev[1]) CUDA_TRY(cudaEventRecord(ev[1], st));
        GE_TRY(ctx->scan(st, ctx->seg_cnt.as<uint32_t>(), devn(&dc->n_iv), iv_cap, ctx->seg_iv_off.as<uint64_t>(), SegTotal{ss, k_off, cap}));
        seg_slot_offsets_kernel<<<ctx->grid_for(slots_cap + 1, 256), 256, 0, st>>>(a, a.xo_off, ctx->seg_iv_off.as<uint64_t>(), off.seg.off.as<uint64_t>());
        GE_TRY(ctx->check_launch("seg_slot_offsets"));
        if (!fixed) GE_TRY(grow());
        if (ev[2]) CUDA_TRY(cudaEventRecord(ev[2], st));
        const unsigned ggrid = ctx->grid_for(iv_cap, SEG_GATHER_IV);
        if (ctx->seg_packed) {
            seg_gather_kernel<uint2, 4><<<ggrid, 256, 0, st>>>(a, ctx->seg_desc.as<uint4>(), ctx->seg_iv_off.as<uint64_t>(), par.seg.seg.as<uint2>(), off.seg.seg.as<uint2>(), cap);
            GE_TRY(ctx->check_launch("seg_gather"));
        } else {
            seg_gather_kernel<uint4, 4><<<ggrid, 256, 0, st>>>(a, ctx->seg_desc.as<uint4>(), ctx->seg_iv_off.as<uint64_t>(), par.seg.seg.as<uint4>(), off.seg.seg.as<uint4>(), cap);
            GE_TRY(ctx->check_launch("seg_gather"));
            seg_verbatim_fill_kernel<<<ctx->n_sm * 8, 128, 0, st>>>(a, ctx->seg_desc.as<uint4>(), ctx->seg_flags.as<unsigned int>(), ctx->seg_verb.as<uint64_t>(), VERB_CAP,
                                                               off.seg.off.as<uint64_t>(), off.seg.seg.as<uint4>(), cap);
            GE_TRY(ctx->check_launch("seg_verbatim_fill"));
        }
    }
    if (ev[3]) {   // 16 (8 packed) B per part: every emitted piece comes from one parental part, read by both passes and written once; the part
        CUDA_TRY(cudaEventRecord(ev[3], st));   // count of THIS generation follows the kernels into a pinned slot
        uint64_t *slot = ctx->pinned_slot();
        CUDA_TRY(cudaMemcpyAsync(slot, &ss->n_seg[k_off], 8, cudaMemcpyDeviceToHost, st));
        ge_ctx::EvPair p1{ev[0], ev[1], GE_KERNEL_RECOMBINE_SEGMENTS, 0, 0, 0}, p2{ev[2], ev[3], GE_KERNEL_RECOMBINE_SEGMENTS, 0, 0, 0};
        p1.count_src = slot; p1.count_scale = esz; p2.count_src = slot; p2.count_scale = 2 * esz;
        ctx->ev_pending.push_back(p1); ctx->ev_pending.push_back(p2);
    }
    if (async) {
        CUDA_TRY(cudaEventRecord(D.bulk_done, st));   // the draw set is read until here
        D.bulk_pending = true;
        // a long copy is in flight: the heavy control kernels of the next generation run on thin grids beside it (as for the bit-packed copy)
        // (measured both ways on the config-5 sample: thin control grids 5.77 ms per generation, full-width 5.70 — the GPU is saturated either way)
        if (!ctx->bits()) ctx->note_bulk((double)par.seg.n_seg * 2.0 * (double)esz);
    }
    return GE_OK;
    };
    // the draws (and whatever the control stream did to the parental lists) are complete: the chain goes to the bulk stream, or runs in place
    if (async) GE_TRY(ctx->to_bulk(P.ev_ready, chain)); else GE_TRY(chain());
    off.seg.valid = true;
    return GE_OK;
}

static int seg_device_tables(ge_ctx *ctx, Buf &tbl, bool cv) {
    std::vector<const void *> ptrs(ctx->cfg.n_pop);
    for (int p = 0; p < ctx->cfg.n_pop; p++) ptrs[p] = cv ? ctx->pop[p].founder_cv.p : ctx->pop[p].founder_rows.p;
    return ctx->upload(tbl, ptrs);
}

static int seg_find_cv(ge_ctx *ctx, int pop) {
    GE_TRY(seg_finish_all(ctx));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists");
    if (ctx->n_cv_tot == 0) return GE_OK;
    Buf tbl, tbl_root;
    GE_TRY(seg_device_tables(ctx, tbl, true));
    const bool rebased_root = ctx->use_root && ctx->n_rebase > 0;
    if (rebased_root) {
        std::vector<const void *> ptrs(ctx->cfg.n_pop);
        for (int p = 0; p < ctx->cfg.n_pop; p++) ptrs[p] = ctx->pop[p].founder_root.p;
        GE_TRY(ctx->upload(tbl_root, ptrs));
    }
    const uint8_t *const *roots = rebased_root ? tbl_root.as<const uint8_t *>() : nullptr;
    uint64_t tot = 2 * S.n * ctx->Wcv;
    const uint64_t *hm_off = S.has_hm ? S.hm_off.as<uint64_t>() : nullptr;
    uint8_t *rootp = ctx->use_root ? S.cv_root.as<uint8_t>() : nullptr;
    if (ctx->seg_packed)
        seg_find_cv_kernel<uint2, true><<<nblk(tot, 256), 256, 0, ctx->stream>>>(ctx->cvset(), 2 * S.n, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint2>(), P.d_cov_hi.as<uint32_t>(), hm_off, S.hm_bp.as<uint32_t>(),
                                                                             tbl.as<const uint8_t *>(), roots, S.cv_allele.as<uint32_t>(), rootp);
    else if (!ctx->seg_per_thread)
        seg_find_cv_kernel<uint4, true><<<nblk(tot, 256), 256, 0, ctx->stream>>>(ctx->cvset(), 2 * S.n, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), P.d_cov_hi.as<uint32_t>(), hm_off, S.hm_bp.as<uint32_t>(),
                                                                             tbl.as<const uint8_t *>(), roots, S.cv_allele.as<uint32_t>(), rootp);
    else
        seg_find_cv_kernel<uint4, false><<<nblk(tot, 256), 256, 0, ctx->stream>>>(ctx->cvset(), 2 * S.n, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), P.d_cov_hi.as<uint32_t>(), hm_off, S.hm_bp.as<uint32_t>(),
                                                                              tbl.as<const uint8_t *>(), roots, S.cv_allele.as<uint32_t>(), rootp);
    GE_TRY(ctx->check_launch("seg_find_cv"));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tbl); ctx->release(tbl_root);
    return GE_OK;
}

// compact CSR of chromosome c over the 2n haplotype rows (all scratch is the caller's or local: downloads never touch a buffer a
// captured graph points at)
static int seg_slice_offsets(ge_ctx *ctx, PopDev &P, GenState &S, int c, Buf &row_off, Buf *hm_row_off, uint64_t *n_seg, uint64_t *n_hm) {
    (void)P;
    const uint64_t n_rows = 2 * S.n;
    Buf cnt_buf;
    GE_TRY(ctx->ensure_exact(cnt_buf, (2 * n_rows + 2) * 4));
    GE_TRY(ctx->ensure(row_off, (n_rows + 1) * 8));
    if (hm_row_off) GE_TRY(ctx->ensure(*hm_row_off, (n_rows + 1) * 8));
    uint32_t *cnt = cnt_buf.as<uint32_t>(), *hcnt = hm_row_off ? cnt + n_rows + 1 : nullptr;
    seg_slice_count_kernel<<<nblk(n_rows, 256), 256, 0, ctx->stream>>>(n_rows, ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.has_hm ? S.hm_off.as<uint64_t>() : nullptr, cnt, hcnt);
    GE_TRY(ctx->check_launch("seg_slice_count"));
    GE_TRY(ctx->exclusive_scan(cnt, n_rows, row_off.as<uint64_t>(), n_seg));
    if (hm_row_off) GE_TRY(ctx->exclusive_scan(hcnt, n_rows, hm_row_off->as<uint64_t>(), n_hm));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(cnt_buf);
    return GE_OK;
}

// packed rows of chromosome c from the segment lists: word w of haplotype row r goes to d_words[r * out_stride + w]
// (out_stride = ceil(n_loci/32): a chromosome matrix for the downloads; = W with d_words advanced to the chromosome's first word:
// whole rows, which is how a re-base builds the next founder panel)
static int seg_materialise_packed(ge_ctx *ctx, int pop, int c, uint32_t *d_words, uint64_t out_stride = 0) {
    GE_TRY(seg_finish_all(ctx));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists");
    for (PopDev &Q : ctx->pop) if (!Q.founder_rows.p) return fail(GE_ERR_INVALID, "haplotypes from segments need the founder panel (ge_set_founder_panel)");
    const uint64_t n_rows = 2 * S.n;
    const uint32_t nl = ctx->chr_nloci[c], nw = (nl + 31) / 32;
    if (n_rows == 0 || nw == 0) return GE_OK;
    if (out_stride == 0) out_stride = nw;
    Buf tbl, row_off, parts;
    GE_TRY(seg_device_tables(ctx, tbl, false));
    uint64_t n_seg = 0;
    GE_TRY(seg_slice_offsets(ctx, P, S, c, row_off, nullptr, &n_seg, nullptr));
    GE_TRY(ctx->ensure_exact(parts, std::max<uint64_t>(n_seg, 1) * sizeof(PartLo)));
    const uint32_t hi_c = (uint32_t)P.rmap_bp[c].back();
    if (ctx->seg_packed) seg_slice_fill_lo_kernel<uint2><<<nblk(n_rows, 128), 128, 0, ctx->stream>>>(ctx->genome(), n_rows, ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint2>(), hi_c, row_off.as<uint64_t>(), parts.as<PartLo>());
    else seg_slice_fill_lo_kernel<uint4><<<nblk(n_rows, 128), 128, 0, ctx->stream>>>(ctx->genome(), n_rows, ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), hi_c, row_off.as<uint64_t>(), parts.as<PartLo>());
    GE_TRY(ctx->check_launch("seg_slice_fill_lo"));
    const auto &L = ctx->loci[c];
    const uint32_t hi_locus = (uint32_t)(std::lower_bound(L.begin(), L.end(), (uint64_t)hi_c) - L.begin());
    const unsigned grid = (unsigned)std::min<uint64_t>(nblk(n_rows * 32, 256), (uint64_t)ctx->n_sm * 32);
    seg_materialise_rows_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->genome(), c, n_rows, row_off.as<uint64_t>(), parts.as<PartLo>(), hi_locus, tbl.as<const uint32_t *>(), nw, d_words, out_stride);
    GE_TRY(ctx->check_launch("seg_materialise_rows"));
    if (S.has_hm) {
        seg_materialise_toggle_kernel<<<nblk(n_rows, 128), 128, 0, ctx->stream>>>(ctx->genome(), c, ctx->cfg.n_chr, n_rows, S.hm_off.as<uint64_t>(), S.hm_bp.as<uint32_t>(), out_stride, d_words);
        GE_TRY(ctx->check_launch("seg_materialise_toggle"));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (Buf *b : {&tbl, &row_off, &parts}) ctx->release(*b);
    return GE_OK;
}

static int seg_materialise(ge_ctx *ctx, int pop, int c, uint8_t *d_alleles) {
    PopDev &P = ctx->pop[pop];
    GenState &S0 = P.st[P.cur];
    if (!ctx->seg_per_thread) {   // sorted tilings: packed rows by the warp walk, then one bit -> one byte
        const uint32_t nl = ctx->chr_nloci[c], nw = (nl + 31) / 32;
        Buf words;
        GE_TRY(ctx->ensure_exact(words, std::max<uint64_t>((uint64_t)2 * S0.n * nw, 1) * 4));
        GE_TRY(seg_materialise_packed(ctx, pop, c, words.as<uint32_t>()));
        const uint64_t tot = (uint64_t)2 * P.st[P.cur].n * nl;
        if (tot) {
            unpack_rows_kernel<<<nblk(tot, 256), 256, 0, ctx->stream>>>(words.as<uint32_t>(), nullptr, nw, 0, (uint32_t)(2 * P.st[P.cur].n), nl, d_alleles);
            GE_TRY(ctx->check_launch("unpack_rows"));
        }
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        ctx->release(words);
        return GE_OK;
    }
    GE_TRY(seg_finish_all(ctx));
    GenState &S = P.st[P.cur];
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists");
    for (PopDev &Q : ctx->pop) if (!Q.founder_rows.p) return fail(GE_ERR_INVALID, "haplotypes from segments need the founder panel (ge_set_founder_panel)");
    Buf tbl;
    GE_TRY(seg_device_tables(ctx, tbl, false));
    uint64_t tot = 2 * S.n * ctx->chr_nloci[c];
    seg_materialise_kernel<uint4><<<nblk(tot, 256), 256, 0, ctx->stream>>>(ctx->genome(), c, 2 * S.n, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), P.d_cov_hi.as<uint32_t>(),
                                                                       S.has_hm ? S.hm_off.as<uint64_t>() : nullptr, S.hm_bp.as<uint32_t>(), tbl.as<const uint32_t *>(), d_alleles);
    GE_TRY(ctx->check_launch("seg_materialise"));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->release(tbl);
    return GE_OK;
}

static int seg_compact(ge_ctx *ctx, int pop, uint64_t *n_before, uint64_t *n_after) {
    GE_TRY(seg_finish_all(ctx));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur], &O = P.st[P.cur ^ 1];   // the other generation's buffers are free between generations
    const bool fixed = ctx->cfg.seg_capacity != 0;
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists (GE_REP_SEGMENTS not enabled)");
    const uint64_t n_slots = S.n * ctx->cfg.n_chr * 2;
    if (n_before) *n_before = S.seg.n_seg;
    GE_TRY(ctx->ensure(P.cnt32, (n_slots + 1) * 4));
    GE_TRY(ctx->ensure(O.seg.off, (n_slots + 1) * 8));
    const unsigned grid = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, nblk(n_slots, 128)), (uint64_t)ctx->n_sm * 64);
    const int C = ctx->cfg.n_chr;
    const uint32_t *chi = P.d_cov_hi.as<uint32_t>();
    if (ctx->seg_packed) seg_compact_kernel<false, uint2><<<grid, 128, 0, ctx->stream>>>(n_slots, C, chi, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint2>(), P.cnt32.as<uint32_t>(), nullptr, nullptr);
    else seg_compact_kernel<false, uint4><<<grid, 128, 0, ctx->stream>>>(n_slots, C, chi, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), P.cnt32.as<uint32_t>(), nullptr, nullptr);
    GE_TRY(ctx->check_launch("seg_compact<count>"));
    uint64_t n_new = 0;
    GE_TRY(ctx->exclusive_scan(P.cnt32.as<uint32_t>(), n_slots, O.seg.off.as<uint64_t>(), &n_new));
    GE_TRY(ctx->ensure(O.seg.seg, std::max<uint64_t>(fixed ? ctx->cfg.seg_capacity : n_new, 1) * ctx->seg_esz()));
    if (ctx->seg_packed) seg_compact_kernel<true, uint2><<<grid, 128, 0, ctx->stream>>>(n_slots, C, chi, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint2>(), nullptr, O.seg.off.as<uint64_t>(), O.seg.seg.as<uint2>());
    else seg_compact_kernel<true, uint4><<<grid, 128, 0, ctx->stream>>>(n_slots, C, chi, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), nullptr, O.seg.off.as<uint64_t>(), O.seg.seg.as<uint4>());
    GE_TRY(ctx->check_launch("seg_compact<fill>"));
    std::swap(S.seg.off, O.seg.off); std::swap(S.seg.seg, O.seg.seg);
    S.seg.n_seg = n_new;
    O.seg.valid = false;
    if (n_after) *n_after = n_new;
    P.hs.n_seg[P.cur] = n_new;
    return ctx->push_state(P, offsetof(StepState, n_seg), 16);
}

// one chromosome's lists for the host (`.int` writer, tests): sliced and converted to the four fields on the device, moved once
static int seg_count(ge_ctx *ctx, int pop, int c, uint64_t *ns, uint64_t *nm) {
    GE_TRY(seg_finish_all(ctx));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists (GE_REP_SEGMENTS not enabled)");
    Buf row_off, hm_row_off;
    uint64_t a = 0, b = 0;
    GE_TRY(seg_slice_offsets(ctx, P, S, c, row_off, &hm_row_off, &a, &b));
    *ns = a; *nm = b;
    ctx->release(row_off); ctx->release(hm_row_off);
    return GE_OK;
}

static int seg_download(ge_ctx *ctx, int pop, int c, uint64_t *o_off, uint64_t *o_seg, uint64_t *o_moff, uint64_t *o_mbp) {
    GE_TRY(seg_finish_all(ctx));
    PopDev &P = ctx->pop[pop];
    GenState &S = P.st[P.cur];
    if (!S.seg.valid) return fail(GE_ERR_INVALID, "no segment lists (GE_REP_SEGMENTS not enabled)");
    const uint64_t n_rows = 2 * S.n;
    Buf row_off, hm_row_off, parts, hm;
    uint64_t n_seg = 0, n_hm = 0;
    GE_TRY(seg_slice_offsets(ctx, P, S, c, row_off, &hm_row_off, &n_seg, &n_hm));
    GE_TRY(ctx->ensure_exact(parts, std::max<uint64_t>(n_seg, 1) * 32));
    const uint32_t hi_c = (uint32_t)P.rmap_bp[c].back();
    if (n_rows) {
        if (ctx->seg_packed) seg_slice_fill_u64_kernel<uint2><<<nblk(n_rows, 128), 128, 0, ctx->stream>>>(n_rows, ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint2>(), hi_c, row_off.as<uint64_t>(), parts.as<uint64_t>());
        else seg_slice_fill_u64_kernel<uint4><<<nblk(n_rows, 128), 128, 0, ctx->stream>>>(n_rows, ctx->cfg.n_chr, c, S.seg.off.as<uint64_t>(), S.seg.seg.as<uint4>(), hi_c, row_off.as<uint64_t>(), parts.as<uint64_t>());
        GE_TRY(ctx->check_launch("seg_slice_fill"));
    }
    CUDA_TRY(cudaMemcpyAsync(o_off, row_off.p, (n_rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_seg) CUDA_TRY(cudaMemcpyAsync(o_seg, parts.p, n_seg * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (o_moff) CUDA_TRY(cudaMemcpyAsync(o_moff, hm_row_off.p, (n_rows + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (o_mbp && n_hm) {
        GE_TRY(ctx->ensure_exact(hm, n_hm * 8));
        seg_slice_fill_hm_kernel<<<nblk(n_rows, 128), 128, 0, ctx->stream>>>(n_rows, ctx->cfg.n_chr, c, S.hm_off.as<uint64_t>(), S.hm_bp.as<uint32_t>(), hm_row_off.as<uint64_t>(), hm.as<uint64_t>());
        GE_TRY(ctx->check_launch("seg_slice_fill_hm"));
        CUDA_TRY(cudaMemcpyAsync(o_mbp, hm.p, n_hm * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (Buf *b : {&row_off, &hm_row_off, &parts, &hm}) ctx->release(*b);
    return GE_OK;
}

// ------------------------------------------------------------------------------------------------
// Re-basing the founder panel (GeneEvolveDocumentation.pdf p.52, limitation #2: the reference's lists only grow —
// it appends clipped parts for ever, :2903-2958).  ge_rebase_founders makes the CURRENT generation the founder panel: every
// haplotype's alleles are materialised once into a new bit-packed panel (from the bit-packed rows when the context carries them,
// else from the lists and the old panel), its causal-variant alleles become the founder CV panel, and every list restarts as one
// part naming the haplotype itself.  Cost and memory of the segment path then depend on the generations since the last re-base,
// not since generation 0.  The lists that were replaced can be kept (keep_history): composing the current lists through them
// gives back the reference's lists against the generation-0 founders (ge_download_segments_gen0).
// ------------------------------------------------------------------------------------------------
namespace gek {
__global__ void rows_to_logical_kernel(const uint32_t *__restrict__ rows, const uint32_t *__restrict__ rowmap, uint64_t n_rows, uint32_t W, uint32_t *__restrict__ out) {
    for (uint64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const uint64_t src = rowmap ? (uint64_t)rowmap[r >> 1] * 2 + (r & 1) : r;
        const uint4 *s = reinterpret_cast<const uint4 *>(rows + src * W);
        uint4 *d = reinterpret_cast<uint4 *>(out + r * W);
        for (uint32_t q = threadIdx.x; q < W / 4; q += blockDim.x) st_stream(d + q, ld_stream(s + q));
    }
}
__global__ void cv_planes_to_bytes_kernel(uint32_t n_cv, uint32_t Wcv, const uint32_t *__restrict__ bitpos, const uint32_t *__restrict__ bits, uint64_t n_rows, uint8_t *__restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * n_cv) return;
    const uint32_t k = (uint32_t)(t % n_cv), bp = bitpos[k];
    out[t] = (bits[(t / n_cv) * Wcv + (bp >> 5)] >> (bp & 31u)) & 1u;
}
__global__ void hm_bake_kernel(uint64_t n, uint32_t *__restrict__ hm_bp) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) hm_bp[e] |= HM_BAKED;
}
}  // namespace gek

static int seg_rebase(ge_ctx *ctx, int keep_history) {
    GE_TRY(seg_finish_all(ctx));
    if (ctx->seg_per_thread) return fail(GE_ERR_UNSUPPORTED, "ge_rebase_founders needs sorted lists (a genetic map with rows closer than bp_dist_in_rmap was given)");
    const int np = ctx->cfg.n_pop, C = ctx->cfg.n_chr;
    const bool have_panel = ctx->pop[0].founder_rows.p != nullptr;
    for (PopDev &P : ctx->pop) {
        if (!P.st[P.cur].seg.valid || P.st[P.cur].n == 0) return fail(GE_ERR_INVALID, "ge_rebase_founders: no segment lists");
        if (2 * P.st[P.cur].n >= (1ull << SEG_ID_BITS)) return fail(GE_ERR_UNSUPPORTED, "ge_rebase_founders: more haplotypes than a part can name");
    }
    // 1. the new panels, all of them from the OLD ones (a migrant's parts name another population's founders)
    std::vector<Buf> new_rows(np), new_cv(np), new_root(np);
    for (int p = 0; p < np; p++) {
        PopDev &P = ctx->pop[p];
        GenState &S = P.st[P.cur];
        const uint64_t n_rows = 2 * S.n;
        if (have_panel) {
            GE_TRY(ctx->ensure_exact(new_rows[p], (size_t)n_rows * ctx->W * 4));
            if (ctx->bits()) {
                rows_to_logical_kernel<<<(unsigned)std::min<uint64_t>(n_rows, 1u << 16), 256, 0, ctx->stream>>>(S.hap.as<uint32_t>(), S.rowmap, n_rows, ctx->W, new_rows[p].as<uint32_t>());
                GE_TRY(ctx->check_launch("rows_to_logical"));
            } else {
                CUDA_TRY(cudaMemsetAsync(new_rows[p].p, 0, (size_t)n_rows * ctx->W * 4, ctx->stream));
                for (int c = 0; c < C; c++) GE_TRY(seg_materialise_packed(ctx, p, c, new_rows[p].as<uint32_t>() + ctx->chr_word_off[c], ctx->W));
            }
        }
        if (ctx->n_cv_tot) {
            GE_TRY(ctx->ensure_exact(new_cv[p], (size_t)n_rows * ctx->n_cv_tot));
            cv_planes_to_bytes_kernel<<<nblk(n_rows * ctx->n_cv_tot, 256), 256, 0, ctx->stream>>>(ctx->n_cv_tot, ctx->Wcv, ctx->d_cv_bitpos.as<uint32_t>(), S.cv_allele.as<uint32_t>(), n_rows, new_cv[p].as<uint8_t>());
            GE_TRY(ctx->check_launch("cv_planes_to_bytes"));
            if (ctx->use_root) {
                GE_TRY(ctx->ensure_exact(new_root[p], (size_t)n_rows * ctx->n_cv_tot));
                CUDA_TRY(cudaMemcpyAsync(new_root[p].p, S.cv_root.p, (size_t)n_rows * ctx->n_cv_tot, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // 2. swap them in, retire the lists, restart every haplotype as one part
    for (int p = 0; p < np; p++) {
        PopDev &P = ctx->pop[p];
        GenState &S = P.st[P.cur];
        if (have_panel) { ctx->release(P.founder_rows); P.founder_rows = new_rows[p]; }
        if (ctx->n_cv_tot) { ctx->release(P.founder_cv); P.founder_cv = new_cv[p]; }
        if (ctx->use_root) { ctx->release(P.founder_root); P.founder_root = new_root[p]; }
        P.n_founder_haps = 2 * S.n;
        if (S.has_hm && S.n_hm) {   // toggles up to here are in the new panel: they stay in the lineage's memory, but are not applied again
            hm_bake_kernel<<<nblk(S.n_hm, 256), 256, 0, ctx->stream>>>(S.n_hm, S.hm_bp.as<uint32_t>());
            GE_TRY(ctx->check_launch("hm_bake"));
        }
        if (keep_history) {
            PopDev::SegSnapshot snap;
            snap.off = S.seg.off; snap.seg = S.seg.seg; snap.n = S.n; snap.n_seg = S.seg.n_seg;
            S.seg.off = Buf(); S.seg.seg = Buf();
            P.history.push_back(snap);
        }
        S.seg.valid = false;
        GE_TRY(seg_init_gen0(ctx, p, S.n));   // (also pushes n_seg to the device-resident step state)
    }
    ctx->n_rebase++;
    ctx->graph_epoch++;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GE_OK;
}

// The current lists of chromosome c expressed against the generation-0 founders: composed on the host through the kept re-basing
// snapshots, newest first.  A current part (L, R, h, q) names haplotype h of population q at the last re-base; what the reference
// would hold there are the parts of that haplotype's snapshot list that recombine's clip rule (:2922-2954) emits for [L, R):
// y > L and (y <= R or x < R), clipped to (max(x, L), min(y, R)).
static int seg_download_gen0(ge_ctx *ctx, int pop, int c, std::vector<uint64_t> &off, std::vector<uint64_t> &seg) {
    PopDev &P = ctx->pop[pop];
    for (PopDev &Q : ctx->pop) if ((int)Q.history.size() != ctx->n_rebase) return fail(GE_ERR_INVALID, "ge_download_segments_gen0: a re-base dropped its history (keep_history = 0)");
    uint64_t ns = 0, nm = 0;
    GE_TRY(seg_count(ctx, pop, c, &ns, &nm));
    const uint64_t n_rows = 2 * P.st[P.cur].n;
    off.assign(n_rows + 1, 0); seg.assign(ns * 4, 0);
    GE_TRY(seg_download(ctx, pop, c, off.data(), seg.data(), nullptr, nullptr));
    const int np = ctx->cfg.n_pop;
    for (int level = ctx->n_rebase - 1; level >= 0; level--) {
        // the snapshot lists of every population at this level (a part may name another population's haplotype)
        std::vector<std::vector<uint64_t>> hoff(np), hseg(np);
        for (int q = 0; q < np; q++) {
            PopDev &Q = ctx->pop[q];
            PopDev::SegSnapshot &H = Q.history[level];
            // view the snapshot through a temporary generation state so that the slicing kernels can be reused
            GenState tmp;
            tmp.n = H.n; tmp.seg.off = H.off; tmp.seg.seg = H.seg; tmp.seg.n_seg = H.n_seg; tmp.seg.valid = true; tmp.has_hm = false;
            Buf row_off, parts;
            uint64_t n_seg = 0;
            GE_TRY(seg_slice_offsets(ctx, Q, tmp, c, row_off, nullptr, &n_seg, nullptr));
            GE_TRY(ctx->ensure_exact(parts, std::max<uint64_t>(n_seg, 1) * 32));
            const uint32_t hi_c = (uint32_t)Q.rmap_bp[c].back();
            const uint64_t hr = 2 * H.n;
            if (ctx->seg_packed) seg_slice_fill_u64_kernel<uint2><<<nblk(hr, 128), 128, 0, ctx->stream>>>(hr, ctx->cfg.n_chr, c, H.off.as<uint64_t>(), H.seg.as<uint2>(), hi_c, row_off.as<uint64_t>(), parts.as<uint64_t>());
            else seg_slice_fill_u64_kernel<uint4><<<nblk(hr, 128), 128, 0, ctx->stream>>>(hr, ctx->cfg.n_chr, c, H.off.as<uint64_t>(), H.seg.as<uint4>(), hi_c, row_off.as<uint64_t>(), parts.as<uint64_t>());
            GE_TRY(ctx->check_launch("seg_slice_fill"));
            hoff[q].resize(hr + 1); hseg[q].resize(n_seg * 4);
            CUDA_TRY(cudaMemcpyAsync(hoff[q].data(), row_off.p, (hr + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
            if (n_seg) CUDA_TRY(cudaMemcpyAsync(hseg[q].data(), parts.p, n_seg * 32, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            ctx->release(row_off); ctx->release(parts);
        }
        std::vector<uint64_t> noff(n_rows + 1, 0), nseg;
        nseg.reserve(seg.size() * 2);
        for (uint64_t r = 0; r < n_rows; r++) {
            noff[r] = nseg.size() / 4;
            for (uint64_t e = off[r]; e < off[r + 1]; e++) {
                const uint64_t L = seg[e * 4], R = seg[e * 4 + 1], h = seg[e * 4 + 2], q = seg[e * 4 + 3];
                if (q >= (uint64_t)np || h + 1 >= hoff[q].size()) return fail(GE_ERR_INVALID, "ge_download_segments_gen0: a part names a haplotype outside the re-base snapshot");
                for (uint64_t k = hoff[q][h]; k < hoff[q][h + 1]; k++) {
                    const uint64_t x = hseg[q][k * 4], y = hseg[q][k * 4 + 1];
                    if (y > L && (y <= R || x < R)) { nseg.push_back(std::max(x, L)); nseg.push_back(std::min(y, R)); nseg.push_back(hseg[q][k * 4 + 2]); nseg.push_back(hseg[q][k * 4 + 3]); }
                }
            }
        }
        noff[n_rows] = nseg.size() / 4;
        off.swap(noff); seg.swap(nseg);
    }
    return GE_OK;
}
