/* geneevolve_b200.h — C-ABI of the B200-native GeneEvolve reproduction hot path.
 *
 * One shared library (libgeneevolve_b200.so, hand-written sm_100a CUDA) replaces the *bodies* of the
 * private per-generation methods of the reference's `class Simulation`.  The reference has
 * no plugin/FFI seam; the seam is cut at `Simulation::sim_next_generation` (src/Simulation.cpp:1890-2082)
 * and the tail of `ras_init_generation0` (:529-679).  Every entry point below cites the reference
 * interface it replaces.  All pointers are HOST pointers unless a name ends in `_dev`; sizes are explicit;
 * no C++/torch types cross the boundary.  All functions return GE_OK (0) or a negative error code and
 * leave a message retrievable with ge_last_error() — the reference's convention is `bool` + message on
 * stdout + `return false` up to main (src/Main.cpp:84-88); the host shim maps non-zero to `return false`.
 *
 * Threading: like the reference, one host thread drives one context.  One context owns one GPU.
 */
#ifndef GENEEVOLVE_B200_H
#define GENEEVOLVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GE_OK 0
#define GE_ERR_INVALID -1      /* bad argument / call order */
#define GE_ERR_CUDA -2         /* CUDA runtime failure (message has the CUDA error string) */
#define GE_ERR_CAPACITY -3     /* a generation outgrew the capacity given at ge_create */
#define GE_ERR_NO_MATES -4     /* "No one can marry" (:2125-2129) / "couples=0" (:2226-2230) */
#define GE_ERR_NAN -5          /* A or D is NaN (:2716-2720) */
#define GE_ERR_MIGRATION -6    /* migration row does not sum to 1 (:891-895) */
#define GE_ERR_UNSUPPORTED -7  /* input outside what the bit-packed representation can express */

/* representations of a chromosome (SURVEY.md §7.2 hard part 1) */
#define GE_REP_BITS 1          /* bit-packed haplotypes, the HBM-bound throughput path */
#define GE_REP_SEGMENTS 2      /* founder segments, the reference's own `class part` lists (src/Population.h:20-51) */

/* where per-generation random draws come from */
#define GE_RNG_PHILOX 0        /* Philox4x32-10 counter-based streams on the device (replaces RasRandomNumber) */
#define GE_RNG_REPLAY 1        /* every draw is supplied by the caller (fixed-draw parity mode) */

/* selection functions, Simulation::ras_selection_func (:3386-3428) */
#define GE_SEL_DEFAULT 0       /* "" -> logit(0,1) */
#define GE_SEL_LOGIT 1
#define GE_SEL_PROBIT 2
#define GE_SEL_STAB 3
#define GE_SEL_THR 4

/* ge_config.flags: measurement and verification aids.  None of them selects a different kernel for the same job: SERIAL changes
 * which stream a kernel is queued on, the SEG flags choose the memory format of a part and the (reference-verbatim) kernel that maps
 * with unsorted crossover lists always use, CV_FROM_SEGMENTS adds the reference's own rescan on top of the carried planes. */
#define GE_FLAG_SERIAL 1            /* queue the bulk copy on the control stream: no overlap of generations (kernel-alone timings) */
#define GE_FLAG_SEG_WIDE_PARTS 2    /* keep the reference's 16-byte parts {st, en, hap_index, root_population} instead of packed 8-byte parts */
#define GE_FLAG_SEG_VERBATIM 4      /* recombine founder segments with the reference's loop verbatim, one thread per gamete (implies 16-byte parts) */
#define GE_FLAG_CV_FROM_SEGMENTS 8  /* ge_compute_AD rebuilds the causal-variant planes from the segment lists like ras_find_cv, every generation */
#define GE_FLAG_NO_GRAPH 16         /* never replay a generation's control chain as a captured CUDA graph */

typedef struct ge_ctx ge_ctx;

/* ge_create: replaces the allocation side of ras_init_parameters (:164-525). */
typedef struct ge_config {
    int32_t device;          /* CUDA device ordinal */
    int32_t n_pop;           /* Simulation::_n_pop */
    int32_t n_chr;           /* Population::_nchr (equal in all populations) */
    int32_t n_phen;          /* Population::_pheno_scheme.size() */
    int32_t vt_type;         /* Parameters::_vt_type: 1 = parents' phenotype, 2 = parents' F (:3122-3131) */
    int32_t representation;  /* GE_REP_BITS | GE_REP_SEGMENTS (bit-or) */
    int32_t rng_mode;        /* GE_RNG_PHILOX or GE_RNG_REPLAY */
    int32_t flags;           /* GE_FLAG_* (bit-or), 0 for production runs */
    uint64_t seed;           /* Parameters::_seed; Philox key */
    uint64_t capacity;       /* max individuals per population in any generation */
    uint64_t seg_capacity;   /* max segments per population per generation (GE_REP_SEGMENTS), 0 = auto (buffers grow, one host read-back
                              * per generation).  When given, the segment path is queued on the bulk stream without a host read-back: a
                              * generation that outgrows it is reported (GE_ERR_CAPACITY) by the first call after that generation. */
    int32_t rank;            /* shard of the offspring axis this context owns */
    int32_t world_size;
} ge_config;

/* one row of the generation-info file, Population::ras_read_generation_info_file (src/Population.cpp:13-96) */
typedef struct ge_gen_params {
    uint64_t pop_size;       /* _pop_size[gen-1] */
    double mat_cor;          /* _mat_cor[gen-1] */
    int32_t offspring_dist;  /* 'p' or 'f' (_offspring_dist[gen-1]) */
    int32_t selection_func;  /* GE_SEL_* (_selection_func[gen-1]) */
    double selection_par1;
    double selection_par2;
} ge_gen_params;

/* Draws of one population for one generation (fixed-draw parity mode, or exported after a Philox
 * generation).  Slot order is the reference's loop order in Simulation::reproduce (:2433-2488):
 * offspring i, chromosome c, gamete g (0 = from the father, 1 = from the mother) -> (i*n_chr + c)*2 + g. */
typedef struct ge_draws {
    uint64_t n_offspring;
    const uint64_t *father;     /* [n_offspring] position of the father in the parent generation (Couples_Info::pos_male) */
    const uint64_t *mother;     /* [n_offspring] */
    const uint8_t *sex;         /* [n_offspring] 1 = male, 2 = female (:2472) */
    const uint64_t *xo_off;     /* [n_offspring*n_chr*2 + 1] CSR offsets into xo_bp */
    const uint64_t *xo_bp;      /* crossover positions in bp, ras_sim_loc_rec's list WITHOUT its two sentinels (:2983,:2993) */
    const uint8_t *start_hap;   /* [n_offspring*n_chr*2] starting haplotype (:2449,:2455) */
    const uint64_t *mut_off;    /* [n_offspring*n_chr + 1] CSR offsets into mut_bp, or NULL (no mutation map) */
    const uint64_t *mut_bp;     /* mutation positions in bp (:2519) */
    const uint8_t *mut_gam;     /* 0 = paternal gamete, 1 = maternal (:2521) */
    const double *e_raw;        /* [n_phen][n_offspring] N(0,1) environment draws (:3102), or NULL */
    const double *common;       /* [n_phen][n_offspring] sibling-common effect per offspring (:2481-2484), or NULL */
    const double *parental0;    /* [n_phen][n] generation-0 parental effect N(0,vf) (:3108-3114); read by ge_init_generation0 only */
} ge_draws;

/* The draws random_mate (:2090-2157) / assort_mate (:2167-2360) consume, for ge_mate_replay (fixed-draw parity of the mating kernels:
 * with the reference's own draws the device must arrive at the reference's `_couples_info`).  Lengths are the reference's. */
typedef struct ge_mate_draws {
    const double *thin_u;            /* [n] the uniform compared with selection_value_func (:2110 / :2191), individual order */
    const double *mm_u;              /* [n] the second uniform of a kept individual, compared with --MM (:2202, :2213); ignored elsewhere; NULL for random mating */
    const uint64_t *trim_order;      /* [n_trim_order] the longer sex list after std::random_shuffle (:2235 / :2242), as positions in the unshuffled */
    uint64_t n_trim_order;           /*   list: its first |n_m - n_f| entries leave.  NULL / 0 when the lists are equal */
    const double *t1, *t2;           /* [n_couples] the two columns of ras_mvnorm's template (:2268) */
    uint64_t n_couples;
    const int32_t *family;           /* [n_couples] ras_rpois's family sizes (:2331), or NULL for the fixed distribution */
    const uint64_t *remainder_order; /* [n_remainder_order] pos_couple_can_marry after std::random_shuffle (:2350): its first pop_size - nf * couples */
    uint64_t n_remainder_order;      /*   entries get one more child.  Fixed distribution only */
    const uint64_t *rm_father_idx;   /* [n_rm] random mating: index of the father in the thinned male list (:2145), mother likewise (:2146) */
    const uint64_t *rm_mother_idx;
    uint64_t n_rm;                   /* = pop_size */
} ge_mate_draws;

/* per-individual arrays, the columns of the `.info` file (Population::ras_save_human_info,
 * src/Population.cpp:510-568).  Caller allocates for ge_get_population_size() individuals. */
typedef struct ge_indiv_soa {
    uint64_t *ids;   /* [n][7] ID, ID_Father, ID_Mother, ID_Fathers_Father, ID_Fathers_Mother, ID_Mothers_Father, ID_Mothers_Mother */
    uint8_t *sex;    /* [n] */
    double *A, *D, *G, *C, *E, *F, *P; /* each [n_phen][n] */
    double *mv, *sv, *svf;             /* each [n] mating_value, selection_value, selection_value_func */
} ge_indiv_soa;

/* the `.summary` row (Simulation::ras_save_res :782-834) */
typedef struct ge_moments {
    double var_A, var_D, var_G, var_C, var_E, var_F, var_P, h2;
} ge_moments;

const char *ge_last_error(void);
int ge_version(void);

int ge_create(const ge_config *cfg, ge_ctx **out);
int ge_destroy(ge_ctx *ctx);

/* ---- inputs: already-parsed flat arrays; the host keeps its text readers (SURVEY.md §5 "config") ---- */

/* Population flags set in ras_init_parameters (:205-214): _avoid_inbreeding, _RM, _MM_percent. */
int ge_set_population(ge_ctx *ctx, int pop, int avoid_inbreeding, int random_mating, double mm_percent);
/* rMap + _recom_prob of one chromosome (src/Population.h:183-189, src/Population.cpp:349-414, 471-507). */
int ge_set_genetic_map(ge_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *recom_prob,
                       uint64_t n_rows, uint64_t bp_dist_in_rmap);
/* MutationMap of one chromosome (src/Population.h:192-197, src/Population.cpp:420-468). */
int ge_set_mutation_map(ge_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rate, uint64_t n_rows);
/* Legend::pos of one chromosome (src/format_hap.h:19-26); shared by all populations (:1186-1230 indexes
 * every population's panel by the same SNP index).  Must be sorted ascending. */
int ge_set_loci(ge_ctx *ctx, int chr, const uint64_t *pos, uint64_t n_loci);
/* Hap_SNP of one population and chromosome (src/format_hap.h:28-32): alleles[h*n_loci + s] in {0,1},
 * hap-major exactly like Hap_SNP::hap.  Packed to bits on the device. */
int ge_set_founder_panel(ge_ctx *ctx, int pop, int chr, const uint8_t *alleles, uint64_t n_founder_haps);
/* Same panel, already bit-packed by the host: words[h*words_per_hap + w], words_per_hap = ceil(n_loci/32),
 * locus s -> word s/32, bit s%32 (the layout ge_download_haplotypes_packed returns). */
int ge_set_founder_panel_packed(ge_ctx *ctx, int pop, int chr, const uint32_t *words, uint64_t n_founder_haps);
/* CV_INFO + CV of one phenotype and chromosome (src/Population.h:201-220, src/Population.cpp:197-343). */
int ge_set_cv(ge_ctx *ctx, int pop, int phen, int chr, const uint64_t *bp, const double *a, const double *d,
              uint64_t n_cv, const uint8_t *founder_cv, uint64_t n_founder_haps);
/* Phenotype_scheme (src/Population.h:223-234). */
int ge_set_pheno_scheme(ge_ctx *ctx, int pop, int phen, double va, double vd, double ve, double vc, double vf,
                        double omega, double beta, double lambda);
/* Simulation::_gamma (:503-508). */
int ge_set_gamma(ge_ctx *ctx, const double *gamma /* [n_phen] */);

/* ---- multi-GPU: one context per GPU, each owning a subset of the chromosomes of EVERY individual (DESIGN.md
 * §Multi-GPU).  Propagation, crossover sampling and allele counts are then rank-local; the only exchange is the
 * sum over ranks of the per-individual partial genetic values inside ge_compute_AD. ----
 * global_ids[c] = index of local chromosome c in the full genome; keys the Philox counters so that draws do
 * not depend on the sharding.  Default: identity. */
int ge_set_chromosome_ids(ge_ctx *ctx, const int32_t *global_ids /* [n_chr] */);
/* Sum-allreduce hook: must leave the element-wise sum over all ranks in dev_buf (count doubles, device memory),
 * ordered after the work already queued on cuda_stream and complete (or stream-ordered) on return.
 * The Python host wires torch.distributed/NCCL (geneevolve_b200/dist.py); a C++ host can wire ncclAllReduce. */
typedef int (*ge_allreduce_fn)(void *user, double *dev_buf, uint64_t count, void *cuda_stream);
int ge_set_allreduce(ge_ctx *ctx, ge_allreduce_fn fn, void *user);
/* The same exchange issued by the library itself: `nccl_comm` is this rank's ncclComm_t, `nccl_all_reduce` the address of
 * ncclAllReduce in the NCCL library the communicator came from (the library does not link NCCL: the host that created the
 * communicator hands the entry point over).  No host code runs per generation, so a sharded generation replays as a
 * captured CUDA graph like an unsharded one — NCCL collectives are capturable.  Replaces a ge_set_allreduce hook. */
int ge_set_allreduce_nccl(ge_ctx *ctx, void *nccl_comm, void *nccl_all_reduce);

/* ---- generation 0: Simulation::ras_init_generation0 (:529-679) + ras_initial_human_gen0 (:3000-3072) ----
 * draws0[pop] (GE_RNG_REPLAY) carries sex, e_raw, common for the founders; NULL in GE_RNG_PHILOX mode. */
int ge_init_generation0(ge_ctx *ctx, const ge_draws *draws0 /* [n_pop] or NULL */);

/* ---- the per-generation seam, one entry point per private method of Simulation (src/Simulation.h:71-128) ---- */

/* bool random_mate(int ipop,int gen_ind) :2090-2157 / bool assort_mate(int ipop,int gen_ind) :2167-2360.
 * Chooses by the population's random_mating flag like sim_next_generation (:1907-1918). */
int ge_mate(ge_ctx *ctx, int pop, int gen, const ge_gen_params *params);
/* The same mating kernels under the reference's own draws (GE_RNG_REPLAY contexts): thinning, the trim of the longer sex list, the
 * sorts by mating value, the ranks of the template, pairing, the inbreeding exclusion and the family sizes run on the device exactly
 * as in ge_mate, with every random number taken from `draws` instead of the Philox streams.  The couples must then equal the
 * reference's (ge_get_couples).  std::sort's order among DISTINCT individuals with equal mating values is the one thing not pinned
 * (the device sorts stably). */
int ge_mate_replay(ge_ctx *ctx, int pop, int gen, const ge_gen_params *params, const ge_mate_draws *draws);
/* Population::_couples_info (src/Population.h:165-180): supply (replay) or read back the couples. */
int ge_set_couples(ge_ctx *ctx, int pop, const uint64_t *pos_male, const uint64_t *pos_female,
                   const uint8_t *inbreed, const int32_t *num_offspring, uint64_t n_couples);
int ge_get_couples_count(ge_ctx *ctx, int pop, uint64_t *n_couples);
int ge_get_couples(ge_ctx *ctx, int pop, uint64_t *pos_male, uint64_t *pos_female, uint8_t *inbreed,
                   int32_t *num_offspring);
/* std::vector<Human> reproduce(int ipop,int gen_num) :2394-2493 with ras_sim_loc_rec :2973-2995,
 * recombine :2903-2958, ras_add_mutation :2497-2552.  draws == NULL -> Philox. */
int ge_reproduce(ge_ctx *ctx, int pop, int gen, const ge_draws *draws);
/* bool ras_compute_AD(int ipop,int gen_num) :2624-2749 with ras_find_cv :2752-2815. */
int ge_compute_AD(ge_ctx *ctx, int pop, int gen);
/* bool ras_scale_AD_compute_GEF(int gen_num,int ipop,int iphen,double,double) :3075-3206.
 * e_raw: [n] N(0,1) draws for replay, NULL -> Philox (or the draws given to ge_reproduce). */
int ge_scale_AD_compute_GEF(ge_ctx *ctx, int pop, int gen, int phen, const double *e_raw);
/* bool sim_environmental_effects_specific_to_each_population(int iphen) :3345-3381 (all populations). */
int ge_environmental_effects_specific_to_each_population(ge_ctx *ctx, int phen);
/* bool ras_compute_mating_value_selection_value(int gen_num,int ipop) :3300-3342. */
int ge_compute_mating_value_selection_value(ge_ctx *ctx, int pop, int gen, const ge_gen_params *params);
/* bool ras_do_migration(int gen_ind) :877-989; row = migration_mat_gen[gen-1], n_pop*n_pop entries. */
int ge_do_migration(ge_ctx *ctx, int gen, const double *migration_row);
/* Fixed-draw mode only: the individuals ras_SampleWithoutReplacement (src/RasRandomNumber.cpp:90-120) picked in
 * population src_pop for the next ge_do_migration (positions in that population, any order). */
int ge_set_migration_sample(ge_ctx *ctx, int src_pop, const uint64_t *positions, uint64_t n);
/* bool ras_save_human_info_to_Pop_info_prev_gen(int ipop) :3211-3236. */
int ge_save_human_info_to_Pop_info_prev_gen(ge_ctx *ctx, int pop);

/* bool sim_next_generation(int gen_num) :1890-2082 — all of the above in the reference's order for every
 * population.  params: [n_pop]; migration_row: n_pop*n_pop or NULL; draws: [n_pop] or NULL. */
int ge_step_generation(ge_ctx *ctx, int gen, const ge_gen_params *params, const double *migration_row,
                       const ge_draws *draws);

/* ---- results back to the host writers ---- */
int ge_get_population_size(ge_ctx *ctx, int pop, uint64_t *n);
int ge_download_individuals(ge_ctx *ctx, int pop, ge_indiv_soa *out);
/* var() of each column like the per-generation report (:2014-2055, src/CommFunc.cpp:57-68) */
int ge_get_moments(ge_ctx *ctx, int pop, int phen, ge_moments *out);
int ge_get_mv_sv_var(ge_ctx *ctx, int pop, double *var_mv, double *var_sv);
/* scaling constants: _var_a_gen0, _var_d_gen0, adjusted _beta, _gen0_SV_mean, _gen0_SV_var */
int ge_get_gen0_constants(ge_ctx *ctx, int pop, int phen, double *var_a0, double *var_d0, double *beta,
                          double *sv_mean0, double *sv_var0);
/* Haplotypes of one chromosome as the reference's Hap_SNP matrix (ras_convert_interval_to_hap_matrix
 * :1186-1230): alleles[(2*i+h)*n_loci + s] in {0,1}.  Works in both representations. */
int ge_download_haplotypes(ge_ctx *ctx, int pop, int chr, uint8_t *alleles);
/* Verification aid: the same matrix built from the founder-segment lists and the founder panel even when the context also carries the
 * bit-packed rows (which ge_download_haplotypes prefers) — the two must agree.  GE_REP_SEGMENTS only. */
int ge_download_haplotypes_from_segments(ge_ctx *ctx, int pop, int chr, uint8_t *alleles);
/* Same, bit-packed: words_per_hap = ceil(n_loci/32) little-endian bit order (locus s -> word s/32, bit s%32). */
int ge_download_haplotypes_packed(ge_ctx *ctx, int pop, int chr, uint32_t *words);
/* Founder segments of one chromosome, the `.int` content (ras_write_hap_to_interval_format :1582-1639).
 * seg_off: [2*n + 1] in (individual, haplotype) order; seg: [n_seg][4] = st, en, hap_index, root_population. */
int ge_get_segment_count(ge_ctx *ctx, int pop, int chr, uint64_t *n_seg, uint64_t *n_mut);
int ge_download_segments(ge_ctx *ctx, int pop, int chr, uint64_t *seg_off, uint64_t *seg,
                         uint64_t *mut_off /* [2*n+1] per haplotype */, uint64_t *mut_bp);
/* Causal-variant alleles, the `--debug` .cvval dump (:2665-2683): out[(i*2+h)*n_cv + k]. */
int ge_download_cv_alleles(ge_ctx *ctx, int pop, int phen, int chr, uint8_t *out);
/* Extension beyond the reference (GeneEvolveDocumentation.pdf p.52, limitation #2: segment lists only
 * grow): merge adjacent parts of a haplotype that continue the same founder haplotype and drop zero-length parts.
 * The materialised haplotypes, causal-variant alleles and all values are unchanged; ge_download_segments then returns
 * the merged lists, which are no longer the reference's `.int` content.  n_before / n_after may be NULL. */
int ge_compact_segments(ge_ctx *ctx, int pop, uint64_t *n_before, uint64_t *n_after);
/* Extension, the other half of SURVEY.md §8f-3: re-base the founder panel to the CURRENT generation (all populations).  Every
 * haplotype's alleles are materialised once into a new bit-packed founder panel (from the bit-packed rows when the context carries
 * them, else from the lists and the old panel; contexts without a panel — BASELINE config 5 — only relabel), its causal-variant
 * alleles become the founder CV panel, and every list restarts as one part {cov_lo, cov_hi, 2*i + h, population}.  Haplotypes, CV
 * alleles and every value are unchanged; cost and memory of the segment path then follow the generations since the last re-base
 * instead of since generation 0 (the reference only ever appends, src/Simulation.cpp:2903-2958).  With keep_history the replaced
 * lists stay on the device and ge_download_segments_gen0 composes the current lists through them back to the generation-0
 * founders — the reference's `.int` content; without it that lineage is dropped and memory stays flat.  ge_ibd_sharing and
 * ge_download_segments afterwards speak of the re-base generation's haplotypes.  Needs sorted lists (every map the bit-packed
 * representation accepts). */
int ge_rebase_founders(ge_ctx *ctx, int keep_history);
/* The lists of one chromosome against the generation-0 founders, composed through every re-base (all of them must have kept
 * their history): same layout as ge_get_segment_count / ge_download_segments. */
int ge_get_segment_count_gen0(ge_ctx *ctx, int pop, int chr, uint64_t *n_seg);
int ge_download_segments_gen0(ge_ctx *ctx, int pop, int chr, uint64_t *seg_off, uint64_t *seg);
/* Verification aid (GE_REP_SEGMENTS): rebuild the causal-variant planes of the current generation by scanning every
 * haplotype's parts exactly like ras_find_cv (:2752-2815).  The hot path never does this — it carries the planes
 * forward by crossover parity — so the planes before and after this call must be identical. */
int ge_recompute_cv_from_segments(ge_ctx *ctx, int pop);
/* Bytes per part in device memory (introspection for the tests and the bench): 16 = the reference's `class part` {st, en, hap_index,
 * root_population}; 8 = packed {st, hap_index | root_population << 27} with the end implied by the next part, which the library
 * uses whenever the lists are sorted tilings (every genetic map whose rows are at least bp_dist_in_rmap apart).  Downloads always
 * return the four fields.  Valid after ge_init_generation0. */
int ge_get_segment_format(ge_ctx *ctx, int *bytes_per_part);
/* Extension (GeneEvolveDocumentation.pdf Example 10 derives the same from the `.int` files with an external
 * tool): identity-by-descent sharing of n_pairs pairs of individuals of the current generation on one chromosome.  Two
 * haplotypes are IBD where their parts name the same founder haplotype (hap_index, root_population); touching pieces are one
 * run.  Over the four haplotype combinations of a pair: shared_bp[k] = total length of the runs of at least min_bp base pairs,
 * n_runs[k] = their number.  ind_a[k] == ind_b[k] is allowed (the two combinations of different haplotypes then measure
 * autozygosity; the two of a haplotype with itself give the covered length each). */
int ge_ibd_sharing(ge_ctx *ctx, int pop, int chr, const uint64_t *ind_a, const uint64_t *ind_b, uint64_t n_pairs, uint64_t min_bp,
                   uint64_t *shared_bp, uint32_t *n_runs);
/* Draws the device generated for the last ge_reproduce of this population (GE_RNG_PHILOX): sizes first,
 * then the arrays (any pointer may be NULL to skip). */
int ge_get_draw_counts(ge_ctx *ctx, int pop, uint64_t *n_offspring, uint64_t *n_xo, uint64_t *n_mut);
int ge_download_draws(ge_ctx *ctx, int pop, uint64_t *father, uint64_t *mother, uint8_t *sex, uint64_t *xo_off,
                      uint64_t *xo_bp, uint8_t *start_hap, uint64_t *mut_off, uint64_t *mut_bp, uint8_t *mut_gam);

/* ---- measurement hooks (bench.py): CUDA-event time of the dominant kernel on the library's stream ---- */
#define GE_KERNEL_PROPAGATE_BITS 0      /* propagate_bits_kernel (bulk stream) */
#define GE_KERNEL_RECOMBINE_SEGMENTS 1  /* seg_recombine count + fill passes */
/* phases of the control chain (control stream, CUDA events around the whole phase: kernels, gaps and read-backs) */
#define GE_PHASE_MATE 2                 /* random_mate / assort_mate */
#define GE_PHASE_SAMPLE 3               /* family sizes, crossover (and mutation) sampling, placement */
#define GE_PHASE_CV_AD 4                /* causal-variant planes, allele counts, genetic values */
#define GE_PHASE_PHENOTYPE 5            /* noise, scaling, phenotypes, mating and selection values */
#define GE_KERNEL_COUNT 8
/* level 0: off; 1: CUDA events around the dominant kernel on its own stream (what `roofline.achieved` is computed from; the control
 * chain may still replay as a graph); 2: also around the phases of the control chain (GE_PHASE_*; the chain is then queued kernel by kernel) */
int ge_set_profiling(ge_ctx *ctx, int level);
int ge_get_kernel_time(ge_ctx *ctx, int kernel, double *total_ms, uint64_t *launches, uint64_t *algorithmic_bytes);
int ge_reset_kernel_times(ge_ctx *ctx);
int ge_get_launch_count(ge_ctx *ctx, uint64_t *launches);   /* every kernel this context launched */
int ge_get_graph_replays(ge_ctx *ctx, uint64_t *replays);   /* generations whose control chain was replayed from a captured CUDA graph */
int ge_synchronize(ge_ctx *ctx);
/* CUDA events on the library's stream around a caller-defined region (everything queued in between) */
int ge_timer_start(ge_ctx *ctx);
int ge_timer_stop(ge_ctx *ctx, double *elapsed_ms);
int ge_device_memory_bytes(ge_ctx *ctx, uint64_t *bytes);    /* device high-water mark (SURVEY.md §5 memory reporting) */

#ifdef __cplusplus
}
#endif
#endif
