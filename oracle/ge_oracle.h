/* ge_oracle.h — C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
 *
 * The oracle is a CPU restatement of the reference's per-generation hot path used as the
 * checker in tests/, in __graft_entry__.smoke() and as bench.py's `cpu_baseline` leg.  The product path
 * (geneevolve_b200/, libgeneevolve_b200.so) never imports, links or executes anything in this directory.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py runs this restatement in reference-stream mode
 * (GO_RNG_REF: std::minstd_rand0 + glibc rand() + the libstdc++ distributions, exactly the engines the
 * reference uses) from `--seed` alone and requires bit-identical couples, crossovers,
 * segments, haplotype matrices and per-individual values against tests/golden (npz files), which were produced
 * by the real reference (oracle/ref_driver.cpp, tests/golden/make_golden.py).
 *
 * The API mirrors include/geneevolve_b200.h one to one with a `go_` prefix so that the parity tests read
 * the same on both sides; the structs are shared from that header.
 */
#ifndef GE_ORACLE_H
#define GE_ORACLE_H
#include "../include/geneevolve_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GO_RNG_REF 2 /* reference streams from --seed (oracle only; the CUDA library has no such mode) */

typedef struct go_ctx go_ctx;

const char *go_last_error(void);
int go_create(const ge_config *cfg, go_ctx **out);
int go_destroy(go_ctx *ctx);
int go_set_population(go_ctx *ctx, int pop, int avoid_inbreeding, int random_mating, double mm_percent);
int go_set_genetic_map(go_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *recom_prob, uint64_t n_rows, uint64_t bp_dist);
int go_set_mutation_map(go_ctx *ctx, int pop, int chr, const uint64_t *bp, const double *rate, uint64_t n_rows);
int go_set_loci(go_ctx *ctx, int chr, const uint64_t *pos, uint64_t n_loci);
int go_set_founder_panel(go_ctx *ctx, int pop, int chr, const uint8_t *alleles, uint64_t n_founder_haps);
int go_set_founder_panel_packed(go_ctx *ctx, int pop, int chr, const uint32_t *words, uint64_t n_founder_haps);
int go_set_cv(go_ctx *ctx, int pop, int phen, int chr, const uint64_t *bp, const double *a, const double *d, uint64_t n_cv, const uint8_t *founder_cv, uint64_t n_founder_haps);
int go_set_pheno_scheme(go_ctx *ctx, int pop, int phen, double va, double vd, double ve, double vc, double vf, double omega, double beta, double lambda);
int go_set_gamma(go_ctx *ctx, const double *gamma);
int go_set_chromosome_ids(go_ctx *ctx, const int32_t *global_ids);
int go_set_allreduce(go_ctx *ctx, ge_allreduce_fn fn, void *user); /* host buffer instead of a device buffer */
int go_init_generation0(go_ctx *ctx, const ge_draws *draws0);
int go_mate(go_ctx *ctx, int pop, int gen, const ge_gen_params *params);
int go_set_couples(go_ctx *ctx, int pop, const uint64_t *pos_male, const uint64_t *pos_female, const uint8_t *inbreed, const int32_t *num_offspring, uint64_t n_couples);
int go_get_couples_count(go_ctx *ctx, int pop, uint64_t *n_couples);
int go_get_couples(go_ctx *ctx, int pop, uint64_t *pos_male, uint64_t *pos_female, uint8_t *inbreed, int32_t *num_offspring);
int go_reproduce(go_ctx *ctx, int pop, int gen, const ge_draws *draws);
int go_compute_AD(go_ctx *ctx, int pop, int gen);
int go_scale_AD_compute_GEF(go_ctx *ctx, int pop, int gen, int phen, const double *e_raw);
int go_environmental_effects_specific_to_each_population(go_ctx *ctx, int phen);
int go_compute_mating_value_selection_value(go_ctx *ctx, int pop, int gen, const ge_gen_params *params);
int go_do_migration(go_ctx *ctx, int gen, const double *migration_row);
int go_set_migration_sample(go_ctx *ctx, int src_pop, const uint64_t *positions, uint64_t n);
int go_save_human_info_to_Pop_info_prev_gen(go_ctx *ctx, int pop);
int go_step_generation(go_ctx *ctx, int gen, const ge_gen_params *params, const double *migration_row, const ge_draws *draws);
int go_get_population_size(go_ctx *ctx, int pop, uint64_t *n);
int go_download_individuals(go_ctx *ctx, int pop, ge_indiv_soa *out);
int go_get_moments(go_ctx *ctx, int pop, int phen, ge_moments *out);
int go_get_mv_sv_var(go_ctx *ctx, int pop, double *var_mv, double *var_sv);
int go_get_gen0_constants(go_ctx *ctx, int pop, int phen, double *var_a0, double *var_d0, double *beta, double *sv_mean0, double *sv_var0);
int go_download_haplotypes(go_ctx *ctx, int pop, int chr, uint8_t *alleles);          /* from the bit-packed state */
int go_download_haplotypes_from_segments(go_ctx *ctx, int pop, int chr, uint8_t *alleles); /* ras_convert_interval_to_hap_matrix restated */
int go_get_segment_count(go_ctx *ctx, int pop, int chr, uint64_t *n_seg, uint64_t *n_mut);
int go_download_segments(go_ctx *ctx, int pop, int chr, uint64_t *seg_off, uint64_t *seg, uint64_t *mut_off, uint64_t *mut_bp);
int go_download_cv_alleles(go_ctx *ctx, int pop, int phen, int chr, uint8_t *out);   /* ras_find_cv restated (segments) */
int go_download_cv_alleles_bits(go_ctx *ctx, int pop, int phen, int chr, uint8_t *out); /* from the propagated CV bit planes */
int go_download_AD_raw(go_ctx *ctx, int pop, double *A_raw, double *D_raw);           /* ras_compute_AD output before scaling, [n_phen][n] */
int go_get_draw_counts(go_ctx *ctx, int pop, uint64_t *n_offspring, uint64_t *n_xo, uint64_t *n_mut);
int go_download_draws(go_ctx *ctx, int pop, uint64_t *father, uint64_t *mother, uint8_t *sex, uint64_t *xo_off, uint64_t *xo_bp, uint8_t *start_hap, uint64_t *mut_off, uint64_t *mut_bp, uint8_t *mut_gam);
int go_download_e_raw(go_ctx *ctx, int pop, double *e_raw);                          /* [n_phen][n] raw N(0,1) of the last phenotype pass */
/* Philox primitive, for known-answer tests */
void go_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
/* Stand-alone timing kernels for bench.py's cpu_baseline (bit-packed propagation of n offspring on one core) */
double go_bench_propagate_bits(uint64_t n_parents, uint64_t n_offspring, uint64_t n_loci, uint64_t n_xo_per_gamete, uint64_t seed, uint64_t *checksum);

#ifdef __cplusplus
}
#endif
#endif
