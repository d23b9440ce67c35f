/* tests/host_oracle_shim.h — TEST INFRASTRUCTURE.  Force-included (-include) when the tests compile host/*.cpp
 * against the CPU oracle instead of the CUDA library, so that the host's readers, writers and generation loop can be
 * exercised without a GPU.  The product binary (host/Makefile) never sees this file. */
#include "../oracle/ge_oracle.h"
#define GE_HOST_NO_COMPACT /* the oracle restates the reference, which never merges segments */
#define ge_ctx go_ctx
#define ge_last_error go_last_error
#define ge_create go_create
#define ge_destroy go_destroy
#define ge_set_gamma go_set_gamma
#define ge_set_chromosome_ids go_set_chromosome_ids
#define ge_set_allreduce go_set_allreduce
#define ge_set_loci go_set_loci
#define ge_set_population go_set_population
#define ge_set_genetic_map go_set_genetic_map
#define ge_set_mutation_map go_set_mutation_map
#define ge_set_cv go_set_cv
#define ge_set_founder_panel_packed go_set_founder_panel_packed
#define ge_set_pheno_scheme go_set_pheno_scheme
#define ge_init_generation0 go_init_generation0
#define ge_step_generation go_step_generation
#define ge_get_population_size go_get_population_size
#define ge_download_individuals go_download_individuals
#define ge_get_moments go_get_moments
#define ge_get_mv_sv_var go_get_mv_sv_var
#define ge_download_haplotypes go_download_haplotypes
#define ge_get_segment_count go_get_segment_count
#define ge_download_segments go_download_segments
