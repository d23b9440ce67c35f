// ge_segments.cuh — founder-segment representation (the reference's `class part` lists, src/Population.h:20-51).
#pragma once
#include "ge_context.cuh"

static void seg_release(SegState &s) { (void)s; }
static int seg_init_gen0(ge_ctx *, int, uint64_t) { return fail(GE_ERR_UNSUPPORTED, "GE_REP_SEGMENTS is not built yet"); }
static int seg_recombine(ge_ctx *, int, uint64_t) { return fail(GE_ERR_UNSUPPORTED, "GE_REP_SEGMENTS is not built yet"); }
static int seg_find_cv(ge_ctx *, int) { return fail(GE_ERR_UNSUPPORTED, "GE_REP_SEGMENTS is not built yet"); }
static int seg_materialise(ge_ctx *, int, int, uint8_t *) { return fail(GE_ERR_UNSUPPORTED, "GE_REP_SEGMENTS is not built yet"); }
static int seg_count(ge_ctx *, int, int, uint64_t *, uint64_t *) { return fail(GE_ERR_UNSUPPORTED, "GE_REP_SEGMENTS is not built yet"); }
static int seg_download(ge_ctx *, int, int, uint64_t *, uint64_t *, uint64_t *, uint64_t *) { return fail(GE_ERR_UNSUPPORTED, "GE_REP_SEGMENTS is not built yet"); }
