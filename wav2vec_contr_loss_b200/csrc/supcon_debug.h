/* TEST-ONLY C entry points of libsupcon_b200_test.so (the product library and include/supcon_b200.h do not
 * carry them).  Used by tests/ and tools/ through tests/debug_lib.py. */
#ifndef SUPCON_DEBUG_H_
#define SUPCON_DEBUG_H_

#include "supcon_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* supcon_debug_last_error(void);

/* tcgen05/TMA building blocks: for 128-row blocks I = row_i.., J = row_j.. of a bf16 matrix z [n][256] writes
 * S = Z_I Z_J^T ([128][128] fp32) and O = bf16(S) Z_J ([128][256] fp32). */
int supcon_debug_tc_tile(const void* z_bf16, int32_t n, int32_t d, int32_t row_i, int32_t row_j, float* s_out,
                         float* o_out, void* stream);

/* Host-only introspection of the tensor path's work distribution (no device work):
 *   supcon_debug_plan  out[0..15] = {fwd CTAs, fwd column tiles, fwd partial-record slots, bwd CTAs, bwd column
 *                      tiles, bwd slots, two-phase eligible, fwd own-column-phase CTAs, fwd other-column-phase
 *                      CTAs, fwd own-column-phase slots, forward (256-row) blocks, backward (128-row) blocks,
 *                      bwd own-column-phase CTAs, bwd other-column-phase CTAs, bwd own-column-phase slots,
 *                      bwd own-column tiles}
 *   supcon_debug_sched the contiguous unit range of one CTA and the first/last CTA touching a row block for a
 *                      flattened (row block, column tile) list of `units` = row_blocks * col_tiles entries */
int supcon_debug_plan(const supcon_problem_t* p, int32_t* out, int32_t n_out);
int supcon_debug_sched(int32_t col_tiles, int32_t ctas, int64_t units, int32_t cta, int32_t row_block,
                       int64_t* range_begin, int64_t* range_end, int32_t* first_cta, int32_t* last_cta);

#ifdef __cplusplus
}
#endif
#endif /* SUPCON_DEBUG_H_ */
