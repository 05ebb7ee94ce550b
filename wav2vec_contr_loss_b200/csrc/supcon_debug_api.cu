// TEST-ONLY entry points (libsupcon_b200_test.so; declared in csrc/supcon_debug.h, not in the public header):
// the tcgen05/TMA building-block diagnostic and host-side introspection of the tensor path's work distribution.
// The product library libsupcon_b200.so does not contain this file.
#include <string>

#include "supcon_debug.h"
#include "supcon_internal.h"

using namespace supcon;

namespace {
thread_local std::string g_dbg_err;
int dbg_fail(int code, const char* msg) {
  g_dbg_err = msg;
  return code;
}
int dbg_validate(const supcon_problem_t* p) {
  if (!p || p->n_total < 2 || p->d < 1 || p->n_rows < 1 || p->row_offset < 0 ||
      p->row_offset + p->n_rows > p->n_total || !(p->tau > 0.f))
    return dbg_fail(SUPCON_E_INVALID, "bad problem");
  return 0;
}
}  // namespace

extern "C" {

const char* supcon_debug_last_error(void) { return g_dbg_err.c_str(); }

int supcon_debug_tc_tile(const void* z_bf16, int32_t n, int32_t d, int32_t row_i, int32_t row_j, float* s_out,
                         float* o_out, void* stream) {
  if (!z_bf16 || !s_out || !o_out) return dbg_fail(SUPCON_E_INVALID, "NULL buffer passed to supcon_debug_tc_tile");
  const char* err = "";
  int rc = tc_debug_tile(z_bf16, n, d, row_i, row_j, s_out, o_out, reinterpret_cast<cudaStream_t>(stream), &err);
  if (rc) return dbg_fail(rc, err);
  return 0;
}

int supcon_debug_plan(const supcon_problem_t* p, int32_t* out, int32_t n_out) {
  if (int rc = dbg_validate(p)) return rc;
  if (!out || n_out < 1) return dbg_fail(SUPCON_E_INVALID, "bad arguments to supcon_debug_plan");
  if (!tc_supported(p)) return dbg_fail(SUPCON_E_UNSUPPORTED, "problem does not take the tensor path");
  return tc_debug_plan(p, out, n_out);
}

int supcon_debug_sched(int32_t col_tiles, int32_t ctas, int64_t units, int32_t cta, int32_t row_block,
                       int64_t* range_begin, int64_t* range_end, int32_t* first_cta, int32_t* last_cta) {
  if (col_tiles < 1 || ctas < 1 || units < ctas || cta < 0 || cta >= ctas || !range_begin || !range_end ||
      !first_cta || !last_cta)
    return dbg_fail(SUPCON_E_INVALID, "bad arguments to supcon_debug_sched");
  long long b = 0, e = 0;
  int f = 0, l = 0;
  tc_debug_sched(col_tiles, ctas, units, cta, row_block, &b, &e, &f, &l);
  *range_begin = b; *range_end = e; *first_cta = f; *last_cta = l;
  return 0;
}

}  // extern "C"
