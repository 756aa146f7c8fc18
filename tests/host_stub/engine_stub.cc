// Test double for the device engine (include/kt_b200.h): lets the host layer's BOOKKEEPING (kt_host.cc: informer events,
// dictionaries, limits, status JSON, gauges) run in CPU tests.  It evaluates NOTHING: every entry point that would need
// a device pass fails with KT_ERR_CUDA, so a test that reaches the device through it fails loudly.  Test infrastructure only;
// never linked into the product library.
// With KT_STUB_PASS_OK=1 in the environment the pass "succeeds" and every result is zero (nothing matched, nothing counted,
// every pod admitted): for timing the host layer's own per-pass work, never for checking a result.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/kt_b200.h"

struct kt_ctx { int64_t n[2]; int32_t m; int32_t R; };
static bool pass_ok() { static const bool v = std::getenv("KT_STUB_PASS_OK") != nullptr; return v; }

extern "C" {
int kt_create(kt_ctx** out, int, const kt_limits* lim) {
  if (!out || !lim) return KT_ERR_INVALID;
  *out = new kt_ctx{{0, 0}, 0, lim->n_resources};
  return KT_OK;
}
void kt_destroy(kt_ctx* c) { delete c; }
const char* kt_last_error(const kt_ctx*) { return "engine stub: no device pass in this test double"; }
int kt_upload_pods(kt_ctx* c, int kind, int64_t n, const int64_t*, const int64_t*, const uint32_t*, const uint32_t*, const int32_t*) { c->n[kind & 1] = n; return KT_OK; }
int kt_update_pod_rows(kt_ctx*, int, int64_t, const int64_t*, const int64_t*, const int64_t*, const uint32_t*, const uint32_t*, const int32_t*) { return KT_OK; }
int kt_upload_namespaces(kt_ctx*, int32_t, const int64_t*) { return KT_OK; }
int kt_upload_throttles(kt_ctx* c, int32_t m, const kt_throttle_cols*, const kt_selector_table*) { c->m = m; return KT_OK; }
int kt_upload_status(kt_ctx*, const kt_status_cols*) { return KT_OK; }
int kt_set_reserved(kt_ctx*, const int64_t*, const uint32_t*, const int64_t*) { return KT_OK; }
static int32_t words(const kt_ctx* c) { const int32_t w = ((c->m + 31) / 32 + 3) / 4 * 4; return w < 4 ? 4 : w; }
int kt_evaluate(kt_ctx*, int64_t, uint32_t) { return pass_ok() ? KT_OK : KT_ERR_CUDA; }
int kt_get_reconcile(kt_ctx* c, const kt_reconcile_out* o) {
  if (!pass_ok()) return KT_ERR_CUDA;
  const size_t M = (size_t)c->m, R = (size_t)c->R;
  if (o->used) std::memset(o->used, 0, R * M * 8);
  if (o->used_present) std::memset(o->used_present, 0, M * 4);
  if (o->used_cnt) std::memset(o->used_cnt, 0, M * 8);
  if (o->throttled) std::memset(o->throttled, 0, M * 4);
  if (o->calc_thr) std::memset(o->calc_thr, 0, R * M * 8);
  if (o->calc_present) std::memset(o->calc_present, 0, M * 4);
  if (o->calc_cnt) std::memset(o->calc_cnt, 0, M * 8);
  if (o->override_active) std::memset(o->override_active, 0, M);
  return KT_OK;
}
int32_t kt_match_words(const kt_ctx* c) { return words(c); }
int kt_get_match_bitmap(kt_ctx* c, int kind, uint32_t* w) {
  if (!pass_ok()) return KT_ERR_CUDA;
  std::memset(w, 0, (size_t)c->n[kind & 1] * words(c) * 4);
  return KT_OK;
}
int kt_get_match_rows(kt_ctx* c, int, int64_t k, const int64_t*, uint32_t* w) {
  if (!pass_ok()) return KT_ERR_CUDA;
  std::memset(w, 0, (size_t)k * words(c) * 4);
  return KT_OK;
}
int kt_get_check(kt_ctx* c, uint32_t* codes, uint8_t* admit) {
  if (!pass_ok()) return KT_ERR_CUDA;
  const size_t P = (size_t)c->n[KT_PODS_PENDING];
  if (codes) std::memset(codes, 0, P * 2 * words(c) * 4);
  if (admit) std::memset(admit, 1, P);
  return KT_OK;
}
int kt_get_check_rows(kt_ctx* c, int64_t k, const int64_t*, uint32_t* codes, uint8_t* admit) {
  if (!pass_ok()) return KT_ERR_CUDA;
  if (codes) std::memset(codes, 0, (size_t)k * 2 * words(c) * 4);
  if (admit) std::memset(admit, 1, (size_t)k);
  return KT_OK;
}
int kt_set_sparse_check(kt_ctx*, int64_t) { return KT_OK; }
int kt_get_changed(kt_ctx* c, int32_t* idx, int64_t cap, int64_t* count, uint8_t*) {
  if (!pass_ok()) return KT_ERR_CUDA;
  const int64_t m = c->m;  // timing mode: everything "changes"
  for (int64_t t = 0; t < m && t < cap; ++t) idx[t] = (int32_t)t;
  *count = m;
  return KT_OK;
}
int kt_get_reconcile_rows(kt_ctx* c, int64_t k, const int32_t*, const kt_reconcile_out* o) {
  if (!pass_ok()) return KT_ERR_CUDA;
  const size_t K = (size_t)k, R = (size_t)c->R;
  if (o->used) std::memset(o->used, 0, R * K * 8);
  if (o->used_present) std::memset(o->used_present, 0, K * 4);
  if (o->used_cnt) std::memset(o->used_cnt, 0, K * 8);
  if (o->throttled) std::memset(o->throttled, 0, K * 4);
  if (o->calc_thr) std::memset(o->calc_thr, 0, R * K * 8);
  if (o->calc_present) std::memset(o->calc_present, 0, K * 4);
  if (o->calc_cnt) std::memset(o->calc_cnt, 0, K * 8);
  if (o->override_active) std::memset(o->override_active, 0, K);
  return KT_OK;
}
int kt_admit_queue(kt_ctx*, int64_t, int64_t, uint32_t, int32_t* rounds, int64_t* admitted) {
  if (!pass_ok()) return KT_ERR_CUDA;
  if (rounds) *rounds = 1;
  if (admitted) *admitted = 0;
  return KT_OK;
}
int kt_get_check_sparse(kt_ctx* c, uint8_t* admit, uint32_t*, int64_t, int64_t* count) {
  if (!pass_ok()) return KT_ERR_CUDA;
  if (admit) std::memset(admit, 1, (size_t)c->n[KT_PODS_PENDING]);
  *count = 0;
  return KT_OK;
}
}
