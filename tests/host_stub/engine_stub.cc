// Test double for the device engine (include/kt_b200.h): lets the host layer's BOOKKEEPING (kt_host.cc: informer events,
// dictionaries, limits, status JSON, gauges) run in CPU tests.  It evaluates NOTHING: every entry point that would need
// a device pass fails with KT_ERR_CUDA, so a test that reaches the device through it fails loudly.  Test infrastructure only;
// never linked into the product library.
#include <cstdint>

#include "../../include/kt_b200.h"

struct kt_ctx { int dummy; };

extern "C" {
int kt_create(kt_ctx** out, int, const kt_limits* lim) {
  if (!out || !lim) return KT_ERR_INVALID;
  *out = new kt_ctx{0};
  return KT_OK;
}
void kt_destroy(kt_ctx* c) { delete c; }
const char* kt_last_error(const kt_ctx*) { return "engine stub: no device pass in this test double"; }
int kt_upload_pods(kt_ctx*, int, int64_t, const int64_t*, const int64_t*, const uint32_t*, const uint32_t*, const int32_t*) { return KT_OK; }
int kt_update_pod_rows(kt_ctx*, int, int64_t, const int64_t*, const int64_t*, const int64_t*, const uint32_t*, const uint32_t*, const int32_t*) { return KT_OK; }
int kt_upload_namespaces(kt_ctx*, int32_t, const int64_t*) { return KT_OK; }
int kt_upload_throttles(kt_ctx*, int32_t, const kt_throttle_cols*, const kt_selector_table*) { return KT_OK; }
int kt_upload_status(kt_ctx*, const kt_status_cols*) { return KT_OK; }
int kt_set_reserved(kt_ctx*, const int64_t*, const uint32_t*, const int64_t*) { return KT_OK; }
int kt_evaluate(kt_ctx*, int64_t, uint32_t) { return KT_ERR_CUDA; }
int kt_get_reconcile(kt_ctx*, const kt_reconcile_out*) { return KT_ERR_CUDA; }
int32_t kt_match_words(const kt_ctx*) { return 4; }
int kt_get_match_bitmap(kt_ctx*, int, uint32_t*) { return KT_ERR_CUDA; }
int kt_get_match_rows(kt_ctx*, int, int64_t, const int64_t*, uint32_t*) { return KT_ERR_CUDA; }
int kt_get_check(kt_ctx*, uint32_t*, uint8_t*) { return KT_ERR_CUDA; }
}
