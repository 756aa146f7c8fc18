// kt_kernels.cuh -- sm_100a kernels of the throttle-admission pass (hand-written CUDA, no libraries).
//
//   k_reconcile   running pods  -> affectedPods bitmap + per-throttle partial sums
//                 (replaces throttle_controller.go:221-246 affectedPods + :116-119 used-sum,
//                  clusterthrottle_controller.go:224-270, for EVERY throttle at once)
//   k_finalize    per throttle  -> CalculateThreshold(now), status.throttled, check constants
//                 (throttle_types.go:65-106, resource_amount.go:127-159, throttle_controller.go:122-133)
//   k_check       pending pods  -> affectedThrottles bitmap + 2-bit CheckThrottleStatus + admit
//                 (throttle_controller.go:248-269,349-397; throttle_types.go:128-153;
//                  clusterthrottle_types.go:30-55; plugin.go:177-180)
//
// HBM-bound integer work: no tensor cores.  One lane = one pod row; the selector match is word-parallel
// (32 throttles per LOP3) over the bit-sliced tables built by kt_tables.cc, so a pod costs
// O(label slots x non-zero namespace words) instead of O(throttles x terms x requirements).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kt_b200.h"

namespace kt {

constexpr int kTile = 256;  // pods per CTA (one lane per pod)

struct PodView {
  const int64_t* labels;    // [L][n]
  const int64_t* req;       // [R][n]
  const uint32_t* present;  // [n]
  const uint32_t* flags;    // [n]
  const int32_t* ns;        // [n]
  int64_t n;
};

struct TableView {
  const ulonglong2* hash;   // {label, row}
  uint32_t hash_mask;
  const uint32_t* table;    // [W][rows][TPpad][2]
  const uint32_t* need;     // [W][TPpad][B]
  const uint32_t* nsmask;   // [NS][W][TPpad]
  const int32_t* nsw_off;   // [NS+1]
  const int32_t* nsw_idx;
  int32_t M, W, Wp, TPpad, B, rows, NS;
};

// Per-throttle constants of the 4-step check, produced by k_finalize, gathered per matched pair.
// Followed in memory by int64 thrv[R] (S1 thresholds) and int64 head[R] (S4: threshold - used - reserved).
struct __align__(16) CheckHdr {
  uint32_t thr_has;   // threshold has resource r
  uint32_t m2;        // status.throttled.resourceRequests[r] == true          (S2)
  uint32_t m3;        // threshold.IsThrottled(used+reserved, E3)[r]           (S3)
  uint32_t cntbits;   // bit0 S1 count, bit1 S2 count, bit2 S3 count, bit3 S4 count, bit4 S4 uses >= instead of >
};

struct ThrottleView {  // device copies of kt_throttle_cols / kt_status_cols / reserved
  const uint8_t* kind;
  const uint8_t* flags;
  const int64_t* thr;          // [R][M]
  const uint32_t* thr_present;
  const int64_t* thr_cnt;
  const int32_t* ovr_off;
  const int64_t* ovr_begin;
  const int64_t* ovr_end;
  const uint8_t* ovr_flags;
  const int64_t* ovr_thr;      // [R][n_ovr]
  const uint32_t* ovr_present;
  const int64_t* ovr_cnt;
  int32_t n_ovr;
  // observed status (GIVEN_STATUS); null otherwise
  const uint8_t* st_calculated;
  const int64_t* st_calc_thr;
  const uint32_t* st_calc_present;
  const int64_t* st_calc_cnt;
  const int64_t* st_used;
  const uint32_t* st_used_present;
  const int64_t* st_used_cnt;
  const uint32_t* st_throttled;
  // reservations (null => nothing reserved)
  const int64_t* reserved;
  const uint32_t* reserved_present;
  const int64_t* reserved_cnt;
};

struct ReconcileView {  // device-resident kt_reconcile_out
  int64_t* used;
  uint32_t* used_present;
  int64_t* used_cnt;
  uint32_t* throttled;
  int64_t* calc_thr;
  uint32_t* calc_present;
  int64_t* calc_cnt;
  uint8_t* override_active;
};

__device__ __forceinline__ uint64_t d_mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27; x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

// label -> table row: exact (key,value) row, else the key's "other value" row, else the neutral row.
__device__ __forceinline__ int32_t lookup_row(const TableView& tb, int64_t label) {
  const int32_t neutral = tb.rows - 1;
  if (label == KT_LABEL_EMPTY) return neutral;
  uint64_t key = (uint64_t)label;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t slot = (uint32_t)d_mix64(key) & tb.hash_mask;
#pragma unroll 1
    while (true) {
      ulonglong2 e = __ldg(&tb.hash[slot]);
      if (e.x == key) return (int32_t)e.y;
      if (e.x == ~0ull) break;
      slot = (slot + 1) & tb.hash_mask;
    }
    key |= 0xffffffffull;  // second pass: the key's "other value" row
  }
  return neutral;
}

// Match word w (32 throttles) of one pod: OR over term planes of
//   AND_labels sat  &  (count of positive keys present == need)  &  namespace mask.
template <int LMAX, int TPC, int B>
__device__ __forceinline__ uint32_t eval_word(const TableView& tb, const int32_t (&rowid)[LMAX], int L, int ns, int w) {
  uint32_t result = 0;
  const size_t row_stride = (size_t)tb.TPpad * 2;
  const uint32_t* wbase = tb.table + (size_t)w * tb.rows * row_stride;
  const uint32_t* nsm = tb.nsmask + ((size_t)ns * tb.W + w) * tb.TPpad;
  const uint32_t* need = tb.need + (size_t)w * tb.TPpad * B;
#pragma unroll 1
  for (int s0 = 0; s0 < tb.TPpad; s0 += TPC) {
    uint32_t sat[TPC], cnt[TPC][B];
#pragma unroll
    for (int s = 0; s < TPC; ++s) {
      sat[s] = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < B; ++b) cnt[s][b] = 0;
    }
#pragma unroll
    for (int i = 0; i < LMAX; ++i) {
      if (i < L) {
        const uint32_t* e = wbase + (size_t)rowid[i] * row_stride + s0 * 2;
        uint32_t v[TPC * 2];
        if constexpr (TPC == 2) {
          uint4 q = __ldg(reinterpret_cast<const uint4*>(e));
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
          uint2 q = __ldg(reinterpret_cast<const uint2*>(e));
          v[0] = q.x; v[1] = q.y;
        }
#pragma unroll
        for (int s = 0; s < TPC; ++s) {
          sat[s] &= v[2 * s];
          uint32_t carry = v[2 * s + 1];  // ripple-add one bit into the bit-sliced counter
#pragma unroll
          for (int b = 0; b < B; ++b) {
            uint32_t t = cnt[s][b] & carry;
            cnt[s][b] ^= carry;
            carry = t;
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < TPC; ++s) {
      uint32_t m = sat[s] & __ldg(&nsm[s0 + s]);
#pragma unroll
      for (int b = 0; b < B; ++b) m &= ~(cnt[s][b] ^ __ldg(&need[(s0 + s) * B + b]));
      result |= m;
    }
  }
  return result;
}

// ------------------------------------------------------------------------------------------------
// k_reconcile: one lane per RUNNING pod.
// ------------------------------------------------------------------------------------------------
template <int LMAX, int TPC, int B>
__global__ void __launch_bounds__(kTile) k_reconcile(PodView pods, TableView tb, int L, int R, uint32_t* __restrict__ bitmap,
                                                     unsigned long long* __restrict__ part /* [2R+1][M]: used, present, cnt */) {
  const int64_t tile0 = (int64_t)blockIdx.x * kTile;
  const int Wp = tb.Wp;
  {  // zero the tile's bitmap rows (contiguous: pod-major)
    int64_t rows_here = pods.n - tile0 < kTile ? pods.n - tile0 : kTile;
    uint4* dst = reinterpret_cast<uint4*>(bitmap + tile0 * Wp);
    int64_t nvec = rows_here * (Wp / 4);
    for (int64_t i = threadIdx.x; i < nvec; i += kTile) dst[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int64_t p = tile0 + threadIdx.x;
  if (p >= pods.n) return;
  const uint32_t flags = __ldg(&pods.flags[p]);
  // shouldCountIn (throttle_controller.go:217-219): schedulerName == target && nodeName != ""
  if ((flags & (KT_POD_SCHEDULER_MATCH | KT_POD_SCHEDULED)) != (KT_POD_SCHEDULER_MATCH | KT_POD_SCHEDULED)) return;
  const int ns = __ldg(&pods.ns[p]);
  if ((unsigned)ns >= (unsigned)tb.NS) return;
  const int lo = __ldg(&tb.nsw_off[ns]), hi = __ldg(&tb.nsw_off[ns + 1]);
  if (lo == hi) return;

  int32_t rowid[LMAX];
#pragma unroll
  for (int i = 0; i < LMAX; ++i) rowid[i] = (i < L) ? lookup_row(tb, __ldg(&pods.labels[(int64_t)i * pods.n + p])) : 0;

  const bool alive = flags & KT_POD_NOT_FINISHED;  // isNotFinished (pod_util.go:26-28)
  const uint32_t present = __ldg(&pods.present[p]);
  const int M = tb.M;
  unsigned long long* part_used = part;
  unsigned long long* part_pres = part + (size_t)R * M;
  unsigned long long* part_cnt = part + (size_t)2 * R * M;

#pragma unroll 1
  for (int j = lo; j < hi; ++j) {
    const int w = __ldg(&tb.nsw_idx[j]);
    uint32_t word = eval_word<LMAX, TPC, B>(tb, rowid, L, ns, w);
    if (!word) continue;
    bitmap[p * Wp + w] = word;
    if (!alive) continue;
    // used = used.Add(ResourceAmountOfPod(p)) for every matched throttle (throttle_controller.go:116-119)
    while (word) {
      const int b = __ffs(word) - 1;
      word &= word - 1;
      const int t = w * 32 + b;
      atomicAdd(&part_cnt[t], 1ull);
      uint32_t pr = present;
      while (pr) {
        const int r = __ffs(pr) - 1;
        pr &= pr - 1;
        const long long v = __ldg(&pods.req[(int64_t)r * pods.n + p]);
        if (v != 0) atomicAdd(&part_used[(size_t)r * M + t], (unsigned long long)v);
        if (__ldcg(&part_pres[(size_t)r * M + t]) == 0ull) part_pres[(size_t)r * M + t] = 1ull;  // idempotent flag
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// k_finalize: one lane per throttle.  Consumes (and re-zeroes) the partial sums.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_finalize(ThrottleView tv, int M, int R, long long now, uint32_t eval_flags,
                                                  unsigned long long* __restrict__ part, ReconcileView out,
                                                  unsigned char* __restrict__ check /* [M][16 + 16R] */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M) return;
  const bool given = eval_flags & KT_EVAL_GIVEN_STATUS;
  const bool on_equal = eval_flags & KT_EVAL_ON_EQUAL;
  const uint32_t tflags = tv.flags[t];
  const bool live = (tflags & KT_THR_RESPONSIBLE) && !(tflags & KT_THR_SELECTOR_ERROR);

  // ---- used (this pass) ----
  long long used[KT_MAX_RESOURCES];
  uint32_t used_present = 0;
  long long used_cnt = (long long)part[(size_t)2 * R * M + t];
  part[(size_t)2 * R * M + t] = 0ull;
  if (used_cnt > 0) used_present |= KT_COUNT_BIT;  // Counts stays nil with zero counted pods (Q3)
  for (int r = 0; r < R; ++r) {
    used[r] = (long long)part[(size_t)r * M + t];
    if (part[(size_t)(R + r) * M + t] != 0ull) used_present |= 1u << r;
    part[(size_t)r * M + t] = 0ull;
    part[(size_t)(R + r) * M + t] = 0ull;
  }

  // ---- CalculateThreshold(now): merged active overrides REPLACE spec.threshold (throttle_types.go:65-106) ----
  long long calc[KT_MAX_RESOURCES];
  uint32_t calc_present = tv.thr_present[t];
  long long calc_cnt = tv.thr_cnt[t];
  for (int r = 0; r < R; ++r) calc[r] = tv.thr[(size_t)r * M + t];
  {
    bool active_found = false;
    long long ov[KT_MAX_RESOURCES];
    uint32_t ov_present = 0;
    long long ov_cnt = 0;
    for (int r = 0; r < R; ++r) ov[r] = 0;
    for (int i = tv.ovr_off[t]; i < tv.ovr_off[t + 1]; ++i) {
      if (tv.ovr_flags[i] & KT_OVR_PARSE_ERROR) continue;          // skipped, reported in Messages by the host
      if (!(tv.ovr_begin[i] <= now && now <= tv.ovr_end[i])) continue;  // IsActive: inclusive both ends
      active_found = true;
      const uint32_t op = tv.ovr_present[i];
      if (!(ov_present & KT_COUNT_BIT) && (op & KT_COUNT_BIT)) { ov_present |= KT_COUNT_BIT; ov_cnt = tv.ovr_cnt[i]; }
      for (int r = 0; r < R; ++r)
        if (((op >> r) & 1) && !((ov_present >> r) & 1)) { ov_present |= 1u << r; ov[r] = tv.ovr_thr[(size_t)r * tv.n_ovr + i]; }
    }
    if (active_found) {
      calc_present = ov_present;
      calc_cnt = ov_cnt;
      for (int r = 0; r < R; ++r) calc[r] = ov[r];
    }
    if (out.override_active) out.override_active[t] = active_found;
  }

  // ---- status.throttled = calculatedThreshold.IsThrottled(used, onEqual=true) (throttle_controller.go:133) ----
  uint32_t throttled = 0;
  if (live) {
    if ((calc_present & KT_COUNT_BIT) && (used_present & KT_COUNT_BIT) && used_cnt >= calc_cnt) throttled |= KT_COUNT_BIT;
    for (int r = 0; r < R; ++r)
      if (((calc_present >> r) & 1) && ((used_present >> r) & 1) && used[r] >= calc[r]) throttled |= 1u << r;
  }
  if (out.used) for (int r = 0; r < R; ++r) out.used[(size_t)r * M + t] = used[r];
  if (out.used_present) out.used_present[t] = used_present;
  if (out.used_cnt) out.used_cnt[t] = used_cnt;
  if (out.throttled) out.throttled[t] = throttled;
  if (out.calc_thr) for (int r = 0; r < R; ++r) out.calc_thr[(size_t)r * M + t] = calc[r];
  if (out.calc_present) out.calc_present[t] = calc_present;
  if (out.calc_cnt) out.calc_cnt[t] = calc_cnt;

  // ---- constants of CheckThrottledFor (throttle_types.go:128-153 / clusterthrottle_types.go:30-55) ----
  // Which status does PreFilter see: this pass's (FRESH) or the informer copy (GIVEN)?
  long long thr[KT_MAX_RESOURCES], su[KT_MAX_RESOURCES];
  uint32_t thr_present, su_present, st_throttled;
  long long thr_cnt, su_cnt;
  if (given) {
    if (tv.st_calculated[t]) {  // calculatedAt != zero => status.calculatedThreshold.threshold
      thr_present = tv.st_calc_present[t];
      thr_cnt = tv.st_calc_cnt[t];
      for (int r = 0; r < R; ++r) thr[r] = tv.st_calc_thr[(size_t)r * M + t];
    } else {
      thr_present = tv.thr_present[t];
      thr_cnt = tv.thr_cnt[t];
      for (int r = 0; r < R; ++r) thr[r] = tv.thr[(size_t)r * M + t];
    }
    su_present = tv.st_used_present[t];
    su_cnt = tv.st_used_cnt[t];
    for (int r = 0; r < R; ++r) su[r] = tv.st_used[(size_t)r * M + t];
    st_throttled = tv.st_throttled[t];
  } else {
    thr_present = calc_present;
    thr_cnt = calc_cnt;
    su_present = used_present;
    su_cnt = used_cnt;
    for (int r = 0; r < R; ++r) { thr[r] = calc[r]; su[r] = used[r]; }
    st_throttled = throttled;
  }
  const uint32_t res_present = tv.reserved_present ? tv.reserved_present[t] : 0u;
  const long long res_cnt = (tv.reserved_cnt && (res_present & KT_COUNT_BIT)) ? tv.reserved_cnt[t] : 0;
  // alreadyUsed = {} + status.used + reserved : nil counts are 0, presence is the union
  const uint32_t au_present = su_present | res_present;
  const long long au_cnt = ((su_present & KT_COUNT_BIT) ? su_cnt : 0) + res_cnt;
  const bool e3 = tv.kind[t] == KT_KIND_THROTTLE ? true : on_equal;  // Q1: Throttle hard-codes true (:143)

  CheckHdr h;
  h.thr_has = thr_present & ~KT_COUNT_BIT;
  h.m2 = st_throttled & ~KT_COUNT_BIT;
  h.m3 = 0;
  h.cntbits = on_equal ? 16u : 0u;
  long long* thrv = reinterpret_cast<long long*>(check + (size_t)t * (16 + 16 * R) + 16);
  long long* head = thrv + R;
  for (int r = 0; r < R; ++r) {
    const long long au = (((su_present >> r) & 1) ? su[r] : 0) + ((tv.reserved && ((res_present >> r) & 1)) ? tv.reserved[(size_t)r * M + t] : 0);
    if (((thr_present >> r) & 1) && ((au_present >> r) & 1) && (e3 ? au >= thr[r] : au > thr[r])) h.m3 |= 1u << r;
    thrv[r] = thr[r];
    head[r] = thr[r] - au;  // S4: used + reserved + pod (>|>=) threshold  <=>  pod (>|>=) head
  }
  if (thr_present & KT_COUNT_BIT) {
    if (1 > thr_cnt) h.cntbits |= 1u;                                              // S1: pod count 1 > threshold (Q4)
    if ((au_present & KT_COUNT_BIT) && (e3 ? au_cnt >= thr_cnt : au_cnt > thr_cnt)) h.cntbits |= 4u;  // S3
    if (on_equal ? au_cnt + 1 >= thr_cnt : au_cnt + 1 > thr_cnt) h.cntbits |= 8u;  // S4 (counts always present: the pod)
  }
  if (st_throttled & KT_COUNT_BIT) h.cntbits |= 2u;                                // S2
  if (!live) { h.thr_has = h.m2 = h.m3 = 0; h.cntbits &= 16u; }
  *reinterpret_cast<CheckHdr*>(check + (size_t)t * (16 + 16 * R)) = h;
}

// ------------------------------------------------------------------------------------------------
// k_check: one lane per PENDING pod.
// ------------------------------------------------------------------------------------------------
template <int LMAX, int TPC, int B>
__global__ void __launch_bounds__(kTile) k_check(PodView pods, TableView tb, int L, int R, const unsigned char* __restrict__ check,
                                                 uint32_t* __restrict__ bitmap, uint32_t* __restrict__ codes,
                                                 unsigned char* __restrict__ admit) {
  const int64_t tile0 = (int64_t)blockIdx.x * kTile;
  const int Wp = tb.Wp;
  {
    int64_t rows_here = pods.n - tile0 < kTile ? pods.n - tile0 : kTile;
    uint4* d0 = reinterpret_cast<uint4*>(bitmap + tile0 * Wp);
    uint4* d1 = reinterpret_cast<uint4*>(codes + tile0 * 2 * Wp);
    int64_t nvec = rows_here * (Wp / 4);
    for (int64_t i = threadIdx.x; i < nvec; i += kTile) d0[i] = make_uint4(0, 0, 0, 0);
    for (int64_t i = threadIdx.x; i < 2 * nvec; i += kTile) d1[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  const int64_t p = tile0 + threadIdx.x;
  if (p >= pods.n) return;
  unsigned char ok = 1;
  const int ns = __ldg(&pods.ns[p]);
  int lo = 0, hi = 0;
  if ((unsigned)ns < (unsigned)tb.NS) { lo = __ldg(&tb.nsw_off[ns]); hi = __ldg(&tb.nsw_off[ns + 1]); }
  if (lo != hi) {
    int32_t rowid[LMAX];
#pragma unroll
    for (int i = 0; i < LMAX; ++i) rowid[i] = (i < L) ? lookup_row(tb, __ldg(&pods.labels[(int64_t)i * pods.n + p])) : 0;
    const uint32_t present = __ldg(&pods.present[p]);
    // ResourceAmountOfPod(pod): the non-zero requests are the only ones IsThrottledFor looks at (Q5)
    uint32_t nz = 0;
    {
      uint32_t pr = present;
      while (pr) {
        const int r = __ffs(pr) - 1;
        pr &= pr - 1;
        if (__ldg(&pods.req[(int64_t)r * pods.n + p]) != 0) nz |= 1u << r;
      }
    }
    const size_t stride = 16 + 16 * (size_t)R;
#pragma unroll 1
    for (int j = lo; j < hi; ++j) {
      const int w = __ldg(&tb.nsw_idx[j]);
      uint32_t word = eval_word<LMAX, TPC, B>(tb, rowid, L, ns, w);
      if (!word) continue;
      bitmap[p * Wp + w] = word;
      uint32_t c0 = 0, c1 = 0;
      while (word) {
        const int b = __ffs(word) - 1;
        word &= word - 1;
        const int t = w * 32 + b;
        const unsigned char* cb = check + (size_t)t * stride;
        const uint4 hq = __ldg(reinterpret_cast<const uint4*>(cb));
        const long long* thrv = reinterpret_cast<const long long*>(cb + 16);
        const long long* head = thrv + R;
        const uint32_t cand = nz & hq.x;
        uint32_t code;
        // S1 threshold.IsThrottled(podAmount, false).IsThrottledFor(pod)
        bool s1 = hq.w & 1u;
        for (uint32_t c = cand; c && !s1;) {
          const int r = __ffs(c) - 1;
          c &= c - 1;
          s1 = __ldg(&pods.req[(int64_t)r * pods.n + p]) > __ldg(&thrv[r]);
        }
        if (s1) code = KT_CHECK_POD_REQUESTS_EXCEEDS_THRESHOLD;
        else if ((hq.w & 2u) || (nz & hq.y)) code = KT_CHECK_ACTIVE;   // S2 status.throttled.IsThrottledFor(pod)
        else if ((hq.w & 4u) || (nz & hq.z)) code = KT_CHECK_ACTIVE;   // S3 used+reserved already over
        else {
          bool s4 = hq.w & 8u;                                         // S4 used+pod+reserved
          const bool ge = hq.w & 16u;
          for (uint32_t c = cand; c && !s4;) {
            const int r = __ffs(c) - 1;
            c &= c - 1;
            const long long v = __ldg(&pods.req[(int64_t)r * pods.n + p]);
            const long long hd = __ldg(&head[r]);
            s4 = ge ? v >= hd : v > hd;
          }
          code = s4 ? KT_CHECK_INSUFFICIENT : KT_CHECK_NOT_THROTTLED;
        }
        if (code) ok = 0;
        if (b < 16) c0 |= code << (2 * b);
        else c1 |= code << (2 * (b - 16));
      }
      if (c0) codes[p * 2 * Wp + 2 * w] = c0;
      if (c1) codes[p * 2 * Wp + 2 * w + 1] = c1;
    }
  }
  admit[p] = ok;
}

// Row-level delta: scatter k packed rows into the resident columns (pod informer Add/Update/Delete).
__global__ void __launch_bounds__(256) k_scatter_rows(int64_t k, const int64_t* __restrict__ rows, int L, int R, int64_t n,
                                                      const int64_t* __restrict__ labels, const int64_t* __restrict__ req,
                                                      const uint32_t* __restrict__ present, const uint32_t* __restrict__ flags,
                                                      const int32_t* __restrict__ ns, int64_t* d_labels, int64_t* d_req,
                                                      uint32_t* d_present, uint32_t* d_flags, int32_t* d_ns) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  const int64_t row = rows[i];
  if (row < 0 || row >= n) return;
  for (int s = 0; s < L; ++s) d_labels[(int64_t)s * n + row] = labels[(int64_t)s * k + i];
  for (int r = 0; r < R; ++r) d_req[(int64_t)r * n + row] = req[(int64_t)r * k + i];
  d_present[row] = present[i];
  d_flags[row] = flags[i];
  d_ns[row] = ns[i];
}

}  // namespace kt
