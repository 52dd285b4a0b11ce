// b2s_table.cu -- device-resident online feature table: entity key -> feature vector (+ imputing), sm_100a.
//
// Replaces the lookup half of real-time feature enrichment: EnrichmentModelRouter / EnrichmentVotingEnsemble.preprocess
// (mlrun/serving/routers.py:1189-1196, 1335-1342) call OnlineVectorService.get (mlrun/feature_store/feature_vector.py:
// 975-1067), which emits every entity row into a storey graph that reads the online (NoSQL) store key by key, then fills
// missing / NaN / Inf values from the impute policy (:1046-1052).  Here the online table lives in HBM: an open-addressing
// hash table of 64-bit entity keys (linear probing, load factor <= 0.5, built once on the host) next to the
// [n_keys][n_features] float32 matrix; one kernel launch resolves a batch of keys: each lane probes for one key, then
// the warp copies the 32 rows it found with coalesced 16-byte accesses, imputing on the way, straight into the row
// matrix the scoring plan reads (no host round trip between enrichment and predict).
// Bound: HBM (random 4*F-byte row reads + the sequential output): algorithmic bytes = 8 (key) + 16 (slot) + 8*F per
// query.
#include <cuda_runtime.h>

#include <exception>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/b200serve.h"
#include "b2s_internal.h"
#include "b2s_hash.cuh"

#define TAB_TRY(expr)                                                                                       \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return b2s_int_fail(B2S_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {

using Slot = b2s::TableSlot;
using b2s::mix64;

struct LookupParams {
  const Slot* slots;
  uint64_t mask;            // capacity - 1
  const float* values;      // [n_keys][n_feat]
  const float* impute;      // [n_feat]; NaN: keep the stored value
  int32_t n_feat;
  int32_t any_impute;
  int32_t lanes_per_row;    // n_feat / 4 when that is a power of two <= 32 (rows are copied by sub-warps), else 0
  const int64_t* keys;      // [n]
  int64_t n;
  float* out;               // row i at out + i * out_stride (bytes)
  int64_t out_stride;
  int32_t* found;           // [n] 1 / 0 (rows of unknown keys are filled with NaN)
};

__global__ void __launch_bounds__(256) table_lookup_kernel(const __grid_constant__ LookupParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (p.n_feat & 3) == 0 && (p.out_stride & 15) == 0;
  for (int64_t base = warp * 32; base < p.n; base += n_warps * 32) {
    const int64_t q = base + lane;
    int64_t row = -1;
    if (q < p.n) {  // every lane probes for its own key
      row = b2s::table_find(p.slots, p.mask, p.keys[q]);
      if (p.found) p.found[q] = row >= 0 ? 1 : 0;
    }
    const int cnt = (int)((p.n - base < 32) ? (p.n - base) : 32);
    if (vec && p.lanes_per_row) {
      // lanes_per_row = n_feat / 4 lanes copy one row with one 16-byte access each, so a warp moves 32 / lanes_per_row rows
      // per step (two for 64 features); kGather steps are loaded before any is stored (more rows in flight)
      const int lpr = p.lanes_per_row, rpi = 32 / lpr;
      const int sub = lane / lpr, c = (lane - sub * lpr) * 4;
      const float4 f = p.any_impute ? *reinterpret_cast<const float4*>(p.impute + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      constexpr int kGather = 4;
      for (int j0 = 0; j0 < cnt; j0 += rpi * kGather) {
        float4 v[kGather];
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
          const int j = j0 + u * rpi + sub;
          const int64_t r = __shfl_sync(0xffffffffu, row, j & 31);
          v[u] = (j < cnt && r >= 0) ? __ldg(reinterpret_cast<const float4*>(p.values + r * p.n_feat + c))
                                     : make_float4(NAN, NAN, NAN, NAN);
        }
#pragma unroll
        for (int u = 0; u < kGather; ++u) {
          const int j = j0 + u * rpi + sub;
          if (j >= cnt) continue;
          float4 w = v[u];
          if (p.any_impute) {  // OnlineVectorService.get (:1046-1052): None / NaN / Inf -> impute value
            w.x = (!(fabsf(w.x) <= 3.402823466e38f) && f.x == f.x) ? f.x : w.x;
            w.y = (!(fabsf(w.y) <= 3.402823466e38f) && f.y == f.y) ? f.y : w.y;
            w.z = (!(fabsf(w.z) <= 3.402823466e38f) && f.z == f.z) ? f.z : w.z;
            w.w = (!(fabsf(w.w) <= 3.402823466e38f) && f.w == f.w) ? f.w : w.w;
          }
          *reinterpret_cast<float4*>(reinterpret_cast<char*>(p.out) + (base + j) * p.out_stride + c * 4) = w;
        }
      }
      continue;
    }
    for (int j = 0; j < cnt; ++j) {  // the warp copies query j's row together
      const int64_t r = __shfl_sync(0xffffffffu, row, j);
      float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(p.out) + (base + j) * p.out_stride);
      const float* src = p.values + r * p.n_feat;
      if (vec) {
        for (int c = lane * 4; c < p.n_feat; c += 128) {
          float4 v = r >= 0 ? *reinterpret_cast<const float4*>(src + c) : make_float4(NAN, NAN, NAN, NAN);
          if (p.any_impute) {  // OnlineVectorService.get (:1046-1052): None / NaN / Inf -> impute value
            const float4 f = *reinterpret_cast<const float4*>(p.impute + c);
            v.x = (!(fabsf(v.x) <= 3.402823466e38f) && f.x == f.x) ? f.x : v.x;
            v.y = (!(fabsf(v.y) <= 3.402823466e38f) && f.y == f.y) ? f.y : v.y;
            v.z = (!(fabsf(v.z) <= 3.402823466e38f) && f.z == f.z) ? f.z : v.z;
            v.w = (!(fabsf(v.w) <= 3.402823466e38f) && f.w == f.w) ? f.w : v.w;
          }
          *reinterpret_cast<float4*>(dst + c) = v;
        }
      } else {
        for (int c = lane; c < p.n_feat; c += 32) {
          float v = r >= 0 ? src[c] : NAN;
          if (p.any_impute) {
            const float f = p.impute[c];
            v = (!(fabsf(v) <= 3.402823466e38f) && f == f) ? f : v;
          }
          dst[c] = v;
        }
      }
    }
  }
}

// status[i] |= B2S_ROW_UNKNOWN_KEY where the key was not in the table (after the scoring plan wrote status)
__global__ void mark_unknown_kernel(const int32_t* __restrict__ found, int32_t* __restrict__ status, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!found[i]) status[i] |= B2S_ROW_UNKNOWN_KEY;
}

}  // namespace

struct b2s_table_s {
  int64_t n_keys = 0;
  int32_t n_feat = 0;
  uint64_t cap = 0;
  int any_impute = 0;
  Slot* d_slots = nullptr;
  float* d_values = nullptr;
  float* d_impute = nullptr;
  std::vector<float> h_impute;  // host copy: the fused gather folds the policy into the scoring kernel's operands
  int grid = 0;
  // host-call staging
  std::mutex mu;
  int64_t cap_rows = 0;
  int64_t* d_keys = nullptr;
  float* d_out = nullptr;
  int32_t* d_found = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // enrich_host staging: votes + status on the device, one pinned block [keys | votes | status] on the host
  int64_t enr_rows = 0;
  int32_t enr_out_cols = 0;
  float* d_votes = nullptr;
  int32_t* d_status = nullptr;
  char* h_pin = nullptr;
};

extern "C" int b2s_table_create(const int64_t* keys, int64_t n_keys, const float* values, int32_t n_features, const float* impute,
                                b2s_table_t* out) {
  try {  // no C++ exception crosses the C boundary
    if (!keys || !values || !out || n_keys <= 0 || n_features <= 0) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
    if (!b2s_int_inited()) return b2s_int_fail(B2S_ERR_STATE, "b2s_init was not called (no CUDA device: there is no CPU fallback)");
    uint64_t cap = 16;
    while (cap < (uint64_t)n_keys * 2) cap <<= 1;
    std::vector<Slot> slots(cap, Slot{0, -1});
    for (int64_t i = 0; i < n_keys; ++i) {
      uint64_t h = mix64((uint64_t)keys[i]) & (cap - 1);
      while (slots[h].row >= 0) {
        if (slots[h].key == keys[i]) return b2s_int_fail(B2S_ERR_INVALID, "duplicate entity key %lld (rows %lld and %lld)", (long long)keys[i], (long long)slots[h].row, (long long)i);
        h = (h + 1) & (cap - 1);
      }
      slots[h] = Slot{keys[i], i};
    }
    auto* t = new b2s_table_s();
    t->n_keys = n_keys;
    t->n_feat = n_features;
    t->cap = cap;
    TAB_TRY(cudaSetDevice(b2s_int_device()));
    TAB_TRY(cudaMalloc(&t->d_slots, cap * sizeof(Slot)));
    TAB_TRY(cudaMemcpy(t->d_slots, slots.data(), cap * sizeof(Slot), cudaMemcpyHostToDevice));
    // one more row than keys: row n_keys is all NaN, what the fused gather copies for an unknown key
    TAB_TRY(cudaMalloc(&t->d_values, ((size_t)n_keys + 1) * n_features * 4));
    TAB_TRY(cudaMemcpy(t->d_values, values, (size_t)n_keys * n_features * 4, cudaMemcpyHostToDevice));
    {
      const std::vector<float> nan_row((size_t)n_features, NAN);
      TAB_TRY(cudaMemcpy(t->d_values + (size_t)n_keys * n_features, nan_row.data(), (size_t)n_features * 4, cudaMemcpyHostToDevice));
    }
    std::vector<float> imp(((size_t)n_features + 3) / 4 * 4, NAN);
    if (impute)
      for (int c = 0; c < n_features; ++c) {
        imp[c] = impute[c];
        if (impute[c] == impute[c]) t->any_impute = 1;
      }
    t->h_impute = imp;
    TAB_TRY(cudaMalloc(&t->d_impute, imp.size() * 4));
    TAB_TRY(cudaMemcpy(t->d_impute, imp.data(), imp.size() * 4, cudaMemcpyHostToDevice));
    int occ = 0;
    TAB_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, table_lookup_kernel, 256, 0));
    t->grid = b2s_int_sm_count() * std::max(occ, 1);
    for (int i = 0; i < 4; ++i) TAB_TRY(cudaEventCreate(&t->ev[i]));
    *out = t;
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

static int launch_lookup(b2s_table_t t, const int64_t* d_keys, int64_t n, float* d_rows, int64_t row_stride, int32_t* d_found, cudaStream_t st) {
  LookupParams p{};
  p.slots = t->d_slots;
  p.mask = t->cap - 1;
  p.values = t->d_values;
  p.impute = t->d_impute;
  p.n_feat = t->n_feat;
  p.any_impute = t->any_impute;
  {
    const int l = t->n_feat / 4;
    p.lanes_per_row = (t->n_feat % 4 == 0 && l >= 1 && l <= 32 && (l & (l - 1)) == 0) ? l : 0;
  }
  p.keys = d_keys;
  p.n = n;
  p.out = d_rows;
  p.out_stride = row_stride;
  p.found = d_found;
  const int64_t warps = (n + 31) / 32;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(t->grid, (warps + 7) / 8));
  b2s_int_count_launches(1);
  table_lookup_kernel<<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return b2s_int_fail(B2S_ERR_CUDA, "table lookup launch failed: %s", cudaGetErrorString(e));
  return B2S_OK;
}

extern "C" int b2s_table_lookup_device(b2s_table_t t, const int64_t* d_keys, int64_t n, float* d_rows, int64_t row_stride_bytes,
                                       int32_t* d_found, void* stream) {
  try {  // no C++ exception crosses the C boundary
    if (!t) return b2s_int_fail(B2S_ERR_INVALID, "null table");
    if (n < 0 || row_stride_bytes < (int64_t)t->n_feat * 4 || (row_stride_bytes & 3)) return b2s_int_fail(B2S_ERR_INVALID, "bad n / row stride");
    if (n == 0) return B2S_OK;
    TAB_TRY(cudaSetDevice(b2s_int_device()));
    return launch_lookup(t, d_keys, n, d_rows, row_stride_bytes, d_found, stream ? (cudaStream_t)stream : b2s_int_stream());
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

extern "C" int b2s_table_lookup_host(b2s_table_t t, const int64_t* keys, int64_t n, float* rows, int32_t* found, b2s_stats* stats) {
  try {  // no C++ exception crosses the C boundary
    if (!t || !keys || !rows || n < 0) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
    if (n == 0) return B2S_OK;
    std::lock_guard<std::mutex> lk(t->mu);
    TAB_TRY(cudaSetDevice(b2s_int_device()));
    if (n > t->cap_rows) {
      if (t->d_keys) { cudaFree(t->d_keys); cudaFree(t->d_out); cudaFree(t->d_found); t->d_keys = nullptr; }
      t->cap_rows = 0;
      const int64_t cap = std::max<int64_t>(n, 4096);
      TAB_TRY(cudaMalloc(&t->d_keys, cap * 8));
      TAB_TRY(cudaMalloc(&t->d_out, (size_t)cap * t->n_feat * 4));
      TAB_TRY(cudaMalloc(&t->d_found, cap * 4));
      t->cap_rows = cap;
    }
    cudaStream_t st = b2s_int_stream();
    const int64_t stride = (int64_t)t->n_feat * 4;
    TAB_TRY(cudaEventRecord(t->ev[0], st));
    TAB_TRY(cudaMemcpyAsync(t->d_keys, keys, n * 8, cudaMemcpyHostToDevice, st));
    TAB_TRY(cudaEventRecord(t->ev[1], st));
    if (int rc = launch_lookup(t, t->d_keys, n, t->d_out, stride, t->d_found, st)) return rc;
    TAB_TRY(cudaEventRecord(t->ev[2], st));
    TAB_TRY(cudaMemcpyAsync(rows, t->d_out, (size_t)n * stride, cudaMemcpyDeviceToHost, st));
    if (found) TAB_TRY(cudaMemcpyAsync(found, t->d_found, n * 4, cudaMemcpyDeviceToHost, st));
    TAB_TRY(cudaEventRecord(t->ev[3], st));
    TAB_TRY(cudaStreamSynchronize(st));
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->rows = n;
      cudaEventElapsedTime(&stats->h2d_ms, t->ev[0], t->ev[1]);
      cudaEventElapsedTime(&stats->kernel_ms, t->ev[1], t->ev[2]);
      cudaEventElapsedTime(&stats->d2h_ms, t->ev[2], t->ev[3]);
      stats->kernels = 1;
    }
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

static int launch_fused(b2s_table_t t, b2s_plan_t plan, const int64_t* d_keys, int64_t n, void* d_out, int32_t* d_status, cudaStream_t st) {
  B2SGather g{};
  g.d_keys = reinterpret_cast<const long long*>(d_keys);
  g.d_slots = t->d_slots;
  g.mask = t->cap - 1;
  g.d_values = t->d_values;
  g.missing_row = t->n_keys;
  g.h_impute = t->h_impute.data();
  g.any_impute = t->any_impute;
  g.n_feat = t->n_feat;
  return b2s_int_launch_gathered(plan, g, n, d_out, d_status, st);
}

extern "C" int b2s_table_enrich_device(b2s_table_t t, b2s_plan_t plan, const int64_t* d_keys, int64_t n, void* d_out,
                                       int32_t* d_status, void* stream) {
  try {  // no C++ exception crosses the C boundary
    if (!t || !plan || !d_keys || !d_out || n < 0) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
    if (n == 0) return B2S_OK;
    TAB_TRY(cudaSetDevice(b2s_int_device()));
    return launch_fused(t, plan, d_keys, n, d_out, d_status, stream ? (cudaStream_t)stream : b2s_int_stream());
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

static bool host_pinned(const void* ptr) {
  cudaPointerAttributes attr{};
  const bool yes = cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return yes;
}

extern "C" int b2s_table_enrich_host(b2s_table_t t, b2s_plan_t plan, const int64_t* keys, int64_t n, void* out, int64_t out_bytes,
                                     int32_t* row_status, b2s_stats* stats) {
  try {  // no C++ exception crosses the C boundary
    if (!t || !keys || !out || n < 0) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
    int n_in = 0, out_cols = 0;
    if (int rc = b2s_int_plan_shape(plan, &n_in, &out_cols)) return rc;
    if (n_in != t->n_feat) return b2s_int_fail(B2S_ERR_INVALID, "the table has %d features, the plan takes %d", t->n_feat, n_in);
    if (out_bytes < n * out_cols * 4) return b2s_int_fail(B2S_ERR_INVALID, "out buffer too small");
    if (n == 0) return B2S_OK;
    std::lock_guard<std::mutex> lk(t->mu);
    TAB_TRY(cudaSetDevice(b2s_int_device()));
    const int64_t stride = (int64_t)t->n_feat * 4;
    if (n > t->cap_rows) {
      if (t->d_keys) { cudaFree(t->d_keys); cudaFree(t->d_out); cudaFree(t->d_found); t->d_keys = nullptr; }
      t->cap_rows = 0;
      const int64_t cap = std::max<int64_t>(n, 4096);
      TAB_TRY(cudaMalloc(&t->d_keys, cap * 8));
      TAB_TRY(cudaMalloc(&t->d_out, (size_t)cap * stride));
      TAB_TRY(cudaMalloc(&t->d_found, cap * 4));
      t->cap_rows = cap;
    }
    if (n > t->enr_rows || out_cols > t->enr_out_cols) {
      if (t->d_votes) { cudaFree(t->d_votes); cudaFree(t->d_status); cudaFreeHost(t->h_pin); t->d_votes = nullptr; }
      t->enr_rows = 0;
      const int64_t cap = std::max<int64_t>(n, 4096);
      const int32_t oc = std::max(out_cols, t->enr_out_cols);
      TAB_TRY(cudaMalloc(&t->d_votes, (size_t)cap * oc * 4));
      TAB_TRY(cudaMalloc(&t->d_status, cap * 4));
      TAB_TRY(cudaMallocHost(&t->h_pin, (size_t)cap * (8 + (size_t)oc * 4 + 4)));
      t->enr_rows = cap;
      t->enr_out_cols = oc;
    }
    cudaStream_t st = b2s_int_stream();
    int64_t* h_keys = (int64_t*)t->h_pin;
    char* h_votes = t->h_pin + (size_t)t->enr_rows * 8;
    int32_t* h_status = (int32_t*)(h_votes + (size_t)t->enr_rows * t->enr_out_cols * 4);
    const size_t votes_sz = (size_t)n * out_cols * 4;
    // pinned caller buffers are used as they are; pageable ones go through the pinned block (one host memcpy each way)
    const void* k_src = keys;
    if (!host_pinned(keys)) {
      memcpy(h_keys, keys, (size_t)n * 8);
      k_src = h_keys;
    }
    void* v_dst = host_pinned(out) ? out : (void*)h_votes;
    int32_t* s_dst = row_status ? (host_pinned(row_status) ? row_status : h_status) : nullptr;
    TAB_TRY(cudaEventRecord(t->ev[0], st));
    TAB_TRY(cudaMemcpyAsync(t->d_keys, k_src, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    TAB_TRY(cudaEventRecord(t->ev[1], st));
    int n_kernels = 1;
    int rc = launch_fused(t, plan, t->d_keys, n, t->d_votes, t->d_status, st);  // gather inside the scoring kernel
    if (rc == B2S_ERR_UNSUPPORTED) {  // plans the gather loader does not cover: gather, score, fold the flags (3 launches)
      n_kernels = 3;
      if ((rc = launch_lookup(t, t->d_keys, n, t->d_out, stride, t->d_found, st))) return rc;
      if ((rc = b2s_run_device(plan, t->d_out, n, stride, t->d_votes, t->d_status, st))) return rc;
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(4 * b2s_int_sm_count(), (n + 255) / 256));
      b2s_int_count_launches(1);
      mark_unknown_kernel<<<grid, 256, 0, st>>>(t->d_found, t->d_status, n);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return b2s_int_fail(B2S_ERR_CUDA, "mark_unknown launch failed: %s", cudaGetErrorString(e));
    } else if (rc) {
      return rc;
    }
    TAB_TRY(cudaEventRecord(t->ev[2], st));
    TAB_TRY(cudaMemcpyAsync(v_dst, t->d_votes, votes_sz, cudaMemcpyDeviceToHost, st));
    if (s_dst) TAB_TRY(cudaMemcpyAsync(s_dst, t->d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    TAB_TRY(cudaEventRecord(t->ev[3], st));
    TAB_TRY(cudaStreamSynchronize(st));
    if (v_dst != out) memcpy(out, h_votes, votes_sz);
    if (s_dst && s_dst != row_status) memcpy(row_status, h_status, (size_t)n * 4);
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->rows = n;
      cudaEventElapsedTime(&stats->h2d_ms, t->ev[0], t->ev[1]);
      cudaEventElapsedTime(&stats->kernel_ms, t->ev[1], t->ev[2]);
      cudaEventElapsedTime(&stats->d2h_ms, t->ev[2], t->ev[3]);
      stats->kernels = n_kernels;  // 1: gather fused into the scoring kernel; 3: gather + the plan + mark_unknown
      if (row_status)
        for (int64_t r = 0; r < n; ++r) stats->nonfinite_rows += (row_status[r] & B2S_ROW_NONFINITE_INPUT) ? 1 : 0;
    }
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

extern "C" int b2s_table_time_device(b2s_table_t t, const int64_t* const* d_keys, int32_t n_bufs, int64_t n, float* d_rows,
                                     int64_t row_stride_bytes, int32_t* d_found, int32_t n_iters, float* total_ms) {
  try {  // no C++ exception crosses the C boundary
    if (!t || !d_keys || n_bufs <= 0 || n_iters <= 0 || !total_ms) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
    TAB_TRY(cudaSetDevice(b2s_int_device()));
    cudaStream_t st = b2s_int_stream();
    std::lock_guard<std::mutex> lk(t->mu);
    TAB_TRY(cudaEventRecord(t->ev[0], st));
    for (int i = 0; i < n_iters; ++i)
      if (int rc = launch_lookup(t, d_keys[i % n_bufs], n, d_rows, row_stride_bytes, d_found, st)) return rc;
    TAB_TRY(cudaEventRecord(t->ev[1], st));
    TAB_TRY(cudaStreamSynchronize(st));
    TAB_TRY(cudaEventElapsedTime(total_ms, t->ev[0], t->ev[1]));
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

extern "C" int b2s_table_info(b2s_table_t t, int64_t* n_keys, int32_t* n_features, int64_t* capacity) {
  try {  // no C++ exception crosses the C boundary
    if (!t) return b2s_int_fail(B2S_ERR_INVALID, "null table");
    if (n_keys) *n_keys = t->n_keys;
    if (n_features) *n_features = t->n_feat;
    if (capacity) *capacity = (int64_t)t->cap;
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

extern "C" int b2s_table_destroy(b2s_table_t t) {
  try {  // no C++ exception crosses the C boundary
    if (!t) return B2S_OK;
    if (t->d_slots) cudaFree(t->d_slots);
    if (t->d_values) cudaFree(t->d_values);
    if (t->d_impute) cudaFree(t->d_impute);
    if (t->d_keys) cudaFree(t->d_keys);
    if (t->d_out) cudaFree(t->d_out);
    if (t->d_found) cudaFree(t->d_found);
    if (t->d_votes) cudaFree(t->d_votes);
    if (t->d_status) cudaFree(t->d_status);
    if (t->h_pin) cudaFreeHost(t->h_pin);
    for (auto& e : t->ev)
      if (e) cudaEventDestroy(e);
    delete t;
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}

// FNV-1a over each string of a packed buffer: the 64-bit entity key of a string-valued entity (host code)
extern "C" int b2s_hash_strings(const char* bytes, const int64_t* offsets, int64_t n, int64_t* keys_out) {
  try {  // no C++ exception crosses the C boundary
    if (!bytes || !offsets || !keys_out || n < 0) return b2s_int_fail(B2S_ERR_INVALID, "bad arguments");
    for (int64_t i = 0; i < n; ++i) {
      uint64_t h = 1469598103934665603ULL;
      for (int64_t j = offsets[i]; j < offsets[i + 1]; ++j) {
        h ^= (unsigned char)bytes[j];
        h *= 1099511628211ULL;
      }
      keys_out[i] = (int64_t)h;
    }
    return B2S_OK;
  } catch (const std::exception& e) {
    return b2s_int_fail(B2S_ERR_INVALID, "%s: %s", __func__, e.what());
  }
}
