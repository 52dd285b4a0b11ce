// b2s_rowthread.cuh -- linear path, one thread per event row, every table operand in the constant bank.
//
//   HBM rows --cp.async 16 B (LDGSTS), STAGES-deep ring--> shared-memory tile, pitch = 16 mod 128 bytes
//   thread r reads row r with conflict-free LDS.128 and runs a fully unrolled column loop whose
//   per-column operands (Imputer fill, copy mask, the NS fp64 weights) are *immediate constant-bank
//   operands* of the FSEL / LOP3 / DFMA instructions: the plan's tables travel as a __grid_constant__
//   kernel parameter, so the inner loop has no table loads and no cross-lane traffic at all:
//        FSETP+FSEL (NaN -> fill)   LOP3 (drop non-copied columns)   F2F   NS x DFMA     per value
//   one-hot columns are a short second loop over the categorical columns (value re-read from the
//   tile, category index by compares against constant-bank categories, weights gathered from shared
//   memory; a zero row stands for "no category matched"); bias, link, vote and the 4-byte store are
//   per-thread.  A non-finite model input surfaces as a non-finite score (NaN/Inf survive fma even
//   with a zero weight), which is what the per-row status tests.
#pragma once
#include "b2s_device.cuh"

namespace b2s {

constexpr int kRTMaxCatCols = 32;
constexpr int kRTMaxCats = 256;

template <int NCH, int NS>
struct RTParams {
  const char* rows;
  int64_t row_stride;
  int64_t n_rows;
  float* out;
  int32_t* status;
  int32_t n_in, out_cols, n_models, vote_kind, out_is_int, fast_epilogue, tile_rows, pitch, stages, vec_ok;
  int32_t n_cat_cols, n_cat;
  const double* wcat;       // [n_cat][NS] (global; copied to shared memory, plus a zero row)
  const double* vote_w_g;   // generic epilogue
  const ModelDesc* models;
  const int32_t* classes;
  double w[NCH * 4][NS];    // constant-bank operands
  float fill[NCH * 4];      // NaN: column not imputed
  uint32_t cmask[NCH * 4];  // all-ones: column feeds a COPY output
  double bias[NS];
  double vote_w[NS];
  int32_t cat_col[kRTMaxCatCols];   // input column of each categorical column
  int32_t cat_base[kRTMaxCatCols];  // first category (index into cat_val / wcat)
  int32_t cat_cnt[kRTMaxCatCols];
  float cat_val[kRTMaxCats];
};

template <int NCH, int NS>
__global__ void __launch_bounds__(128) rowthread_kernel(const __grid_constant__ RTParams<NCH, NS> p) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* s_wcat = reinterpret_cast<double*>(smem);
  float* s_tiles = reinterpret_cast<float*>(smem + (((size_t)(p.n_cat + 1) * NS * 8 + 15) / 16) * 16);

  const int tid = threadIdx.x;
  const int TR = p.tile_rows;
  const int S = p.stages;
  const int tile_words = TR * p.pitch;
  const int64_t n_tiles = (p.n_rows + TR - 1) / TR;

  const int cprv = p.n_in >> 2;  // 16-byte chunks per row
  const int r0v = cprv ? tid / cprv : 0, c0v = cprv ? tid - r0v * cprv : 0;
  const int drv = cprv ? (int)blockDim.x / cprv : 0, dcv = cprv ? (int)blockDim.x - drv * cprv : 0;
  const int r0s = tid / p.n_in, c0s = tid - r0s * p.n_in;
  const int drs = (int)blockDim.x / p.n_in, dcs = (int)blockDim.x - drs * p.n_in;
  auto issue = [&](float* tile, int64_t row0) {
    int64_t left = p.n_rows - row0;
    const int rows = left < TR ? (left < 0 ? 0 : (int)left) : TR;
    const char* base = p.rows + row0 * p.row_stride;
    // (row, chunk) walk without per-iteration division: thread i starts at i and advances by blockDim.x
    if (p.vec_ok) {
      int r = r0v, c = c0v;
      while (r < rows) {
        cp_async16(tile + r * p.pitch + c * 4, base + (int64_t)r * p.row_stride + c * 16);
        r += drv;
        c += dcv;
        if (c >= cprv) {
          c -= cprv;
          ++r;
        }
      }
    } else {
      int r = r0s, c = c0s;
      while (r < rows) {
        cp_async4(tile + r * p.pitch + c, base + (int64_t)r * p.row_stride + c * 4);
        r += drs;
        c += dcs;
        if (c >= p.n_in) {
          c -= p.n_in;
          ++r;
        }
      }
    }
  };

  for (int s = 0; s < S - 1; ++s) {
    const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
    if (t < n_tiles) issue(s_tiles + s * tile_words, t * TR);
    cp_async_commit();
  }
  for (int i = tid; i < p.n_cat * NS; i += blockDim.x) s_wcat[i] = p.wcat[i];
  for (int i = tid; i < NS; i += blockDim.x) s_wcat[p.n_cat * NS + i] = 0.0;

  int stage = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    if (S == 2) cp_async_wait<0>();
    else if (S == 3) cp_async_wait<1>();
    else cp_async_wait<2>();
    __syncthreads();
    {
      const int64_t tn = t + (int64_t)(S - 1) * gridDim.x;
      int sn = stage + S - 1;
      if (sn >= S) sn -= S;
      if (tn < n_tiles) issue(s_tiles + sn * tile_words, tn * TR);
      cp_async_commit();
    }
    const float* tile = s_tiles + stage * tile_words;
    const int64_t row = t * TR + tid;
    if (tid < TR && row < p.n_rows) {
      const float* xr = tile + tid * p.pitch;
      double acc[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) acc[k] = p.bias[k];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const float4 v = *reinterpret_cast<const float4*>(xr + ch * 4);
        const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = ch * 4 + u;
          float x = xs[u];
          x = (x != x) ? p.fill[c] : x;                                           // Imputer
          const double xd = (double)__uint_as_float(__float_as_uint(x) & p.cmask[c]);  // non-copied -> +0
#pragma unroll
          for (int k = 0; k < NS; ++k) acc[k] = fma(p.w[c][k], xd, acc[k]);
        }
      }
      // one-hot columns: onehot(x) . w  ==  w[cat_base + index_of(x)]
      for (int cc = 0; cc < p.n_cat_cols; ++cc) {
        const int col = p.cat_col[cc];
        float x = xr[col];
        x = (x != x) ? p.fill[col] : x;
        const int b0 = p.cat_base[cc], n = p.cat_cnt[cc];
        int j = p.n_cat;  // the zero row: no category matched
        for (int q = 0; q < n; ++q) j = (x == p.cat_val[b0 + q]) ? b0 + q : j;
        const double* wc = s_wcat + (size_t)j * NS;
#pragma unroll
        for (int k = 0; k < NS; ++k) acc[k] += wc[k];
      }
      uint32_t st = 0;
#pragma unroll
      for (int k = 0; k < NS; ++k) st |= (fabs(acc[k]) <= 1.7976931348623157e308) ? 0u : 1u;
      if (p.fast_epilogue) {
        if (p.vote_kind == 1) {  // VotingEnsemble._mean_vote: sum_m w[m] * pred[m], model order
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < NS; ++k) s = __dadd_rn(s, __dmul_rn(acc[k], p.vote_w[k]));
          p.out[row] = (float)s;
        } else {
#pragma unroll
          for (int k = 0; k < NS; ++k)
            if (k < p.n_models) p.out[row * p.out_cols + k] = (float)acc[k];
        }
        if (p.status) p.status[row] = (int32_t)st;
      } else {
        double sl[NS];
#pragma unroll
        for (int k = 0; k < NS; ++k) sl[k] = acc[k];
        double pred[kMaxModels];
        for (int m = 0; m < p.n_models; ++m) {
          const ModelDesc md = p.models[m];
          pred[m] = apply_link(md, sl + md.score_off, p.classes);
        }
        KParams kp;  // vote_and_store only reads these fields
        kp.out = p.out;
        kp.out_cols = p.out_cols;
        kp.n_models = p.n_models;
        kp.vote_kind = p.vote_kind;
        kp.out_is_int = p.out_is_int;
        kp.vote_w = p.vote_w_g;
        kp.status = p.status;
        vote_and_store(kp, pred, row, st);
      }
    }
    ++stage;
    if (stage == S) stage = 0;
  }
  cp_async_wait<0>();
}

}  // namespace b2s
