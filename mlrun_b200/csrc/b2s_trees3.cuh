// b2s_trees3.cuh -- tree-ensemble scorer, round 2 (sm_100a): "parts" resident in shared memory.
//
// What bounds a root->leaf walk (measured on B200, profiles/r2/trees_lab_r2a*.{txt,csv}): with the model in shared memory
// and the event tile transposed (xt[feature][row], lanes = 32 consecutive rows walking the same tree) every LDS is
// conflict free, and the kernel runs exactly at the LSU limit of one 128-byte shared-memory wavefront per cycle and
// SM -- issue slots are 35-40 % busy.  So the design minimises *wavefronts per visit*, not instructions:
//   * 8-byte heap nodes {x offset, threshold}: one LDS.64 (2 wavefronts, same as 2 x LDS.32, one instruction fewer);
//   * the top two levels of every tree are read once per tree with warp-uniform loads (one broadcast wavefront each)
//     and kept in registers for the warp's RPT row blocks: levels 0-1 cost only their x gathers;
//   * node addresses are carried as absolute shared-window addresses, a' = 2a + ((right ? 8 : 0) - tree_base): a visit is
//     LDS.64, IADD, LDS, FSETP, SEL, IADD3.
//
// Work decomposition.  A *part* is what one CTA keeps resident: the trees of one (model, score slot) -- split further when
// they do not fit -- re-packed on the host as complete heap-ordered depth-D trees (early leaves are padded: +inf threshold,
// both children carry the leaf), or ALL linear models of the ensemble (one part: fp64 weights).  Parts own CTAs in
// proportion to their cost; each CTA streams the tiles t = rank, rank + n_ctas, ... of the batch:
//     TMA boxes (32 floats x 64 rows, 128-byte swizzle) / cp.async  ->  landing tile  ->  transpose (+ Imputer, + the
//     non-finite test, + order-preserving integer keys when NaN routing is on)  ->  next tile's load is issued  ->  walk
//     ->  per-warp partial sums combined in a fixed order  ->  partial[column][row] (fp64, coalesced).
// `t3_vote_kernel` then adds each model's columns to its init scores in column order, applies the link and the
// VotingEnsemble reduce (serving/routers.py:708-741).  Multi-class GradientBoosting (n_classes x n_estimators trees) and
// ensembles mixing linear and tree scorers (BASELINE configs[3]) therefore run on this path too.
//
// Missing values (xgboost / LightGBM / scikit-learn >= 1.3 trees route NaN to a per-node default child): with MISS the
// transposed tile holds order-preserving int32 keys in TWO copies -- NaN = INT_MAX in copy A (compares greater than every
// threshold: goes right), NaN = INT_MIN in copy B (goes left) -- and a node's x offset points into the copy that matches
// its default direction, so the walk itself is unchanged (ISETP instead of FSETP).  `x < t` (xgboost) is `x <= prev(t)`:
// thresholds are converted when the model is added, not in the kernel.
#pragma once
#include "b2s_device.cuh"

namespace b2s {

constexpr int kT3RPT = 2;            // row blocks (of 32 rows) per warp: the tile is 64 rows
constexpr int kT3TR = 32 * kT3RPT;   // rows per tile
constexpr int kT3U = 2;              // trees in flight per warp (x RPT rows = 4 independent walks per thread)
constexpr int kT3MaxLin = 8;         // score columns of the linear part
constexpr int kT3MaxDepth = 8;

struct T3Part {           // one per part, in global memory
  const uint2* nodes;     // trees: [n_trees][1 << D] heap nodes (slot 0 unused) {x byte offset in the tile, threshold bits}
  const double* leaves;   // trees: [n_trees][1 << D] tree_scale * leaf value;  linear part: weights [n_cols][n_in]
  int32_t n_trees;        // 0: the linear part
  int32_t n_cols;         // columns of `partial` this part writes (trees: 1)
  int32_t col0;
  int32_t cta0, n_ctas;   // the CTAs [cta0, cta0 + n_ctas) of the grid work on this part
  int32_t flags_rows;     // != 0: this part's CTAs also write the per-row "non-finite input" flags
};

struct T3Params {
  const char* rows;
  int64_t row_stride;
  int64_t n_rows;
  double* partial;        // [n_cols_total][col_stride]
  int64_t col_stride;
  int32_t* row_bad;       // [n_rows]
  const T3Part* parts;
  const float* fill;      // [n_in] Imputer values (NaN: column not imputed)
  int32_t n_in, n_parts, warps, use_tmap, vec_ok, pitch, any_fill;
  int32_t sm_leaf, sm_fill, sm_part, sm_xt, sm_land, sm_bad, sm_bar;  // byte offsets into dynamic shared memory
  int32_t xt_words;       // words of one transposed tile copy (n_in rounded up to 4, times TR)
};

// explicit shared-window loads (32-bit addresses: no generic->shared conversion in the address arithmetic)
__device__ __forceinline__ uint2 t3_lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 t3_lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t t3_lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ double t3_ldsd(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

// order-preserving int32 key of a float (monotone for every non-NaN value; -0 and +0 share a key)
__device__ __forceinline__ int32_t t3_key(float x) {
  const int32_t b = __float_as_int(x + 0.0f);  // -0 -> +0
  return b ^ ((b >> 31) & 0x7fffffff);
}

template <bool MISS>
__device__ __forceinline__ bool t3_right(uint32_t x, uint32_t thr) {
  // floats: sklearn's rule "left when x <= threshold"; keys: the same order on integers (NaN keys sit at the ends)
  return MISS ? ((int32_t)x > (int32_t)thr) : !(__uint_as_float(x) <= __uint_as_float(thr));
}

template <int D, bool MISS>
__global__ void __launch_bounds__(1024) trees3_kernel(const __grid_constant__ T3Params p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) unsigned char smem3[];
  unsigned char* const smem = smem3;
  constexpr int NN = 1 << D;  // node slots per tree (1-based heap) == leaves per tree
  constexpr int TR = kT3TR, RPT = kT3RPT, U = kT3U;
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int W = p.warps, nthr = W * 32;

  int pi = 0;
  while (pi + 1 < p.n_parts && (int)blockIdx.x >= p.parts[pi].cta0 + p.parts[pi].n_ctas) ++pi;
  const T3Part part = p.parts[pi];
  const int cta = (int)blockIdx.x - part.cta0;
  if ((int64_t)cta * kT3TR >= p.n_rows) return;  // small batch: this CTA has no tile (decided before the tables are loaded)
  const int NT = part.n_trees;
  const bool is_lin = NT == 0;

  unsigned char* s_nodes = smem;
  double* s_leaf = reinterpret_cast<double*>(smem + p.sm_leaf);
  float* s_fill = reinterpret_cast<float*>(smem + p.sm_fill);
  double* s_part = reinterpret_cast<double*>(smem + p.sm_part);  // trees: [W][TR]; linear part: [n_cols][TR]
  uint32_t* s_xt = reinterpret_cast<uint32_t*>(smem + p.sm_xt);  // [n_in][TR] (x2 with MISS)
  float* s_land = reinterpret_cast<float*>(smem + p.sm_land);    // 1024-byte aligned (TMA swizzle atom)
  int* s_bad = reinterpret_cast<int*>(smem + p.sm_bad);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + p.sm_bar);

  // ---- the part's tables -> shared memory (once per CTA)
  if (!is_lin) {
    uint2* sn = reinterpret_cast<uint2*>(s_nodes);
    for (int i = tid; i < NT * NN; i += nthr) {
      sn[i] = part.nodes[i];
      s_leaf[i] = part.leaves[i];
    }
  } else {
    double* sw = reinterpret_cast<double*>(s_nodes);
    for (int i = tid; i < part.n_cols * p.n_in; i += nthr) sw[i] = part.leaves[i];
  }
  for (int i = tid; i < p.n_in; i += nthr) s_fill[i] = p.fill[i];
  if (tid < TR) s_bad[tid] = 0;
  const bool tma = p.use_tmap != 0;
  if (tma && tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t tma_phase = 0;

  const int64_t n_tiles = (p.n_rows + TR - 1) / TR;
  const int cprv = p.vec_ok ? (p.n_in >> 2) : p.n_in;
  auto issue = [&](int64_t row0) {
    if (tma) {  // one thread: n_in / 32 box copies of (32 floats x TR rows); rows past the end arrive as zeros
      if (tid == 0) {
        const int boxes = p.n_in >> 5;
        mbar_expect_tx(s_bar, (uint32_t)boxes * (uint32_t)TR * 128u);
        for (int b = 0; b < boxes; ++b) tensor_load_2d(s_land + b * TR * 32, &tmap, b * 32, (int)row0, s_bar);
      }
      return;
    }
    const int64_t left = p.n_rows - row0;
    const int rows = left < TR ? (left < 0 ? 0 : (int)left) : TR;
    const char* base = p.rows + row0 * p.row_stride;
    for (int i = tid; i < rows * cprv; i += nthr) {
      const int rr = i / cprv, cc = i - rr * cprv;
      if (p.vec_ok)
        cp_async16(s_land + rr * p.pitch + cc * 4, base + (int64_t)rr * p.row_stride + cc * 16);
      else
        cp_async4(s_land + rr * p.pitch + cc, base + (int64_t)rr * p.row_stride + cc * 4);
    }
  };

  __syncthreads();  // tables and the barrier are initialised before anybody uses them
  if ((int64_t)cta < n_tiles) issue((int64_t)cta * TR);
  cp_async_commit();

  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t xls = sbase + (uint32_t)p.sm_xt + (uint32_t)lane * 4u;  // this lane's column of the transposed tile
  const uint32_t leaf0 = (uint32_t)p.sm_leaf - (uint32_t)(NN * 8);       // leaf address = node address + leaf0
  const int TPW = (NT + W - 1) / W;                                      // trees per warp
  const int ncols = part.n_cols;

  for (int64_t t = cta; t < n_tiles; t += part.n_ctas) {
    const int64_t row0 = t * TR;
    if (tma) {
      mbar_wait(s_bar, tma_phase);
      tma_phase ^= 1u;
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();  // landing tile visible; everybody is done with the previous tile (walk and combine)
    {  // ---- transpose (+ Imputer, non-finite test, keys): lanes take consecutive rows, LDS.128 and STS are conflict free
      const int64_t left = p.n_rows - row0;
      const int rows = left < TR ? (int)left : TR;
      if (p.vec_ok) {
        for (int i = tid; i < (p.n_in >> 2) * TR; i += nthr) {
          const int c = i / TR, rr = i - c * TR;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rr < rows)
            v = tma ? *reinterpret_cast<const float4*>(s_land + (c >> 3) * (TR * 32) + rr * 32 + (((c & 7) ^ (rr & 7)) << 2))
                    : *reinterpret_cast<const float4*>(s_land + rr * p.pitch + c * 4);
          float xs[4] = {v.x, v.y, v.z, v.w};
          bool bad = false;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float x = xs[u];
            if (p.any_fill) {
              const float f = s_fill[c * 4 + u];
              x = (x != x) ? f : x;  // Imputer._impute (feature_store/steps.py:397-406); f is NaN where nothing is imputed
            }
            // what scikit-learn's check_array refuses: Inf always, NaN unless every model routes missing values
            bad |= MISS ? (fabsf(x) == __int_as_float(0x7f800000)) : !is_finite_f(x);
            uint32_t* o = s_xt + (size_t)(c * 4 + u) * TR + rr;
            if (MISS) {
              const bool isn = x != x;
              const int32_t k = t3_key(x);
              o[0] = (uint32_t)(isn ? 0x7fffffff : k);
              o[p.xt_words] = (uint32_t)(isn ? (int32_t)0x80000000 : k);
            } else {
              o[0] = __float_as_uint(x);
            }
          }
          if (part.flags_rows && bad) atomicOr(&s_bad[rr], 1);
        }
      } else {
        for (int i = tid; i < p.n_in * TR; i += nthr) {
          const int f = i / TR, rr = i - f * TR;
          float x = rr < rows ? s_land[rr * p.pitch + f] : 0.0f;
          if (p.any_fill) {
            const float fv = s_fill[f];
            x = (x != x) ? fv : x;
          }
          const bool bad = MISS ? (fabsf(x) == __int_as_float(0x7f800000)) : !is_finite_f(x);
          if (MISS) {
            const bool isn = x != x;
            const int32_t k = t3_key(x);
            s_xt[i] = (uint32_t)(isn ? 0x7fffffff : k);
            s_xt[i + p.xt_words] = (uint32_t)(isn ? (int32_t)0x80000000 : k);
          } else {
            s_xt[i] = __float_as_uint(x);
          }
          if (part.flags_rows && bad) atomicOr(&s_bad[rr], 1);
        }
      }
    }
    __syncthreads();  // transposed tile visible; landing tile free
    {
      const int64_t tn = t + part.n_ctas;
      if (tn < n_tiles) issue(tn * TR);  // lands while this tile is walked
      cp_async_commit();
    }

    if (!is_lin) {
      // ---- walk: warp g takes the trees g, g + W, ...; lane = row (+ 32 j)
      double acc[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) acc[j] = 0.0;
      for (int i = 0; i < TPW; i += U) {
        uint32_t tba[U], cl[U], cr[U], a[U][RPT];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int tr = g + (i + u) * W;
          valid[u] = (i + u) < TPW && tr < NT;
          tba[u] = sbase + (uint32_t)((valid[u] ? tr : 0) * (NN * 8));
          cl[u] = 0u - tba[u];
          cr[u] = 8u - tba[u];
        }
        uint2 nd[U][RPT];
        uint32_t x[U][RPT];
        {  // levels 0 and 1: nodes 1..3 of the tree, one warp-uniform LDS.64 + LDS.128 for all RPT row blocks
          uint2 n1[U], n2[U], n3[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            n1[u] = t3_lds64(tba[u] + 8);
            const uint4 q = t3_lds128(tba[u] + 16);
            n2[u] = make_uint2(q.x, q.y);
            n3[u] = make_uint2(q.z, q.w);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            x[u][0] = t3_lds32<0>(xls + n1[u].x);
            if (RPT > 1) x[u][1] = t3_lds32<128>(xls + n1[u].x);
          }
          bool r0[U][RPT];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              r0[u][j] = t3_right<MISS>(x[u][j], n1[u].y);
              nd[u][j].x = r0[u][j] ? n3[u].x : n2[u].x;
              nd[u][j].y = r0[u][j] ? n3[u].y : n2[u].y;
            }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            x[u][0] = t3_lds32<0>(xls + nd[u][0].x);
            if (RPT > 1) x[u][1] = t3_lds32<128>(xls + nd[u][1].x);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              const bool r1 = t3_right<MISS>(x[u][j], nd[u][j].y);
              a[u][j] = tba[u] + 32u + (r0[u][j] ? 16u : 0u) + (r1 ? 8u : 0u);
            }
        }
#pragma unroll
        for (int d = 2; d < D; ++d) {
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < RPT; ++j) nd[u][j] = t3_lds64(a[u][j]);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            x[u][0] = t3_lds32<0>(xls + nd[u][0].x);
            if (RPT > 1) x[u][1] = t3_lds32<128>(xls + nd[u][1].x);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < RPT; ++j) a[u][j] = a[u][j] + a[u][j] + (t3_right<MISS>(x[u][j], nd[u][j].y) ? cr[u] : cl[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const double v = t3_ldsd(a[u][j] + leaf0);
            if (valid[u]) acc[j] = __dadd_rn(acc[j], v);
          }
      }
#pragma unroll
      for (int j = 0; j < RPT; ++j) s_part[g * TR + j * 32 + lane] = acc[j];
    } else if (g < ncols) {
      // ---- the linear part: warp s computes score column s of the tile's rows, features in order (fp64 products of
      // float32 inputs are exact; two interleaved chains per row hide the DFMA latency)
      const double* sw = reinterpret_cast<const double*>(s_nodes) + (size_t)g * p.n_in;
      double a0[RPT], a1[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) a0[j] = a1[j] = 0.0;
      int f = 0;
      for (; f + 1 < p.n_in; f += 2) {
        const double w0 = sw[f], w1 = sw[f + 1];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          a0[j] = fma(w0, (double)__uint_as_float(s_xt[(size_t)f * TR + j * 32 + lane]), a0[j]);
          a1[j] = fma(w1, (double)__uint_as_float(s_xt[(size_t)(f + 1) * TR + j * 32 + lane]), a1[j]);
        }
      }
      if (f < p.n_in) {
        const double w0 = sw[f];
#pragma unroll
        for (int j = 0; j < RPT; ++j) a0[j] = fma(w0, (double)__uint_as_float(s_xt[(size_t)f * TR + j * 32 + lane]), a0[j]);
      }
#pragma unroll
      for (int j = 0; j < RPT; ++j) s_part[g * TR + j * 32 + lane] = a0[j] + a1[j];
    }
    __syncthreads();
    // ---- combine the warps' partial sums in a fixed order (deterministic fp64) and store column-major
    for (int i = tid; i < ncols * TR; i += nthr) {
      const int s = i / TR, r = i - s * TR;
      if (row0 + r < p.n_rows) {
        double sum = 0.0;
        const int n_sum = is_lin ? 1 : W;  // tree parts: one partial per warp; the linear part: warp s wrote column s
        for (int gg = 0; gg < n_sum; ++gg) sum = __dadd_rn(sum, s_part[(gg * ncols + s) * TR + r]);
        p.partial[(int64_t)(part.col0 + s) * p.col_stride + row0 + r] = sum;
      }
    }
    if (part.flags_rows && tid < TR) {  // flags gathered during the transpose; reset for the next tile
      if (row0 + tid < p.n_rows) p.row_bad[row0 + tid] = s_bad[tid];
      s_bad[tid] = 0;
    }
    // (the barrier at the top of the next iteration orders these reads before the next tile's writes)
  }
  cp_async_wait<0>();
}

// Per row: scores = init + the model's columns of `partial` in column order, link, then the VotingEnsemble reduce.
__global__ void __launch_bounds__(256) t3_vote_kernel(KParams kp, const double* __restrict__ partial, int64_t col_stride,
                                                      const int32_t* __restrict__ col_score, int n_cols,
                                                      const int32_t* __restrict__ row_bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < kp.n_rows; row += stride) {
    double sc[kMaxScores];
    for (int k = 0; k < kp.n_scores; ++k) sc[k] = kp.bias[k];
    for (int c = 0; c < n_cols; ++c) {
      const int k = col_score[c];
      sc[k] = __dadd_rn(sc[k], partial[(int64_t)c * col_stride + row]);
    }
    double pred[kMaxModels];
    for (int m = 0; m < kp.n_models; ++m) {
      const ModelDesc md = kp.models[m];
      pred[m] = apply_link(md, sc + md.score_off, kp.classes);
    }
    vote_and_store(kp, pred, row, row_bad[row] ? 1u : 0u);
  }
}

}  // namespace b2s
