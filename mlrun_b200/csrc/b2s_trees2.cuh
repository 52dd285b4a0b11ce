// b2s_trees2.cuh -- tree-ensemble scorer with the model resident in shared memory (sm_100a).
//
// A root->leaf walk is a chain of dependent gathers; from L2 (the generic kernel, tables ~1 MB for
// 4 x 100 depth-6 trees) every level costs ~250 cycles.  Here each persistent CTA owns ONE model of the
// router: its trees are re-packed on the host into complete (heap-ordered) depth-D trees -- 8-byte
// internal nodes {feature, float32 threshold}, children implicit (2i+1 + go_right), fp64 leaves -- about
// 1 KB per depth-6 tree, ~100 KB per 100-tree model, copied to shared memory once per kernel.  Rows
// stream through a cp.async ring (same padded tile as the linear kernels); thread (g, r) walks the trees
// t = g, g+G, ... of row r, so one level is two LDS (node, feature value) + compare + index update.
// The G partial sums of a row are combined in shared memory in a fixed order; the link function
// (identity / >0 / >=0 / argmax) is applied per model and the prediction goes to a small fp64 buffer
// pred[row][model]; `vote_kernel` then applies the VotingEnsemble reduce (routers.py:708-741) and the
// row status.  Rows are read once per model (M x 4*n_in bytes/event, mostly from L2): the path is
// bound by shared-memory gather throughput, not by HBM.
#pragma once
#include "b2s_device.cuh"

namespace b2s {

struct HeapNode {
  int32_t feature;
  float threshold;  // go right when !(x <= threshold)
};

struct T2Model {  // one per model, in global memory
  const HeapNode* nodes;   // [n_trees][n_internal]
  const double* leaves;    // [n_trees][n_leaves]
  const int32_t* slot;     // [n_trees]
  const double* scale;     // [n_trees]
  int32_t n_trees, depth, n_internal, n_leaves;
};

struct T2Params {
  const char* rows;
  int64_t row_stride;
  int64_t n_rows;
  double* pred;        // [n_rows][n_models]
  int32_t* row_bad;    // [n_rows] non-finite input flags (written by the CTAs of model 0)
  int32_t n_in, n_models, tile_rows, pitch, stages, vec_ok, groups;
  const T2Model* t2;   // [n_models]
  const ModelDesc* models;
  const int32_t* classes;
  const double* bias;  // init scores, indexed by ModelDesc.score_off
  int32_t sm_tables, sm_part, sm_tiles;  // byte offsets
};

template <int NS>
__global__ void __launch_bounds__(512) trees_model_kernel(const __grid_constant__ T2Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x;
  const int m = blockIdx.x % p.n_models;
  const int part = blockIdx.x / p.n_models;
  const int nparts = (gridDim.x - m + p.n_models - 1) / p.n_models;  // CTAs working on model m
  const T2Model tm = p.t2[m];
  const ModelDesc md = p.models[m];

  // ---- the model's tables -> shared memory (once)
  HeapNode* s_nodes = reinterpret_cast<HeapNode*>(smem + p.sm_tables);
  const int n_nodes = tm.n_trees * tm.n_internal;
  double* s_leaves = reinterpret_cast<double*>(smem + p.sm_tables + (((size_t)n_nodes * 8 + 15) / 16) * 16);
  const int n_leaves = tm.n_trees * tm.n_leaves;
  {
    const int2* src = reinterpret_cast<const int2*>(tm.nodes);
    int2* dst = reinterpret_cast<int2*>(s_nodes);
    for (int i = tid; i < n_nodes; i += blockDim.x) dst[i] = __ldg(src + i);
    for (int i = tid; i < n_leaves; i += blockDim.x) s_leaves[i] = __ldg(tm.leaves + i);
  }
  double* s_part = reinterpret_cast<double*>(smem + p.sm_part);  // [groups][tile_rows][NS]
  float* s_tiles = reinterpret_cast<float*>(smem + p.sm_tiles);

  const int TR = p.tile_rows;
  const int G = p.groups;
  const int S = p.stages;
  const int tile_words = TR * p.pitch;
  const int64_t n_tiles = (p.n_rows + TR - 1) / TR;
  const int g = tid / TR;  // tree group (warp-uniform: TR is a multiple of 32)
  const int r = tid - g * TR;

  const int cprv = p.vec_ok ? (p.n_in >> 2) : p.n_in;
  const int r0 = tid / cprv, c0 = tid - r0 * cprv;
  const int dr = (int)blockDim.x / cprv, dc = (int)blockDim.x - dr * cprv;
  auto issue = [&](float* tile, int64_t row0) {
    int64_t left = p.n_rows - row0;
    const int rows = left < TR ? (left < 0 ? 0 : (int)left) : TR;
    const char* base = p.rows + row0 * p.row_stride;
    int rr = r0, cc = c0;
    while (rr < rows) {
      if (p.vec_ok)
        cp_async16(tile + rr * p.pitch + cc * 4, base + (int64_t)rr * p.row_stride + cc * 16);
      else
        cp_async4(tile + rr * p.pitch + cc, base + (int64_t)rr * p.row_stride + cc * 4);
      rr += dr;
      cc += dc;
      if (cc >= cprv) {
        cc -= cprv;
        ++rr;
      }
    }
  };

  for (int s = 0; s < S - 1; ++s) {
    const int64_t t = (int64_t)part + (int64_t)s * nparts;
    if (t < n_tiles) issue(s_tiles + s * tile_words, t * TR);
    cp_async_commit();
  }
  int stage = 0;
  for (int64_t t = part; t < n_tiles; t += nparts) {
    if (S == 2) cp_async_wait<0>();
    else cp_async_wait<1>();
    __syncthreads();  // tile (and, first time round, the tables) visible; previous tile fully consumed
    {
      const int64_t tn = t + (int64_t)(S - 1) * nparts;
      int sn = stage + S - 1;
      if (sn >= S) sn -= S;
      if (tn < n_tiles) issue(s_tiles + sn * tile_words, tn * TR);
      cp_async_commit();
    }
    const float* xr = s_tiles + stage * tile_words + r * p.pitch;
    const int64_t row = t * TR + r;
    const bool live = row < p.n_rows;
    double acc[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) acc[k] = 0.0;
    if (live) {
      for (int tr = g; tr < tm.n_trees; tr += G) {
        const HeapNode* tn = s_nodes + tr * tm.n_internal;
        int node = 0;
        for (int d = 0; d < tm.depth; ++d) {
          const HeapNode nd = tn[node];
          const float x = xr[nd.feature];
          node = 2 * node + 1 + ((x <= nd.threshold) ? 0 : 1);  // sklearn: left when x <= threshold
        }
        const double v = __dmul_rn(tm.scale[tr], s_leaves[tr * tm.n_leaves + (node - tm.n_internal)]);
        const int slot = tm.slot[tr];
#pragma unroll
        for (int k = 0; k < NS; ++k)
          if (k == slot) acc[k] = __dadd_rn(acc[k], v);
      }
    }
    if (g > 0) {
      double* o = s_part + ((size_t)(g - 1) * TR + r) * NS;
#pragma unroll
      for (int k = 0; k < NS; ++k) o[k] = acc[k];
    }
    __syncthreads();
    if (g == 0 && live) {
      double sc[NS];
#pragma unroll
      for (int k = 0; k < NS; ++k) sc[k] = (k < md.n_scores ? p.bias[md.score_off + k] : 0.0) + acc[k];
      for (int gg = 1; gg < G; ++gg) {
        const double* o = s_part + ((size_t)(gg - 1) * TR + r) * NS;
#pragma unroll
        for (int k = 0; k < NS; ++k) sc[k] += o[k];
      }
      p.pred[row * p.n_models + m] = apply_link(md, sc, p.classes);
      if (m == 0) {
        int bad = 0;
        for (int j = 0; j < p.n_in; ++j) bad |= is_finite_f(xr[j]) ? 0 : 1;
        p.row_bad[row] = bad;
      }
    }
    ++stage;
    if (stage == S) stage = 0;
  }
  cp_async_wait<0>();
}

// VotingEnsemble reduce over pred[row][model] -> out (+ status)
__global__ void __launch_bounds__(256) vote_kernel(KParams kp, const double* __restrict__ pred_buf,
                                                   const int32_t* __restrict__ row_bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < kp.n_rows; row += stride) {
    double pred[kMaxModels];
    for (int m = 0; m < kp.n_models; ++m) pred[m] = pred_buf[row * kp.n_models + m];
    vote_and_store(kp, pred, row, row_bad[row] ? 1u : 0u);
  }
}

}  // namespace b2s
