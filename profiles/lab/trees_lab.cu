// trees_lab.cu -- round-2 microbenchmark for the tree-ensemble walk (not product code; the winner is
// integrated as csrc/b2s_trees3.cuh).  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
//
// Workload = BASELINE configs[2]: 4 models x 100 complete depth-6 trees, 128 float32 features, 256 Ki rows.
// Every CTA owns one "part" (here: one model) resident in shared memory; a warp is 32 consecutive rows (x RPT row
// blocks) walking the same tree, tile transposed in shared memory (xt[feature][row]) so the x gather is conflict free.
//
// Variants (VAR):
//   -1  loader + transpose only (no walk): the fixed cost every variant pays
//    0  split node arrays: 4-byte feature offset + 4-byte threshold (what round 1 shipped): 3 LDS per visit
//    1  8-byte heap nodes {foff, thr}: LDS.64 + LDS per visit
//    2  VAR 1 + the top two levels from warp-uniform LDS (3 nodes loaded once per tree for RPT rows)
//    3  VAR 1 + the top two levels as constant-bank operands (__grid_constant__ table)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

constexpr int kMaxTopTrees = 400;
struct TopNode { int32_t foff; float thr; };
struct TopTable { TopNode n[kMaxTopTrees][3]; };  // nodes 1, 2, 3 of every tree (9.6 KB)

struct LabParams {
  const float* X;
  int64_t n_rows;
  int n_in;
  const uint2* nodes;    // [n_parts][NT][2^D]  (heap, slot 0 unused)
  const int32_t* foff;   // [n_parts][NT][2^D]
  const float* thr;      // [n_parts][NT][2^D]
  const double* leaves;  // [n_parts][NT][2^D]
  double* pred;          // [n_rows][n_parts]
  int n_parts, NT, W;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// explicit shared-window loads: volatile keeps them in program order (and on their side of the barriers), so the
// walk is written level by level in issue order: all node loads, all x loads, all index updates
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
template <int OFF>
__device__ __forceinline__ float ldsf(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ double ldsd(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
template <int J, int RPT>
struct XLoad {  // x of row block J for every tree in flight (compile-time immediate offset J * 128)
  template <int U>
  static __device__ __forceinline__ void run(uint32_t xl, const uint2 (&nd)[U][RPT], float (&x)[U][RPT]) {
#pragma unroll
    for (int u = 0; u < U; ++u) x[u][J] = ldsf<J * 128>(xl + nd[u][J].x);
    XLoad<J + 1, RPT>::run(xl, nd, x);
  }
};
template <int RPT>
struct XLoad<RPT, RPT> {
  template <int U>
  static __device__ __forceinline__ void run(uint32_t, const uint2 (&)[U][RPT], float (&)[U][RPT]) {}
};

template <int D, int RPT, int U, int VAR>
__global__ void __launch_bounds__(1024) trees_lab_kernel(const __grid_constant__ LabParams p, const __grid_constant__ TopTable top) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int NN = 1 << D;        // node slots per tree (heap, 1-based) == leaves per tree
  constexpr int TR = 32 * RPT;      // rows per tile
  constexpr int LP = 36;            // landing pitch (floats): 32-feature slab + 4 pad
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int W = p.W, NT = p.NT;
  const int part = blockIdx.x % p.n_parts;
  const int cta = blockIdx.x / p.n_parts;
  const int nctas = (gridDim.x - part + p.n_parts - 1) / p.n_parts;

  // ---- shared memory: node table | leaves | partial sums | transposed tile | landing slab
  unsigned char* s_nodes = smem;                                     // NT*NN*8
  double* s_leaf = reinterpret_cast<double*>(smem + (size_t)NT * NN * 8);  // NT*NN*8
  double* s_part = s_leaf + (size_t)NT * NN;                         // W*TR
  float* s_xt = reinterpret_cast<float*>(s_part + (size_t)W * TR);   // n_in*TR
  float* s_land = s_xt + (size_t)p.n_in * TR;                        // TR*LP

  if (VAR == 0) {
    int32_t* sf = reinterpret_cast<int32_t*>(s_nodes);
    float* st = reinterpret_cast<float*>(s_nodes + (size_t)NT * NN * 4);
    for (int i = tid; i < NT * NN; i += blockDim.x) {
      sf[i] = p.foff[(size_t)part * NT * NN + i];
      st[i] = p.thr[(size_t)part * NT * NN + i];
    }
  } else {
    uint2* sn = reinterpret_cast<uint2*>(s_nodes);
    for (int i = tid; i < NT * NN; i += blockDim.x) sn[i] = p.nodes[(size_t)part * NT * NN + i];
  }
  for (int i = tid; i < NT * NN; i += blockDim.x) s_leaf[i] = p.leaves[(size_t)part * NT * NN + i];
  __syncthreads();

  const int64_t n_tiles = (p.n_rows + TR - 1) / TR;
  const int TPW = (NT + W - 1) / W;  // trees per warp
  const char* xl = reinterpret_cast<const char*>(s_xt + lane);
  const int n_slabs = p.n_in / 32;

  for (int64_t t = cta; t < n_tiles; t += nctas) {
    const int64_t row0 = t * TR;
    // ---- loader: 32-feature slabs -> landing -> transposed tile
    for (int s = 0; s < n_slabs; ++s) {
      for (int i = tid; i < TR * 8; i += blockDim.x) {
        const int r = i >> 3, c = i & 7;
        const int64_t row = row0 + r;
        if (row < p.n_rows) cp_async16(s_land + r * LP + c * 4, p.X + row * p.n_in + s * 32 + c * 4);
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
      for (int i = tid; i < TR * 8; i += blockDim.x) {
        const int c = i / TR, r = i - c * TR;  // lanes = consecutive rows
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < p.n_rows) v = *reinterpret_cast<const float4*>(s_land + r * LP + c * 4);
        float* o = s_xt + (size_t)(s * 32 + c * 4) * TR + r;
        o[0] = v.x;
        o[TR] = v.y;
        o[2 * TR] = v.z;
        o[3 * TR] = v.w;
      }
      __syncthreads();
    }
    // ---- walk: warp g takes trees g, g+W, ...; lane = row (+32 j)
    double acc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] = 0.0;
    if (VAR >= 0) {
      for (int i = 0; i < TPW; i += U) {
        const unsigned char* tb[U];
        const double* lb[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int tr = g + (i + u) * W;
          valid[u] = (i + u) < TPW && tr < NT;
          const int tt = valid[u] ? tr : 0;
          tb[u] = s_nodes + (size_t)tt * NN * 8;
          lb[u] = s_leaf + (size_t)tt * NN - NN;  // leaf index = node - NN
        }
        int n8[U][RPT];
        if (VAR == 0) {
          const int32_t* sf = reinterpret_cast<const int32_t*>(s_nodes);
          const float* st = reinterpret_cast<const float*>(s_nodes + (size_t)NT * NN * 4);
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < RPT; ++j) n8[u][j] = 4;
#pragma unroll
          for (int d = 0; d < D; ++d) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int tr = valid[u] ? g + (i + u) * W : 0;
#pragma unroll
              for (int j = 0; j < RPT; ++j) {
                const int at = tr * NN * 4 + n8[u][j];
                const int fo = *reinterpret_cast<const int32_t*>(reinterpret_cast<const char*>(sf) + at);
                const float th = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(st) + at);
                const float x = *reinterpret_cast<const float*>(xl + j * 128 + fo);
                n8[u][j] = 2 * n8[u][j] + ((x <= th) ? 0 : 4);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < RPT; ++j) n8[u][j] *= 2;  // byte offset of an 8-byte slot
        } else {
          // absolute shared-window addresses: a = tree_base + 8 * node, so one visit is
          //   LDS.64 node <- [a] ; IADD xa = foff + lane_base ; LDS x <- [xa] ; FSETP ; SEL ; IADD3 a = a + a + sel
          // with sel = (go right ? 8 : 0) - tree_base
          constexpr int D0 = (VAR >= 2) ? 2 : 0;  // levels handled with warp-uniform node data
          const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
          const uint32_t xls = (uint32_t)__cvta_generic_to_shared(xl);
          uint32_t tba[U], cl[U], cr[U], a[U][RPT];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            tba[u] = sbase + (uint32_t)(tb[u] - smem);
            cl[u] = 0u - tba[u];
            cr[u] = 8u - tba[u];
          }
          uint2 nd[U][RPT];
          float x[U][RPT];
          if (VAR >= 2) {
            uint2 n1[U], n2[U], n3[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (VAR == 2) {
                n1[u] = lds64(tba[u] + 8);
                const uint4 q = lds128(tba[u] + 16);
                n2[u] = make_uint2(q.x, q.y);
                n3[u] = make_uint2(q.z, q.w);
              } else {
                const int tr = valid[u] ? g + (i + u) * W : 0;
                const TopNode* tn = top.n[part * NT + tr];
                n1[u] = make_uint2((uint32_t)tn[0].foff, __float_as_uint(tn[0].thr));
                n2[u] = make_uint2((uint32_t)tn[1].foff, __float_as_uint(tn[1].thr));
                n3[u] = make_uint2((uint32_t)tn[2].foff, __float_as_uint(tn[2].thr));
              }
#pragma unroll
              for (int j = 0; j < RPT; ++j) nd[u][j] = n1[u];
            }
            XLoad<0, RPT>::run(xls, nd, x);
            bool r0[U][RPT];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int j = 0; j < RPT; ++j) {
                r0[u][j] = !(x[u][j] <= __uint_as_float(n1[u].y));
                nd[u][j].x = r0[u][j] ? n3[u].x : n2[u].x;
                nd[u][j].y = r0[u][j] ? n3[u].y : n2[u].y;
              }
            XLoad<0, RPT>::run(xls, nd, x);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int j = 0; j < RPT; ++j) {
                const bool r1 = !(x[u][j] <= __uint_as_float(nd[u][j].y));
                a[u][j] = tba[u] + 32u + (r0[u][j] ? 16u : 0u) + (r1 ? 8u : 0u);
              }
          } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int j = 0; j < RPT; ++j) a[u][j] = tba[u] + 8u;
          }
#pragma unroll
          for (int d = D0; d < D; ++d) {
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int j = 0; j < RPT; ++j) nd[u][j] = lds64(a[u][j]);
            XLoad<0, RPT>::run(xls, nd, x);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int j = 0; j < RPT; ++j)
                a[u][j] = a[u][j] + a[u][j] + ((x[u][j] <= __uint_as_float(nd[u][j].y)) ? cl[u] : cr[u]);
          }
          // leaves: s_leaf + tree * NN * 8 + (a - tree_base - NN * 8)
          const uint32_t leaf0 = (uint32_t)__cvta_generic_to_shared(s_leaf) - sbase - (uint32_t)(NN * 8);
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              const double v = ldsd(a[u][j] + leaf0);
              if (valid[u]) acc[j] += v;
            }
          }
          continue;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (valid[u]) {
#pragma unroll
            for (int j = 0; j < RPT; ++j)
              acc[j] += *reinterpret_cast<const double*>(reinterpret_cast<const char*>(lb[u]) + n8[u][j]);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < RPT; ++j) acc[j] = (double)*reinterpret_cast<const float*>(xl + j * 128 + g * TR * 4);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) s_part[g * TR + j * 32 + lane] = acc[j];
    __syncthreads();
    if (tid < TR && row0 + tid < p.n_rows) {
      double s = 0.0;
      for (int gg = 0; gg < W; ++gg) s += s_part[gg * TR + tid];
      p.pred[(row0 + tid) * p.n_parts + part] = s;
    }
    // (the loader's first barrier of the next tile orders these reads before the next partial-sum writes)
  }
}

struct Host {
  int D = 6, NT = 100, n_parts = 4, n_in = 128;
  int64_t n_rows = 262144;
  std::vector<int32_t> feat;  // [parts][NT][NN]
  std::vector<float> thr;
  std::vector<double> leaves;
  std::vector<float> X;
};

static double cpu_ref(const Host& h, int part, const float* x) {
  const int NN = 1 << h.D;
  double s = 0.0;
  for (int t = 0; t < h.NT; ++t) {
    const size_t base = ((size_t)part * h.NT + t) * NN;
    int n = 1;
    for (int d = 0; d < h.D; ++d) n = 2 * n + ((x[h.feat[base + n]] <= h.thr[base + n]) ? 0 : 1);
    s += h.leaves[base + n - NN];
  }
  return s;
}

template <int RPT, int U, int VAR>
static void run(const char* name, const Host& h, LabParams p, const TopTable& top, int W, int iters, const float* dX2) {
  constexpr int D = 6;
  constexpr int TR = 32 * RPT;
  const int NN = 1 << D;
  p.W = W;
  const size_t smem = (size_t)h.NT * NN * 16 + (size_t)W * TR * 8 + (size_t)h.n_in * TR * 4 + (size_t)TR * 36 * 4;
  auto kern = trees_lab_kernel<D, RPT, U, VAR>;
  cudaFuncAttributes fa;
  CK(cudaFuncGetAttributes(&fa, kern));
  if (smem > 227 * 1024) {
    printf("%-34s W=%2d  smem %zu too large\n", name, W, smem);
    return;
  }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = (sms / h.n_parts) * h.n_parts;
  CK(cudaMemset(p.pred, 0, (size_t)h.n_rows * h.n_parts * 8));
  for (int i = 0; i < 3; ++i) kern<<<grid, W * 32, smem>>>(p, top);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  // check the first rows against the CPU walk
  const int n_chk = 2048;
  std::vector<double> got((size_t)n_chk * h.n_parts);
  CK(cudaMemcpy(got.data(), p.pred, got.size() * 8, cudaMemcpyDeviceToHost));
  double max_err = 0.0;
  if (VAR >= 0)
    for (int r = 0; r < n_chk; ++r)
      for (int m = 0; m < h.n_parts; ++m)
        max_err = std::max(max_err, fabs(got[(size_t)r * h.n_parts + m] - cpu_ref(h, m, h.X.data() + (size_t)r * h.n_in)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const float* bufs[2] = {p.X, dX2};
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) {
    p.X = bufs[i & 1];
    kern<<<grid, W * 32, smem>>>(p, top);
  }
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("%-34s W=%2d regs=%3d smem=%6zu  %.4f ms  %.3f G events/s  max|err|=%.2e\n", name, W, fa.numRegs, smem, ms,
         h.n_rows / ms / 1e6, max_err);
  fflush(stdout);
}

int main(int argc, char** argv) {
  Host h;
  const int iters = argc > 1 ? atoi(argv[1]) : 20;
  if (argc > 2) h.n_rows = atoll(argv[2]);
  const int NN = 1 << h.D;
  std::mt19937_64 rng(1234);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<double> ud(-1.0, 1.0);
  h.feat.resize((size_t)h.n_parts * h.NT * NN);
  h.thr.resize(h.feat.size());
  h.leaves.resize(h.feat.size());
  for (size_t i = 0; i < h.feat.size(); ++i) {
    // like a fitted GBT on y = f(x0..x3) + noise: half the splits on the informative features
    h.feat[i] = (rng() & 1) ? (int)(rng() % 4) : (int)(rng() % h.n_in);
    h.thr[i] = nd(rng);
    h.leaves[i] = ud(rng);
  }
  h.X.resize((size_t)h.n_rows * h.n_in);
  for (auto& v : h.X) v = nd(rng);

  float *dX, *dX2;
  CK(cudaMalloc(&dX, h.X.size() * 4));
  CK(cudaMalloc(&dX2, h.X.size() * 4));
  CK(cudaMemcpy(dX, h.X.data(), h.X.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX2, dX, h.X.size() * 4, cudaMemcpyDeviceToDevice));
  LabParams p{};
  p.X = dX;
  p.n_rows = h.n_rows;
  p.n_in = h.n_in;
  p.n_parts = h.n_parts;
  p.NT = h.NT;
  CK(cudaMalloc((void**)&p.pred, (size_t)h.n_rows * h.n_parts * 8));
  CK(cudaMalloc((void**)&p.leaves, h.leaves.size() * 8));
  CK(cudaMemcpy((void*)p.leaves, h.leaves.data(), h.leaves.size() * 8, cudaMemcpyHostToDevice));

  auto upload_nodes = [&](int TR) {  // feature offsets depend on the tile pitch
    std::vector<uint2> nodes(h.feat.size());
    std::vector<int32_t> foff(h.feat.size());
    for (size_t i = 0; i < h.feat.size(); ++i) {
      foff[i] = h.feat[i] * TR * 4;
      uint32_t tb;
      memcpy(&tb, &h.thr[i], 4);
      nodes[i] = make_uint2((uint32_t)foff[i], tb);
    }
    if (p.nodes) {
      cudaFree((void*)p.nodes);
      cudaFree((void*)p.foff);
      cudaFree((void*)p.thr);
    }
    CK(cudaMalloc((void**)&p.nodes, nodes.size() * 8));
    CK(cudaMalloc((void**)&p.foff, foff.size() * 4));
    CK(cudaMalloc((void**)&p.thr, h.thr.size() * 4));
    CK(cudaMemcpy((void*)p.nodes, nodes.data(), nodes.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void*)p.foff, foff.data(), foff.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy((void*)p.thr, h.thr.data(), h.thr.size() * 4, cudaMemcpyHostToDevice));
    TopTable top;
    memset(&top, 0, sizeof(top));
    for (int m = 0; m < h.n_parts; ++m)
      for (int t = 0; t < h.NT; ++t)
        for (int k = 0; k < 3; ++k) {
          const size_t i = ((size_t)m * h.NT + t) * NN + 1 + k;
          top.n[m * h.NT + t][k] = TopNode{foff[i], h.thr[i]};
        }
    return top;
  };

  printf("rows=%lld features=%d parts=%d trees=%d depth=%d iters=%d\n", (long long)h.n_rows, h.n_in, h.n_parts, h.NT, h.D, iters);
  {
    TopTable top = upload_nodes(32);
    run<1, 4, -1>("loader only        RPT=1", h, p, top, 25, iters, dX2);
    run<1, 4, 0>("split 4+4          RPT=1 U=4", h, p, top, 25, iters, dX2);
    run<1, 4, 1>("node8              RPT=1 U=4", h, p, top, 25, iters, dX2);
    run<1, 4, 2>("node8 top2-lds     RPT=1 U=4", h, p, top, 25, iters, dX2);
    run<1, 4, 3>("node8 top2-const   RPT=1 U=4", h, p, top, 25, iters, dX2);
  }
  {
    TopTable top = upload_nodes(64);
    run<2, 2, -1>("loader only        RPT=2", h, p, top, 25, iters, dX2);
    run<2, 2, 0>("split 4+4          RPT=2 U=2", h, p, top, 25, iters, dX2);
    run<2, 2, 1>("node8              RPT=2 U=2", h, p, top, 25, iters, dX2);
    run<2, 2, 2>("node8 top2-lds     RPT=2 U=2", h, p, top, 25, iters, dX2);
    run<2, 2, 3>("node8 top2-const   RPT=2 U=2", h, p, top, 25, iters, dX2);
    run<2, 4, 1>("node8              RPT=2 U=4", h, p, top, 25, iters, dX2);
    run<2, 4, 3>("node8 top2-const   RPT=2 U=4", h, p, top, 25, iters, dX2);
    run<2, 1, 1>("node8              RPT=2 U=1", h, p, top, 20, iters, dX2);
    run<2, 5, 1>("node8              RPT=2 U=5", h, p, top, 20, iters, dX2);
    run<2, 5, 3>("node8 top2-const   RPT=2 U=5", h, p, top, 20, iters, dX2);
    run<2, 2, 1>("node8              RPT=2 U=2", h, p, top, 16, iters, dX2);
    run<2, 2, 1>("node8              RPT=2 U=2", h, p, top, 32, iters, dX2);
    run<2, 2, 3>("node8 top2-const   RPT=2 U=2", h, p, top, 32, iters, dX2);
    run<2, 5, 3>("node8 top2-const   RPT=2 U=5", h, p, top, 10, iters, dX2);
  }
  {
    TopTable top = upload_nodes(128);
    run<4, 1, -1>("loader only        RPT=4", h, p, top, 20, iters, dX2);
    run<4, 1, 1>("node8              RPT=4 U=1", h, p, top, 20, iters, dX2);
    run<4, 1, 2>("node8 top2-lds     RPT=4 U=1", h, p, top, 20, iters, dX2);
    run<4, 1, 3>("node8 top2-const   RPT=4 U=1", h, p, top, 20, iters, dX2);
    run<4, 2, 3>("node8 top2-const   RPT=4 U=2", h, p, top, 25, iters, dX2);
    run<4, 5, 3>("node8 top2-const   RPT=4 U=5", h, p, top, 10, iters, dX2);
  }
  return 0;
}
