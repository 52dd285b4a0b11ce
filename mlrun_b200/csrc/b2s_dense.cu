// b2s_dense.cu -- dense linear head on the 5th-generation tensor cores (sm_100a: tcgen05.mma + TMEM).
//
// north_star: "tensor cores only on the dense linear-predict path".  A linear / logistic scorer (or an ensemble of them)
// with many scores per event -- a 16-class LogisticRegression is scores = X (B x K) . W^T (K x 16) + b, then argmax
// (sklearn decision_function + predict behind PickleModelServer.predict, frameworks/_ml_common/pkl_model_server.py:52-60)
// -- is a GEMM with a skinny N.  The fp64 FMA path of the row kernels does K x N DFMAs per event and turns FP64-pipe bound
// beyond ~8 scores; here the products run on the tensor cores and the kernel goes back to being HBM bound:
//
//   HBM rows --TMA boxes (32 floats x 128 rows, 128-byte swizzle)--> shared memory, 2 stages
//     split   every value x (after the Imputer) becomes xh = tf32(x) and xl = tf32(x - xh), written in place / to a second
//             tile in the same K-major SWIZZLE_128B layout (what the TMA produced is already the UMMA operand layout)
//     mma     ONE thread issues  D  = xh . wh ;  D += xh . wl ;  D += xl . wh   (tcgen05.mma.cta_group::1.kind::tf32, M = 128,
//             N = 16 | 32, K = 8 per instruction; weights split on the host the same way) -- the "3xTF32" scheme: every
//             partial product is exact in the fp32 accumulator's input, the dropped xl . wl term is < 2^-22 |x w|
//     TMEM    the 128 x N fp32 accumulator lives in 32 TMEM columns; tcgen05.commit -> mbarrier tells the CTA it is complete
//     epilogue  thread r reads row r with tcgen05.ld (32 lanes x 32 bit x N), adds the intercepts in fp64, applies each
//             model's link (argmax / > 0 / identity) and the VotingEnsemble reduce, stores votes + status (to every merge
//             target when sharded) -- the same epilogue functions as the other kernels.
// Scores are within ~1e-6 relative of the fp64 path (tests: rtol 1e-5); labels are exact unless two class scores tie to
// that precision.  Evidence to look for: SASS UTCHMMA / UTCQMMA-family + LDTM, ncu sm__pipe_tensor_cycles_active > 0.
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>

#include "b2s_rowthread.cuh"  // mbarrier / TMA helpers, KParams, epilogue functions
#include "b2s_dense.cuh"

namespace b2s {

constexpr int kDM = kDenseTileRows;  // rows per tile == UMMA M == threads per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused: one swizzle atom along K) [16,30), stride byte offset = 8 rows x 128 B >> 4 in [32,46),
// version 1 in [46,48), layout type SWIZZLE_128B (2) in [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A and B tf32 (2 << 7, 2 << 10), both K-major,
// N >> 3 in [17,23), M >> 4 in [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kDM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }

template <int NP>  // padded score count: 16 | 32
__global__ void __launch_bounds__(kDM) dense_head_kernel(const __grid_constant__ DenseParams p, const __grid_constant__ KParams kp,
                                                        const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(1024) unsigned char smem_dense[];
  unsigned char* const smem = smem_dense;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = p.n_in, boxes = K >> 5;
  const uint32_t tile_bytes = (uint32_t)boxes * kDM * 128u;  // one A tile: boxes x (128 rows x 128 B)
  // shared memory: A stage 0 | A stage 1 | A residual (xl) | B hi | B lo | bias, fill, flags, barriers, TMEM address
  unsigned char* s_a[2] = {smem, smem + tile_bytes};
  unsigned char* s_al = smem + 2 * tile_bytes;
  unsigned char* s_bh = smem + 3 * tile_bytes;
  const uint32_t b_bytes = (uint32_t)boxes * NP * 128u;
  unsigned char* s_bl = s_bh + b_bytes;
  unsigned char* s_misc = s_bl + b_bytes;
  float* s_fill = reinterpret_cast<float*>(s_misc);               // [K]
  int* s_bad = reinterpret_cast<int*>(s_misc + 4 * 128);          // [128]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(s_misc + 1024);  // [2]
  uint64_t* s_mma = s_full + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_full + 4);

  // ---- one-time setup: TMEM columns, barriers, the weights in the UMMA layout (rows = scores, K-major, 128-byte swizzle)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(s_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < NP * K; i += kDM) {
    const int n = i / K, k = i - n * K;
    const int b = k >> 5, c = (k & 31) >> 2, e = k & 3;
    const uint32_t off = (uint32_t)b * NP * 128u + (uint32_t)n * 128u + (uint32_t)((c ^ (n & 7)) << 4) + (uint32_t)e * 4u;
    *reinterpret_cast<float*>(s_bh + off) = p.wh[i];
    *reinterpret_cast<float*>(s_bl + off) = p.wl[i];
  }
  for (int i = tid; i < K; i += kDM) s_fill[i] = p.fill[i];
  s_bad[tid] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores above -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;
  const uint32_t idesc = umma_idesc(NP);

  const int64_t n_tiles = (p.n_rows + kDM - 1) / kDM;
  auto issue = [&](int64_t t, int stage) {  // one thread: the tile's boxes; rows past the end arrive as zeros
    mbar_expect_tx(&s_full[stage], tile_bytes);
    for (int b = 0; b < boxes; ++b)
      tensor_load_2d(s_a[stage] + (size_t)b * kDM * 128, &tmap, b * 32, (int)(t * kDM), &s_full[stage]);
  };
  if (tid == 0 && (int64_t)blockIdx.x < n_tiles) issue(blockIdx.x, 0);

  int it = 0;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int stage = it & 1;
    const int64_t row0 = t * kDM;
    mbar_wait(&s_full[stage], (uint32_t)(it >> 1) & 1u);
    // the other stage was this CTA's previous tile: its MMAs completed before that tile's epilogue ran, so it is free
    if (tid == 0 && t + gridDim.x < n_tiles) issue(t + gridDim.x, stage ^ 1);
    {  // ---- split: x -> (xh, xl), Imputer and the non-finite test on the way.  Lanes take consecutive rows: conflict free.
      const int64_t left = p.n_rows - row0;
      const int rows = left < kDM ? (int)left : kDM;
      const int n_chunks = (K >> 2) * kDM;
      for (int i0 = tid; i0 < n_chunks; i0 += kDM * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kDM;
          const int c = i / kDM, rr = i - c * kDM;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < n_chunks)
            v[u] = *reinterpret_cast<const float4*>(s_a[stage] + (size_t)(c >> 3) * (kDM * 128) + rr * 128 + (((c & 7) ^ (rr & 7)) << 4));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kDM;
          if (i >= n_chunks) break;
          const int c = i / kDM, rr = i - c * kDM;
          float xs[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          uint32_t hi[4], lo[4];
          bool bad = false;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float x = xs[e];
            if (p.any_fill) {
              const float f = s_fill[c * 4 + e];
              x = (x != x) ? f : x;  // Imputer._impute (feature_store/steps.py:397-406); NaN where nothing is imputed
            }
            bad |= !is_finite_f(x) && rr < rows;
            x = is_finite_f(x) ? x : 0.0f;  // a flagged row must not poison the accumulator of its neighbours' columns
            hi[e] = tf32_hi(x);
            lo[e] = tf32_hi(x - __uint_as_float(hi[e]));  // exact difference, then its own 11 significant bits
          }
          const size_t off = (size_t)(c >> 3) * (kDM * 128) + rr * 128 + (((c & 7) ^ (rr & 7)) << 4);
          *reinterpret_cast<uint4*>(s_a[stage] + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(s_al + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          if (bad) atomicOr(&s_bad[rr], 1);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {  // ---- one thread feeds the tensor core: 3 x K / 8 instructions, then the completion arrives on s_mma
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(s_a[stage]), a_lo = smem_u32(s_al), b_hi = smem_u32(s_bh), b_lo = smem_u32(s_bl);
      bool acc = false;
      for (int b = 0; b < boxes; ++b)
        for (int kk = 0; kk < 4; ++kk) {  // 8 tf32 = 32 bytes per instruction inside the 128-byte swizzle atom
          const uint32_t ao = (uint32_t)b * kDM * 128u + (uint32_t)kk * 32u, bo = (uint32_t)b * NP * 128u + (uint32_t)kk * 32u;
          umma_tf32(tmem, umma_desc(a_hi + ao), umma_desc(b_hi + bo), idesc, acc);
          umma_tf32(tmem, umma_desc(a_hi + ao), umma_desc(b_lo + bo), idesc, true);
          umma_tf32(tmem, umma_desc(a_lo + ao), umma_desc(b_hi + bo), idesc, true);
          acc = true;
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(s_mma)) : "memory");
    }
    mbar_wait(s_mma, (uint32_t)it & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {  // ---- epilogue: thread r owns row r (TMEM lane r: warp w reads lanes 32 w .. 32 w + 31)
      uint32_t r[NP];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < NP; c0 += 16)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[c0 + 0]), "=r"(r[c0 + 1]), "=r"(r[c0 + 2]), "=r"(r[c0 + 3]), "=r"(r[c0 + 4]), "=r"(r[c0 + 5]), "=r"(r[c0 + 6]),
              "=r"(r[c0 + 7]), "=r"(r[c0 + 8]), "=r"(r[c0 + 9]), "=r"(r[c0 + 10]), "=r"(r[c0 + 11]), "=r"(r[c0 + 12]),
              "=r"(r[c0 + 13]), "=r"(r[c0 + 14]), "=r"(r[c0 + 15])
            : "r"(taddr + (uint32_t)c0));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int64_t row = row0 + tid;
      const uint32_t st = s_bad[tid] ? 1u : 0u;
      s_bad[tid] = 0;
      if (row < p.n_rows) {
        if (p.epi != DENSE_EPI_GENERIC && kp.n_peers == 0) {
          // ---- fast epilogues: float32 registers only (compile-time indices; the intercepts are constant-bank operands)
          float sc[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) sc[k] = __uint_as_float(r[k]) + p.biasf[k];
          if (p.epi == DENSE_EPI_SCORES) {  // out_cols == n_scores consecutive floats per row
            float* o = kp.out + row * kp.out_cols;
            if ((kp.out_cols & 3) == 0) {
#pragma unroll
              for (int k = 0; k < NP; k += 4)
                if (k < p.n_scores) *reinterpret_cast<float4*>(o + k) = make_float4(sc[k], sc[k + 1], sc[k + 2], sc[k + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < NP; ++k)
                if (k < p.n_scores) o[k] = sc[k];
            }
          } else if (p.epi == DENSE_EPI_MEAN) {  // VotingEnsemble._mean_vote: sum_m w[m] * pred[m], model order
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < NP; ++k) v = fmaf(p.votewf[k], sc[k], v);  // the padding has zero weight
            kp.out[row] = v;
          } else {  // one multi-class linear classifier: np.argmax (first maximum), then classes_[index]
            int best = 0;
            float bv = sc[0];
#pragma unroll
            for (int k = 1; k < NP; ++k)
              if (k < p.n_scores && sc[k] > bv) {
                bv = sc[k];
                best = k;
              }
            int lab = p.labels[0];
#pragma unroll
            for (int k = 1; k < NP; ++k) lab = best == k ? p.labels[k] : lab;
            reinterpret_cast<int32_t*>(kp.out)[row] = lab;
          }
          if (kp.status) kp.status[row] = (int32_t)st;
        } else {
          double sc[NP];
#pragma unroll
          for (int k = 0; k < NP; ++k) sc[k] = k < p.n_scores ? (double)__uint_as_float(r[k]) + p.bias[k] : 0.0;
          double pred[kMaxModels];
          for (int m = 0; m < kp.n_models; ++m) {
            const ModelDesc md = kp.models[m];
            pred[m] = apply_link(md, sc + md.score_off, kp.classes);
          }
          vote_and_store(kp, pred, row, st);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // the accumulator, the residual tile and the flags are free for the next tile
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
  merge_signal(kp.sig);
}

cudaError_t dense_launch(const DenseParams& p, const KParams& kp, const CUtensorMap& tmap, int grid, int smem, int smem_optin,
                         cudaStream_t st) {
  static std::atomic<bool> attr{false};
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dense_head_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dense_head_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  if (p.n_pad == 16) dense_head_kernel<16><<<grid, kDM, smem, st>>>(p, kp, tmap);
  else dense_head_kernel<32><<<grid, kDM, smem, st>>>(p, kp, tmap);
  return cudaGetLastError();
}

int dense_smem_bytes(int n_in, int n_pad) {
  const int boxes = n_in / 32;
  return 3 * boxes * kDM * 128 + 2 * boxes * n_pad * 128 + 2048;
}

}  // namespace b2s
