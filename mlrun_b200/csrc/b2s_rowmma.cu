// b2s_rowmma.cu -- instantiations and launcher of rowmma_kernel (b2s_rowmma.cuh); kept out of b2s_runtime.cu for build time.
#include <cuda_runtime.h>

#include "b2s_rowmma.cuh"

namespace b2s {

size_t rowmma_smem_bytes(int nch, int ns, int n_cat, int warps, int stages) {
  const int nsp = ns < 2 ? 2 : ns;
  size_t b = (size_t)kRMMaxWarps * kRMMaxStages * 16 + 16;
  b += (((size_t)(n_cat + 1) * ns * 8 + 15) / 16) * 16;
  b += (size_t)warps * 32 * nsp * 8;
  b += 1024;  // alignment slack of the swizzled tiles
  b += (size_t)warps * stages * nch * 512;
  return b;
}

template <int NCH, int NS>
static cudaError_t launch_t(const void* params, const CUtensorMap* tmap, int grid, int warps, size_t smem, cudaStream_t st) {
  rowmma_kernel<NCH, NS><<<grid, warps * 32, smem, st>>>(*reinterpret_cast<const RTParams<NCH, NS>*>(params), *tmap);
  return cudaGetLastError();
}
template <int NCH, int NS>
static cudaError_t prepare_t(int max_smem) {
  return cudaFuncSetAttribute(rowmma_kernel<NCH, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
}

#define RM_DISPATCH(FN, ...)                                      \
  do {                                                            \
    if (nch == 8 && ns == 1) return FN<8, 1>(__VA_ARGS__);        \
    if (nch == 8 && ns == 2) return FN<8, 2>(__VA_ARGS__);        \
    if (nch == 8 && ns == 4) return FN<8, 4>(__VA_ARGS__);        \
    if (nch == 8 && ns == 8) return FN<8, 8>(__VA_ARGS__);        \
    if (nch == 16 && ns == 1) return FN<16, 1>(__VA_ARGS__);      \
    if (nch == 16 && ns == 2) return FN<16, 2>(__VA_ARGS__);      \
    if (nch == 16 && ns == 4) return FN<16, 4>(__VA_ARGS__);      \
    if (nch == 16 && ns == 8) return FN<16, 8>(__VA_ARGS__);      \
  } while (0)

cudaError_t rowmma_launch(int nch, int ns, const void* params, const CUtensorMap* tmap, int grid, int warps, size_t smem, cudaStream_t st) {
  RM_DISPATCH(launch_t, params, tmap, grid, warps, smem, st);
  return cudaErrorInvalidValue;
}
cudaError_t rowmma_prepare(int nch, int ns, int max_smem) {
  RM_DISPATCH(prepare_t, max_smem);
  return cudaErrorInvalidValue;
}

}  // namespace b2s
