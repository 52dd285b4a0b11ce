// b2s_hash.cuh -- the online table's open-addressing layout, shared by the gather kernel (b2s_table.cu) and the fused
// gather loader of the scoring kernel (b2s_rowthread.cuh).
#pragma once
#include <cstdint>

namespace b2s {

struct TableSlot {
  long long key;
  long long row;  // -1: empty
};

__host__ __device__ inline uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ULL;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebULL;
  x ^= x >> 31;
  return x;
}

#ifdef __CUDACC__
// linear probing; the table is at most half full, so an empty slot always ends the walk.  -1: unknown key
__device__ __forceinline__ long long table_find(const TableSlot* __restrict__ slots, uint64_t mask, long long key) {
  uint64_t h = mix64((uint64_t)key) & mask;
  for (;;) {
    const longlong2 s = __ldg(reinterpret_cast<const longlong2*>(slots) + h);
    if (s.y < 0) return -1;
    if (s.x == key) return s.y;
    h = (h + 1) & mask;
  }
}
#endif

}  // namespace b2s
