// common.cuh — shared helpers for the sparseconv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "../../include/sparseconv_b200.h"

namespace spc {

// ---- error plumbing -------------------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<long long> g_launch_count;

inline int fail(const char* what, const char* detail) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, detail);
  return -1;
}
inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail(what, cudaGetErrorString(e));
}
#define SPC_CUDA(expr)                                    \
  do {                                                    \
    int _rc = ::spc::check_cuda((expr), #expr);           \
    if (_rc) return _rc;                                  \
  } while (0)
// call right after a kernel launch
#define SPC_LAUNCHED(name)                                        \
  do {                                                            \
    ::spc::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    int _rc = ::spc::check_cuda(cudaGetLastError(), name);        \
    if (_rc) return _rc;                                          \
  } while (0)
#define SPC_REQUIRE(cond, msg)                 \
  do {                                         \
    if (!(cond)) return ::spc::fail(__func__, msg); \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- coordinate keys --------------------------------------------------------
// 64-bit key = batch:10 | x:18 | y:18 | z:18 (offset binary).  All-ones = empty.
constexpr unsigned long long kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr int kCoordBias = 131072;         // 2^17
constexpr unsigned kCoordRange = 262144u;  // 2^18
constexpr unsigned kBatchMax = 1023u;      // batch index must be < 1023

struct __align__(16) Slot {
  unsigned long long key;
  unsigned int first;  // smallest source row that produced this key
  unsigned int row;    // row in the coordinate map
};
static_assert(sizeof(Slot) == SPC_SLOT_BYTES, "slot size");

__device__ __forceinline__ bool pack_key(int b, int x, int y, int z, unsigned long long& key) {
  unsigned ux = (unsigned)(x + kCoordBias), uy = (unsigned)(y + kCoordBias),
           uz = (unsigned)(z + kCoordBias);
  bool ok = ((unsigned)b < kBatchMax) & (ux < kCoordRange) & (uy < kCoordRange) & (uz < kCoordRange);
  key = ((unsigned long long)(unsigned)b << 54) | ((unsigned long long)ux << 36) |
        ((unsigned long long)uy << 18) | (unsigned long long)uz;
  return ok;
}
__device__ __forceinline__ int4 unpack_key(unsigned long long key) {
  int4 c;
  c.x = (int)(key >> 54);
  c.y = (int)((key >> 36) & 0x3FFFFu) - kCoordBias;
  c.z = (int)((key >> 18) & 0x3FFFFu) - kCoordBias;
  c.w = (int)(key & 0x3FFFFu) - kCoordBias;
  return c;
}
// splitmix64 finaliser; buckets are pairs of slots (one 32-byte sector).
// Tables larger than the L2 (>= 2^22 buckets = 128 MB: maps of more than ~2 M voxels) leave the lowest bit of z out of
// the hash, so that the voxels (.., z even) and (.., z + 1) share a home bucket: the reference's loaders deliver
// voxels in raster order (z fastest) and neighbouring lanes' probes then read the same DRAM sector half of the time
// (3^3 map at 10 M voxels 2.36 -> 1.91 ms).  A pair fills its bucket, so look-ups take 1.44 instead of 1.17 probes
// (simulated on the 1 M-voxel room) — on an L2-resident table that costs more than the shared sectors save (1 M
// voxels: 0.136 -> 0.171 ms), hence the size rule.  Measured: profiles/r2_row_order_experiments.md.
constexpr unsigned long long kPairHashBuckets = 1ull << 22;
__device__ __forceinline__ unsigned long long hash_key(unsigned long long k, unsigned long long bucket_mask) {
  if (bucket_mask >= kPairHashBuckets - 1) k >>= 1;
  k ^= k >> 30;
  k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 27;
  k *= 0x94d049bb133111ebull;
  k ^= k >> 31;
  return k & bucket_mask;
}

// One 256-bit read-only load (LDG.E.256, sm_100+): a whole 32-byte bucket = both slots in ONE instruction and
// one L1 tag look-up.  The probes are fully divergent, so the kernel-map builder is bound by L1 tag throughput
// (one sector per clock and SM): two LDG.128 per probe cost twice as much.
__device__ __forceinline__ void ldg256(const void* p, unsigned long long (&v)[4]) {
  asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
               : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3])
               : "l"(p));
}

// Read-only probe of a 2-slot bucket table.  Returns map row or -1.
__device__ __forceinline__ int table_lookup(const Slot* __restrict__ slots,
                                            unsigned long long bucket_mask,
                                            unsigned long long key) {
  unsigned long long b = hash_key(key, bucket_mask);
  for (;;) {
    unsigned long long v[4];  // {key0, first0 | row0 << 32, key1, first1 | row1 << 32}
    ldg256(slots + 2 * b, v);
    if (v[0] == key) return (int)(v[1] >> 32);
    if (v[2] == key) return (int)(v[3] >> 32);
    // slots of a bucket fill in order and nothing is ever deleted
    if (v[0] == kEmptyKey || v[2] == kEmptyKey) return -1;
    b = (b + 1) & bucket_mask;
  }
}

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

}  // namespace spc
