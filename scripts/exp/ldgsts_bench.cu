// Microbenchmark: LSU cost of cp.async (LDGSTS.128) row gathers as a function of how many distinct rows one
// instruction touches (8 rows x 64 B, 4 rows x 128 B, 2.67 rows x 192 B, 2 x 256 B, 1 x 512 B).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_bench ldgsts_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t n) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// PPR = 16-byte pieces per row visit (4: 64 B ... 32: 512 B).  Every warp issues, per iteration, 16 LDGSTS = 8 KB
// into one of its 3 shared-memory slots; wait_group keeps 2 iterations in flight.
template <int PPR>
__global__ void __launch_bounds__(256, 1)
k(const char* __restrict__ table, int row_bytes, const int* __restrict__ idx, int n_idx, int iters, int conflict_free) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int RPI_NUM = 32;  // rows per instruction = 32 / PPR (may be fractional: flattened piece order)
  long long pos = ((long long)blockIdx.x * 8 + warp) * 4096;
  for (int it = 0; it < iters; ++it) {
    const uint32_t slot = base + (warp * 3 + it % 3) * 8192;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int f = q * 32 + lane;
      const int r = f / PPR, piece = f % PPR;
      const int row = idx[(pos + r) % n_idx];
      const char* src = table + (size_t)(row >= 0 ? row : 0) * row_bytes + piece * 16;
      // destination: linear (bank-conflict free per 8 lanes) or the 64-B-row swizzled block layout
      uint32_t dst = slot + f * 16;
      if (!conflict_free) dst = slot + ((piece >> 2) * 2048 + r * 64 + (((piece & 3) ^ ((r >> 1) & 3)) << 4)) % 8192;
      cp_async_16(dst, src, row >= 0 ? 16u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    pos += 512 / PPR + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  (void)RPI_NUM;
}

template <int PPR>
void run(const char* table, int row_bytes, const int* idx, int n_idx, const char* name, int cf) {
  const size_t smem = 8 * 3 * 8192 + 1024;
  CK(cudaFuncSetAttribute(k<PPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<PPR><<<148, 256, smem>>>(table, row_bytes, idx, n_idx, 200, cf);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k<PPR><<<148, 256, smem>>>(table, row_bytes, idx, n_idx, iters, cf);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double instr = 8.0 * 16 * iters;  // per SM
  printf("%-28s pieces/row=%2d dst=%s: %.3f ms  %.2f cycles/LDGSTS/SM @1.9GHz  %.1f B/clk/SM\n", name, PPR,
         cf ? "linear  " : "swizzled", ms, ms * 1e-3 * 1.9e9 / instr, instr * 512 / (ms * 1e-3 * 1.9e9));
}

int main() {
  const int M = 200000, row_bytes = 512;  // 512-B rows so that every visit width up to 512 B is contiguous
  char* table;
  CK(cudaMalloc(&table, (size_t)M * row_bytes));
  CK(cudaMemset(table, 1, (size_t)M * row_bytes));
  const int n_idx = 1 << 22;
  std::vector<int> h(n_idx);
  for (int pattern = 0; pattern < 2; ++pattern) {
    uint64_t st = 777;
    int cur = 0;
    for (int i = 0; i < n_idx; ++i) {
      st = st * 6364136223846793005ull + 1442695040888963407ull;
      unsigned u = (unsigned)(st >> 33);
      if (u % 100 < 26) cur = (int)(u % M); else cur = (cur + 1) % M;   // 74 % consecutive rows, like the kernel maps
      h[i] = (pattern == 1 && (u >> 8) % 100 < 36) ? -1 : cur;            // pattern 1: 36 % missing (zero fill)
    }
    int* idx;
    CK(cudaMalloc(&idx, n_idx * 4));
    CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
    const char* name = pattern ? "neighbour-like, 36% missing" : "neighbour-like, all present";
    for (int cf = 1; cf >= 0; --cf) {
      run<4>(table, row_bytes, idx, n_idx, name, cf);
      run<8>(table, row_bytes, idx, n_idx, name, cf);
      run<12>(table, row_bytes, idx, n_idx, name, cf);
      run<16>(table, row_bytes, idx, n_idx, name, cf);
      run<32>(table, row_bytes, idx, n_idx, name, cf);
    }
    cudaFree(idx);
  }
  return 0;
}
