// Microbenchmark (r2, second version): THROUGHPUT of the cp.async (LDGSTS.128) row gather of the convolution kernels
// as a function of the width of one row visit (64 / 128 / 192 / 256 B) and of the shared-memory destination layout.
// Unlike ldgsts_bench.cu nothing here is latency bound: row numbers are arithmetic (no index loads), 8 producer
// warps per SM each keep DEPTH commit groups of 8 KB in flight (the kernel: 8 ring slots of 16-22 KB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldgsts_tput ldgsts_tput.cu
// Row pattern: runs of RUN consecutive table rows starting at a pseudo-random row (kernel maps: ~74 % of the
// neighbours of consecutive output rows are consecutive), a fraction `miss` of the visits zero-filled (src-size 0).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t n) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// PPR = 16-byte pieces per row visit.  One "stage" of a warp = 8 KB = 16 LDGSTS.  LAYOUT 0: destination rows of
// 64 B (SWIZZLE_64B image: piece ^ ((r >> 1) & 3)), a visit wider than 64 B is split over PPR/4 tiles 2 KB apart
// (what template G of conv_umma.cu does); LAYOUT 1: destination rows of PPR*16 B contiguous, 16-byte pieces XORed
// with (r & 7) (SWIZZLE_128B image for PPR = 8).
template <int PPR, int LAYOUT, int DEPTH, int RUN>
__global__ void __launch_bounds__(512, 1)
k(const char* __restrict__ table, int n_rows, int row_bytes, int iters, int miss64, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int ROWS = 512 / PPR;   // rows per stage (8 KB / visit width)
  uint32_t pos = (blockIdx.x * 16 + warp) * 1000003u;
  for (int it = 0; it < iters; ++it) {
    const uint32_t slot = base + (uint32_t)((warp * DEPTH + it % DEPTH) * 8192);
    // one hash per stage (uniform); the stage's rows are runs of RUN consecutive rows 977 rows apart
    const uint32_t start = mix(pos) % (uint32_t)(n_rows - 977 * (ROWS / RUN + 2));
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int f = q * 32 + lane;
      const int r = f / PPR, piece = f - r * PPR;        // r in [0, ROWS)
      const uint32_t row = start + (uint32_t)((r / RUN) * 977 + r % RUN);
      const bool miss = (uint32_t)((r * 37 + it * 11) & 63) < (uint32_t)miss64;
      const char* src = table + (size_t)row * row_bytes + piece * 16;
      uint32_t dst;
      if (LAYOUT == 0) dst = slot + (uint32_t)((piece >> 2) * (ROWS * 64) + r * 64 + (((piece & 3) ^ ((r >> 1) & 3)) << 4));
      else dst = slot + (uint32_t)(r * (PPR * 16) + ((piece ^ (r & 7)) % PPR) * 16);
      cp_async_16(dst, src, miss ? 0u : 16u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    pos += ROWS;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (threadIdx.x == 0 && iters < 0) *sink = smem[0];
}

template <int PPR, int LAYOUT, int DEPTH, int RUN>
void run(const char* table, int n_rows, int row_bytes, int miss, unsigned long long* sink, int NW = 8) {
  const int runlen = RUN;
  const int miss64 = miss * 64 / 100;
  const size_t smem = NW * DEPTH * 8192 + 1024;
  auto kern = k<PPR, LAYOUT, DEPTH, RUN>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 6000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, NW * 32, smem>>>(table, n_rows, row_bytes, 300, miss64, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  kern<<<148, NW * 32, smem>>>(table, n_rows, row_bytes, iters, miss64, sink);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clk = ms * 1e-3 * 1.9e9;                 // nominal 1.9 GHz
  const double visits = (double)NW * iters * (512.0 / PPR);    // per SM
  printf("warps %2d visit %3d B layout %d depth %d run %2d miss %2d%%: %7.3f ms  %5.2f clk/visit  %5.2f clk/LDGSTS  %6.1f B/clk/SM (smem bytes)\n",
         NW, PPR * 16, LAYOUT, DEPTH, runlen, miss, ms, clk / visits, clk / ((double)NW * iters * 16), (double)NW * iters * 8192 / clk);
}

int main() {
  const int n_rows = 150000, row_bytes = 256;   // 38 MB table: L2 resident like the feature rows of a level
  char* table;
  CK(cudaMalloc(&table, (size_t)n_rows * row_bytes));
  CK(cudaMemset(table, 1, (size_t)n_rows * row_bytes));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  for (int miss = 0; miss <= 36; miss += 36) {
    run<4, 0, 3, 8>(table, n_rows, row_bytes, miss, sink);
    run<8, 0, 3, 8>(table, n_rows, row_bytes, miss, sink);
    run<8, 1, 3, 8>(table, n_rows, row_bytes, miss, sink);
    run<12, 0, 3, 8>(table, n_rows, row_bytes, miss, sink);
    run<16, 0, 3, 8>(table, n_rows, row_bytes, miss, sink);
    run<16, 1, 3, 8>(table, n_rows, row_bytes, miss, sink);
    run<4, 0, 3, 1>(table, n_rows, row_bytes, miss, sink);
    run<8, 1, 3, 1>(table, n_rows, row_bytes, miss, sink);
    run<16, 1, 3, 1>(table, n_rows, row_bytes, miss, sink);
  }
  // warp-count sweep: is the rate per warp or per SM?
  for (int nw = 1; nw <= 16; nw *= 2) {
    if (nw * 3 * 8192 + 1024 > 227 * 1024) { run<4, 0, 1, 8>(table, n_rows, row_bytes, 36, sink, nw); continue; }
    run<4, 0, 3, 8>(table, n_rows, row_bytes, 36, sink, nw);
  }
  for (int nw = 1; nw <= 16; nw *= 2) run<4, 0, 1, 8>(table, n_rows, row_bytes, 0, sink, nw);
  // depth sweep on the two candidates
  run<4, 0, 1, 8>(table, n_rows, row_bytes, 36, sink);
  run<4, 0, 2, 8>(table, n_rows, row_bytes, 36, sink);
  run<8, 1, 1, 8>(table, n_rows, row_bytes, 36, sink);
  run<8, 1, 2, 8>(table, n_rows, row_bytes, 36, sink);
  return 0;
}
