// Microbenchmark: TMA tile::gather4 row-gather throughput on sm_100a (is the TMA engine a faster
// way than LDGSTS to lay gathered feature rows into swizzled UMMA tiles?).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_bench gather4_bench.cu
//   run  : ./gather4_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void gather4(uint32_t dst, const void* tmap, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :: "r"(dst), "l"(tmap), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// Every warp owns NS slots of 32 x 4 rows; each lane issues one gather4 per slot use.
template <int NS, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
bench_kernel(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx, int n_idx, int iters,
             int row_bytes, int ncol_chunks, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot_bytes = 128 * row_bytes;  // 32 lanes x 4 rows
  const uint32_t bar0 = base + NW * NS * slot_bytes;
  if (threadIdx.x == 0)
    for (int i = 0; i < NW * NS; ++i) mbar_init(bar0 + 8 * i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t phase_bits = 0;
  int s = 0;
  long long pos = ((long long)blockIdx.x * NW + warp) * 128;
  for (int it = 0; it < iters; ++it) {
    const uint32_t bar = bar0 + 8 * (warp * NS + s);
    if (it >= NS) {  // previous use of this slot must have landed
      while (!mbar_try_wait(bar, (phase_bits >> s) & 1u)) {}
      phase_bits ^= 1u << s;
    }
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)slot_bytes);
    __syncwarp();
    const int p = (int)((pos + lane * 4) % n_idx);
    const int4 r = *reinterpret_cast<const int4*>(idx + p);
    const uint32_t dst = base + (warp * NS + s) * slot_bytes + lane * 4 * row_bytes;
    gather4(dst, &tmap, bar, (it % ncol_chunks) * (row_bytes / 2), r.x, r.y, r.z, r.w);
    pos += (long long)gridDim.x * NW * 128;
    if (++s == NS) s = 0;
  }
  // drain
  for (int d = 0; d < NS && d < iters; ++d) {
    const uint32_t bar = bar0 + 8 * (warp * NS + s);
    while (!mbar_try_wait(bar, (phase_bits >> s) & 1u)) {}
    if (++s == NS) s = 0;
  }
  if (sink && threadIdx.x == 0) sink[blockIdx.x] = *reinterpret_cast<unsigned long long*>(smem + 1024);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Variant B: the 32 gather4 of a slot are issued by ONE elected thread per warp from indices staged in
// shared memory (no per-lane election waterfall around the UTMALDG).
template <int NS, int NW>
__global__ void __launch_bounds__(NW * 32, 1)
bench_kernel_elect(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx, int n_idx, int iters,
                   int row_bytes, int ncol_chunks, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot_bytes = 128 * row_bytes;
  const uint32_t bar0 = base + NW * NS * slot_bytes;
  const uint32_t idx0 = bar0 + 8 * NW * NS + 64;  // per warp 2 x 128 ints
  if (threadIdx.x == 0)
    for (int i = 0; i < NW * NS; ++i) mbar_init(bar0 + 8 * i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  uint32_t phase_bits = 0;
  int s = 0;
  long long pos = ((long long)blockIdx.x * NW + warp) * 128;
  const bool leader = elect_one();
  for (int it = 0; it < iters; ++it) {
    const uint32_t bar = bar0 + 8 * (warp * NS + s);
    // stage this slot's 128 row indices in shared memory (all lanes), double-buffered by parity
    const int p = (int)((pos + lane * 4) % n_idx);
    const int4 r = *reinterpret_cast<const int4*>(idx + p);
    const uint32_t ib = idx0 + (warp * 2 + (it & 1)) * 512;
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(ib + lane * 16), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w) : "memory");
    __syncwarp();
    if (leader) {
      if (it >= NS) {
        while (!mbar_try_wait(bar, (phase_bits >> s) & 1u)) {}
        phase_bits ^= 1u << s;
      }
      mbar_expect_tx(bar, (uint32_t)slot_bytes);
      const uint32_t dst0 = base + (warp * NS + s) * slot_bytes;
      const int col = (it % ncol_chunks) * (row_bytes / 2);
#pragma unroll 8
      for (int g = 0; g < 32; ++g) {
        int a, b, c, d;
        asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(ib + g * 16) : "memory");
        gather4(dst0 + g * 4 * row_bytes, &tmap, bar, col, a, b, c, d);
      }
    }
    __syncwarp();
    pos += (long long)gridDim.x * NW * 128;
    if (++s == NS) s = 0;
  }
  if (leader) {
    for (int d = 0; d < NS && d < iters; ++d) {
      const uint32_t bar = bar0 + 8 * (warp * NS + s);
      while (!mbar_try_wait(bar, (phase_bits >> s) & 1u)) {}
      if (++s == NS) s = 0;
    }
  }
  if (sink && threadIdx.x == 0) sink[blockIdx.x] = *reinterpret_cast<unsigned long long*>(smem + 1024);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NS, int NW, int VAR>
void run(const CUtensorMap& tmap, const int* idx, int n_idx, int row_bytes, int M, int pattern) {
  auto kern = VAR ? bench_kernel_elect<NS, NW> : bench_kernel<NS, NW>;
  const int slot_bytes = 128 * row_bytes;
  const size_t smem = (size_t)NW * NS * slot_bytes + 1024 + 8 * NW * NS + 64 + NW * 1024 + 64;
  if (smem > 227 * 1024) return;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int iters = 2000;
  const int ncc = row_bytes == 64 ? 3 : 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, NW * 32, smem>>>(tmap, idx, n_idx, 200, row_bytes, ncc, nullptr);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  kern<<<148, NW * 32, smem>>>(tmap, idx, n_idx, iters, row_bytes, ncc, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double rows = 148.0 * NW * 128 * iters;
  printf("M=%7d pattern=%d row_bytes=%3d NS=%d NW=%2d var=%d: %.3f ms  %.1f Grows/s  %.0f GB/s (slot bytes)  %.2f cycles/gather4/SM @1.9GHz\n", M,
         pattern, row_bytes, NS, NW, VAR, ms, rows / ms / 1e6, rows * row_bytes / ms / 1e6, ms * 1e-3 * 1.9e9 / (NW * 32.0 * iters));
}

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int C = 96;
  for (int M : {200000}) {
    uint16_t* table;
    CK(cudaMalloc(&table, (size_t)M * C * 2));
    CK(cudaMemset(table, 1, (size_t)M * C * 2));
    const int n_idx = 1 << 22;
    std::vector<int> h(n_idx);
    for (int pattern = 0; pattern < 3; ++pattern) {
      uint64_t st = 12345;
      for (int i = 0; i < n_idx; ++i) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        int rnd = (int)((st >> 33) % (uint64_t)M);
        if (pattern == 0) h[i] = i % M;                               // sequential rows
        else if (pattern == 1) h[i] = rnd;                            // uniform random rows
        else h[i] = ((st >> 20) % 100 < 36) ? -1 : (int)(((long long)i + (rnd % 4096)) % M);  // neighbour-like, 36 % missing
      }
      int* idx;
      CK(cudaMalloc(&idx, n_idx * 4));
      CK(cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice));
      for (int row_bytes : {64, 128}) {
        CUtensorMap tmap;
        cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)M};
        cuuint64_t gstr[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {(cuuint32_t)(row_bytes / 2), 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, table, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        run<2, 8, 0>(tmap, idx, n_idx, row_bytes, M, pattern);
        run<2, 8, 1>(tmap, idx, n_idx, row_bytes, M, pattern);
        run<1, 8, 1>(tmap, idx, n_idx, row_bytes, M, pattern);
        run<1, 16, 1>(tmap, idx, n_idx, row_bytes, M, pattern);
        run<3, 4, 1>(tmap, idx, n_idx, row_bytes, M, pattern);
      }
      cudaFree(idx);
    }
    cudaFree(table);
  }
  return 0;
}
