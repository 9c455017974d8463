// conv_umma.cu — sparse convolution forward / dgrad as an output-stationary implicit GEMM on
// the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   out[o, :] = sum_k  A[nbr[k, o], :] @ B_k          A: [*, Ck] fp32 rows, B_k: [Ck, Cn]
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-3  producers : gather A rows with 16-byte cp.async (LDGSTS, zero-fill for missing
//                          neighbours) straight into the 128B-swizzled K-major UMMA layout; one
//                          elected thread also streams the pre-swizzled weight slab of the stage
//                          with a single bulk copy on the TMA engine (UBLKCP);
//   warp  8    MMA       : one elected lane issues tcgen05.mma (M=128, N=cn_tile, K=8 per
//                          instruction) for every pipeline stage and commits stage release /
//                          accumulator completion to mbarriers;
//   warps 4-7  epilogue  : tcgen05.ld the fp32 accumulator (32 lanes per warp), add bias, store.
// A CTA tile is MT x 128 output rows (MT accumulators share every weight slab, halving weight
// traffic at MT=2); accumulators are double-buffered in TMEM when they fit in 512 columns so the
// epilogue of tile i overlaps the main loop of tile i+1.  A pipeline stage is one (offset k,
// 32-channel chunk) pair: MT*16 KB of gathered rows + cn_tile*128 B of weights.
// Offsets with no neighbour inside a tile are skipped using the per-tile offset mask built
// with the kernel map (spc_tile_mask).  Output rows are owned by one CTA: no atomics.
//
// The same kernel computes dgrad (A = dOut gathered by the transposed map, B_k = W[k]^T).
#include "common.cuh"
#include "ptx.cuh"

namespace spc {
using namespace ptx;

constexpr int kTileM = 128;          // rows per accumulator (UMMA M)
constexpr int kChunkBytes = 128;     // one swizzle row = 32 fp32 channels
constexpr int kAStageBytes = kTileM * kChunkBytes;  // 16 KB per sub-tile per stage
constexpr int kMaxStages = 8;
constexpr int kNumProducerThreads = 128;
constexpr int kNumEpilogueThreads = 128;
constexpr int kNumThreads = 288;     // 4 producer + 4 epilogue + 1 MMA warp
constexpr int kSmemLimit = 227 * 1024;

struct UmmaConvParams {
  const float* A;            // [*, Ck]
  const float* Bp;           // packed weights [K][kc][nt][cn_tile][32] (swizzled rows)
  const float* bias;         // [Cn] or null
  const int* nbr;            // [K, m_out]
  const uint32_t* tile_mask; // [ceil(m_out/128)] or null (all offsets active)
  float* out;                // [m_out, Cn]
  int m_out, Ck, Cn, K;
  int cn_tile, n_ntiles, kc_count;
  int stages, acc_bufs, tmem_cols;
  int n_work;                // m_tiles * n_ntiles
};

static inline int pick_cn_tile(int Cn) {
  for (int t = 256; t >= 16; t -= 16)
    if (Cn % t == 0) return t;
  return 0;
}

bool umma_fwd_supported(int c_in, int c_out) {
  return c_in >= 32 && c_in % 32 == 0 && c_out % 16 == 0 && pick_cn_tile(c_out) >= 16;
}
int64_t umma_fwd_workspace(int K, int c_in, int c_out) {
  return align_up((int64_t)K * c_in * c_out * 4, 1024) + 1024;
}

// W [K][Ck][Cn] (or [K][Cn][Ck] when transposed) -> per (k, chunk, n-tile) slab of cn_tile rows x
// 128 B, 16-byte chunks XOR-swizzled with (row & 7): the exact shared-memory image UMMA expects
// for a K-major SWIZZLE_128B B operand, so a stage's weights arrive with ONE bulk copy.
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ W, float* __restrict__ Wp, int K, int Ck, int Cn,
                    int cn_tile, int transpose) {
  const long long total = (long long)K * Ck * Cn;
  const int n_ntiles = Cn / cn_tile, kc_count = Ck / 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int jj = (int)(e % 32);
    long long t = e / 32;
    int n = (int)(t % cn_tile); t /= cn_tile;
    int nt = (int)(t % n_ntiles); t /= n_ntiles;
    int kc = (int)(t % kc_count);
    int k = (int)(t / kc_count);
    int c = kc * 32 + jj, col = nt * cn_tile + n;
    float v = transpose ? W[((long long)k * Cn + col) * Ck + c] : W[((long long)k * Ck + c) * Cn + col];
    int j = jj >> 2, w = jj & 3;
    long long slab = (((long long)k * kc_count + kc) * n_ntiles + nt) * cn_tile * 32;
    Wp[slab + n * 32 + ((j ^ (n & 7)) << 2) + w] = v;
  }
}

template <int MT, int LOOKAHEAD>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_umma_kernel(const UmmaConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = p.cn_tile * kChunkBytes;
  const int stage_bytes = MT * kAStageBytes + b_stage_bytes;
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), kNumProducerThreads + 1);  // 128 gather threads + 1 expect_tx
      mbar_init(empty_bar(s), 1);                       // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kNumEpilogueThreads);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int rows_per_work = kTileM * MT;

  if (warp < 4) {
    // ============================ producers ============================
    int stage = 0;
    uint32_t phase = 0;
    int arr_stage = 0, outstanding = 0;
    const int sub = lane >> 3;  // row within a group of 4
    const int j = lane & 7;     // 16-byte chunk within the 128-byte row
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int mtile = w / p.n_ntiles, ntile = w - mtile * p.n_ntiles;
      const int o0 = mtile * rows_per_work;
      uint32_t mask = 0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int t = mtile * MT + mt;
        if ((long long)t * kTileM < p.m_out) mask |= p.tile_mask ? p.tile_mask[t] : 0xFFFFFFFFu;
      }
      if (p.K < 32) mask &= (1u << p.K) - 1u;
      // prefetch the first active offset's indices
      int idx_next[MT];
      int k = mask ? __ffs(mask) - 1 : -1;
      auto load_idx = [&](int kk, int* idx) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          int o = o0 + mt * kTileM + (int)threadIdx.x;
          idx[mt] = (kk >= 0 && o < p.m_out) ? __ldg(p.nbr + (size_t)kk * p.m_out + o) : -1;
        }
      };
      load_idx(k, idx_next);
      while (k >= 0) {
        int idx[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) idx[mt] = idx_next[mt];
        uint32_t rest = mask & ~((2u << k) - 1u);
        const int k_next = rest ? __ffs(rest) - 1 : -1;
        load_idx(k_next, idx_next);
        for (int kc = 0; kc < p.kc_count; ++kc) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t stage_addr = smem_base + (uint32_t)stage * stage_bytes;
          if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)b_stage_bytes);
            const float* src = p.Bp + (((size_t)k * p.kc_count + kc) * p.n_ntiles + ntile) * (size_t)p.cn_tile * 32;
            bulk_g2s(stage_addr + MT * kAStageBytes, src, (uint32_t)b_stage_bytes, full_bar(stage));
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int rl = q * 4 + sub;            // row within this warp's 32 rows
              const int r = warp * 32 + rl;          // row within the 128-row sub-tile
              const int src_row = __shfl_sync(0xffffffffu, idx[mt], rl);
              const float* src = p.A + (size_t)(src_row >= 0 ? src_row : 0) * p.Ck + kc * 32 + j * 4;
              const uint32_t dst = stage_addr + mt * kAStageBytes + r * kChunkBytes + ((j ^ (r & 7)) << 4);
              cp_async_16(dst, src, src_row >= 0 ? 16u : 0u);
            }
          }
          cp_async_commit();
          ++outstanding;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          if (outstanding > LOOKAHEAD) {
            cp_async_wait<LOOKAHEAD>();
            fence_proxy_async_smem();
            mbar_arrive(full_bar(arr_stage));
            if (++arr_stage == p.stages) arr_stage = 0;
            --outstanding;
          }
        }
        k = k_next;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    while (outstanding > 0) {
      mbar_arrive(full_bar(arr_stage));
      if (++arr_stage == p.stages) arr_stage = 0;
      --outstanding;
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t idesc = make_idesc_tf32(kTileM, (uint32_t)p.cn_tile, 0, 0);
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int mtile = w / p.n_ntiles;
      uint32_t mask = 0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int t = mtile * MT + mt;
        if ((long long)t * kTileM < p.m_out) mask |= p.tile_mask ? p.tile_mask[t] : 0xFFFFFFFFu;
      }
      if (p.K < 32) mask &= (1u << p.K) - 1u;
      const int n_iters = __popc(mask) * p.kc_count;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t stage_addr = smem_base + (uint32_t)stage * stage_bytes;
          const uint64_t bdesc = make_desc_sw128(stage_addr + MT * kAStageBytes, 16, 1024);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = make_desc_sw128(stage_addr + mt * kAStageBytes, 16, 1024);
            const uint32_t d = tmem_base + (uint32_t)((acc * MT + mt) * p.cn_tile);
#pragma unroll
            for (int q = 0; q < 4; ++q)  // 4 x (K = 8 tf32 = 32 bytes) per 128-byte swizzle row
              mma_tf32(d, adesc + 2u * q, bdesc + 2u * q, idesc, (it > 0 || q > 0) ? 1u : 0u);
          }
          mma_commit(empty_bar(stage));  // stage reusable once these MMAs have read it
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (lane == 0) mma_commit(tfull_bar(acc));
      __syncwarp();
      if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ============================ epilogue ============================
    const int ew = warp & 3;  // TMEM lane group this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int mtile = w / p.n_ntiles, ntile = w - mtile * p.n_ntiles;
      const int o0 = mtile * rows_per_work;
      uint32_t mask = 0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int t = mtile * MT + mt;
        if ((long long)t * kTileM < p.m_out) mask |= p.tile_mask ? p.tile_mask[t] : 0xFFFFFFFFu;
      }
      if (p.K < 32) mask &= (1u << p.K) - 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int row = o0 + mt * kTileM + ew * 32 + lane;
        const bool row_ok = row < p.m_out;
        float* dst = p.out + (size_t)(row_ok ? row : 0) * p.Cn + ntile * p.cn_tile;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((acc * MT + mt) * p.cn_tile);
        for (int c0 = 0; c0 < p.cn_tile; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (mask == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + ntile * p.cn_tile + c0 + i);
          }
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int MT, int LOOKAHEAD>
static int launch_conv_umma(const UmmaConvParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = conv_umma_kernel<MT, LOOKAHEAD>;
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kNumThreads, smem, stream>>>(p);
  SPC_LAUNCHED("conv_umma_kernel");
  return 0;
}

static int g_umma_force_mt = 0;  // test hook: 0 = auto

int conv_fwd_umma(const float* in, const float* w, const float* bias, const int* nbr,
                  const uint32_t* tile_mask, int64_t m_out, int c_in, int c_out, int K,
                  bool transpose_w, float* out, void* workspace, int64_t workspace_bytes,
                  cudaStream_t stream) {
  if (m_out == 0) return 0;
  SPC_REQUIRE(umma_fwd_supported(c_in, c_out), "shape not supported by the tcgen05 path");
  SPC_REQUIRE(K <= 32, "tcgen05 path supports kernel volume <= 32");
  SPC_REQUIRE(workspace && workspace_bytes >= umma_fwd_workspace(K, c_in, c_out), "workspace too small");
  SPC_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0, "feature rows must be 16-byte aligned");
  float* Wp = (float*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);

  UmmaConvParams p;
  p.A = in; p.Bp = Wp; p.bias = bias; p.nbr = nbr; p.tile_mask = tile_mask; p.out = out;
  p.m_out = (int)m_out; p.Ck = c_in; p.Cn = c_out; p.K = K;
  p.cn_tile = pick_cn_tile(c_out);
  p.n_ntiles = c_out / p.cn_tile;
  p.kc_count = c_in / 32;

  {
    long long total = (long long)K * c_in * c_out;
    int grid = (int)std::min<long long>(ceil_div(total, 256), kNumSMs * 8);
    pack_weights_kernel<<<grid, 256, 0, stream>>>(w, Wp, K, c_in, c_out, p.cn_tile, transpose_w ? 1 : 0);
    SPC_LAUNCHED("pack_weights_kernel");
  }

  // MT = 2 halves weight traffic; only worth it when there are enough tiles to fill the GPU.
  int mt = (2 * p.cn_tile <= 512 && ceil_div(m_out, 256) * p.n_ntiles >= 2 * kNumSMs) ? 2 : 1;
  if (g_umma_force_mt == 1 || g_umma_force_mt == 2) mt = g_umma_force_mt;
  if (mt * p.cn_tile > 512) mt = 1;
  const int rows_per_work = kTileM * mt;
  p.n_work = (int)ceil_div(m_out, rows_per_work) * p.n_ntiles;
  p.acc_bufs = (2 * mt * p.cn_tile <= 512) ? 2 : 1;
  int cols = p.acc_bufs * mt * p.cn_tile;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const int stage_bytes = mt * kAStageBytes + p.cn_tile * kChunkBytes;
  int stages = (kSmemLimit - 1024 - 256) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  SPC_REQUIRE(stages >= 2, "tile does not fit in shared memory");
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  const int grid = p.n_work < kNumSMs ? p.n_work : kNumSMs;
  // LOOKAHEAD < stages
  const int la = stages >= 8 ? 7 : (stages >= 6 ? 5 : (stages >= 4 ? 3 : 1));
#define SPC_LAUNCH_UMMA(MTV)                                                          \
  switch (la) {                                                                       \
    case 7: return launch_conv_umma<MTV, 7>(p, grid, smem, stream);                   \
    case 5: return launch_conv_umma<MTV, 5>(p, grid, smem, stream);                   \
    case 3: return launch_conv_umma<MTV, 3>(p, grid, smem, stream);                   \
    default: return launch_conv_umma<MTV, 1>(p, grid, smem, stream);                  \
  }
  if (mt == 2) { SPC_LAUNCH_UMMA(2) } else { SPC_LAUNCH_UMMA(1) }
#undef SPC_LAUNCH_UMMA
}

void umma_set_force_mt(int mt) { g_umma_force_mt = mt; }

bool umma_wgrad_supported(int, int) { return false; }
int64_t umma_wgrad_workspace(int, int, int) { return 0; }
int conv_wgrad_umma(const float*, const float*, const int*, int64_t, int, int, int, float*, void*,
                    int64_t, cudaStream_t) {
  return fail("conv_wgrad_umma", "not built");
}

}  // namespace spc
