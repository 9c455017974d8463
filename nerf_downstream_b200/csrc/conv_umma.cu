// conv_umma.cu — sparse convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators
// in TMEM): forward / dgrad as an output-stationary implicit GEMM, wgrad as a reduction GEMM.
//
//   fwd / dgrad:  out[o, :] = sum_k  A[nbr[k, o], :] @ B_k        A: [*, Ck] rows, B_k: [Ck, Cn]
//   wgrad      :  dW[k]     = sum_o  in[nbr[k, o], :]^T  dout[o, :]
//
// Two operand precisions share the code (template BF16):
//   false  kind::tf32 — fp32 rows in shared memory (the tensor core uses the upper 19 bits),
//          32 channels = 128-byte swizzle rows, 4 MMAs (K=8) per row chunk;
//   true   kind::f16 with bf16 operands — rows converted once per layer (spc_to_bf16), 32 channels =
//          64-byte rows (SWIZZLE_64B), 2 MMAs (K=16) per row chunk: half the gather bytes and twice
//          the MMA rate.  Accumulation is fp32 in TMEM in both.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-7   producers : ONE WARP PER PIPELINE STAGE.  Stage n is filled entirely by producer warp
//                           n mod 8: it prefetches the stage's neighbour indices, waits for its ring slot,
//                           gathers the rows with 16-byte cp.async (LDGSTS, zero-fill for missing
//                           neighbours) straight into the swizzled UMMA layout and brings the stage's
//                           weights as ONE TMA bulk copy (UBLKCP) of a pre-swizzled slab.  "Full"
//                           barriers are signalled by the hardware (cp.async.mbarrier.arrive.noinc):
//                           no wait_group / proxy fence on the producer side, it only waits for slots.
//                           (wgrad: one warp per A stage + dedicated warps that bring the contiguous
//                           dout row blocks with TMA tile loads.)
//   warps 8-11  epilogue  : tcgen05.ld the fp32 accumulator (32 TMEM lanes per warp), bias, 16-byte
//                           conflict-free stores into a 128B-swizzled shared-memory tile, TMA tensor
//                           store (reduce-add on offset-split items) — no st.global, the LSU belongs to
//                           the gather (wgrad: fp32 red.global.add.v4 of the partial dW);
//   warp  12    MMA       : one ELECTED thread (elect.sync) runs the whole issue loop: tcgen05.mma from
//                           uniform registers, tcgen05.commit for stage release / accumulator completion.
// Forward: a work item is MT x 128 output rows x cn_tile channels (MT accumulators share every weight
// slab), on small maps x one group of kernel offsets; accumulators are double-buffered in TMEM when they
// fit in 512 columns so the epilogue of tile i overlaps the main loop of tile i+1.  A pipeline stage is
// one (offset k, 32-channel chunk) pair.  Offsets with no neighbour inside a tile are skipped using the
// per-tile offset mask (spc_tile_mask).  On large maps output rows are owned by one CTA: no atomics in
// forward / dgrad.  What bounds it (profiles/): the LSU / shared-memory pipe shared by the LDGSTS gather,
// any epilogue stores and even UTCHMMA issue — which is why everything except the gather stays off it.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace spc {
using namespace ptx;

constexpr int kTileM = 128;  // rows per accumulator (UMMA M)
constexpr int kMaxStages = 8;
constexpr int kNumProducerWarps = 8;
constexpr int kNumEpilogueThreads = 128;
constexpr int kMmaWarp = kNumProducerWarps + 4;
constexpr int kNumThreads = (kMmaWarp + 1) * 32;  // 8 producer + 4 epilogue + 1 MMA warp
constexpr int kSmemLimit = 227 * 1024;

// cycle counters of the wgrad roles (test hook spc_debug_read): per CTA 8 x int64
__device__ long long g_wg_counters[kNumSMs * 8];

template <bool BF16>
struct Prec {
  static constexpr int kElt = BF16 ? 2 : 4;             // bytes per element
  static constexpr int kRowBytes = 32 * kElt;           // one 32-channel chunk of a row
  static constexpr int kLanesPerRow = kRowBytes / 16;   // 16-byte cp.async pieces per row chunk
  static constexpr int kRowsPerInstr = 32 / kLanesPerRow;
  static constexpr int kMmaPerRow = BF16 ? 2 : 4;       // K = 16 bf16 / 8 tf32 = 32 bytes each
  // K-major operand (forward): SWIZZLE_64B for 64-byte rows, SWIZZLE_128B for 128-byte rows
  static constexpr uint32_t kLayoutK = BF16 ? 4u : 2u;
  static constexpr uint32_t kSboK = 8 * kRowBytes;
  // MN-major operand (wgrad): 32-bit types must use SWIZZLE_128B_BASE32B, bf16 uses SWIZZLE_64B
  static constexpr uint32_t kLayoutMN = BF16 ? 4u : 1u;
  // byte offset of 16-byte piece j of row r inside its row chunk (K-major forward layout)
  __device__ static __forceinline__ uint32_t swz_k(int j, int r) {
    return BF16 ? (uint32_t)((j ^ ((r >> 1) & 3)) << 4) : (uint32_t)((j ^ (r & 7)) << 4);
  }
  // same for the MN-major wgrad layout (tf32: 32-byte chunks XOR row&3)
  __device__ static __forceinline__ uint32_t swz_mn(int j, int r) {
    return BF16 ? (uint32_t)((j ^ ((r >> 1) & 3)) << 4)
                : (uint32_t)((((j >> 1) ^ (r & 3)) << 5) | ((j & 1) << 4));
  }
  __device__ static __forceinline__ uint32_t idesc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return BF16 ? make_idesc_bf16(M, N, a_mn, b_mn) : make_idesc_tf32(M, N, a_mn, b_mn);
  }
  __device__ static __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    if (BF16) mma_bf16(d, a, b, id, acc);
    else mma_tf32(d, a, b, id, acc);
  }
};

struct UmmaConvParams {
  const void* A;             // [*, Ck] fp32 or bf16
  const void* Bp;            // packed weights [K][kc][nt][cn_tile][32] (swizzled rows)
  const float* bias;         // [Cn] or null
  const int* nbr;            // [K, m_out]
  const uint32_t* tile_mask; // [ceil(m_out/128)] or null (all offsets active)
  float* out;                // [m_out, Cn]
  double* stats;             // optional [2][Cn]: per-channel sum and sum of squares of `out` (BatchNorm fusion)
  int m_out, Ck, Cn, K;
  int cn_tile, n_ntiles, kc_count;
  int stages, acc_bufs, tmem_cols;
  int dbg_skip_store;        // timing experiment only: epilogue does not write the output
  int tma_store;             // epilogue stages the tile in shared memory and writes it with TMA tensor stores
  int ksplit, k_per;         // offsets split over ksplit work items of k_per offsets each (small maps)
  int n_work;                // m_tiles * n_ntiles * ksplit
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
// one row chunk, 16-byte pieces XOR-swizzled: the exact shared-memory image UMMA expects for a
// K-major swizzled B operand, so a stage's weights arrive with ONE bulk copy.
template <bool BF16>
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ W, void* __restrict__ Wp_, int K, int Ck, int Cn,
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
    long long slab = (((long long)k * kc_count + kc) * n_ntiles + nt) * cn_tile * 32;
    if (BF16) {
      int j = jj >> 3, w = jj & 7;  // 8 bf16 per 16-byte piece
      reinterpret_cast<__nv_bfloat16*>(Wp_)[slab + n * 32 + ((j ^ ((n >> 1) & 3)) << 3) + w] = __float2bfloat16_rn(v);
    } else {
      int j = jj >> 2, w = jj & 3;  // 4 fp32 per 16-byte piece
      reinterpret_cast<float*>(Wp_)[slab + n * 32 + ((j ^ (n & 7)) << 2) + w] = v;
    }
  }
}

// fp32 rows (row pitch `src_pitch` elements, c_src valid columns) -> dense bf16 rows of c_dst >= c_src columns
// (columns past c_src are zero): conversion, de-striding of a column slice and channel padding in one pass.
__global__ void __launch_bounds__(256)
to_bf16_kernel(const float* __restrict__ src, long long rows, int c_src, long long src_pitch, int c_dst,
               __nv_bfloat16* __restrict__ dst, int vec_ok) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {  // 4 columns per thread: c_src, c_dst, src_pitch multiples of 4, pointers aligned
    const int q_per_row = c_dst / 4;
    const long long nq = rows * q_per_row;
    for (long long q = i0; q < nq; q += stride) {
      const long long r = q / q_per_row;
      const int c = (int)(q - r * q_per_row) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < c_src) v = *reinterpret_cast<const float4*>(src + r * src_pitch + c);
      __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(dst + r * c_dst + c) = o;
    }
  } else {
    const long long n = rows * c_dst;
    for (long long e = i0; e < n; e += stride) {
      const long long r = e / c_dst;
      const int c = (int)(e - r * c_dst);
      dst[e] = __float2bfloat16_rn(c < c_src ? src[r * src_pitch + c] : 0.f);
    }
  }
}

int to_bf16(const float* src, int64_t rows, int c_src, int64_t src_pitch, int c_dst, void* dst, cudaStream_t stream) {
  if (rows == 0 || c_dst == 0) return 0;
  const int vec_ok = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 8 == 0) && c_src % 4 == 0 && c_dst % 4 == 0 &&
                     src_pitch % 4 == 0;
  const int64_t n = rows * c_dst;
  int64_t want = ceil_div(vec_ok ? ceil_div(n, 4) : n, 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  to_bf16_kernel<<<grid, 256, 0, stream>>>(src, rows, c_src, src_pitch, c_dst, (__nv_bfloat16*)dst, vec_ok);
  SPC_LAUNCHED("to_bf16_kernel");
  return 0;
}

template <int MT, bool BF16>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_umma_kernel(const UmmaConvParams p, const __grid_constant__ CUtensorMap tmap_out) {
  using PR = Prec<BF16>;
  constexpr int kAStage = kTileM * PR::kRowBytes;  // one sub-tile of one stage
  extern __shared__ uint8_t smem_raw[];
  // swizzled operands need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = p.cn_tile * PR::kRowBytes;
  const int stage_bytes = MT * kAStage + b_stage_bytes;
  // output staging of the TMA-store epilogue: cn_tile / 32 blocks of [128 rows x 32 fp32], SWIZZLE_128B
  const uint32_t out_stage = smem_base + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = out_stage + (p.tma_store ? (uint32_t)(p.cn_tile / 32) * 16384u : 0u);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // BatchNorm-statistics fusion (p.stats): per epilogue warp a [16][33] transposition scratch and
  // double accumulators [2][Cn] behind the barrier block
  uint8_t* stats_smem = smem_raw + (bar_base + 256u - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 32 + 1);                   // the 32 lanes of the owning warp + 1 expect_tx
      mbar_init(empty_bar(s), 1);                       // one tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kNumEpilogueThreads);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int rows_per_work = kTileM * MT;
  // work item w -> (m tile, n tile, offset group); w = (mtile * n_ntiles + ntile) * ksplit + kg
  const int items_per_mtile = p.n_ntiles * p.ksplit;
  auto work_mask = [&](int w) -> uint32_t {
    const int mtile = w / items_per_mtile;
    const int kg = w % p.ksplit;
    uint32_t mask = 0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      int t = mtile * MT + mt;
      if ((long long)t * kTileM < p.m_out) mask |= p.tile_mask ? p.tile_mask[t] : 0xFFFFFFFFu;
    }
    if (p.K < 32) mask &= (1u << p.K) - 1u;
    if (p.ksplit > 1) {  // this item's share of the offsets
      const int k0 = kg * p.k_per, k1 = min(k0 + p.k_per, p.K);
      const uint32_t hi = k1 >= 32 ? 0xFFFFFFFFu : ((1u << k1) - 1u);
      mask &= hi & ~((1u << k0) - 1u);
    }
    return mask;
  };

  if (warp < kNumProducerWarps) {
    // ============================ producers ============================
    // Stage n of this CTA's stage sequence (work item, active offset k, channel chunk kc) is filled
    // entirely by warp n % nprod: a warp waits for ITS ring slot, issues every gather copy of the
    // stage (and the weight slab) and moves on, so the address arithmetic of up to eight stages runs
    // concurrently on the four schedulers instead of all warps marching through one stage in
    // lock-step.  nprod <= ring slots: a warp is then never two ring revolutions ahead of the MMA
    // warp, which the one-bit phase parity of the "empty" barriers could not tell apart.
    const int nprod = p.stages < kNumProducerWarps ? p.stages : kNumProducerWarps;
    constexpr int R = PR::kRowsPerInstr;  // rows one LDGSTS instruction covers
    constexpr int NQ = kTileM / R;        // instructions per 128-row sub-tile
    const int sub = lane / PR::kLanesPerRow;  // row within the rows one instruction covers
    const int j = lane % PR::kLanesPerRow;    // 16-byte piece within the row chunk
    const char* Abase = reinterpret_cast<const char*>(p.A) + j * 16;
    const char* Bbase = reinterpret_cast<const char*>(p.Bp);
    const size_t row_pitch = (size_t)p.Ck * PR::kElt;
    const bool leader = elect_one();

    // iterator over the stages this warp owns
    struct It { int w, s, S, ord, n0; uint32_t mask, rest; bool ok; };  // n0 = (first stage of item) % nprod
    auto open_item = [&](It& it) {  // first owned stage of item it.w or of a later item
      for (;;) {
        if (it.w >= p.n_work || warp >= nprod) { it.ok = false; return; }
        it.mask = work_mask(it.w);
        it.S = __popc(it.mask) * p.kc_count;
        it.s = warp - it.n0;
        if (it.s < 0) it.s += nprod;
        if (it.s < it.S) { it.rest = it.mask; it.ord = 0; return; }
        it.n0 = (it.n0 + it.S) % nprod;
        it.w += gridDim.x;
      }
    };
    auto advance = [&](It& it) {
      it.s += nprod;
      if (it.s >= it.S) {
        it.n0 = (it.n0 + it.S) % nprod;
        it.w += gridDim.x;
        open_item(it);
      }
    };
    // (k, kc) of the current stage; `rest` / `ord` walk the set bits of the offset mask
    auto locate = [&](It& it, int& k, int& kc) {
      const int ord = it.s / p.kc_count;
      kc = it.s - ord * p.kc_count;
      while (it.ord < ord) { it.rest &= it.rest - 1u; ++it.ord; }
      k = __ffs(it.rest) - 1;
    };
    // lane l holds the neighbour rows of tile rows l, l+32, l+64, l+96 of every sub-tile
    auto load_idx = [&](const It& it, int k, int* idx) {
      const int o0 = (it.w / items_per_mtile) * rows_per_work;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int o = o0 + mt * kTileM + i * 32 + lane;
          idx[mt * 4 + i] = o < p.m_out ? __ldg(p.nbr + (size_t)k * p.m_out + o) : -1;
        }
      }
    };

    It cur;
    cur.w = blockIdx.x; cur.n0 = 0; cur.ok = true;
    open_item(cur);
    int k = 0, kc = 0;
    int idx[MT * 4];
    if (cur.ok) { locate(cur, k, kc); load_idx(cur, k, idx); }
    int slot = warp;  // ring slot / phase of sequence number warp + nprod * i
    uint32_t phase = 0;
    while (cur.ok) {
      // indices of the NEXT owned stage: their latency hides behind this stage's slot wait
      It nxt = cur;
      advance(nxt);
      int k_n = 0, kc_n = 0;
      int idx_n[MT * 4];
      if (nxt.ok) { locate(nxt, k_n, kc_n); load_idx(nxt, k_n, idx_n); }

      const int ntile = (cur.w / p.ksplit) % p.n_ntiles;
      mbar_wait(empty_bar(slot), phase ^ 1u);
      const uint32_t stage_addr = smem_base + (uint32_t)slot * stage_bytes;
      if (leader) {
        mbar_arrive_expect_tx(full_bar(slot), (uint32_t)b_stage_bytes);
        const char* src = Bbase + (((size_t)k * p.kc_count + kc) * p.n_ntiles + ntile) * (size_t)b_stage_bytes;
        bulk_g2s(stage_addr + MT * kAStage, src, (uint32_t)b_stage_bytes, full_bar(slot));
      }
      __syncwarp();
      const char* Akc = Abase + (size_t)kc * PR::kRowBytes;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int r = q * R + sub;  // row within the 128-row sub-tile
          const int src_row = __shfl_sync(0xffffffffu, idx[mt * 4 + ((q * R) >> 5)], r & 31);
          const char* src = Akc + (size_t)(src_row >= 0 ? src_row : 0) * row_pitch;
          const uint32_t dst = stage_addr + mt * kAStage + r * PR::kRowBytes + PR::swz_k(j, r);
          cp_async_16(dst, src, src_row >= 0 ? 16u : 0u);
        }
      }
      // the stage's "full" barrier is signalled by the hardware when this lane's copies have landed:
      // no wait_group, no fence, nothing blocks here
      cp_async_mbar_arrive_noinc(full_bar(slot));
      slot += nprod;
      if (slot >= p.stages) { slot -= p.stages; phase ^= 1u; }
      cur = nxt; k = k_n; kc = kc_n;
#pragma unroll
      for (int i = 0; i < MT * 4; ++i) idx[i] = idx_n[i];
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ============================
    // ONE elected thread runs the whole loop: under elect.sync the compiler keeps descriptors in
    // uniform registers and emits back-to-back UTCHMMA (a `lane == 0` branch costs an ELECT /
    // BRA.U.ANY waterfall around every tcgen05 instruction).
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = PR::idesc(kTileM, (uint32_t)p.cn_tile, 0, 0);
      const uint64_t desc_hi = make_desc(0, 16, PR::kSboK, PR::kLayoutK);
      for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
        const uint32_t mask = work_mask(w);
        const int n_iters = __popc(mask) * p.kc_count;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          fence_proxy_async_smem();  // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
          tc_fence_after();
          const uint32_t stage_addr = smem_base + (uint32_t)stage * stage_bytes;
          const uint64_t bdesc = desc_hi | (uint64_t)(((stage_addr + MT * kAStage) >> 4) & 0x3FFFu);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = desc_hi | (uint64_t)(((stage_addr + mt * kAStage) >> 4) & 0x3FFFu);
            const uint32_t d = tmem_base + (uint32_t)((acc * MT + mt) * p.cn_tile);
#pragma unroll
            for (int q = 0; q < PR::kMmaPerRow; ++q)  // 32 bytes of K per MMA
              PR::mma(d, adesc + 2u * q, bdesc + 2u * q, idesc, (it > 0 || q > 0) ? 1u : 0u);
          }
          mma_commit(empty_bar(stage));  // stage reusable once these MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        mma_commit(tfull_bar(acc));
        if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue ============================
    const int ew = warp & 3;  // TMEM lane group this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    // statistics fusion: this warp's scratch; lanes 0-15 own the sums, lanes 16-31 the sums of squares of
    // 16 columns at a time (host guarantees n_ntiles == 1 and ksplit == 1 when p.stats is set)
    float* s_tr = reinterpret_cast<float*>(stats_smem) + ew * (16 * 33);
    double* s_dacc = reinterpret_cast<double*>(stats_smem + 4 * 16 * 33 * sizeof(float)) + ew * 2 * p.Cn;
    if (p.stats) {
      for (int c = lane; c < 2 * p.Cn; c += 32) s_dacc[c] = 0.0;
      __syncwarp();
    }
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int mtile = w / items_per_mtile, ntile = (w / p.ksplit) % p.n_ntiles, kg = w % p.ksplit;
      const int o0 = mtile * rows_per_work;
      const uint32_t mask = work_mask(w);
      const bool add_bias = p.bias != nullptr && kg == 0;
      mbar_wait_sleep(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (p.tma_store) {
        // The accumulator tile goes TMEM -> registers -> shared memory (128-byte-swizzled rows: conflict-free
        // 16-byte stores) -> global memory with TMA tensor stores (full 128-byte row segments; rows past
        // m_out are clipped by the TMA unit).  A thread-per-row st.global from the 32x32b TMEM layout is
        // 32 half-written sectors per instruction and competes with the row gather for the LSU: with the
        // stores removed the kernel ran 23 % faster (96->96, 1 M voxels).
        const bool store_leader = ew == 0 && elect_one();
        const bool skip = mask == 0 && !add_bias && p.ksplit > 1;  // nothing to add (uniform over the CTA)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int r = ew * 32 + lane;  // row within the sub-tile
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((acc * MT + mt) * p.cn_tile);
          // the previous tensor stores must have finished READING the staging buffer
          if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (!skip) {
            for (int c0 = 0; c0 < p.cn_tile; c0 += 16) {
              float v[16];
              tmem_ld16(taddr + c0, v);
              if (mask == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
              }
              if (add_bias) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + ntile * p.cn_tile + c0 + i);
              }
              const uint32_t blk = out_stage + (uint32_t)(c0 >> 5) * 16384u + (uint32_t)r * 128u;
              const int j0 = (c0 >> 4 & 1) * 4;  // 16-byte piece of the 128-byte row
#pragma unroll
              for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (uint32_t)(((j0 + i) ^ (r & 7)) << 4)),
                             "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                             : "memory");
            }
            fence_proxy_async_smem();  // generic-proxy stores -> visible to the TMA (async proxy) reads
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (store_leader && !skip && !p.dbg_skip_store) {
            const int row0 = o0 + mt * kTileM;
            if (row0 < p.m_out) {
              for (int cb = 0; cb < p.cn_tile / 32; ++cb) {
                const int col0 = ntile * p.cn_tile + cb * 32;
                if (p.ksplit > 1) tma_reduce_add_2d(&tmap_out, out_stage + cb * 16384u, col0, row0);
                else tma_store_2d(&tmap_out, out_stage + cb * 16384u, col0, row0);
              }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else {
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
          if (add_bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + ntile * p.cn_tile + c0 + i);
          }
          if (row_ok && !p.dbg_skip_store) {
            if (p.ksplit > 1) {  // partial sums of several offset groups meet in the (zeroed) output
              if (mask != 0 || add_bias) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(v[i]),
                               "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                               : "memory");
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(dst + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
          if (p.stats) {
            // column sums over this warp's 32 rows: transpose through shared memory (conflict-free, pitch 33),
            // one lane per (column, sum | sum of squares), fp32 over 32 rows, double across tiles
#pragma unroll
            for (int i = 0; i < 16; ++i) s_tr[i * 33 + lane] = row_ok ? v[i] : 0.f;
            __syncwarp();
            const int col = lane & 15;
            float a = 0.f;
            if (lane < 16) {
#pragma unroll 8
              for (int r = 0; r < 32; ++r) a += s_tr[col * 33 + r];
            } else {
#pragma unroll 8
              for (int r = 0; r < 32; ++r) { const float x = s_tr[col * 33 + r]; a = fmaf(x, x, a); }
            }
            s_dacc[(lane >> 4) * p.Cn + c0 + col] += (double)a;
            __syncwarp();
          }
        }
      }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.stats) {  // one double atomic per (warp, channel, moment) and CTA
      for (int c = lane; c < 2 * p.Cn; c += 32) atomicAdd(p.stats + c, s_dacc[c]);
    }
    if (p.tma_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // (only the issuing thread has groups)
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int MT, bool BF16>
static int launch_conv_umma(const UmmaConvParams& p, const CUtensorMap& tmap_out, int grid, size_t smem,
                            cudaStream_t stream) {
  auto kern = conv_umma_kernel<MT, BF16>;
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kNumThreads, smem, stream>>>(p, tmap_out);
  SPC_LAUNCHED("conv_umma_kernel");
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TensorMapEncodeFn)ptr;
  }
  return fn;
}
// fp32 [rows, C] row-major output, box = 32 columns x 128 rows, SWIZZLE_128B: the epilogue's store target
static bool make_out_tile_map(CUtensorMap* map, float* base, int64_t rows, int C) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)C * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int g_umma_force_mt = 0;  // test hook: 0 = auto
static int g_dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // test hook spc_debug_set: 0 = default

// `in` is fp32 (bf16 == false) or bf16 (bf16 == true) rows; weights are always fp32 and packed here.
// `stats` (optional, [2][c_out] doubles): per-channel sum / sum of squares of the output rows, accumulated
// in the epilogue when the launch has one n tile and no offset split; *stats_fused says whether it was.
int conv_fwd_umma(const void* in, const float* w, const float* bias, const int* nbr,
                  const uint32_t* tile_mask, int64_t m_out, int c_in, int c_out, int K,
                  bool transpose_w, bool bf16, float* out, double* stats, int* stats_fused, void* workspace,
                  int64_t workspace_bytes, cudaStream_t stream) {
  if (stats_fused) *stats_fused = 0;
  if (m_out == 0) return 0;
  SPC_REQUIRE(umma_fwd_supported(c_in, c_out), "shape not supported by the tcgen05 path");
  SPC_REQUIRE(K <= 32, "tcgen05 path supports kernel volume <= 32");
  SPC_REQUIRE(workspace && workspace_bytes >= umma_fwd_workspace(K, c_in, c_out), "workspace too small");
  SPC_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0, "feature rows must be 16-byte aligned");
  void* Wp = (void*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const int row_bytes = bf16 ? 64 : 128;
  const int a_stage = kTileM * row_bytes;

  UmmaConvParams p;
  p.A = in; p.Bp = Wp; p.bias = bias; p.nbr = nbr; p.tile_mask = tile_mask; p.out = out; p.stats = nullptr;
  p.dbg_skip_store = g_dbg[3];
  p.m_out = (int)m_out; p.Ck = c_in; p.Cn = c_out; p.K = K;
  p.kc_count = c_in / 32;
  // Tile shape: MT sub-tiles of 128 rows x cn_tile output channels per work item, and on small maps
  // (deep UNet levels: too few row tiles for 148 SMs) the K offsets split over `ksplit` items whose
  // partial sums meet in the zeroed output through fp32 red.global.add.  Chosen to minimise the
  // critical path = waves x bytes a CTA stages per item (gather + weight slabs, what bounds the
  // kernel); large maps always come out as ksplit = 1 (atomic-free, one owner per output row).
  int mt = 1;
  p.cn_tile = pick_cn_tile(c_out);
  p.ksplit = 1;
  {
    double best = 1e300;
    const int cn_max = p.cn_tile;
    for (int cand_mt = 1; cand_mt <= 2; ++cand_mt) {
      if (g_umma_force_mt && cand_mt != g_umma_force_mt) continue;
      for (int cn = cn_max; cn >= 16; cn -= 16) {
        if (c_out % cn || cand_mt * cn > 512) continue;
        const int64_t stage_b = (int64_t)cand_mt * a_stage + (int64_t)cn * row_bytes;
        if ((kSmemLimit - 1024 - 256) / stage_b < 2) continue;
        const int64_t items1 = ceil_div(m_out, kTileM * cand_mt) * (c_out / cn);
        const int ks_max = items1 >= kNumSMs ? 1 : K;
        for (int ks = 1; ks <= ks_max; ++ks) {
          const int k_per = (int)ceil_div(K, ks);
          if (ceil_div(K, k_per) != ks) continue;  // same split as a smaller ks
          const int64_t items = items1 * ks;
          double per_item = (double)k_per * p.kc_count * stage_b;
          if (ks > 1) per_item += 3.0 * cand_mt * kTileM * cn * 4;  // zero-fill + atomic epilogue
          per_item += 20000.0;                                       // fixed per-item latency
          const double cost = (double)ceil_div(items, kNumSMs) * per_item + 1e-6 * items * per_item;
          if (cost < best) { best = cost; mt = cand_mt; p.cn_tile = cn; p.ksplit = ks; }
        }
      }
    }
    SPC_REQUIRE(best < 1e300, "tile does not fit in shared memory");
  }
  p.k_per = (int)ceil_div(K, p.ksplit);
  p.n_ntiles = c_out / p.cn_tile;

  {
    long long total = (long long)K * c_in * c_out;
    int grid = (int)std::min<long long>(ceil_div(total, 256), kNumSMs * 8);
    if (bf16) pack_weights_kernel<true><<<grid, 256, 0, stream>>>(w, Wp, K, c_in, c_out, p.cn_tile, transpose_w ? 1 : 0);
    else pack_weights_kernel<false><<<grid, 256, 0, stream>>>(w, Wp, K, c_in, c_out, p.cn_tile, transpose_w ? 1 : 0);
    SPC_LAUNCHED("pack_weights_kernel");
  }
  if (p.ksplit > 1) SPC_CUDA(cudaMemsetAsync(out, 0, (size_t)m_out * c_out * sizeof(float), stream));

  const int rows_per_work = kTileM * mt;
  p.n_work = (int)ceil_div(m_out, rows_per_work) * p.n_ntiles * p.ksplit;
  p.acc_bufs = (2 * mt * p.cn_tile <= 512) ? 2 : 1;
  int cols = p.acc_bufs * mt * p.cn_tile;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const int stage_bytes = mt * a_stage + p.cn_tile * row_bytes;
  // BatchNorm-statistics fusion: one n tile, no offset split (every output value is final in the epilogue)
  int extra = 0;
  if (stats && p.n_ntiles == 1 && p.ksplit == 1) {
    extra = 4 * 16 * 33 * (int)sizeof(float) + 4 * 2 * c_out * (int)sizeof(double) + 16;
    if ((kSmemLimit - 1024 - 256 - extra) / stage_bytes >= 4) {
      p.stats = stats;
      SPC_CUDA(cudaMemsetAsync(stats, 0, (size_t)2 * c_out * sizeof(double), stream));
      if (stats_fused) *stats_fused = 1;
    } else {
      extra = 0;
    }
  }
  // TMA-store epilogue: needs 32-column blocks and room for the staged tile next to >= 4 ring slots
  CUtensorMap tmap_out;
  memset(&tmap_out, 0, sizeof(tmap_out));
  p.tma_store = 0;
  const int out_stage_bytes = (p.cn_tile / 32) * 16384;
  // (tf32 with narrow tiles measured slower with it: the staged tile costs a ring slot of 36 KB stages)
  if (g_dbg[2] != 1 && !p.stats && (bf16 || p.cn_tile >= 64) && p.cn_tile % 32 == 0 && ((uintptr_t)out % 16) == 0 &&
      (kSmemLimit - 1024 - 256 - extra - out_stage_bytes) / stage_bytes >= 4 &&
      make_out_tile_map(&tmap_out, out, m_out, c_out)) {
    p.tma_store = 1;
    extra += out_stage_bytes;
  }
  int stages = (kSmemLimit - 1024 - 256 - extra) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  SPC_REQUIRE(stages >= 2, "tile does not fit in shared memory");
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + extra;
  const int grid = p.n_work < kNumSMs ? p.n_work : kNumSMs;
  if (bf16) return mt == 2 ? launch_conv_umma<2, true>(p, tmap_out, grid, smem, stream)
                           : launch_conv_umma<1, true>(p, tmap_out, grid, smem, stream);
  return mt == 2 ? launch_conv_umma<2, false>(p, tmap_out, grid, smem, stream)
                 : launch_conv_umma<1, false>(p, tmap_out, grid, smem, stream);
}

void umma_set_force_mt(int mt) { g_umma_force_mt = mt; }
int umma_debug_read(long long* host, int n) {
  if (n > kNumSMs * 8) n = kNumSMs * 8;
  return (int)cudaMemcpyFromSymbol(host, g_wg_counters, (size_t)n * sizeof(long long));
}
void umma_debug_set(int idx, int val) { if (idx >= 0 && idx < 8) g_dbg[idx] = val; }

// =====================================================================================
// wgrad on tcgen05:  dW[k][ci][co] = sum_o in[nbr[k,o]][ci] * dout[o][co]
// =====================================================================================
// GEMM view: D[M x N] += A[M x Kred] * B[Kred x N] with the REDUCTION over out rows o:
//   M = 128 = four 32-channel "chunks", each chunk = (kernel offset k, channel group cc) — so for
//       Cin = 32 four different offsets share one MMA, for Cin = 128 one offset fills it;
//   N = Cout;  Kred = 8 (tf32) / 16 (bf16) rows per tcgen05.mma.
// Both operands are MN-major: a gathered input row IS 32 consecutive M elements, a dout row IS
// Cout consecutive N elements, so rows are laid down as [rows x row-chunk] blocks like in the forward
// kernel.  MN-major 32-bit operands must use SWIZZLE_128B_BASE32B (32-byte chunks XOR row & 3), bf16
// uses SWIZZLE_64B; LBO = distance between 32-element column groups, SBO = distance between 4- (tf32)
// / 8-row (bf16) groups = 512 B either way.
// A pipeline step covers kRows = 64 (tf32) / 128 (bf16) out rows of ONE M block: 4 chunk blocks of
// 8 KB (halving the bf16 step to 64 rows was measured 27 % slower: the per-step costs dominate).  TMEM holds floor(512 / Cout) accumulators; a work item = (row range, pass over a group of
// M blocks) and ends with an fp32 red.global.add of its partial dW — the only atomics of the
// convolution path.
constexpr int kWgChunkBlock = 8192;              // [kRows x row chunk]
constexpr int kWgAStage = 4 * kWgChunkBlock;     // 32 KB: four chunks = one 128-row M block
constexpr int kWgMaxAStages = 6;
constexpr int kWgBStages = 2;
constexpr int kWgIdxBytes = 0;                   // (the per-warp neighbour-index rings of the lock-step producer are gone)
constexpr int kWgMaxMb = 128;    // M blocks (K * Cin / 128): 27 offsets x 512 channels = 108
constexpr int kWgTabBytes = kWgMaxMb * 4 + kWgMaxMb * 4 * 2;  // per-M-block offset masks, per-chunk (k, cc)

struct UmmaWgradParams {
  const void* in;             // [m_in, Cin] fp32 or bf16
  const void* dout;           // [m_out, Cout] fp32 or bf16
  const int* nbr;             // [K, m_out]
  const uint32_t* tile_mask;  // [ceil(m_out/128)] or null
  float* dw;                  // [K, Cin, Cout], zeroed
  int m_out, Cin, Cout, K;
  int ncc, nq;                // chunks per offset, total chunks
  int n_mb, mb_per_pass, n_pass;
  int n_rb, rb_per_split, n_split;
  int a_stages, b_warps;      // A ring slots (= A producer warps), warps per B ring slot
  int b_tma;                  // dout row blocks arrive by TMA tile loads (no LSU work)
  int dbg_skip_mma, dbg_skip_gather, dbg_skip;  // timing experiments only (results are wrong when set)
  int n_work;
};

#define WG_TIMED_WAIT(slot, call)                         \
  do {                                                    \
    long long _t0 = clock64();                            \
    call;                                                 \
    wg_cnt[slot] += clock64() - _t0;                      \
  } while (0)

// offsets (bitmask) that the chunks of M block `mb` (global index) belong to
__device__ __forceinline__ uint32_t mblock_taps(int mb, int ncc, int nq) {
  uint32_t t = 0;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    int q = mb * 4 + s;
    if (q < nq) t |= 1u << (q / ncc);
  }
  return t;
}

template <bool BF16>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_wgrad_umma_kernel(const UmmaWgradParams p, const __grid_constant__ CUtensorMap tmap_dout) {
  using PR = Prec<BF16>;
  constexpr int kRows = kWgChunkBlock / PR::kRowBytes;      // 64 (tf32) / 128 (bf16) rows per step
  constexpr int kMmaPerStep = kRows / (BF16 ? 16 : 8);      // 8 either way, 1024 B of rows each
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = (p.Cout / 32) * kWgChunkBlock;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)p.a_stages * kWgAStage;
  const uint32_t bar_base = b_base + (uint32_t)kWgBStages * b_stage_bytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kWgMaxAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kWgMaxAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kWgMaxAStages + kWgBStages + s); };
  const uint32_t t_full = bar_base + 8u * (2 * kWgMaxAStages + 2 * kWgBStages);
  const uint32_t t_empty = t_full + 8u;
  const uint32_t tmem_slot = t_full + 16u;
  const uint32_t idx_base = bar_base + 256u;
  // lookup tables (no integer division inside the per-step loops: every role is a single warp per
  // scheduler, so long dependent instruction chains cost their full latency)
  uint32_t* s_mbtaps = reinterpret_cast<uint32_t*>(smem_raw + (idx_base + kWgIdxBytes - smem_u32(smem_raw)));
  uint8_t* s_qk = reinterpret_cast<uint8_t*>(s_mbtaps + kWgMaxMb);
  uint8_t* s_qcc = s_qk + kWgMaxMb * 4;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // full barriers: the 32 lanes of the one warp that fills the stage (cp.async ... arrive.noinc)
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full(s), 32); mbar_init(a_empty(s), 1); }
    // b_full: one expect_tx arrival (TMA tile loads of the dout rows), or the lanes of the B warps (LDGSTS)
    for (int s = 0; s < kWgBStages; ++s) { mbar_init(b_full(s), p.b_tma ? 1 : 32 * p.b_warps); mbar_init(b_empty(s), 1); }
    mbar_init(t_full, 1);
    mbar_init(t_empty, kNumEpilogueThreads);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512u); tmem_relinquish(); }
  for (int mb = threadIdx.x; mb < p.n_mb; mb += blockDim.x) s_mbtaps[mb] = mblock_taps(mb, p.ncc, p.nq);
  for (int q = threadIdx.x; q < p.n_mb * 4; q += blockDim.x) {
    s_qk[q] = (uint8_t)(q < p.nq ? q / p.ncc : 255);
    s_qcc[q] = (uint8_t)(q % p.ncc);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t all_taps = p.K >= 32 ? 0xFFFFFFFFu : ((1u << p.K) - 1u);
  long long wg_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long wg_t0 = clock64();

  // (row block -> offset mask); tile masks are per 128 rows
  auto rb_mask = [&](int rb) -> uint32_t {
    return p.tile_mask ? (p.tile_mask[(rb * kRows) >> 7] & all_taps) : all_taps;
  };

  if (warp < kNumProducerWarps) {
    // ============================ producers ============================
    // One warp per pipeline stage (see conv_umma_kernel): A stage number n (the n-th active (work item,
    // row block, M block) step of this CTA) is gathered entirely by warp n mod a_stages, which always
    // fills ring slot = its own rank; the dout rows of the m-th active row block are loaded by B warp
    // m mod 2 (warps a_stages, a_stages + 1).  Every warp walks the same step sequence and acts on its own.
    constexpr int R = PR::kRowsPerInstr;       // rows one LDGSTS instruction covers
    constexpr int NQ = kRows / R;              // instructions per [kRows x 32-channel] chunk block (16)
    constexpr int NI = kRows / 32;             // neighbour indices per lane and chunk (rows lane + 32 i)
    const int sub = lane / PR::kLanesPerRow;   // row within the rows one instruction covers
    const int j = lane % PR::kLanesPerRow;     // 16-byte piece within the row chunk
    const char* in_base = reinterpret_cast<const char*>(p.in) + j * 16;
    const char* dout_base = reinterpret_cast<const char*>(p.dout);
    const size_t in_pitch = (size_t)p.Cin * PR::kElt;

    // flat iterator over the active (work item, row block, M block) steps of this CTA; `first` marks the
    // first active M block of a row block (= one B stage)
    struct Step { int w, rb, mb, rb1, mb0, mb1; uint32_t mask; bool ok, first; };
    auto begin_work = [&](Step& s) {
      const int split = s.w / p.n_pass, pass = s.w - split * p.n_pass;
      s.mb0 = pass * p.n_mb / p.n_pass;
      s.mb1 = (pass + 1) * p.n_mb / p.n_pass;
      s.rb = split * p.rb_per_split - 1;
      s.rb1 = min((split + 1) * p.rb_per_split, p.n_rb);
      s.mb = s.mb1;  // forces the first next() onto (rb0, mb0)
      s.mask = 0;
    };
    auto next = [&](Step& s) {
      s.first = false;
      for (;;) {
        if (++s.mb >= s.mb1) {
          s.mb = s.mb0;
          if (++s.rb >= s.rb1) {
            s.w += gridDim.x;
            if (s.w >= p.n_work) { s.ok = false; return; }
            begin_work(s);
            continue;
          }
          s.mask = rb_mask(s.rb);
          s.first = true;  // stays set until an active M block of this row block is found
        }
        if (s_mbtaps[s.mb] & s.mask) return;
      }
    };
    Step cur;
    cur.w = blockIdx.x; cur.ok = cur.w < p.n_work; cur.first = false;
    if (cur.ok) {
      begin_work(cur);
      // first step: `first` must survive skipped M blocks, which next() guarantees (it only clears it on entry)
      next(cur);
    }

    if (warp < p.a_stages) {
      // ---- A stages: gathered input rows of the four (offset, channel group) chunks of one M block ----
      auto load_idx = [&](const Step& st, int* idx) {
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          const int q = st.mb * 4 + sl;
          const int k = q < p.nq ? (int)s_qk[q] : -1;
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int o = st.rb * kRows + i * 32 + lane;
            idx[sl * NI + i] = (k >= 0 && o < p.m_out) ? __ldg(p.nbr + (size_t)k * p.m_out + o) : -1;
          }
        }
      };
      // advance to this warp's first owned step (step number == warp)
      for (int skip = 0; skip < warp && cur.ok; ++skip) next(cur);
      int idx[4 * NI];
      if (cur.ok) load_idx(cur, idx);
      uint32_t a_phase = 0;
      const uint32_t stage_addr = a_base + (uint32_t)warp * kWgAStage;
      while (cur.ok) {
        Step nxt = cur;
        for (int skip = 0; skip < p.a_stages && nxt.ok; ++skip) next(nxt);
        int idx_n[4 * NI];
        if (nxt.ok) load_idx(nxt, idx_n);  // latency hides behind this stage's slot wait

        WG_TIMED_WAIT(0, mbar_wait(a_empty(warp), a_phase ^ 1u));
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          const int q = cur.mb * 4 + sl;
          if (q < p.nq && !p.dbg_skip_gather) {  // (padding chunk of the last M block: rows stay as they are)
            const char* src_c = in_base + (size_t)s_qcc[q] * PR::kRowBytes;
            const uint32_t dst_c = stage_addr + sl * kWgChunkBlock;
#pragma unroll
            for (int qi = 0; qi < NQ; ++qi) {
              const int r = qi * R + sub;  // row within the row block
              const int src_row = __shfl_sync(0xffffffffu, idx[sl * NI + ((qi * R) >> 5)], r & 31);
              const char* src = src_c + (size_t)(src_row >= 0 ? src_row : 0) * in_pitch;
              cp_async_16(dst_c + r * PR::kRowBytes + PR::swz_mn(j, r), src, src_row >= 0 ? 16u : 0u);
            }
          }
        }
        cp_async_mbar_arrive_noinc(a_full(warp));
        a_phase ^= 1u;
        cur = nxt;
#pragma unroll
        for (int i = 0; i < 4 * NI; ++i) idx[i] = idx_n[i];
      }
    } else if (warp < p.a_stages + kWgBStages * p.b_warps) {
      // ---- B stages: the dout rows of one row block (contiguous rows, all Cout channels), b_warps warps
      // per ring slot (wide Cout: a B stage is as large as two A stages) ----
      const int bi = warp - p.a_stages;
      const int bw = bi / p.b_warps, part = bi - bw * p.b_warps;  // ring slot / share of its pieces
      uint32_t b_phase = 0;
      int m = 0;  // number of the active row block
      while (cur.ok) {
        if (cur.first) {
          if ((m & 1) == bw) {
            const uint32_t dstb = b_base + (uint32_t)bw * b_stage_bytes;
            const int o0 = cur.rb * kRows;
            if (p.b_tma) {
              // contiguous rows: Cout / 32 tile loads (32 channels x kRows rows, hardware swizzle = the MN-major
              // UMMA layout) issued by one lane; rows past m_out are zero-filled by the TMA unit
              if (lane == 0) {
                WG_TIMED_WAIT(1, mbar_wait(b_empty(bw), b_phase ^ 1u));
                mbar_arrive_expect_tx(b_full(bw), (uint32_t)b_stage_bytes);
                for (int cbk = 0; cbk < p.Cout / 32 && !(p.dbg_skip & 2); ++cbk)
                  tma_load_2d(dstb + cbk * kWgChunkBlock, &tmap_dout, b_full(bw), cbk * 32, o0);
              }
              __syncwarp();
            } else {
              WG_TIMED_WAIT(1, mbar_wait(b_empty(bw), b_phase ^ 1u));
              const int n16 = (p.Cout / 32) * kRows * PR::kLanesPerRow;  // 16-byte pieces of the dout block
              for (int e = part * 32 + lane; e < n16 && !(p.dbg_skip & 2); e += 32 * p.b_warps) {
                const int jj = e % PR::kLanesPerRow, r = (e / PR::kLanesPerRow) % kRows, cbk = e / (PR::kLanesPerRow * kRows);
                const int o = o0 + r;
                const bool ok = o < p.m_out;
                const char* src = dout_base + ((size_t)(ok ? o : 0) * p.Cout + cbk * 32) * PR::kElt + jj * 16;
                cp_async_16(dstb + cbk * kWgChunkBlock + r * PR::kRowBytes + PR::swz_mn(jj, r), src, ok ? 16u : 0u);
              }
              cp_async_mbar_arrive_noinc(b_full(bw));
            }
            b_phase ^= 1u;
          }
          ++m;
        }
        next(cur);
      }
    }
    cp_async_wait<0>();
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ============================
    // one elected thread runs the whole loop (see conv_umma_kernel)
    if (elect_one()) {
    int a_stage = 0, b_stage = 0;
    uint32_t a_phase = 0, b_phase = 0, t_phase = 0;
    const uint32_t idesc = PR::idesc(128, (uint32_t)p.Cout, 1, 1);  // both operands MN-major
    const uint64_t desc_hi = make_desc(0, kWgChunkBlock, 512, PR::kLayoutMN);
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int split = w / p.n_pass, pass = w - split * p.n_pass;
      const int mb0 = pass * p.n_mb / p.n_pass, mb1 = (pass + 1) * p.n_mb / p.n_pass;  // balanced
      const int rb0 = split * p.rb_per_split, rb1 = min(rb0 + p.rb_per_split, p.n_rb);
      WG_TIMED_WAIT(5, mbar_wait(t_empty, t_phase ^ 1u));  // epilogue drained the previous item's accumulators
      tc_fence_after();
      uint32_t touched = 0;
      uint32_t mask_next = rb0 < rb1 ? rb_mask(rb0) : 0u;
      for (int rb = rb0; rb < rb1; ++rb) {
        const uint32_t mask = mask_next;
        mask_next = rb + 1 < rb1 ? rb_mask(rb + 1) : 0u;  // prefetched: hidden behind this row block
        bool b_ready = false;
        int b_used = -1;
        for (int mb = mb0; mb < mb1; ++mb) {
          if (!(s_mbtaps[mb] & mask)) continue;
          if (!b_ready) {
            WG_TIMED_WAIT(4, mbar_wait(b_full(b_stage), b_phase));
            b_ready = true;
            b_used = b_stage;
          }
          WG_TIMED_WAIT(3, mbar_wait(a_full(a_stage), a_phase));
          fence_proxy_async_smem();  // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
          tc_fence_after();
          {
            const uint32_t a_addr = a_base + (uint32_t)a_stage * kWgAStage;
            const uint32_t b_addr = b_base + (uint32_t)b_used * b_stage_bytes;
            const uint32_t d = tmem_base + (uint32_t)((mb - mb0) * p.Cout);
            const uint32_t was = (touched >> (mb - mb0)) & 1u;
#pragma unroll
            for (int r8 = 0; r8 < kMmaPerStep; ++r8) {
              // descriptors differ only in the start address field: 1024 B of rows per MMA = 64 units
              const uint64_t adesc = desc_hi | (uint64_t)(((a_addr >> 4) + 64u * r8) & 0x3FFFu);
              const uint64_t bdesc = desc_hi | (uint64_t)(((b_addr >> 4) + 64u * r8) & 0x3FFFu);
              if (!p.dbg_skip_mma) PR::mma(d, adesc, bdesc, idesc, (was || r8 > 0) ? 1u : 0u);
            }
            mma_commit(a_empty(a_stage));
          }
          touched |= 1u << (mb - mb0);
          if (++a_stage == p.a_stages) { a_stage = 0; a_phase ^= 1u; }
        }
        if (b_ready) {
          mma_commit(b_empty(b_used));
          if (++b_stage == kWgBStages) { b_stage = 0; b_phase ^= 1u; }
        }
      }
      mma_commit(t_full);
      t_phase ^= 1u;
    }
    }
    __syncwarp();
  } else {
    // ============================ epilogue ============================
    const int ew = warp & 3;
    uint32_t t_phase = 0;
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int split = w / p.n_pass, pass = w - split * p.n_pass;
      const int mb0 = pass * p.n_mb / p.n_pass, mb1 = (pass + 1) * p.n_mb / p.n_pass;  // balanced
      const int rb0 = split * p.rb_per_split, rb1 = min(rb0 + p.rb_per_split, p.n_rb);
      // which accumulators received at least one MMA (same rule as the issuer)
      uint32_t seen = 0;
      for (int rb = rb0 + lane; rb < rb1; rb += 32) seen |= rb_mask(rb);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, d);
      mbar_wait_sleep(t_full, t_phase);
      tc_fence_after();
      for (int mb = mb0; mb < mb1; ++mb) {
        if (!(s_mbtaps[mb] & seen)) continue;
        const int q = mb * 4 + ew;
        if (q >= p.nq) continue;  // padding chunk of the last M block (warp-uniform)
        const int k = s_qk[q], cc = s_qcc[q];
        if (!((seen >> k) & 1u)) continue;  // this offset has no pair in the row range: stays zero
        float* dst = p.dw + ((size_t)k * p.Cin + cc * 32 + lane) * p.Cout;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((mb - mb0) * p.Cout);
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (p.dbg_skip & 4) continue;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i), "f"(v[i]),
                         "f"(v[i + 1]), "f"(v[i + 2]), "f"(v[i + 3])
                         : "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(t_empty);
      t_phase ^= 1u;
    }
  }
  if (warp == kMmaWarp) {  // the counters live in the elected lane: make them visible to lane 0
#pragma unroll
    for (int i = 3; i < 6; ++i)
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, wg_cnt[i], d);
        wg_cnt[i] = o > wg_cnt[i] ? o : wg_cnt[i];
      }
  }
  if (blockIdx.x < kNumSMs && (threadIdx.x == 0 || threadIdx.x == kMmaWarp * 32)) {
    const int base = threadIdx.x == 0 ? 0 : 3, n = 3;
    for (int i = 0; i < n; ++i) g_wg_counters[blockIdx.x * 8 + base + i] = wg_cnt[base + i];
    g_wg_counters[blockIdx.x * 8 + (threadIdx.x == 0 ? 6 : 7)] = clock64() - wg_t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

bool umma_wgrad_supported(int c_in, int c_out) {
  return c_in >= 32 && c_in % 32 == 0 && c_out >= 32 && c_out % 32 == 0 && c_out <= 256;
}
int64_t umma_wgrad_workspace(int, int, int) { return 256; }

// [rows, C] row-major tensor, box = 32 channels x box_rows rows, written in the MN-major UMMA layout of the
// wgrad operands: bf16 SWIZZLE_64B, fp32 SWIZZLE_128B with 32-byte atoms
static bool make_rows_tile_map(CUtensorMap* map, const void* base, int64_t rows, int C, int box_rows, bool bf16) {
  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)C * (bf16 ? 2 : 4)};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
             const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool BF16>
static int launch_wgrad_umma(const UmmaWgradParams& p, const CUtensorMap& tmap, int grid, size_t smem,
                             cudaStream_t stream) {
  auto kern = conv_wgrad_umma_kernel<BF16>;
  SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kNumThreads, smem, stream>>>(p, tmap);
  SPC_LAUNCHED("conv_wgrad_umma_kernel");
  return 0;
}

int conv_wgrad_umma(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask,
                    int64_t m_out, int c_in, int c_out, int K, bool bf16, float* dw, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream) {
  (void)workspace; (void)workspace_bytes;
  SPC_REQUIRE(umma_wgrad_supported(c_in, c_out), "shape not supported by the tcgen05 wgrad path");
  SPC_REQUIRE(K <= 32, "tcgen05 path supports kernel volume <= 32");
  SPC_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dw % 16) == 0,
              "rows must be 16-byte aligned");
  SPC_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (m_out == 0) return 0;
  const int rows = bf16 ? 128 : 64;  // out rows per pipeline step
  UmmaWgradParams p;
  p.in = in; p.dout = dout; p.nbr = nbr; p.tile_mask = tile_mask; p.dw = dw;
  p.m_out = (int)m_out; p.Cin = c_in; p.Cout = c_out; p.K = K;
  p.ncc = c_in / 32;
  p.nq = K * p.ncc;
  p.n_mb = (p.nq + 3) / 4;
  SPC_REQUIRE(p.n_mb <= kWgMaxMb, "too many M blocks");
  int cap = 512 / c_out;                       // accumulators that fit in TMEM
  p.n_pass = (p.n_mb + cap - 1) / cap;
  p.mb_per_pass = (p.n_mb + p.n_pass - 1) / p.n_pass;
  p.n_rb = (int)ceil_div(m_out, rows);
  int want_split = (2 * kNumSMs) / p.n_pass;  // <= 2 work items per CTA (static round-robin)
  if (want_split > p.n_rb) want_split = p.n_rb;
  if (want_split < 1) want_split = 1;
  p.rb_per_split = (p.n_rb + want_split - 1) / want_split;
  if (!bf16 && (p.rb_per_split & 1)) ++p.rb_per_split;  // keep splits aligned to 128-row mask tiles
  p.n_split = (p.n_rb + p.rb_per_split - 1) / p.rb_per_split;
  p.n_work = p.n_split * p.n_pass;
  const int b_stage_bytes = (c_out / 32) * kWgChunkBlock;
  int a_stages = (kSmemLimit - 1024 - 256 - kWgIdxBytes - kWgTabBytes - kWgBStages * b_stage_bytes) / kWgAStage;
  if (a_stages > kWgMaxAStages) a_stages = kWgMaxAStages;  // one producer warp per A stage + two B warps <= 8 warps
  SPC_REQUIRE(a_stages >= 2, "wgrad tile does not fit in shared memory");
  p.a_stages = a_stages;
  p.b_warps = (kNumProducerWarps - a_stages) / kWgBStages >= 2 && c_out >= 128 ? 2 : 1;
  p.dbg_skip_mma = g_dbg[4];
  p.dbg_skip_gather = g_dbg[5];
  p.dbg_skip = g_dbg[6];
  const size_t smem = (size_t)a_stages * kWgAStage + (size_t)kWgBStages * b_stage_bytes + 1024 + 256 + kWgIdxBytes + kWgTabBytes;
  const int grid = p.n_work < kNumSMs ? p.n_work : kNumSMs;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  p.b_tma = (g_dbg[1] != 1 && make_rows_tile_map(&tmap, dout, m_out, c_out, rows, bf16)) ? 1 : 0;
  if (p.b_tma) p.b_warps = 1;  // one issuing lane per B ring slot
  return bf16 ? launch_wgrad_umma<true>(p, tmap, grid, smem, stream) : launch_wgrad_umma<false>(p, tmap, grid, smem, stream);
}

}  // namespace spc
