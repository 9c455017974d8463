// umma_common.cuh — pieces shared by the tcgen05 convolution kernels (conv_umma.cu: forward / dgrad,
// conv_wgrad_umma.cu: wgrad): role layout of the 416-thread CTA, operand-precision traits, tensor-map helpers.
#pragma once
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

template <bool BF16>
struct Prec {
  static constexpr int kElt = BF16 ? 2 : 4;             // bytes per element
  static constexpr int kRowBytes = 32 * kElt;           // one 32-channel chunk of a row
  static constexpr int kLanesPerRow = kRowBytes / 16;   // 16-byte cp.async pieces per row chunk
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
  __device__ static __forceinline__ void mma_p(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc,
                                               uint32_t issue) {
    if (BF16) mma_bf16_p(d, a, b, id, acc, issue);
    else mma_tf32_p(d, a, b, id, acc, issue);
  }
};

// ---- CTA pairs (tcgen05 cta_group::2; conv_umma_pair.cu, conv_wgrad_umma_pair.cu) ----------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// all threads of both CTAs (also orders shared-memory / mbarrier initialisation across the pair)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// (relaxed: a release at cluster scope is a MEMBAR.ALL.GPU in front of every arrival — the first version of this
// kernel ran 4x slower than the single-CTA kernel with cluster-scope acquires / releases and an unqualified
// fence.proxy.async in the per-stage paths: L1 invalidations and GPU-wide membars in the producer and relay warps.  What
// has to be ordered here is shared memory against the async proxy (fence.proxy.async.shared::cta in front of the
// arrival) and tcgen05 operations (tcgen05.fence), not generic global memory.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows of each CTA] * B[N/2 rows of each CTA], bf16 operands, issue predicate as in
// ptx.cuh (mma_bf16_p)
__device__ __forceinline__ void mma_bf16_pair_p(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once the tcgen05.mma issued so far have completed
__device__ __forceinline__ void mma_commit_pair_p(uint32_t bar, uint32_t issue) {
  const uint16_t both = 3;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}" ::"r"(bar),
      "r"(issue), "h"(both)
      : "memory");
}


// test / measurement knobs (spc_debug_set), all 0 = default:
//   [0] forward: force the number of producer groups (1, 2, 4, 8)      [1] wgrad: dout rows by LDGSTS instead of TMA
//   [2] forward: 1 = st.global epilogue instead of TMA stores          [3] forward: epilogue writes nothing (timing only)
//   [4] forward: 1 = one 32-channel chunk per stage (no contiguous multi-chunk row visits)
//   [5] wgrad: 1 = one row visit per chunk (no grouped visits)
//   [6] forward: producer warps 8 / 16 (0 = by tile width)             [7] forward: 1 = no shared ring slots (8 groups x 2 warps)
//   [8] forward: 1 = never the CTA-pair kernel (conv_umma_pair.cu), 2 = always where it is eligible
//   [9] wgrad: the same for conv_wgrad_umma_pair.cu
extern int g_umma_dbg[16];
extern int g_umma_force_mt;

struct UmmaConvParams {
  const void* A;             // [*, Ck] fp32 or bf16
  const void* Bp;            // packed weights [K][Ck/32][Cn][32] (swizzled rows)
  const float* bias;         // [Cn] or null
  const int* nbr;            // [K, m_out]
  const uint32_t* tile_mask; // [ceil(m_out/128)] or null (all offsets active)
  float* out;                // [m_out, Cn]
  int m_out, Ck, Cn, K;
  int cn_tile, n_ntiles, kc_count, kg_count;  // 32-channel chunks / chunk groups (stages) per offset
  int stages, acc_bufs, tmem_cols;
  int out_bufs;              // 16 KB staging blocks of the TMA-store epilogue (0: st.global epilogue)
  int ngroups, wps;          // producer groups, warps per group
  int dbg_skip_store;        // timing experiment only: epilogue does not write the output
  int dbg_flags;             // timing experiments (-DSPC_EXPERIMENTS): 1 no gather copies, 2 no MMAs, 4 no weight slabs
  int ksplit, k_per;         // offsets split over ksplit work items of k_per offsets each (small maps)
  int reduce_out;            // the epilogue ADDS to `out` (offset-split items, or accumulate: out += result)
  double* stats;             // [2 * Cn] per-column sum / sum of squares of `out` for the BatchNorm that follows, or null
  int n_work;                // m_tiles * n_ntiles * ksplit
};


// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encoder();
// fp32 [rows, C] row-major output, box = 32 columns x 128 rows, SWIZZLE_128B: the forward epilogue's store target
bool make_out_tile_map(CUtensorMap* map, float* base, int64_t rows, int C);
// [rows, C] row-major tensor, box = 32 channels x box_rows rows, written in the MN-major UMMA layout of the
// wgrad operands: bf16 SWIZZLE_64B, fp32 SWIZZLE_128B with 32-byte atoms
bool make_rows_tile_map(CUtensorMap* map, const void* base, int64_t rows, int C, int box_rows, bool bf16);

}  // namespace spc
