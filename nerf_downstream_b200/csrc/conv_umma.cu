// conv_umma.cu — sparse convolution forward / dgrad on the 5th-generation tensor cores (tcgen05.mma,
// accumulators in TMEM) as an output-stationary implicit GEMM:
//
//   out[o, :] = sum_k  A[nbr[k, o], :] @ B_k        A: [*, Ck] rows, B_k: [Ck, Cn]
//   (dgrad = the same kernel on the transposed map with W^T; wgrad: conv_wgrad_umma.cu)
//
// Two operand precisions share the code (template BF16):
//   false  kind::tf32 — fp32 rows in shared memory (the tensor core uses the upper 19 bits),
//          32 channels = 128-byte swizzle rows, 4 MMAs (K=8) per row chunk;
//   true   kind::f16 with bf16 operands — rows converted once per layer (spc_to_bf16 / BatchNorm side copies),
//          32 channels = 64-byte rows (SWIZZLE_64B), 2 MMAs (K=16) per row chunk: half the gather bytes and
//          twice the MMA rate.  Accumulation is fp32 in TMEM in both.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-7   producers : a pipeline stage is one (kernel offset k, 32-channel chunk) of MT x 128 output rows.
//                           The rows are gathered with 16-byte cp.async (LDGSTS, zero-fill for missing
//                           neighbours) straight into the swizzled UMMA layout.  The producer warps form
//                           `ngroups` groups of WPS warps; group g fills stages g, g + ngroups, ... (each warp a
//                           slice of the stage's rows), so `ngroups` stages are in flight and the per-stage fixed
//                           costs (slot wait, index prefetch, address arithmetic) run concurrently: 8 groups of
//                           one warp when 8 ring slots fit, 4 x 2 or 2 x 4 for larger stages.  The stage's weights
//                           arrive as a TMA bulk copy (UBLKCP) of a pre-swizzled slab.  "Full" barriers are
//                           signalled by the hardware (cp.async.mbarrier.arrive.noinc): no wait_group / proxy
//                           fence on the producer side, it only waits for slots.  (Template G > 1 = stages of G
//                           chunks with one contiguous G x 64-byte row visit: built, measured slower, compiled
//                           only with -DSPC_EXPERIMENTS.)
//   warps 8-11  epilogue  : tcgen05.ld the fp32 accumulator (32 TMEM lanes per warp), bias, 16-byte
//                           conflict-free stores into 128B-swizzled 16 KB staging blocks, TMA tensor store
//                           (reduce-add on offset-split items) — no st.global, the LSU belongs to the gather;
//   warp  12    MMA       : one ELECTED thread (elect.sync) runs the whole issue loop: tcgen05.mma from
//                           uniform registers, tcgen05.commit for stage release / accumulator completion.
// A work item is MT x 128 output rows x cn_tile channels (MT accumulators share every weight slab), on small
// maps x one group of kernel offsets; accumulators are double-buffered in TMEM when they fit in 512 columns so
// the epilogue of tile i overlaps the main loop of tile i+1.  Offsets with no neighbour inside a tile are
// skipped using the per-tile offset mask (spc_tile_mask).  On large maps output rows are owned by one CTA: no
// atomics in forward / dgrad.
#include "umma_common.cuh"

namespace spc {

int g_umma_dbg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#ifdef SPC_EXPERIMENTS
// role timing of the forward kernel (clock64 sums over all CTAs): [0] producer warp 0 loop, [1] its empty-slot waits,
// [2] its copy issue, [3] its bookkeeping + index prefetch, [4] MMA thread loop, [5] its full-stage waits, [6] its
// accumulator waits, [7] stages, [8] epilogue warp 0 accumulator waits, [9] its body
__device__ unsigned long long g_role_cycles[16];
#define SPC_T(var) const long long var = clock64()
#define SPC_ACC(i, v) atomicAdd(&g_role_cycles[i], (unsigned long long)(v))
#else
#define SPC_T(var)
#define SPC_ACC(i, v)
#endif
int g_umma_force_mt = 0;  // test hook: 0 = auto
bool conv_umma_pair_eligible(int64_t m_out, int c_in, int c_out, int K, bool bf16, const float* bias, const void* packed,
                             const float* out);
int conv_fwd_umma_pair(const void* in, const void* packed, const int* nbr, const uint32_t* tile_mask, int64_t m_out,
                       int c_in, int c_out, int K, float* out, cudaStream_t stream, bool accumulate, double* stats,
                       int* stats_fused);

static inline int pick_cn_tile(int Cn) {
  for (int t = 256; t >= 16; t -= 16)
    if (Cn % t == 0) return t;
  return 0;
}

bool umma_fwd_supported(int c_in, int c_out) {
  return c_in >= 32 && c_in % 32 == 0 && c_out % 16 == 0 && pick_cn_tile(c_out) >= 16;
}
int64_t umma_packed_bytes(int K, int c_in, int c_out) { return align_up((int64_t)K * c_in * c_out * 4, 1024); }
int64_t umma_fwd_workspace(int K, int c_in, int c_out) { return umma_packed_bytes(K, c_in, c_out) + 1024; }

// W [K][Ck][Cn] (or [K][Cn][Ck] when transposed) -> per (k, 32-channel chunk) a slab of Cn rows x one row
// chunk, 16-byte pieces XOR-swizzled by the row: the exact shared-memory image UMMA expects for a K-major
// swizzled B operand.  An n tile of cn_tile rows (a multiple of 16, so the swizzle phase is the same) is a
// contiguous piece of the slab, and the G chunks of a stage are G such pieces (one bulk copy when the tile
// spans all Cn): the layout does not depend on the tile shape, so a layer packs its weights ONCE per
// optimiser step for every launch that uses them.
template <bool BF16>
__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ W, void* __restrict__ Wp_, int K, int Ck, int Cn, int flags) {
  // flags: bit 0 = W is [K][Cn][Ck] (dgrad contracts over Cout with W^T); bit 1 = slab k holds W[K-1-k]: on a centrally
  // symmetric self map nbr_t[k] == nbr[K-1-k], so dgrad reads the forward map with the offsets reversed
  const int transpose = flags & 1, reverse = flags & 2;
  const long long total = (long long)K * Ck * Cn;
  const int kc_count = Ck / 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int jj = (int)(e % 32);
    long long t = e / 32;
    int n = (int)(t % Cn); t /= Cn;
    int kc = (int)(t % kc_count);
    int k = (int)(t / kc_count);
    int c = kc * 32 + jj;
    const int ks = reverse ? K - 1 - k : k;
    float v = transpose ? W[((long long)ks * Cn + n) * Ck + c] : W[((long long)ks * Ck + c) * Cn + n];
    long long row = ((long long)k * kc_count + kc) * Cn + n;
    if (BF16) {
      int j = jj >> 3, w = jj & 7;  // 8 bf16 per 16-byte piece
      reinterpret_cast<__nv_bfloat16*>(Wp_)[row * 32 + ((j ^ ((n >> 1) & 3)) << 3) + w] = __float2bfloat16_rn(v);
    } else {
      int j = jj >> 2, w = jj & 3;  // 4 fp32 per 16-byte piece
      reinterpret_cast<float*>(Wp_)[row * 32 + ((j ^ (n & 7)) << 2) + w] = v;
    }
  }
}

// One launch for every layer of a network (blockIdx.y = layer): the descriptors live in device memory
// ([n][8] int64: w, packed, K, Ck, Cn, flags (bit 0 transpose, bit 1 reversed offsets), bf16, unused) and are rebuilt only when the set of layers changes.
__global__ void __launch_bounds__(256)
pack_weights_batch_kernel(const long long* __restrict__ desc) {
  const long long* d = desc + (size_t)blockIdx.y * 8;
  const float* W = reinterpret_cast<const float*>(d[0]);
  void* Wp_ = reinterpret_cast<void*>(d[1]);
  const int K = (int)d[2], Ck = (int)d[3], Cn = (int)d[4], transpose = (int)d[5] & 1, reverse = (int)d[5] & 2, bf16 = (int)d[6];
  const long long total = (long long)K * Ck * Cn;
  const int kc_count = Ck / 32;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int jj = (int)(e % 32);
    long long t = e / 32;
    int n = (int)(t % Cn); t /= Cn;
    int kc = (int)(t % kc_count);
    int k = (int)(t / kc_count);
    int c = kc * 32 + jj;
    const int ks = reverse ? K - 1 - k : k;
    float v = transpose ? W[((long long)ks * Cn + n) * Ck + c] : W[((long long)ks * Ck + c) * Cn + n];
    long long row = ((long long)k * kc_count + kc) * Cn + n;
    if (bf16) {
      int j = jj >> 3, w = jj & 7;
      reinterpret_cast<__nv_bfloat16*>(Wp_)[row * 32 + ((j ^ ((n >> 1) & 3)) << 3) + w] = __float2bfloat16_rn(v);
    } else {
      int j = jj >> 2, w = jj & 3;
      reinterpret_cast<float*>(Wp_)[row * 32 + ((j ^ (n & 7)) << 2) + w] = v;
    }
  }
}

int conv_pack_weights_batch(const long long* desc_dev, int n_layers, cudaStream_t stream) {
  if (n_layers <= 0) return 0;
  SPC_REQUIRE(n_layers <= 65535, "too many layers for one launch");
  dim3 grid(96, (unsigned)n_layers);  // 96 x 256 threads per layer: a 27 x 256 x 256 kernel takes 72 iterations
  pack_weights_batch_kernel<<<grid, 256, 0, stream>>>(desc_dev);
  SPC_LAUNCHED("pack_weights_batch_kernel");
  return 0;
}

int conv_pack_weights(const float* w, int K, int Ck, int Cn, int flags, bool bf16, void* packed,
                      cudaStream_t stream) {
  SPC_REQUIRE(((uintptr_t)packed % 1024) == 0, "packed weights must be 1024-byte aligned");
  long long total = (long long)K * Ck * Cn;
  int grid = (int)std::min<long long>(ceil_div(total, 256), kNumSMs * 8);
  if (bf16) pack_weights_kernel<true><<<grid, 256, 0, stream>>>(w, packed, K, Ck, Cn, flags);
  else pack_weights_kernel<false><<<grid, 256, 0, stream>>>(w, packed, K, Ck, Cn, flags);
  SPC_LAUNCHED("pack_weights_kernel");
  return 0;
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

// NPW = producer warps (8 or 16).  The producers are bound by instruction latency, not by bytes: a stage costs a warp
// ~450 dependent instructions (index shuffle, address arithmetic, LDGSTS) and with 8 producer warps a scheduler holds
// two of them (0.37 IPC per scheduler, 330 clocks per stage with neither copies nor MMAs issued —
// profiles/r2_conv_breakdown.md); 16 warps split every stage in two halves and give each scheduler four.
constexpr int kFwdMaxThreads = (16 + 4 + 1) * 32;
template <int MT, bool BF16, int G, int WPS, int NPW>
__global__ void __launch_bounds__(kFwdMaxThreads, 1)
conv_umma_kernel(const UmmaConvParams p, const __grid_constant__ CUtensorMap tmap_out) {
  using PR = Prec<BF16>;
  constexpr int kNumProducerWarps = NPW;      // (shadow the 8-producer layout of umma_common.cuh)
  constexpr int kMmaWarp = NPW + 4;
  constexpr int kAStage = kTileM * PR::kRowBytes;  // one 32-channel chunk of one 128-row sub-tile
  extern __shared__ uint8_t smem_raw[];
  // swizzled operands need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_chunk_bytes = p.cn_tile * PR::kRowBytes;
  const int stage_bytes = G * (MT * kAStage + b_chunk_bytes);
  // output staging of the TMA-store epilogue: out_bufs blocks of [128 rows x 32 fp32], SWIZZLE_128B
  const uint32_t out_stage = smem_base + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = out_stage + (uint32_t)p.out_bufs * 16384u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kMaxStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 4);
  auto turn_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 5 + s); };  // (shared ring slots, see producers)
  // column sums of the output tiles this CTA writes (p.stats): [4 epilogue warps][2 * cn_tile] doubles behind the
  // barrier area — every epilogue thread owns its slots (plain loads / stores, no atomics)
  double* s_stats = reinterpret_cast<double*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // (REDUX: the warp index as a value ptxas KNOWS to be warp-uniform, so that the role branches below are uniform
  // control flow and the MMA warp's loop state can live in uniform registers)
  const int warp = (int)__reduce_min_sync(0xffffffffu, threadIdx.x >> 5), lane = threadIdx.x & 31;

  if (p.stats != nullptr)
    for (int c = threadIdx.x; c < 8 * p.cn_tile; c += blockDim.x) s_stats[c] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
#ifdef SPC_EXPERIMENTS
      if (p.dbg_flags & 128) mbar_init(full_bar(s), WPS + 1); else   // timing experiment: one arrival per warp
#endif
      mbar_init(full_bar(s), 32 * WPS + 1);             // the lanes of the group's warps + 1 expect_tx
      mbar_init(empty_bar(s), 1);                       // one tcgen05.commit
      mbar_init(turn_bar(s), 1);
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
  const uint32_t tmem_base = __reduce_min_sync(0xffffffffu, *tmem_slot_ptr);   // (REDUX: a provably uniform value)

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
    // Stage n of this CTA's stage sequence (work item, active offset k, chunk group cg) is filled by producer
    // group n % ngroups; inside a group warp slice `sl` gathers rows [sl * RW, (sl + 1) * RW) of the item's
    // MT x 128 rows.  A group waits for ITS ring slot, issues the stage's copies (slice 0 also the weight slabs)
    // and moves on.  ngroups <= ring slots: a group is then never two ring revolutions ahead of the MMA warp,
    // which the one-bit phase parity of the "empty" barriers could not tell apart.
    const int ngroups = p.ngroups;
    const int grp = warp / WPS, sl = warp - grp * WPS;
    constexpr int LPR = PR::kLanesPerRow;   // 16-byte pieces per 32-channel chunk of a row
    constexpr int PPR = G * LPR;            // pieces per row visit; PPR instructions cover 32 rows
    constexpr int RW = kTileM * MT / WPS;   // rows per warp slice (multiple of 32)
    constexpr int NI = RW / 32;             // 32-row groups of a slice
    static_assert(RW >= 32 && RW % 32 == 0, "a warp slice holds whole 32-row groups");
    const char* Bbase = reinterpret_cast<const char*>(p.Bp);
    const bool leader = elect_one();

    // lane-constant pieces of the copy addresses: instruction q of a 32-row group covers flat pieces
    // q * 32 + lane -> (row rr of the group, piece of the row visit)
    int rr_tab[PPR];
    uint32_t dst_tab[PPR], src_tab[PPR];
#pragma unroll
    for (int q = 0; q < PPR; ++q) {
      const int f = q * 32 + lane;
      const int rr = f / PPR, piece = f - rr * PPR;
      const int g = piece / LPR, j = piece - g * LPR;
      rr_tab[q] = rr;
      dst_tab[q] = (uint32_t)(g * kAStage + rr * PR::kRowBytes) + PR::swz_k(j, rr);  // (group bases are multiples of 32 rows)
      src_tab[q] = (uint32_t)(piece * 16);
    }

    // Iterator over the stages this group owns.  Everything per STAGE is incremental — (offset, chunk group) move
    // by `ngroups` stages with compare / subtract on the item's remaining offset mask; the divisions (work item ->
    // m tile / n tile, the offset share of a split item) happen once per ITEM.  (r2: the first version located every
    // stage with four runtime integer divisions; with neither copies nor MMAs issued the kernel still took 0.375 of
    // its 0.57 ms at 96->96 on 1 M voxels — 330 clocks per stage of dependent integer bookkeeping in eight warps.)
    struct It { int w, mtile, ntile, cg; uint32_t rest; bool ok; };   // lowest set bit of `rest` = current offset
    auto open_item = [&](It& it, int skip) {  // stage `skip` (0-based) of item it.w, carried over into later items
      for (;;) {
        if (it.w >= p.n_work || grp >= ngroups) { it.ok = false; return; }
        uint32_t rest = work_mask(it.w);
        int cg = skip;
        while (rest != 0u && cg >= p.kg_count) { cg -= p.kg_count; rest &= rest - 1u; }
        if (rest != 0u) {
          it.rest = rest; it.cg = cg;
          it.mtile = it.w / items_per_mtile;
          it.ntile = (it.w / p.ksplit) % p.n_ntiles;
          return;
        }
        skip = cg;            // the item has fewer stages than that: the rest of the step lands in the next item
        it.w += gridDim.x;
      }
    };
    auto advance = [&](It& it) {
      int cg = it.cg + ngroups;
      uint32_t rest = it.rest;
      while (rest != 0u && cg >= p.kg_count) { cg -= p.kg_count; rest &= rest - 1u; }
      if (rest != 0u) { it.cg = cg; it.rest = rest; return; }
      it.w += gridDim.x;
      open_item(it, cg);
    };
    // lane l holds the neighbour row of row l of each 32-row group of this warp's slice
    auto load_idx = [&](const It& it, int* idx) {
      const int k = __ffs(it.rest) - 1;
      const int o0 = it.mtile * rows_per_work + sl * RW + lane;
      const int* row = p.nbr + (size_t)k * p.m_out;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int o = o0 + i * 32;
#ifdef SPC_EXPERIMENTS
        if (p.dbg_flags & 32) { idx[i] = o < p.m_out ? o : -1; continue; }
#endif
        idx[i] = o < p.m_out ? __ldg(row + o) : -1;
      }
    };
    const uint32_t row_pitch32 = (uint32_t)p.Ck * PR::kElt;

    It cur;
    cur.w = blockIdx.x; cur.ok = true;
    open_item(cur, grp);
    int idx[NI];
    if (cur.ok) load_idx(cur, idx);
    // ring slot and revolution of this group's stage sequence number n = grp + ngroups * i (slot = n % stages).
    // ngroups <= stages: a slot is always refilled by the group that filled it, which waits for the MMA warp to
    // release it.  ngroups > stages (16 one-warp groups on 8 slots: twice as many warps doing their bookkeeping and
    // index prefetch while others issue copies): two groups alternate on a slot.  One parity bit cannot tell "the
    // revolution before mine is not even filled" from "it is already released", so the groups hand the slot to each
    // other through a third barrier: a group arrives on turn_bar(slot) when it has issued its revolution, and waits
    // for the other group's arrival before it looks at the slot's "empty" barrier.
    int slot = grp % p.stages, rev = grp / p.stages;
    const bool shared_slots = ngroups > p.stages;
#ifdef SPC_EXPERIMENTS
    long long t_wait = 0, t_copy = 0, t_book = 0;
    const long long t_loop0 = clock64();
#endif
    while (cur.ok) {
      // indices of the NEXT owned stage: their latency hides behind this stage's slot wait
      SPC_T(ta);
      It nxt = cur;
      advance(nxt);
      int idx_n[NI];
      if (nxt.ok) load_idx(nxt, idx_n);

      const int k = __ffs(cur.rest) - 1, cg = cur.cg, ntile = cur.ntile;
      SPC_T(tb);
      const uint32_t par = (uint32_t)(rev - 1) & 1u;
      if (shared_slots) mbar_wait(turn_bar(slot), par);
      mbar_wait(empty_bar(slot), par);
      SPC_T(tc);
      const uint32_t stage_addr = smem_base + (uint32_t)slot * stage_bytes;
#ifdef SPC_EXPERIMENTS
      if (sl == 0 && leader && (p.dbg_flags & 4)) mbar_arrive(full_bar(slot));
      else
#endif
      if (sl == 0 && leader) {
        mbar_arrive_expect_tx(full_bar(slot), (uint32_t)(G * b_chunk_bytes));
        // slab rows [ntile * cn_tile, +cn_tile) of chunks cg * G .. cg * G + G - 1 of offset k
        const char* src = Bbase + (((size_t)k * p.kc_count + (size_t)cg * G) * p.Cn + (size_t)ntile * p.cn_tile) * PR::kRowBytes;
        const uint32_t dstb = stage_addr + G * MT * kAStage;
        if (p.n_ntiles == 1) {
          bulk_g2s(dstb, src, (uint32_t)(G * b_chunk_bytes), full_bar(slot));
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g)
            bulk_g2s(dstb + g * b_chunk_bytes, src + (size_t)g * p.Cn * PR::kRowBytes, (uint32_t)b_chunk_bytes, full_bar(slot));
        }
      }
      __syncwarp();
      // per-lane source bases of the PPR instruction slots of a 32-row group: row address = base + row * pitch
      const char* src_q[PPR];
#pragma unroll
      for (int q = 0; q < PPR; ++q)
        src_q[q] = reinterpret_cast<const char*>(p.A) + (size_t)cg * (G * PR::kRowBytes) + src_tab[q];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int R0 = sl * RW + i * 32;  // first row of this 32-row group among the item's MT x 128 rows
        const uint32_t dst_base = stage_addr + (uint32_t)((R0 >> 7) * (G * kAStage) + (R0 & 127) * PR::kRowBytes);
#pragma unroll
        for (int q = 0; q < PPR; ++q) {
          const int src_row = __shfl_sync(0xffffffffu, idx[i], rr_tab[q]);
          const char* src = src_q[q] + (size_t)(uint32_t)max(src_row, 0) * row_pitch32;   // (IMAD.WIDE.U32)
#ifdef SPC_EXPERIMENTS
          if (p.dbg_flags & 1) continue;
#endif
          cp_async_16(dst_base + dst_tab[q], src, src_row >= 0 ? 16u : 0u);
        }
      }
      // the stage's "full" barrier is signalled by the hardware when this lane's copies have landed:
      // no wait_group, no fence, nothing blocks here
#ifdef SPC_EXPERIMENTS
      if (p.dbg_flags & 128) { if (leader) mbar_arrive(full_bar(slot)); } else
#endif
      cp_async_mbar_arrive_noinc(full_bar(slot));
#ifdef SPC_EXPERIMENTS
      { const long long td = clock64(); t_book += tb - ta; t_wait += tc - tb; t_copy += td - tc; }
#endif
      if (shared_slots && sl == 0 && leader) mbar_arrive(turn_bar(slot));
      slot += ngroups;
      while (slot >= p.stages) { slot -= p.stages; ++rev; }
      cur = nxt;
#pragma unroll
      for (int i = 0; i < NI; ++i) idx[i] = idx_n[i];
    }
#ifdef SPC_EXPERIMENTS
    if (warp == 0 && lane == 0) {
      SPC_ACC(0, clock64() - t_loop0); SPC_ACC(1, t_wait); SPC_ACC(2, t_copy); SPC_ACC(3, t_book);
    }
#endif
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ============================
    // The WHOLE warp runs the loop in uniform control flow and the tcgen05 instructions carry an issue predicate
    // that is true in one elected lane (ptx.cuh: mma_*_p).  Shared-memory descriptors and barrier addresses advance
    // incrementally with the ring slot.  (r2: the single-thread `if (elected)` loop kept every counter in vector
    // registers and moved six of them to uniform registers per stage; with the slot wait and the proxy fence the
    // issuing thread needed ~390 clocks of its own per stage against 192 clocks of tensor work at 96->96 and was
    // never waiting for a full stage — profiles/r2_conv_breakdown.md.)
    {
      const uint32_t issue = elect_one() ? 1u : 0u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = PR::idesc(kTileM, (uint32_t)p.cn_tile, 0, 0);
      const uint64_t desc_hi = make_desc(0, 16, PR::kSboK, PR::kLayoutK);
      const uint32_t sb16 = (uint32_t)stage_bytes >> 4;            // (all shared-memory offsets in 16-byte units)
      const uint32_t lo0 = smem_base >> 4;
      constexpr uint32_t kA16 = (uint32_t)kAStage >> 4;
      const uint32_t bc16 = (uint32_t)b_chunk_bytes >> 4;
      uint32_t lo = lo0;                                            // slot `stage` starts at lo << 4
      uint32_t fbar = full_bar(0);
      for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
        const uint32_t mask = work_mask(w);
        const int n_iters = (int)__reduce_max_sync(0xffffffffu, (unsigned)(__popc(mask) * p.kg_count));  // (uniform)
#ifdef SPC_EXPERIMENTS
        long long m_full = 0;
        const long long ma = clock64();
#endif
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
#ifdef SPC_EXPERIMENTS
        if (issue) { SPC_ACC(6, clock64() - ma); SPC_ACC(7, n_iters); }
        const long long m_loop0 = clock64();
#endif
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * MT * p.cn_tile);
        for (int it = 0; it < n_iters; ++it) {
          SPC_T(mb);
          mbar_wait(fbar, phase);
#ifdef SPC_EXPERIMENTS
          m_full += clock64() - mb;
          if (!(p.dbg_flags & 8))
#endif
          fence_proxy_async_smem();  // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
          tc_fence_after();
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d = d0 + (uint32_t)(mt * p.cn_tile);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint64_t adesc = desc_hi | (uint64_t)(lo + (uint32_t)(mt * G + g) * kA16);
              const uint64_t bdesc = desc_hi | (uint64_t)(lo + (uint32_t)(G * MT) * kA16 + (uint32_t)g * bc16);
#pragma unroll
              for (int q = 0; q < PR::kMmaPerRow; ++q) {  // 32 bytes of K per MMA
#ifdef SPC_EXPERIMENTS
                if (p.dbg_flags & 2) continue;
#endif
                PR::mma_p(d, adesc + 2u * q, bdesc + 2u * q, idesc, (it > 0 || g > 0 || q > 0) ? 1u : 0u, issue);
              }
            }
          }
          mma_commit_p(fbar + 8u * kMaxStages, issue);  // the slot's "empty" barrier: reusable once these MMAs have read it
          lo += sb16; fbar += 8u;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; lo = lo0; fbar = full_bar(0); }
        }
        mma_commit_p(tfull_bar(acc), issue);
#ifdef SPC_EXPERIMENTS
        if (issue) { SPC_ACC(4, clock64() - m_loop0); SPC_ACC(5, m_full); }
#endif
        if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue ============================
    const int ew = warp & 3;  // TMEM lane group this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    int ob = 0;  // staging block of the next round
    const bool store_leader = ew == 0 && elect_one();
    for (int w = blockIdx.x; w < p.n_work; w += gridDim.x) {
      const int mtile = w / items_per_mtile, ntile = (w / p.ksplit) % p.n_ntiles, kg = w % p.ksplit;
      const int o0 = mtile * rows_per_work;
      const uint32_t mask = work_mask(w);
      const bool add_bias = p.bias != nullptr && kg == 0;
      SPC_T(ea);
      mbar_wait_sleep(tfull_bar(acc), acc_phase);
      SPC_T(eb);
#ifdef SPC_EXPERIMENTS
      if (ew == 0 && lane == 0) SPC_ACC(8, eb - ea);
#endif
      tc_fence_after();
#ifdef SPC_EXPERIMENTS
      if (p.dbg_flags & 16) {
        tc_fence_before();
        mbar_arrive(tempty_bar(acc));
        if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
        continue;
      }
#endif
      if (p.out_bufs > 0) {
        // The accumulator tile goes TMEM -> registers -> shared memory (128-byte-swizzled rows: conflict-free
        // 16-byte stores) -> global memory with TMA tensor stores (full 128-byte row segments; rows past
        // m_out are clipped by the TMA unit), one [128 rows x 32 columns] block per round through a ring of
        // out_bufs staging blocks.  A thread-per-row st.global from the 32x32b TMEM layout is 32 half-written
        // sectors per instruction and competes with the row gather for the LSU (r1: 23 % of the kernel).
        const bool skip = mask == 0 && !add_bias && p.reduce_out;  // nothing to add (uniform over the CTA)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int r = ew * 32 + lane;  // row within the sub-tile
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((acc * MT + mt) * p.cn_tile);
          const int row0 = o0 + mt * kTileM;
          for (int cb = 0; cb < p.cn_tile / 32; ++cb) {
            const uint32_t blk = out_stage + (uint32_t)ob * 16384u;
            // the tensor store that used this staging block must have finished READING it
            if (store_leader) {
              if (p.out_bufs == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (!skip) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int c0 = cb * 32 + h * 16;
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
                const uint32_t rowp = blk + (uint32_t)r * 128u;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (uint32_t)(((h * 4 + i) ^ (r & 7)) << 4)),
                               "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                               : "memory");
              }
              fence_proxy_async_smem();  // generic-proxy stores -> visible to the TMA (async proxy) reads
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (p.stats != nullptr && !skip && mask != 0) {
              // BatchNorm statistics of this [128 rows x 32 columns] block, read back from the staging block the
              // tensor store is about to read as well: thread (ew, lane) sums column `lane` over rows 32 ew .. + 31
              // (a warp reads the 32 words of one row: conflict-free), one private double slot per thread and sum.
              // Rows past m_out are zero (their gathers were zero-filled, there is no bias in this mode).
              float s0 = 0.f, s1 = 0.f;
              const int jj = lane >> 2, ww = lane & 3;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const int rw = ew * 32 + rr;
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(blk + (uint32_t)rw * 128u + (uint32_t)(((jj ^ (rw & 7)) << 4) + ww * 4)));
                s0 += v;
                s1 = fmaf(v, v, s1);
              }
              double* mine = s_stats + (size_t)ew * 2 * p.cn_tile + cb * 32 + lane;
              mine[0] += (double)s0;
              mine[p.cn_tile] += (double)s1;
            }
            if (store_leader) {
              if (!skip && !p.dbg_skip_store && row0 < p.m_out) {
                const int col0 = ntile * p.cn_tile + cb * 32;
                if (p.reduce_out) tma_reduce_add_2d(&tmap_out, blk, col0, row0);
                else tma_store_2d(&tmap_out, blk, col0, row0);
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // (possibly empty: keeps the group count in step)
            }
            if (++ob == p.out_bufs) ob = 0;
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
              if (p.reduce_out) {  // partial sums of several offset groups meet in the (zeroed) output
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
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
#ifdef SPC_EXPERIMENTS
      if (ew == 0 && lane == 0) SPC_ACC(9, clock64() - eb);
#endif
      if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
    }
    if (p.out_bufs > 0 && store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.stats != nullptr) {   // this CTA's column sums -> the layer's (n_ntiles == 1: the tile spans all columns)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = (warp & 3) * 32 + lane; c < 2 * p.cn_tile; c += 128) {
        const double v = s_stats[c] + s_stats[2 * p.cn_tile + c] + s_stats[4 * p.cn_tile + c] + s_stats[6 * p.cn_tile + c];
        if (v != 0.0) atomicAdd(p.stats + c, v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

template <int MT, bool BF16, int G, int WPS, int NPW>
static int launch_conv_umma(const UmmaConvParams& p, const CUtensorMap& tmap_out, int grid, size_t smem,
                            cudaStream_t stream) {
  auto kern = conv_umma_kernel<MT, BF16, G, WPS, NPW>;
  static int smem_set = 0;  // (per instantiation) the attribute only ever needs to grow
  if ((int)smem > smem_set) {
    SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = (int)smem;
  }
  kern<<<grid, (NPW + 5) * 32, smem, stream>>>(p, tmap_out);
  SPC_LAUNCHED("conv_umma_kernel");
  return 0;
}

TensorMapEncodeFn tensor_map_encoder() {
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
bool make_out_tile_map(CUtensorMap* map, float* base, int64_t rows, int C) {
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
bool make_rows_tile_map(CUtensorMap* map, const void* base, int64_t rows, int C, int box_rows, bool bf16) {
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

// Path / precision bookkeeping for the bench line: how many launches took which route since the last reset
// (spc_conv_path_counts): [0] tcgen05 bf16, [1] tcgen05 tf32, [2] CUDA-core fp32 (conv_api.cu adds [2]).
std::atomic<long long> g_conv_path_counts[4];

// `in` is fp32 (bf16 == false) or bf16 (bf16 == true) rows.  Weights: `packed` (conv_pack_weights image, 1024-byte
// aligned) or, when null, fp32 `w` packed here into `workspace`.
int conv_fwd_umma(const void* in, const float* w, const void* packed, const float* bias, const int* nbr,
                  const uint32_t* tile_mask, int64_t m_out, int c_in, int c_out, int K,
                  bool transpose_w, bool bf16, float* out, void* workspace,
                  int64_t workspace_bytes, cudaStream_t stream, bool accumulate, double* stats, int* stats_fused) {
  if (stats_fused) *stats_fused = 0;
  if (m_out == 0) return 0;
  SPC_REQUIRE(umma_fwd_supported(c_in, c_out), "shape not supported by the tcgen05 path");
  SPC_REQUIRE(K <= 32, "tcgen05 path supports kernel volume <= 32");
  SPC_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0, "feature rows must be 16-byte aligned");
  const void* Wp = packed;
  if (!Wp) {
    SPC_REQUIRE(w && workspace && workspace_bytes >= umma_fwd_workspace(K, c_in, c_out), "workspace too small");
    void* dstp = (void*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    int rc = conv_pack_weights(w, K, c_in, c_out, transpose_w ? 1 : 0, bf16, dstp, stream);
    if (rc) return rc;
    Wp = dstp;
  }
  SPC_REQUIRE(((uintptr_t)Wp % 1024) == 0, "packed weights must be 1024-byte aligned");
  // large bf16 maps: CTA pairs (cta_group::2), each CTA stages and reads half of every weight slab (conv_umma_pair.cu)
  if (packed != nullptr && g_umma_force_mt != 1 && conv_umma_pair_eligible(m_out, c_in, c_out, K, bf16, bias, Wp, out))
    return conv_fwd_umma_pair(in, Wp, nbr, tile_mask, m_out, c_in, c_out, K, out, stream, accumulate, stats, stats_fused);
  const int row_bytes = bf16 ? 64 : 128;
  const int a_stage = kTileM * row_bytes;

  UmmaConvParams p;
  p.A = in; p.Bp = Wp; p.bias = bias; p.nbr = nbr; p.tile_mask = tile_mask; p.out = out;
#ifdef SPC_EXPERIMENTS
  p.dbg_skip_store = g_umma_dbg[3];
  p.dbg_flags = g_umma_dbg[5];
#else
  p.dbg_skip_store = 0;   // (timing experiments, compiled out of the shipped library)
  p.dbg_flags = 0;
#endif
  p.m_out = (int)m_out; p.Ck = c_in; p.Cn = c_out; p.K = K;
  p.kc_count = c_in / 32;
  // Chunks per stage: one row visit brings G x 64 contiguous bytes (bf16).  128-byte-aligned pairs when the row
  // is a multiple of 128 B; whole 192-byte rows for Ck % 96 == 0 (2 line visits instead of 3).  tf32 rows are
  // 128 B per chunk already (G = 1).
  // Measured on the 1 M-voxel map (bf16, r2): G = 3 at 96->96 0.622 ms against 0.526 with G = 1, G = 2 at 128->128
  // 0.822 against 0.759, at 64->64 0.386 against 0.361 — the larger stages leave 3-4 ring slots instead of 8 and that
  // costs more than the line visits save.  G > 1 is therefore compiled only with -DSPC_EXPERIMENTS (knob 4 = 2).
  int G = 1;
#ifdef SPC_EXPERIMENTS
  if (bf16 && g_umma_dbg[4] == 2) G = (c_in % 64 == 0) ? 2 : (c_in % 96 == 0 ? 3 : 1);
#endif
  // (4 KB behind the barrier area for the epilogue's column sums when statistics are asked for)
  const int stats_smem = (stats != nullptr && c_out <= 128) ? 64 * c_out : 0;   // [4 warps][2 C] doubles
  const int budget = kSmemLimit - 1024 - 256 - stats_smem;
  // Tile shape: MT sub-tiles of 128 rows x cn_tile output channels per work item, and on small maps
  // (deep UNet levels: too few row tiles for 148 SMs) the K offsets split over `ksplit` items whose
  // partial sums meet in the zeroed output through fp32 reduce-adds.  Chosen to minimise the
  // critical path = waves x bytes a CTA stages per item (gather + weight slabs, what bounds the
  // kernel); large maps always come out as ksplit = 1 (atomic-free, one owner per output row).
  int mt = 1;
  for (;;) {
    p.cn_tile = pick_cn_tile(c_out);
    p.ksplit = 1;
    mt = 1;
    double best = 1e300;
    const int cn_max = p.cn_tile;
    for (int cand_mt = 1; cand_mt <= 2; ++cand_mt) {
      if (g_umma_force_mt && cand_mt != g_umma_force_mt) continue;
      for (int cn = cn_max; cn >= 16; cn -= 16) {
        if (c_out % cn || cand_mt * cn > 512) continue;
        const int64_t stage_b = (int64_t)G * ((int64_t)cand_mt * a_stage + (int64_t)cn * row_bytes);
        if (budget / stage_b < (G > 1 ? 3 : 2)) continue;
        const int64_t items1 = ceil_div(m_out, kTileM * cand_mt) * (c_out / cn);
        const int ks_max = items1 >= kNumSMs ? 1 : K;
        for (int ks = 1; ks <= ks_max; ++ks) {
          const int k_per = (int)ceil_div(K, ks);
          if (ceil_div(K, k_per) != ks) continue;  // same split as a smaller ks
          const int64_t items = items1 * ks;
          double per_item = (double)k_per * (p.kc_count / G) * stage_b;
          if (ks > 1) per_item += 3.0 * cand_mt * kTileM * cn * 4;  // zero-fill + atomic epilogue
          per_item += 20000.0;                                       // fixed per-item latency
          const double cost = (double)ceil_div(items, kNumSMs) * per_item + 1e-6 * items * per_item;
          if (cost < best) { best = cost; mt = cand_mt; p.cn_tile = cn; p.ksplit = ks; }
        }
      }
    }
    if (best < 1e300) break;
    SPC_REQUIRE(G > 1, "tile does not fit in shared memory");
    G = 1;  // multi-chunk stages do not leave three ring slots for this shape
  }
  p.kg_count = p.kc_count / G;
  p.k_per = (int)ceil_div(K, p.ksplit);
  p.n_ntiles = c_out / p.cn_tile;
  // accumulate: `out` already holds the other summand (the gradient that reached the same rows through a residual
  // connection), every item reduce-adds into it and nothing is cleared
  if (p.ksplit > 1 && !accumulate) SPC_CUDA(cudaMemsetAsync(out, 0, (size_t)m_out * c_out * sizeof(float), stream));
  p.reduce_out = (p.ksplit > 1 || accumulate) ? 1 : 0;

  const int rows_per_work = kTileM * mt;
  p.n_work = (int)ceil_div(m_out, rows_per_work) * p.n_ntiles * p.ksplit;
  p.acc_bufs = (2 * mt * p.cn_tile <= 512) ? 2 : 1;
  int cols = p.acc_bufs * mt * p.cn_tile;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const int stage_bytes = G * (mt * a_stage + p.cn_tile * row_bytes);
  // TMA-store epilogue: 32-column blocks through one or two 16 KB staging blocks, as long as enough ring
  // slots remain (two blocks when that still leaves four slots)
  CUtensorMap tmap_out;
  memset(&tmap_out, 0, sizeof(tmap_out));
  p.out_bufs = 0;
  const int min_slots = G > 1 ? 3 : 4;
  if (g_umma_dbg[2] != 1 && (bf16 || p.cn_tile >= 64) && p.cn_tile % 32 == 0 && ((uintptr_t)out % 16) == 0 &&
      (budget - 16384) / stage_bytes >= min_slots && make_out_tile_map(&tmap_out, out, m_out, c_out)) {
    p.out_bufs = (budget - 32768) / stage_bytes >= std::min(4, budget / stage_bytes) ? 2 : 1;
  }
  int stages = (budget - p.out_bufs * 16384) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  SPC_REQUIRE(stages >= 2, "tile does not fit in shared memory");
  p.stages = stages;
  // producer groups: the largest power of two <= min(ring slots, 8); every warp slice holds >= 32 rows
  // producer warps: 16 split every stage over two warps (more loads in flight while one of them does its
  // bookkeeping): 32->32 0.173 -> 0.163 ms, 64->64 0.332 -> 0.316, 96->96 0.505 -> 0.493 on 1 M voxels; wider tiles
  // (stages of >= 24 KB) are better off with 8 (128->128 0.705 vs 0.746).  Knob 6 forces 8 / 16.
  const int npw = g_umma_dbg[6] == 8 ? 8 : (g_umma_dbg[6] == 16 ? 16 : (p.cn_tile <= 96 ? 16 : 8));
  int ngroups = 1;
  while (ngroups * 2 <= stages && ngroups * 2 <= 8) ngroups *= 2;
  if (g_umma_dbg[0] >= 1 && g_umma_dbg[0] <= stages && g_umma_dbg[0] <= 8 && (g_umma_dbg[0] & (g_umma_dbg[0] - 1)) == 0)
    ngroups = g_umma_dbg[0];
  int wps = npw / ngroups;
  if (wps > 4) wps = 4;        // (instantiated: 1, 2, 4 warps per group; every slice holds >= 32 rows)
  // 16 one-warp groups alternating on 8 ring slots (see the producer loop); knob 7 = 1 keeps 8 groups of two warps
  if (npw == 16 && ngroups == 8 && stages == 8 && g_umma_dbg[7] != 1) { ngroups = 16; wps = 1; }
  if (kTileM * mt / wps < 32) wps = kTileM * mt / 32;
  p.ngroups = ngroups;
  p.wps = wps;
  // BatchNorm statistics in the epilogue: one owner per output row (no offset split, no accumulation), the tile spans
  // all columns, no bias, TMA-store epilogue (the sums are read back from its staging blocks)
  p.stats = nullptr;
  if (stats_smem && p.ksplit == 1 && !accumulate && p.n_ntiles == 1 && bias == nullptr && p.out_bufs > 0) {
    SPC_CUDA(cudaMemsetAsync(stats, 0, (size_t)2 * c_out * sizeof(double), stream));
    p.stats = stats;
    if (stats_fused) *stats_fused = 1;
  }
  const size_t smem = (size_t)stages * stage_bytes + (size_t)p.out_bufs * 16384 + 1024 + 256 + stats_smem;
  const int grid = p.n_work < kNumSMs ? p.n_work : kNumSMs;
  g_conv_path_counts[bf16 ? 0 : 1].fetch_add(1, std::memory_order_relaxed);
  // (MT, precision, G, warps per producer group) -> instantiation
#define SPC_CONV_CASE(MT_, BF_, G_, W_) \
  if (mt == MT_ && bf16 == BF_ && G == G_ && wps == W_) \
    return npw == 8 ? launch_conv_umma<MT_, BF_, G_, W_, 8>(p, tmap_out, grid, smem, stream) \
                    : launch_conv_umma<MT_, BF_, G_, W_, 16>(p, tmap_out, grid, smem, stream);
#define SPC_CONV_WPS(MT_, BF_, G_) SPC_CONV_CASE(MT_, BF_, G_, 1) SPC_CONV_CASE(MT_, BF_, G_, 2) SPC_CONV_CASE(MT_, BF_, G_, 4)
  SPC_CONV_WPS(1, true, 1) SPC_CONV_WPS(2, true, 1) SPC_CONV_WPS(1, false, 1) SPC_CONV_WPS(2, false, 1)
#ifdef SPC_EXPERIMENTS   // multi-chunk stages (contiguous 128 / 192-byte row visits): measured slower, see below
  SPC_CONV_WPS(1, true, 2) SPC_CONV_WPS(2, true, 2) SPC_CONV_WPS(1, true, 3) SPC_CONV_WPS(2, true, 3)
#endif
#undef SPC_CONV_WPS
#undef SPC_CONV_CASE
  return fail("conv_fwd_umma", "no kernel instantiation for this tile configuration");
}

#ifdef SPC_EXPERIMENTS
extern "C" int spc_debug_role_cycles(long long* out16, int reset) {
  unsigned long long h[16];
  if (cudaMemcpyFromSymbol(h, g_role_cycles, sizeof(h)) != cudaSuccess) return 1;
  for (int i = 0; i < 16; ++i) out16[i] = (long long)h[i];
  if (reset) { memset(h, 0, sizeof(h)); if (cudaMemcpyToSymbol(g_role_cycles, h, sizeof(h)) != cudaSuccess) return 1; }
  return 0;
}
#endif
void umma_set_force_mt(int mt) { g_umma_force_mt = mt; }
void umma_debug_set(int idx, int val) { if (idx >= 0 && idx < 16) g_umma_dbg[idx] = val; }

}  // namespace spc
