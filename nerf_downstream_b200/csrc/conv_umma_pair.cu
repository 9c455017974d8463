// conv_umma_pair.cu — the forward / dgrad implicit GEMM of conv_umma.cu on CTA PAIRS (tcgen05 cta_group::2).
//
// Why: at N = cn_tile <= 128 output channels the single-CTA kernel is bound by shared-memory bandwidth, not by the
// tensor pipe — one tcgen05.mma (M = 128, N = 96, K = 16, bf16) reads 4 KB of A and 3 KB of B from shared memory in
// its 48 tensor clocks (146 B/clk against the 128 B/clk the SM delivers), and the gather writes the A bytes a
// second time (profiles/r2_conv_breakdown.md: tensor 192, copy port 256, shared memory 390 of the ~440 clocks a
// 96->96 stage takes).  A CTA pair issues ONE M = 256 instruction for two row tiles: each CTA gathers its own 128
// rows per sub-tile as before, but stages and reads only HALF of every weight slab — the B operand of a
// cta_group::2 instruction is split over the two CTAs' shared memories and each half feeds both tensor cores.
// Shared-memory bytes per stage and CTA at 96->96 (MT = 2): 16 KB gathered + 3 KB slab written, 16 KB + 6 KB read by
// the MMAs = 41 KB instead of 50 KB.
//
// What changes against conv_umma_kernel (everything else — producer groups, ring slots, TMA-store epilogue through
// staging blocks, epilogue statistics — is the same code):
//   * cluster of 2 CTAs; a work item is 2 x MT x 128 output rows, CTA `rank` owns the rank-th half; both CTAs walk
//     the same (offset, chunk) stage sequence (offset mask = OR over the four 128-row tiles);
//   * only rank 0's MMA warp issues tcgen05.mma (cta_group::2, M = 256).  It waits for ITS stage and for the peer's:
//     rank 1's MMA warp is a relay — it waits for its own "full" barrier (signalled by the hardware when the cp.async
//     rows and the slab half have landed), fences the generic-proxy writes for the async proxy, and arrives on
//     rank 0's `peer_full` barrier through the cluster address space;
//   * tcgen05.commit ... multicast::cluster releases the ring slot (and publishes the accumulators) in BOTH CTAs;
//   * the epilogue threads of both CTAs hand the accumulator buffer back on rank 0's "tmem empty" barrier;
//   * TMEM is allocated / freed with the cta_group::2 forms by one warp of each CTA, cluster barriers around set-up
//     and tear-down.
// bf16 operands, MT = 2, one n tile (cn_tile = Cn <= 256), no offset split: the shapes of the large maps, where the
// single-CTA kernel spends its time.  Everything else stays on conv_umma_kernel.
#include "umma_common.cuh"

namespace spc {

namespace {

constexpr int kPairBarBytes = 512;   // barrier area (the single-CTA kernel's 256 B + the peer_full barriers)
constexpr int kPairMaxStages = 12;   // ring slots: the pair's stages are smaller (half a slab) and the relay hop wants depth

}  // namespace

constexpr int kPairMaxThreads = (16 + 4 + 1) * 32;
template <int NPW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairMaxThreads, 1)
conv_umma_pair_kernel(const UmmaConvParams p, const __grid_constant__ CUtensorMap tmap_out) {
  using PR = Prec<true>;
  constexpr int MT = 2;
  constexpr int kMmaWarpP = NPW + 4;
  constexpr int kAStage = kTileM * PR::kRowBytes;  // one 32-channel chunk of one 128-row sub-tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_half_bytes = (p.cn_tile / 2) * PR::kRowBytes;   // this CTA's half of a weight slab
  const int stage_bytes = MT * kAStage + b_half_bytes;
  const uint32_t out_stage = smem_base + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = out_stage + (uint32_t)p.out_bufs * 16384u;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPairMaxStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kPairMaxStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kPairMaxStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kPairMaxStages + 4);
  auto turn_bar = [&](int s) { return bar_base + 8u * (2 * kPairMaxStages + 5 + s); };
  auto peer_full_bar = [&](int s) { return bar_base + 8u * (3 * kPairMaxStages + 5 + s); };   // used in rank 0
  double* s_stats = reinterpret_cast<double*>(smem_raw + (bar_base + kPairBarBytes - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = (int)__reduce_min_sync(0xffffffffu, threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)(blockIdx.x >> 1), n_cl = (int)(gridDim.x >> 1);

  if (p.stats != nullptr)
    for (int c = threadIdx.x; c < 8 * p.cn_tile; c += blockDim.x) s_stats[c] = 0.0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 32 + 1);                   // the lanes of the stage's producer warp + 1 expect_tx
      mbar_init(empty_bar(s), 1);                       // one multicast tcgen05.commit
      mbar_init(turn_bar(s), 1);
      mbar_init(peer_full_bar(s), 1);                   // the peer's relay
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * kNumEpilogueThreads);   // the epilogue threads of BOTH CTAs
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarpP) {
    tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();      // barriers of both CTAs are initialised before anybody arrives on the peer's
  tc_fence_after();
  const uint32_t tmem_base = __reduce_min_sync(0xffffffffu, *tmem_slot_ptr);

  constexpr int rows_per_cta = kTileM * MT;   // rows of a work item owned by one CTA
  // work item w = rows [512 w, 512 w + 512): offsets with a neighbour in ANY of its four 128-row tiles
  auto work_mask = [&](int w) -> uint32_t {
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < 2 * MT; ++q) {
      const int t = w * (2 * MT) + q;
      if ((long long)t * kTileM < p.m_out) mask |= p.tile_mask ? p.tile_mask[t] : 0xFFFFFFFFu;
    }
    if (p.K < 32) mask &= (1u << p.K) - 1u;
    return mask;
  };

  if (warp < NPW) {
    // ============================ producers (one warp per stage, as conv_umma_kernel with WPS = 1) ============
    const int ngroups = p.ngroups;
    const int grp = warp;
    constexpr int LPR = PR::kLanesPerRow;
    constexpr int NI = rows_per_cta / 32;
    const char* Bbase = reinterpret_cast<const char*>(p.Bp);
    const bool leader = elect_one();
    int rr_tab[LPR];
    uint32_t dst_tab[LPR], src_tab[LPR];
#pragma unroll
    for (int q = 0; q < LPR; ++q) {
      const int f = q * 32 + lane;
      const int rr = f / LPR, j = f - rr * LPR;
      rr_tab[q] = rr;
      dst_tab[q] = (uint32_t)(rr * PR::kRowBytes) + PR::swz_k(j, rr);
      src_tab[q] = (uint32_t)(j * 16);
    }
    struct It { int w, cg; uint32_t rest; bool ok; };   // lowest set bit of `rest` = current offset
    auto open_item = [&](It& it, int skip) {
      for (;;) {
        if (it.w >= p.n_work || grp >= ngroups) { it.ok = false; return; }
        uint32_t rest = work_mask(it.w);
        int cg = skip;
        while (rest != 0u && cg >= p.kg_count) { cg -= p.kg_count; rest &= rest - 1u; }
        if (rest != 0u) { it.rest = rest; it.cg = cg; return; }
        skip = cg;
        it.w += n_cl;
      }
    };
    auto advance = [&](It& it) {
      int cg = it.cg + ngroups;
      uint32_t rest = it.rest;
      while (rest != 0u && cg >= p.kg_count) { cg -= p.kg_count; rest &= rest - 1u; }
      if (rest != 0u) { it.cg = cg; it.rest = rest; return; }
      it.w += n_cl;
      open_item(it, cg);
    };
    auto load_idx = [&](const It& it, int* idx) {
      const int k = __ffs(it.rest) - 1;
      const int o0 = (it.w * 2 + (int)rank) * rows_per_cta + lane;
      const int* row = p.nbr + (size_t)k * p.m_out;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int o = o0 + i * 32;
        idx[i] = o < p.m_out ? __ldg(row + o) : -1;
      }
    };
    const uint32_t row_pitch32 = (uint32_t)p.Ck * PR::kElt;

    It cur;
    cur.w = cid; cur.ok = true; cur.cg = 0; cur.rest = 0;
    open_item(cur, grp);
    int idx[NI];
    if (cur.ok) load_idx(cur, idx);
    int slot = grp % p.stages, rev = grp / p.stages;
    const bool shared_slots = ngroups > p.stages;
    while (cur.ok) {
      It nxt = cur;
      advance(nxt);
      int idx_n[NI];
      if (nxt.ok) load_idx(nxt, idx_n);

      const int k = __ffs(cur.rest) - 1, cg = cur.cg;
      const uint32_t par = (uint32_t)(rev - 1) & 1u;
      if (shared_slots) mbar_wait(turn_bar(slot), par);
      mbar_wait(empty_bar(slot), par);             // (released by rank 0's multicast commit)
      const uint32_t stage_addr = smem_base + (uint32_t)slot * stage_bytes;
      if (leader) {
        mbar_arrive_expect_tx(full_bar(slot), (uint32_t)b_half_bytes);
        // rows [rank * cn_tile / 2, + cn_tile / 2) of the slab of (offset k, chunk cg)
        const char* src = Bbase + (((size_t)k * p.kc_count + (size_t)cg) * p.Cn + (size_t)rank * (p.cn_tile / 2)) * PR::kRowBytes;
        bulk_g2s(stage_addr + MT * kAStage, src, (uint32_t)b_half_bytes, full_bar(slot));
      }
      __syncwarp();
      const char* src_q[LPR];
#pragma unroll
      for (int q = 0; q < LPR; ++q)
        src_q[q] = reinterpret_cast<const char*>(p.A) + (size_t)cg * PR::kRowBytes + src_tab[q];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int R0 = i * 32;
        const uint32_t dst_base = stage_addr + (uint32_t)((R0 >> 7) * kAStage + (R0 & 127) * PR::kRowBytes);
#pragma unroll
        for (int q = 0; q < LPR; ++q) {
          const int src_row = __shfl_sync(0xffffffffu, idx[i], rr_tab[q]);
          const char* src = src_q[q] + (size_t)(uint32_t)max(src_row, 0) * row_pitch32;
          cp_async_16(dst_base + dst_tab[q], src, src_row >= 0 ? 16u : 0u);
        }
      }
      cp_async_mbar_arrive_noinc(full_bar(slot));
      if (shared_slots && leader) mbar_arrive(turn_bar(slot));
      slot += ngroups;
      while (slot >= p.stages) { slot -= p.stages; ++rev; }
      cur = nxt;
#pragma unroll
      for (int i = 0; i < NI; ++i) idx[i] = idx_n[i];
    }
  } else if (warp == kMmaWarpP) {
    const uint32_t issue = elect_one() ? 1u : 0u;
    int stage = 0;
    uint32_t phase = 0;
    if (rank == 0) {
      // ============================ MMA issuer (rank 0) ============================
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = PR::idesc(2 * kTileM, (uint32_t)p.cn_tile, 0, 0);   // M = 256 over the pair
      const uint64_t desc_hi = make_desc(0, 16, PR::kSboK, PR::kLayoutK);
      const uint32_t sb16 = (uint32_t)stage_bytes >> 4;
      const uint32_t lo0 = smem_base >> 4;
      constexpr uint32_t kA16 = (uint32_t)kAStage >> 4;
      uint32_t lo = lo0;
      uint32_t fbar = full_bar(0);
      for (int w = cid; w < p.n_work; w += n_cl) {
        const uint32_t mask = work_mask(w);
        const int n_iters = (int)__reduce_max_sync(0xffffffffu, (unsigned)(__popc(mask) * p.kg_count));
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)(acc * MT * p.cn_tile);
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(fbar, phase);                                       // this CTA's rows and slab half
          mbar_wait(fbar + 8u * (3 * kPairMaxStages + 5), phase);           // the peer's (peer_full_bar)
          fence_proxy_async_smem();
          tc_fence_after();
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t d = d0 + (uint32_t)(mt * p.cn_tile);
            const uint64_t adesc = desc_hi | (uint64_t)(lo + (uint32_t)mt * kA16);
            const uint64_t bdesc = desc_hi | (uint64_t)(lo + (uint32_t)MT * kA16);
#pragma unroll
            for (int q = 0; q < PR::kMmaPerRow; ++q)
              mma_bf16_pair_p(d, adesc + 2u * q, bdesc + 2u * q, idesc, (it > 0 || q > 0) ? 1u : 0u, issue);
          }
          mma_commit_pair_p(fbar + 8u * kPairMaxStages, issue);   // "empty" of this slot in both CTAs
          lo += sb16; fbar += 8u;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; lo = lo0; fbar = full_bar(0); }
        }
        mma_commit_pair_p(tfull_bar(acc), issue);
        if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
      }
    } else {
      // ============================ relay (rank 1) ============================
      // "my rows and my half of the slab of this stage are in shared memory, visible to the tensor cores"
      uint32_t fbar = full_bar(0);
      for (int w = cid; w < p.n_work; w += n_cl) {
        const uint32_t mask = work_mask(w);
        const int n_iters = (int)__reduce_max_sync(0xffffffffu, (unsigned)(__popc(mask) * p.kg_count));
        for (int it = 0; it < n_iters; ++it) {
          mbar_wait(fbar, phase);
          fence_proxy_async_smem();
          if (issue) mbar_arrive_cluster(map_to_rank(fbar + 8u * (3 * kPairMaxStages + 5), 0u));
          fbar += 8u;
          if (++stage == p.stages) { stage = 0; phase ^= 1u; fbar = full_bar(0); }
        }
      }
    }
    __syncwarp();
  } else {
    // ============================ epilogue (as conv_umma_kernel, TMA-store path) ============================
    const int ew = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ob = 0;
    const bool store_leader = ew == 0 && elect_one();
    for (int w = cid; w < p.n_work; w += n_cl) {
      const int o0 = (w * 2 + (int)rank) * rows_per_cta;
      const uint32_t mask = work_mask(w);
      mbar_wait_sleep(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const bool skip = mask == 0 && p.reduce_out;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int r = ew * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((acc * MT + mt) * p.cn_tile);
        const int row0 = o0 + mt * kTileM;
        for (int cb = 0; cb < p.cn_tile / 32; ++cb) {
          const uint32_t blk = out_stage + (uint32_t)ob * 16384u;
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
              const uint32_t rowp = blk + (uint32_t)r * 128u;
#pragma unroll
              for (int i = 0; i < 4; ++i)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (uint32_t)(((h * 4 + i) ^ (r & 7)) << 4)),
                             "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                             : "memory");
            }
            fence_proxy_async_smem();
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (p.stats != nullptr && !skip && mask != 0) {
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
            if (!skip && row0 < p.m_out) {
              const int col0 = cb * 32;
              if (p.reduce_out) tma_reduce_add_2d(&tmap_out, blk, col0, row0);
              else tma_store_2d(&tmap_out, blk, col0, row0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (++ob == p.out_bufs) ob = 0;
        }
      }
      tc_fence_before();
      // the accumulator buffer goes back to rank 0's MMA warp: 128 arrivals from each CTA
      if (rank == 0) mbar_arrive(tempty_bar(acc));
      else mbar_arrive_cluster(map_to_rank(tempty_bar(acc), 0u));
      if (++acc == p.acc_bufs) { acc = 0; acc_phase ^= 1u; }
    }
    if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = (warp & 3) * 32 + lane; c < 2 * p.cn_tile; c += 128) {
        const double v = s_stats[c] + s_stats[2 * p.cn_tile + c] + s_stats[4 * p.cn_tile + c] + s_stats[6 * p.cn_tile + c];
        if (v != 0.0) atomicAdd(p.stats + c, v);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();    // nobody leaves while the peer may still read its shared memory or arrive on its barriers
  if (warp == kMmaWarpP) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
  }
}

bool make_out_tile_map(CUtensorMap* map, float* base, int64_t rows, int C);
extern std::atomic<long long> g_conv_path_counts[4];

// Eligible: bf16 rows and packed weights; one n tile of <= 256 output channels that is a multiple of 32 (32-column
// store blocks; each CTA's half of a slab is then a multiple of 16 rows, which keeps the 64-byte swizzle phase and is
// a legal N / 2 of a cta_group::2 instruction); no bias (what reaches this path are convolutions in front of a
// BatchNorm); a map large enough that every pair gets at least two work items.
bool conv_umma_pair_eligible(int64_t m_out, int c_in, int c_out, int K, bool bf16, const float* bias, const void* packed,
                             const float* out) {
  constexpr bool kPairByDefault = true;    // (knob 8: 1 = never, 2 = on every eligible shape whatever the map size)
  if (g_umma_dbg[8] == 1 || (g_umma_dbg[8] == 0 && !kPairByDefault)) return false;
  if (!bf16 || bias != nullptr || packed == nullptr) return false;
  if (c_in < 32 || c_in % 32 || c_out % 32 || c_out > 256 || K > 32) return false;
  if (((uintptr_t)out % 16) != 0) return false;
  if (g_umma_dbg[8] == 2) return m_out >= 1;
  return m_out >= (int64_t)2 * 512 * (kNumSMs / 2);   // at least two work items per pair
}

template <int NPW>
static int launch_pair(const UmmaConvParams& p, const CUtensorMap& tmap_out, int grid, size_t smem, cudaStream_t stream) {
  auto kern = conv_umma_pair_kernel<NPW>;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = (int)smem;
  }
  kern<<<grid, (NPW + 5) * 32, smem, stream>>>(p, tmap_out);
  SPC_LAUNCHED("conv_umma_pair_kernel");
  return 0;
}

// same contract as conv_fwd_umma (conv_umma.cu) for the eligible shapes; *stats_fused as there
int conv_fwd_umma_pair(const void* in, const void* packed, const int* nbr, const uint32_t* tile_mask, int64_t m_out,
                       int c_in, int c_out, int K, float* out, cudaStream_t stream, bool accumulate, double* stats,
                       int* stats_fused) {
  if (stats_fused) *stats_fused = 0;
  if (m_out == 0) return 0;
  SPC_REQUIRE(((uintptr_t)packed % 1024) == 0, "packed weights must be 1024-byte aligned");
  SPC_REQUIRE(((uintptr_t)in % 16) == 0, "feature rows must be 16-byte aligned");
  UmmaConvParams p;
  memset(&p, 0, sizeof(p));
  p.A = in; p.Bp = packed; p.bias = nullptr; p.nbr = nbr; p.tile_mask = tile_mask; p.out = out;
  p.m_out = (int)m_out; p.Ck = c_in; p.Cn = c_out; p.K = K;
  p.kc_count = c_in / 32; p.kg_count = p.kc_count;
  p.cn_tile = c_out; p.n_ntiles = 1; p.ksplit = 1; p.k_per = K;
  p.reduce_out = accumulate ? 1 : 0;
  p.n_work = (int)ceil_div(m_out, 4 * kTileM);
  p.acc_bufs = (2 * 2 * c_out <= 512) ? 2 : 1;
  const int cols = p.acc_bufs * 2 * c_out;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const int stats_smem = (stats != nullptr && c_out <= 128) ? 64 * c_out : 0;
  const int budget = kSmemLimit - 1024 - kPairBarBytes - stats_smem;
  const int stage_bytes = 2 * kTileM * 64 + (c_out / 2) * 64;
  CUtensorMap tmap_out;
  memset(&tmap_out, 0, sizeof(tmap_out));
  SPC_REQUIRE(make_out_tile_map(&tmap_out, out, m_out, c_out), "cuTensorMapEncodeTiled unavailable");
  p.out_bufs = (budget - 32768) / stage_bytes >= std::min(4, budget / stage_bytes) ? 2 : 1;
  if (g_umma_dbg[11] == 1) p.out_bufs = 1;   // (knob 11: one staging block, one more ring slot)
  int stages = (budget - p.out_bufs * 16384) / stage_bytes;
  if (stages > kPairMaxStages) stages = kPairMaxStages;
  SPC_REQUIRE(stages >= 2, "pair tile does not fit in shared memory");
  p.stages = stages;
  const int npw = g_umma_dbg[6] == 8 ? 8 : (g_umma_dbg[6] == 16 ? 16 : (c_out <= 96 ? 16 : 8));
  if (g_umma_dbg[10] >= 2 && g_umma_dbg[10] < stages) { stages = g_umma_dbg[10]; p.stages = stages; }   // (knob 10: ring slots)
  // one-warp groups: as many as ring slots (a group is never two revolutions ahead), or 16 alternating on 8 slots
  int ngroups = std::min(stages, npw);
  if (npw == 16 && stages == 8 && g_umma_dbg[7] != 1) ngroups = 16;   // two groups alternate on a slot
  p.ngroups = ngroups; p.wps = 1;
  p.stats = nullptr;
  if (stats_smem && !accumulate) {
    SPC_CUDA(cudaMemsetAsync(stats, 0, (size_t)2 * c_out * sizeof(double), stream));
    p.stats = stats;
    if (stats_fused) *stats_fused = 1;
  }
  const size_t smem = (size_t)stages * stage_bytes + (size_t)p.out_bufs * 16384 + 1024 + kPairBarBytes + stats_smem;
  int grid = 2 * std::min(p.n_work, kNumSMs / 2);
  g_conv_path_counts[0].fetch_add(1, std::memory_order_relaxed);
  return npw == 8 ? launch_pair<8>(p, tmap_out, grid, smem, stream) : launch_pair<16>(p, tmap_out, grid, smem, stream);
}

}  // namespace spc
