// conv_wgrad_umma_pair.cu — the weight-gradient GEMM of conv_wgrad_umma.cu on CTA PAIRS (tcgen05 cta_group::2).
//
//     dW[k][ci][co] = sum_o in[nbr[k,o]][ci] * dout[o][co]        D[M x N] += A[M x rows] * B[rows x N]
//
// The single-CTA kernel re-reads the dout tile (B, 128 rows x Cout) from shared memory for every M block (128 (offset,
// channel) rows of dW) it multiplies with: at Cout = 96 a step reads 32 KB of A and 24 KB of B for 384 tensor clocks —
// 56 KB against the 48 KB the SM delivers in that time, before the gather has written a byte.  A pair issues M = 256:
// CTA `rank` gathers the rows of M block (mb0 + 2 i + rank), and each CTA stages and reads only HALF of the dout
// columns (the B operand of a cta_group::2 instruction is split over the two CTAs' shared memories).  Each CTA's TMEM
// holds floor(512 / Cout) accumulators, so a pass covers twice as many M blocks and the dout rows are re-read by
// ceil(n_mb / (2 cap)) passes instead of ceil(n_mb / cap) (96->96: 3 instead of 5).
//
// Roles, step tables and ring slots are those of conv_wgrad_umma_kernel; what changes:
//   * a step is (row block, PAIR of M blocks); it exists when either M block has a neighbour in the row block (the
//     other CTA's gather is then all zero-fill, which is what its accumulator must see);
//   * rank 0's MMA warp issues for both; rank 1's MMA warp is a relay: it waits for its own A / B "full" barriers in
//     step order and arrives on rank 0's peer barriers through the cluster address space;
//   * tcgen05.commit multicast releases A / B slots and publishes the accumulators in both CTAs; both CTAs' epilogue
//     threads hand the accumulators back on rank 0's barrier; cta_group::2 TMEM allocation, cluster barriers around
//     set-up and tear-down.
// bf16 operands only (the mode the step runs in); everything else stays on conv_wgrad_umma_kernel.
#include "umma_common.cuh"

namespace spc {

constexpr int kWpChunkBlock = 8192;              // [128 rows x 64-byte row chunk]
constexpr int kWpAStage = 4 * kWpChunkBlock;     // one 128-row M block
constexpr int kWpMaxAStages = 6;
constexpr int kWpBStages = 2;
constexpr int kWpMaxMb = 128;
constexpr int kWpActBytes = 384;
constexpr int kWpTabBytes = kWpMaxMb * 4 * 2 + 2 * kWpActBytes;

struct UmmaWgradPairParams {
  const void* in;             // [m_in, Cin] bf16
  const int* nbr;             // [K, m_out]
  const uint32_t* tile_mask;  // [ceil(m_out/128)] or null
  float* dw;                  // [K, Cin, dw_pitch] (+ column offset applied)
  int m_out, Cin, Cout, K;    // Cout = width of THIS launch's column group (<= 256)
  int dw_pitch, col0;
  int ncc, nq;                // chunks per offset, total chunks
  int n_mb, n_pass;
  int n_rb, rb_per_split, n_split;
  int a_stages;
  int nb_half;                // 32-column tile loads per CTA and row block: ceil(Cout / 2 / 32)
  int act16;                  // step-table entries are 16 bits (more than 8 M-block pairs per pass)
  int n_work;
};

template <int RM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
conv_wgrad_pair_kernel(const UmmaWgradPairParams p, const __grid_constant__ CUtensorMap tmap_dout) {
  using PR = Prec<true>;
  constexpr int kRows = 128;
  constexpr int kMmaPerStep = kRows / 16;
  constexpr int LPR = PR::kLanesPerRow;
  constexpr int NI = kRows / 32;
  constexpr int NR = 4 / RM;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_stage_bytes = p.nb_half * kWpChunkBlock;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)p.a_stages * kWpAStage;
  const uint32_t bar_base = b_base + (uint32_t)kWpBStages * b_stage_bytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kWpMaxAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kWpMaxAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kWpMaxAStages + kWpBStages + s); };
  const uint32_t t_full = bar_base + 8u * (2 * kWpMaxAStages + 2 * kWpBStages);
  const uint32_t t_empty = t_full + 8u;
  const uint32_t tmem_slot = t_full + 16u;
  auto peer_a_full = [&](int s) { return t_full + 24u + 8u * s; };                    // (rank 0)
  auto peer_b_full = [&](int s) { return t_full + 24u + 8u * (kWpMaxAStages + s); };  // (rank 0)
  uint8_t* tab_raw = smem_raw + (bar_base + 256u - smem_u32(smem_raw));
  uint16_t* s_run = reinterpret_cast<uint16_t*>(tab_raw);             // [kWpMaxMb][4] row visits: k | first chunk << 6
  uint8_t* s_act = tab_raw + kWpMaxMb * 4 * 2;                        // [2][kWpActBytes] active M-block pairs per row block
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = (int)__reduce_min_sync(0xffffffffu, threadIdx.x >> 5), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)(blockIdx.x >> 1), n_cl = (int)(gridDim.x >> 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(a_full(s), 32);
      mbar_init(a_empty(s), 1);
      mbar_init(peer_a_full(s), 1);
    }
    for (int s = 0; s < kWpBStages; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
      mbar_init(peer_b_full(s), 1);
    }
    mbar_init(t_full, 1);
    mbar_init(t_empty, 2 * kNumEpilogueThreads);   // the epilogue threads of both CTAs
    fence_mbar_init();
  }
  if (warp == kMmaWarp) { tmem_alloc_pair(tmem_slot, 512u); tmem_relinquish_pair(); }
  for (int e = threadIdx.x; e < p.n_mb * 4; e += blockDim.x) {
    const int mb = e >> 2, r = e & 3;
    const int q = mb * 4 + r * RM;
    s_run[e] = (uint16_t)((r < NR && q < p.nq) ? ((q / p.ncc) | ((q % p.ncc) << 6)) : 63);  // 63: nothing to gather
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_min_sync(0xffffffffu, *tmem_slot_ptr);
  const uint32_t all_taps = p.K >= 32 ? 0xFFFFFFFFu : ((1u << p.K) - 1u);

  constexpr int PPR = RM * LPR;
  uint32_t tab[PPR];
#pragma unroll
  for (int q = 0; q < PPR; ++q) {
    const int f = q * 32 + lane;
    const int rr = f / PPR, piece = f - rr * PPR;
    const int c = piece / LPR, j = piece - c * LPR;
    const uint32_t dsto = (uint32_t)(c * kWpChunkBlock + rr * PR::kRowBytes) + PR::swz_mn(j, rr);
    tab[q] = (uint32_t)rr | ((uint32_t)piece << 5) | (dsto << 12);
  }
  // offsets (bit mask) of M block mb; 0 for mb >= n_mb
  auto mb_taps = [&](int mb) -> uint32_t {
    uint32_t t = 0;
    if (mb >= p.n_mb) return 0u;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const uint32_t d = s_run[mb * 4 + r];
      if ((d & 63u) != 63u) t |= 1u << (d & 63u);
    }
    return t;
  };
  auto act_at = [&](const uint8_t* act, int rbi) -> uint32_t {
    return p.act16 ? (uint32_t)reinterpret_cast<const uint16_t*>(act)[rbi] : (uint32_t)act[rbi];
  };

  uint32_t a_phase = 0, b_phase = 0, t_phase = 0;
  uint32_t m_a_phase = 0, m_b_phase = 0, m_t_phase = 0;
  int local_item = 0;
  for (int w = cid; w < p.n_work; w += n_cl, ++local_item) {
    const int split = w / p.n_pass, pass = w - split * p.n_pass;
    const int mb0 = pass * p.n_mb / p.n_pass, mb1 = (pass + 1) * p.n_mb / p.n_pass;  // balanced passes
    const int n_pairs = (mb1 - mb0 + 1) >> 1;                                         // pair i = M blocks mb0 + 2 i, + 1
    const int rb0 = split * p.rb_per_split, rb1 = min(rb0 + p.rb_per_split, p.n_rb);
    const int n_rbi = rb1 - rb0;
    uint8_t* act = s_act + (local_item & 1) * kWpActBytes;
    // ---- step table: bit i of act[rbi] = pair i has a neighbour in row block rb0 + rbi (identical in both CTAs) ----
    const int filler = warp == kMmaWarp ? kNumProducerWarps * 32 + lane : threadIdx.x;
    for (int rbi = filler; rbi < n_rbi && (warp < kNumProducerWarps || warp == kMmaWarp); rbi += 288) {
      const uint32_t m = p.tile_mask ? (p.tile_mask[rb0 + rbi] & all_taps) : all_taps;
      uint32_t a = 0;
      for (int i = 0; i < n_pairs; ++i) {
        const int mb = mb0 + 2 * i;
        uint32_t taps = mb_taps(mb);
        if (mb + 1 < mb1) taps |= mb_taps(mb + 1);
        if (taps & m) a |= 1u << i;
      }
      if (p.act16) reinterpret_cast<uint16_t*>(act)[rbi] = (uint16_t)a;
      else act[rbi] = (uint8_t)a;
    }
    if (warp < kNumProducerWarps || warp == kMmaWarp) asm volatile("bar.sync 2, 288;" ::: "memory");

    if (warp < p.a_stages) {
      // ============================ A producers ============================
      const char* in_base = reinterpret_cast<const char*>(p.in);
      const size_t in_pitch = (size_t)p.Cin * PR::kElt;
      const uint32_t stage_addr = a_base + (uint32_t)warp * kWpAStage;
      struct Walk { int rbi, c, j; bool ok; int mb; };   // mb: THIS CTA's M block of the step (>= mb1: none)
      auto seek = [&](Walk& s) {
        while (s.rbi < n_rbi) {
          const uint32_t a = act_at(act, s.rbi);
          const int cnt = __popc(a);
          if (s.j < s.c + cnt) { s.mb = mb0 + 2 * (int)__fns(a, 0, s.j - s.c + 1) + (int)rank; s.ok = true; return; }
          s.c += cnt;
          ++s.rbi;
        }
        s.ok = false;
      };
      auto load_idx = [&](const Walk& s, int* idx) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const uint32_t d = s.mb < mb1 ? s_run[s.mb * 4 + r] : 63u;
          const int k = (int)(d & 63u);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int o = (rb0 + s.rbi) * kRows + i * 32 + lane;
            idx[r * NI + i] = (k != 63 && o < p.m_out) ? __ldg(p.nbr + (size_t)k * p.m_out + o) : -1;
          }
        }
      };
      auto gather = [&](int r, const int* idx_r, uint32_t d) {
        if ((d & 63u) == 63u) return;  // padding chunks: rows stay as they are (their accumulator rows are never read)
        const char* src_c = in_base + (size_t)(d >> 6) * PR::kRowBytes;
        const uint32_t dst_c = stage_addr + (uint32_t)(r * RM) * kWpChunkBlock;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
#pragma unroll
          for (int q = 0; q < PPR; ++q) {
            const uint32_t e = tab[q];
            const int src_row = __shfl_sync(0xffffffffu, idx_r[i], (int)(e & 31u));
            const char* src = src_c + (size_t)(src_row >= 0 ? src_row : 0) * in_pitch + ((e >> 5) & 127u) * 16u;
            cp_async_16(dst_c + (uint32_t)(i * 32 * PR::kRowBytes) + (e >> 12), src, src_row >= 0 ? 16u : 0u);
          }
        }
      };
      Walk cur;
      cur.rbi = 0; cur.c = 0; cur.j = warp; cur.ok = false; cur.mb = 0;
      seek(cur);
      int idx[NR * NI];
      if (cur.ok) load_idx(cur, idx);
      while (cur.ok) {
        Walk nxt = cur;
        nxt.j += p.a_stages;
        seek(nxt);
        int idx_n[NR * NI];
        if (nxt.ok) load_idx(nxt, idx_n);

        mbar_wait(a_empty(warp), a_phase ^ 1u);
#pragma unroll
        for (int r = 0; r < NR; ++r) gather(r, idx + r * NI, cur.mb < mb1 ? (uint32_t)s_run[cur.mb * 4 + r] : 63u);
        cp_async_mbar_arrive_noinc(a_full(warp));
        a_phase ^= 1u;
        cur = nxt;
#pragma unroll
        for (int i = 0; i < NR * NI; ++i) idx[i] = idx_n[i];
      }
    } else if (warp == p.a_stages) {
      // ============================ B producer: this CTA's half of the dout columns ============================
      if (lane == 0) {
        int bs = 0;
        const int c_first = p.col0 + (int)rank * (p.Cout / 2);
        for (int rbi = 0; rbi < n_rbi; ++rbi) {
          if (!act_at(act, rbi)) continue;
          const uint32_t dstb = b_base + (uint32_t)bs * b_stage_bytes;
          mbar_wait(b_empty(bs), ((b_phase >> bs) & 1u) ^ 1u);
          mbar_arrive_expect_tx(b_full(bs), (uint32_t)b_stage_bytes);
          for (int cbk = 0; cbk < p.nb_half; ++cbk)   // (the last tile may reach past this half: columns nobody multiplies)
            tma_load_2d(dstb + cbk * kWpChunkBlock, &tmap_dout, b_full(bs), c_first + cbk * 32, (rb0 + rbi) * kRows);
          b_phase ^= 1u << bs;
          bs ^= 1;
        }
      }
      __syncwarp();
    } else if (warp == kMmaWarp) {
      const uint32_t issue = elect_one() ? 1u : 0u;
      if (rank == 0) {
        // ============================ MMA issuer (rank 0) ============================
        const uint32_t idesc = PR::idesc(256, (uint32_t)p.Cout, 1, 1);  // M = 256 over the pair, both operands MN-major
        const uint64_t desc_hi = make_desc(0, kWpChunkBlock, 512, PR::kLayoutMN);
        mbar_wait(t_empty, m_t_phase ^ 1u);
        m_t_phase ^= 1u;
        tc_fence_after();
        uint32_t touched = 0;
        int as = 0, bs = 0;
        for (int rbi = 0; rbi < n_rbi; ++rbi) {
          uint32_t a = __reduce_or_sync(0xffffffffu, act_at(act, rbi));
          if (!a) continue;
          mbar_wait(b_full(bs), (m_b_phase >> bs) & 1u);
          mbar_wait(peer_b_full(bs), (m_b_phase >> bs) & 1u);
          const uint32_t b16 = (b_base + (uint32_t)bs * b_stage_bytes) >> 4;
          while (a) {
            const int i = __ffs(a) - 1;
            a &= a - 1u;
            mbar_wait(a_full(as), (m_a_phase >> as) & 1u);
            mbar_wait(peer_a_full(as), (m_a_phase >> as) & 1u);
            fence_proxy_async_smem();
            tc_fence_after();
            const uint32_t a16 = (a_base + (uint32_t)as * kWpAStage) >> 4;
            const uint32_t d = tmem_base + (uint32_t)(i * p.Cout);
            const uint32_t was = (touched >> i) & 1u;
#pragma unroll
            for (int r8 = 0; r8 < kMmaPerStep; ++r8) {
              const uint64_t adesc = desc_hi | (uint64_t)(a16 + 64u * r8);
              const uint64_t bdesc = desc_hi | (uint64_t)(b16 + 64u * r8);
              mma_bf16_pair_p(d, adesc, bdesc, idesc, (was || r8 > 0) ? 1u : 0u, issue);
            }
            mma_commit_pair_p(a_empty(as), issue);
            m_a_phase ^= 1u << as;
            if (++as == p.a_stages) as = 0;
            touched |= 1u << i;
          }
          mma_commit_pair_p(b_empty(bs), issue);
          m_b_phase ^= 1u << bs;
          bs ^= 1;
        }
        mma_commit_pair_p(t_full, issue);
      } else {
        // ============================ relay (rank 1) ============================
        int as = 0, bs = 0;
        for (int rbi = 0; rbi < n_rbi; ++rbi) {
          uint32_t a = __reduce_or_sync(0xffffffffu, act_at(act, rbi));
          if (!a) continue;
          mbar_wait(b_full(bs), (m_b_phase >> bs) & 1u);
          if (issue) mbar_arrive_cluster(map_to_rank(peer_b_full(bs), 0u));
          while (a) {
            a &= a - 1u;
            mbar_wait(a_full(as), (m_a_phase >> as) & 1u);
            fence_proxy_async_smem();
            if (issue) mbar_arrive_cluster(map_to_rank(peer_a_full(as), 0u));
            m_a_phase ^= 1u << as;
            if (++as == p.a_stages) as = 0;
          }
          m_b_phase ^= 1u << bs;
          bs ^= 1;
        }
      }
      __syncwarp();
    } else if (warp >= kNumProducerWarps && warp < kMmaWarp) {
      // ============================ epilogue: this CTA's M blocks ============================
      const int ew = warp & 3;
      uint32_t seen = 0;
      for (int rbi = lane; rbi < n_rbi; rbi += 32)
        seen |= p.tile_mask ? (p.tile_mask[rb0 + rbi] & all_taps) : all_taps;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, d);
      uint32_t touched = 0;   // accumulators of this CTA whose OWN M block met a neighbour (others hold zeros or nothing)
      for (int i = 0; i < n_pairs; ++i)
        if (mb_taps(mb0 + 2 * i + (int)rank) & seen && mb0 + 2 * i + (int)rank < mb1) touched |= 1u << i;
      mbar_wait_sleep(t_full, t_phase);
      tc_fence_after();
      while (touched) {
        const int i = __ffs(touched) - 1;
        touched &= touched - 1u;
        const int q = (mb0 + 2 * i + (int)rank) * 4 + ew;
        if (q >= p.nq) continue;
        const int k = q / p.ncc, cc = q - k * p.ncc;
        float* dst = p.dw + ((size_t)k * p.Cin + cc * 32 + lane) * p.dw_pitch + p.col0;
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(i * p.Cout);
        for (int c0 = 0; c0 < p.Cout; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + e), "f"(v[e]),
                         "f"(v[e + 1]), "f"(v[e + 2]), "f"(v[e + 3])
                         : "memory");
        }
      }
      tc_fence_before();
      if (rank == 0) mbar_arrive(t_empty);
      else mbar_arrive_cluster(map_to_rank(t_empty, 0u));
    }
    if (warp >= kNumProducerWarps && warp < kMmaWarp) t_phase ^= 1u;
  }
  cp_async_wait<0>();
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512u);
  }
}

bool make_rows_tile_map(CUtensorMap* map, const void* base, int64_t rows, int C, int box_rows, bool bf16);
extern std::atomic<long long> g_conv_path_counts[4];

// large bf16 maps, column groups whose halves are whole 16-column MMA steps (Cout % 32 == 0)
bool conv_wgrad_pair_eligible(int64_t m_out, int c_in, int c_grp, int K, bool bf16) {
  constexpr bool kPairByDefault = false;   // (knob 9 = 2 asks for it until the kernel is validated on the hardware)
  if (g_umma_dbg[9] == 1 || (g_umma_dbg[9] == 0 && !kPairByDefault)) return false;
  if (!bf16 || c_in % 32 || c_grp % 32 || c_grp > 256 || K > 32) return false;
  if (g_umma_dbg[9] == 2) return m_out >= 1;
  return m_out >= 128 * 4 * (kNumSMs / 2);
}

template <int RM>
static int launch_wgrad_pair(const UmmaWgradPairParams& p, const CUtensorMap& tmap, int grid, size_t smem, cudaStream_t stream) {
  auto kern = conv_wgrad_pair_kernel<RM>;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = (int)smem;
  }
  kern<<<grid, kNumThreads, smem, stream>>>(p, tmap);
  SPC_LAUNCHED("conv_wgrad_pair_kernel");
  return 0;
}

// one column group [col0, col0 + c_grp) of the c_out_full output channels (contract of conv_wgrad_umma_cols)
int conv_wgrad_pair_cols(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask, int64_t m_out,
                         int c_in, int c_out_full, int col0, int c_grp, int K, float* dw, cudaStream_t stream) {
  const int c_out = c_grp;
  UmmaWgradPairParams p;
  memset(&p, 0, sizeof(p));
  p.in = in; p.nbr = nbr; p.tile_mask = tile_mask; p.dw = dw;
  p.m_out = (int)m_out; p.Cin = c_in; p.Cout = c_out; p.K = K;
  p.dw_pitch = c_out_full; p.col0 = col0;
  p.ncc = c_in / 32;
  p.nq = K * p.ncc;
  p.n_mb = (p.nq + 3) / 4;
  SPC_REQUIRE(p.n_mb <= kWpMaxMb, "too many M blocks");
  const int cap = 512 / c_out;                       // accumulators per CTA: a pass covers 2 * cap M blocks
  p.n_pass = (p.n_mb + 2 * cap - 1) / (2 * cap);
  p.n_rb = (int)ceil_div(m_out, 128);
  int want_split = (2 * (kNumSMs / 2)) / p.n_pass;   // <= 2 work items per pair (static round-robin)
  if (want_split > p.n_rb) want_split = p.n_rb;
  if (want_split < 1) want_split = 1;
  p.rb_per_split = (p.n_rb + want_split - 1) / want_split;
  const int pairs_per_pass = ((p.n_mb + p.n_pass - 1) / p.n_pass + 1 + 1) / 2;   // (balanced passes differ by one M block)
  p.act16 = pairs_per_pass > 8 ? 1 : 0;
  const int max_rb = p.act16 ? kWpActBytes / 2 : kWpActBytes;
  if (p.rb_per_split > max_rb) p.rb_per_split = max_rb;
  p.n_split = (p.n_rb + p.rb_per_split - 1) / p.rb_per_split;
  p.n_work = p.n_split * p.n_pass;
  p.nb_half = (c_out / 2 + 31) / 32;
  const int b_stage_bytes = p.nb_half * kWpChunkBlock;
  int a_stages = (kSmemLimit - 1024 - 256 - kWpTabBytes - kWpBStages * b_stage_bytes) / kWpAStage;
  if (a_stages > kWpMaxAStages) a_stages = kWpMaxAStages;
  SPC_REQUIRE(a_stages >= 2, "wgrad pair tile does not fit in shared memory");
  p.a_stages = a_stages;
  const size_t smem = (size_t)a_stages * kWpAStage + (size_t)kWpBStages * b_stage_bytes + 1024 + 256 + kWpTabBytes;
  const int grid = 2 * std::min(p.n_work, kNumSMs / 2);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  SPC_REQUIRE(make_rows_tile_map(&tmap, dout, m_out, c_out_full, 128, true), "cuTensorMapEncodeTiled unavailable");
  g_conv_path_counts[0].fetch_add(1, std::memory_order_relaxed);
  int rm = 1;
  if (g_umma_dbg[5] != 1) rm = (p.ncc % 4 == 0) ? 4 : (p.ncc % 2 == 0 ? 2 : 1);
  switch (rm) {
    case 4: return launch_wgrad_pair<4>(p, tmap, grid, smem, stream);
    case 2: return launch_wgrad_pair<2>(p, tmap, grid, smem, stream);
    default: return launch_wgrad_pair<1>(p, tmap, grid, smem, stream);
  }
}

}  // namespace spc
