// conv_wgrad_umma.cu — sparse-convolution weight gradient on tcgen05:
//
//     dW[k][ci][co] = sum_o in[nbr[k,o]][ci] * dout[o][co]
//
// GEMM view: D[M x N] += A[M x Kred] * B[Kred x N] with the REDUCTION over out rows o:
//   M = 128 = four 32-channel "chunks", chunk q = (kernel offset k = q / ncc, channel group cc = q % ncc) — so for
//       Cin = 32 four different offsets share one MMA, for Cin = 128 one offset fills it;
//   N = Cout;  Kred = 8 (tf32) / 16 (bf16) rows per tcgen05.mma.
// Both operands are MN-major: a gathered input row IS 32 consecutive M elements, a dout row IS Cout consecutive N
// elements, so rows are laid down as [rows x row-chunk] blocks like in the forward kernel.  MN-major 32-bit
// operands must use SWIZZLE_128B_BASE32B (32-byte chunks XOR row & 3), bf16 uses SWIZZLE_64B; LBO = distance
// between 32-element column groups, SBO = distance between 4- (tf32) / 8-row (bf16) groups = 512 B either way.
// A pipeline step covers kRows = 64 (tf32) / 128 (bf16) out rows of ONE M block: 4 chunk blocks of 8 KB.  TMEM holds
// floor(512 / Cout) accumulators; a work item = (row range, pass over a group of M blocks) and ends with an fp32
// red.global.add of its partial dW — the only atomics of the convolution path.
//
// Roles (512 threads, one persistent CTA per SM; see the role-layout constants below — since the end of r2 two producer
// warps share an A ring slot when a step has two or more row visits):
//   warps 0..             A producers: step j of an item is gathered by the warp(s) of slot j % a_stages.  Chunks of
//                        an M block that belong to the same offset are consecutive channels of the same input row,
//                        so they are fetched by ONE row visit of RM x 64 contiguous bytes (bf16; the L1 handles an
//                        LDGSTS one 128-byte line at a time: a 128-byte visit costs one line where two 64-byte
//                        visits cost two);
//   warp 10              dout row blocks by TMA tile loads (hardware swizzle = the MN-major UMMA layout), 2 slots;
//   warps 11-14          epilogue: tcgen05.ld + red.global.add.v4 of the item's partial dW;
//   warp 15              the MMA issuer (tcgen05.mma / commit under an issue predicate true in one lane).
// Which (row block, M block) steps exist is data (the per-128-row offset masks: offsets without a neighbour in a
// row block are skipped — on the faithful ScanNet geometry 26 of 27).  r1 walked the masks from global memory in
// every role (each producer warp walked ALL steps to find its own: ~3000 cycles of dependent loads per step with
// the gather and the MMAs switched off: 0.33 of the kernel's 0.73 ms at 96->96, 1 M voxels); now the item's masks
// are reduced ONCE, by the producer and MMA warps, to a table of "active M blocks per row block" in shared memory
// and every role enumerates its steps from that table with a popcount / find-nth-set-bit per row block
// (96->96: 0.726 -> 0.554 ms, 64->64 0.427 -> 0.343, 32->32 0.228 -> 0.175).
#include "umma_common.cuh"

namespace spc {

constexpr int kWgChunkBlock = 8192;              // [kRows x row chunk]
constexpr int kWgAStage = 4 * kWgChunkBlock;     // 32 KB: four chunks = one 128-row M block
constexpr int kWgMaxAStages = 6;
constexpr int kWgBStages = 2;
constexpr int kWgMaxMb = 128;    // M blocks (K * Cin / 128): 27 offsets x 512 channels = 108
// shared memory behind the rings: barriers (256 B), run descriptors (uint16 [kWgMaxMb][4]) and two step tables of
// kWgActBytes each (one byte per row block when a pass holds <= 8 M blocks, i.e. Cout >= 64, else two).  1792 B in
// all: with more, Cout = 128 loses its fifth and Cout = 256 its third A ring slot (measured: 256->256 2.68 -> 3.03 ms).
constexpr int kWgActBytes = 384;
constexpr int kWgTabBytes = kWgMaxMb * 4 * 2 + 2 * kWgActBytes;
// role layout (16 warps = 512 threads x 128 registers): warps 0..9 A producers — TWO per ring slot when a step has at
// least two row visits (each warp gathers half of the visits; at most 5 slots then), warp 10 the dout producer, warps
// 11..14 the epilogue (warp & 3 = its TMEM lane group), warp 15 the MMA issuer.  r2 had one warp per slot: a producer's
// step is latency — index loads, 64 dependent LDGSTS, the slot wait — and five warps in flight needed ~950 clocks per
// step at 96->96 where shared memory allows ~730.  Measured (1 M voxels, bf16, profiles/r2_row_order_experiments.md
// section 6): 96->96 0.533 -> 0.497 ms, 32->32 0.167 -> 0.161, tf32 96->96 1.054 -> 0.972.  A first version with 18
// warps was capped at 96 registers, spilled, and was 7 % SLOWER; splitting the ROWS of every visit between the two
// warps instead of the visits was 1-3 % slower than this.
constexpr int kWgAWarps = 10;
constexpr int kWgBWarp = kWgAWarps;
constexpr int kWgEpiWarp0 = kWgBWarp + 1;
constexpr int kWgMmaWarp = kWgEpiWarp0 + 4;
constexpr int kWgThreads = (kWgMmaWarp + 1) * 32;
constexpr int kWgTableThreads = (kWgBWarp + 2) * 32;   // producers + MMA warp build and read the step tables

struct UmmaWgradParams {
  const void* in;             // [m_in, Cin] fp32 or bf16
  const int* nbr;             // [K, m_out]
  const uint32_t* tile_mask;  // [ceil(m_out/128)] or null
  float* dw;                  // [K, Cin, dw_pitch] (+ column offset applied), zeroed (or holding the gradient to add to)
  int m_out, Cin, Cout, K;    // Cout = width of THIS launch's column group (<= 256)
  int dw_pitch, col0;         // full output-channel count / first column of the group (dout tile loads start there)
  int ncc, nq;                // chunks per offset, total chunks
  int n_mb, n_pass;
  int n_rb, rb_per_split, n_split;
  int a_stages;               // A ring slots (= A producer warps)
  int act16;                  // step-table entries are 16 bits (more than 8 M blocks per pass)
  int n_work;
};

// RM: chunks per row visit (1, 2 or 4; ncc % RM == 0): RM x 64 contiguous bytes of a bf16 input row per LDGSTS
// row visit.  Measured (1 M voxels, bf16): 64->64 0.351 -> 0.343 ms with RM = 2, 128->128 0.773 -> 0.757 with RM = 4;
// mixed visits for Cin = 96 ((3 + 1) / (2 + 2) / (1 + 3) chunks of two offsets) were SLOWER (0.554 -> 0.698) and
// are not built.
template <bool BF16, int RM>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_umma_kernel(const UmmaWgradParams p, const __grid_constant__ CUtensorMap tmap_dout) {
  using PR = Prec<BF16>;
  constexpr int kRows = kWgChunkBlock / PR::kRowBytes;      // 64 (tf32) / 128 (bf16) rows per step
  constexpr int kMmaPerStep = kRows / (BF16 ? 16 : 8);      // 8 either way, 1024 B of rows each
  constexpr int LPR = PR::kLanesPerRow;
  constexpr int NI = kRows / 32;                            // neighbour indices per lane and row visit
  constexpr int NR = 4 / RM;                                // row visits per M block
  constexpr int WPA = NR >= 2 ? 2 : 1;                      // producer warps per A ring slot
  constexpr int NRW = NR / WPA;                             // row visits per producer warp and step
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
  uint8_t* tab_raw = smem_raw + (bar_base + 256u - smem_u32(smem_raw));
  uint16_t* s_run = reinterpret_cast<uint16_t*>(tab_raw);             // [kWgMaxMb][4] row visits: k | first chunk << 6
  uint8_t* s_act = tab_raw + kWgMaxMb * 4 * 2;                        // [2][kWgActBytes] active M blocks per row block
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // (REDUX: a warp index ptxas knows to be warp-uniform — the role branches are then uniform control flow and the MMA
  // warp's loop state lives in uniform registers, see conv_umma_kernel)
  const int warp = (int)__reduce_min_sync(0xffffffffu, threadIdx.x >> 5), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // full barriers: the 32 lanes of the one warp that fills the stage (cp.async ... arrive.noinc)
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full(s), 32 * WPA); mbar_init(a_empty(s), 1); }
    // b_full: one expect_tx arrival (TMA tile loads of the dout rows)
    for (int s = 0; s < kWgBStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    mbar_init(t_full, 1);
    mbar_init(t_empty, kNumEpilogueThreads);
    fence_mbar_init();
  }
  if (warp == kWgMmaWarp) { tmem_alloc(tmem_slot, 512u); tmem_relinquish(); }
  // row visit r of M block mb covers chunks 4 mb + r RM .. + RM - 1, all of one offset (ncc % RM == 0)
  for (int e = threadIdx.x; e < p.n_mb * 4; e += blockDim.x) {
    const int mb = e >> 2, r = e & 3;
    const int q = mb * 4 + r * RM;
    s_run[e] = (uint16_t)((r < NR && q < p.nq) ? ((q / p.ncc) | ((q % p.ncc) << 6)) : 63);  // 63: nothing to gather
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __reduce_min_sync(0xffffffffu, *tmem_slot_ptr);   // (REDUX: a provably uniform value)
  const uint32_t all_taps = p.K >= 32 ? 0xFFFFFFFFu : ((1u << p.K) - 1u);

  // ---- per-lane constants of the row visits (producer warps): instruction q of a 32-row group covers flat
  // pieces q * 32 + lane -> row rr of the group | piece of the visit | shared-memory offset, packed in one register
  constexpr int PPR = RM * LPR;
  uint32_t tab[PPR];
#pragma unroll
  for (int q = 0; q < PPR; ++q) {
    const int f = q * 32 + lane;
    const int rr = f / PPR, piece = f - rr * PPR;
    const int c = piece / LPR, j = piece - c * LPR;
    const uint32_t dsto = (uint32_t)(c * kWgChunkBlock + rr * PR::kRowBytes) + PR::swz_mn(j, rr);
    tab[q] = (uint32_t)rr | ((uint32_t)piece << 5) | (dsto << 12);
  }
  // offsets (bit mask) of M block mb = union over its row visits
  auto mb_taps = [&](int mb) -> uint32_t {
    uint32_t t = 0;
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

  uint32_t a_phase = 0;        // producer warp: parity of ITS ring slot
  uint32_t b_phase = 0;        // B warp: bit s = parity of slot s
  uint32_t t_phase = 0;        // epilogue warps
  // the MMA warp's own copies (one parity bit per A slot / B slot): kept apart from the other roles' so that they are
  // provably warp-uniform values
  uint32_t m_a_phase = 0, m_b_phase = 0, m_t_phase = 0;
  int local_item = 0;
  for (int w = blockIdx.x; w < p.n_work; w += gridDim.x, ++local_item) {
    const int split = w / p.n_pass, pass = w - split * p.n_pass;
    const int mb0 = pass * p.n_mb / p.n_pass, mb1 = (pass + 1) * p.n_mb / p.n_pass;  // balanced passes
    const int rb0 = split * p.rb_per_split, rb1 = min(rb0 + p.rb_per_split, p.n_rb);
    const int n_rbi = rb1 - rb0;
    uint8_t* act = s_act + (local_item & 1) * kWgActBytes;
    // ---- step table of the item: bit i of act[rbi] = M block mb0 + i has a neighbour in row block rb0 + rbi ----
    const bool tabler = warp <= kWgBWarp || warp == kWgMmaWarp;
    const int filler = warp == kWgMmaWarp ? (kWgBWarp + 1) * 32 + lane : threadIdx.x;  // 0..383 among the 12 warps
    for (int rbi = filler; rbi < n_rbi && tabler; rbi += kWgTableThreads) {
      const uint32_t m = p.tile_mask ? (p.tile_mask[((rb0 + rbi) * kRows) >> 7] & all_taps) : all_taps;
      uint32_t a = 0;
      for (int mb = mb0; mb < mb1; ++mb)
        if (mb_taps(mb) & m) a |= 1u << (mb - mb0);
      if (p.act16) reinterpret_cast<uint16_t*>(act)[rbi] = (uint16_t)a;
      else act[rbi] = (uint8_t)a;
    }
    // (filled and read by the producer and MMA warps only — the epilogue warps take no part, so the next item's
    // gather starts while they still drain this item's accumulators.  Two tables alternate: a role that runs ahead
    // fills the next item's table while a slower role may still read this one; every one of the 288 threads has
    // passed this barrier before any of them reaches the one after the next.)
    if (tabler) asm volatile("bar.sync 2, %0;" ::"n"(kWgTableThreads) : "memory");

    if (warp < kWgAWarps && warp / WPA < p.a_stages) {
      const int aslot = warp / WPA, ahalf = warp - aslot * WPA;   // this warp gathers row visits ahalf * NRW .. + NRW - 1
      // ============================ A producers ============================
      const char* in_base = reinterpret_cast<const char*>(p.in);
      const size_t in_pitch = (size_t)p.Cin * PR::kElt;
      const uint32_t stage_addr = a_base + (uint32_t)aslot * kWgAStage;
      // walker over this warp's steps j = warp, warp + a_stages, ... of the item
      struct Walk { int rbi, c, j; bool ok; int mb; };
      auto seek = [&](Walk& s) {  // position on step s.j (>= the steps before row block s.rbi = s.c)
        while (s.rbi < n_rbi) {
          const uint32_t a = act_at(act, s.rbi);
          const int cnt = __popc(a);
          if (s.j < s.c + cnt) { s.mb = mb0 + (int)__fns(a, 0, s.j - s.c + 1); s.ok = true; return; }
          s.c += cnt;
          ++s.rbi;
        }
        s.ok = false;
      };
      auto load_idx = [&](const Walk& s, int* idx) {
#pragma unroll
        for (int rw = 0; rw < NRW; ++rw) {
          const uint32_t d = s_run[s.mb * 4 + ahalf * NRW + rw];
          const int k = (int)(d & 63u);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int o = (rb0 + s.rbi) * kRows + i * 32 + lane;
            idx[rw * NI + i] = (k != 63 && o < p.m_out) ? __ldg(p.nbr + (size_t)k * p.m_out + o) : -1;
          }
        }
      };
      // row visit r of the step: rows of the row block x RM x 32 channels into chunk blocks r RM .. r RM + RM - 1
      auto gather = [&](int r, const int* idx_r, uint32_t d) {
        if ((d & 63u) == 63u) return;  // padding chunks of the last M block: rows stay as they are
        const char* src_c = in_base + (size_t)(d >> 6) * PR::kRowBytes;
        const uint32_t dst_c = stage_addr + (uint32_t)(r * RM) * kWgChunkBlock;
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
      cur.rbi = 0; cur.c = 0; cur.j = aslot; cur.ok = false; cur.mb = 0;
      seek(cur);
      int idx[NRW * NI];
      if (cur.ok) load_idx(cur, idx);
      while (cur.ok) {
        Walk nxt = cur;
        nxt.j += p.a_stages;
        seek(nxt);
        int idx_n[NRW * NI];
        if (nxt.ok) load_idx(nxt, idx_n);  // latency hides behind this stage's slot wait

        mbar_wait(a_empty(aslot), a_phase ^ 1u);
#pragma unroll
        for (int rw = 0; rw < NRW; ++rw) {
          const int r = ahalf * NRW + rw;
          gather(r, idx + rw * NI, s_run[cur.mb * 4 + r]);
        }
        cp_async_mbar_arrive_noinc(a_full(aslot));
        a_phase ^= 1u;
        cur = nxt;
#pragma unroll
        for (int i = 0; i < NRW * NI; ++i) idx[i] = idx_n[i];
      }
    } else if (warp == kWgBWarp) {
      // ============================ B producer ============================
      // the dout rows of one row block (contiguous rows, all Cout channels): Cout / 32 tile loads (32 channels x
      // kRows rows, hardware swizzle = the MN-major UMMA layout) issued by one lane; rows past m_out are zero-filled
      if (lane == 0) {
        int bs = 0;
        for (int rbi = 0; rbi < n_rbi; ++rbi) {
          if (!act_at(act, rbi)) continue;
          const uint32_t dstb = b_base + (uint32_t)bs * b_stage_bytes;
          mbar_wait(b_empty(bs), ((b_phase >> bs) & 1u) ^ 1u);
          mbar_arrive_expect_tx(b_full(bs), (uint32_t)b_stage_bytes);
          for (int cbk = 0; cbk < p.Cout / 32; ++cbk)
            tma_load_2d(dstb + cbk * kWgChunkBlock, &tmap_dout, b_full(bs), p.col0 + cbk * 32, (rb0 + rbi) * kRows);
          b_phase ^= 1u << bs;
          bs ^= 1;
        }
      }
      __syncwarp();
    } else if (warp == kWgMmaWarp) {
      // ============================ MMA issuer ============================
      // the whole warp runs the loop in uniform control flow; the tcgen05 instructions carry an issue predicate that
      // is true in one elected lane (see conv_umma_kernel)
      {
        const uint32_t issue = elect_one() ? 1u : 0u;
        const uint32_t idesc = PR::idesc(128, (uint32_t)p.Cout, 1, 1);  // both operands MN-major
        const uint64_t desc_hi = make_desc(0, kWgChunkBlock, 512, PR::kLayoutMN);
        mbar_wait(t_empty, m_t_phase ^ 1u);  // epilogue drained the previous item's accumulators
        m_t_phase ^= 1u;
        tc_fence_after();
        uint32_t touched = 0;
        int as = 0, bs = 0;
        for (int rbi = 0; rbi < n_rbi; ++rbi) {
          uint32_t a = __reduce_or_sync(0xffffffffu, act_at(act, rbi));   // (the same value in every lane: uniform)
          if (!a) continue;
          mbar_wait(b_full(bs), (m_b_phase >> bs) & 1u);
          const uint32_t b16 = (b_base + (uint32_t)bs * b_stage_bytes) >> 4;
          while (a) {
            const int i = __ffs(a) - 1;
            a &= a - 1u;
            mbar_wait(a_full(as), (m_a_phase >> as) & 1u);
            fence_proxy_async_smem();  // cp.async (generic proxy) writes -> tensor-core (async proxy) reads
            tc_fence_after();
            const uint32_t a16 = (a_base + (uint32_t)as * kWgAStage) >> 4;
            const uint32_t d = tmem_base + (uint32_t)(i * p.Cout);
            const uint32_t was = (touched >> i) & 1u;
#pragma unroll
            for (int r8 = 0; r8 < kMmaPerStep; ++r8) {
              // descriptors differ only in the start address field: 1024 B of rows per MMA = 64 units
              const uint64_t adesc = desc_hi | (uint64_t)(a16 + 64u * r8);
              const uint64_t bdesc = desc_hi | (uint64_t)(b16 + 64u * r8);
              PR::mma_p(d, adesc, bdesc, idesc, (was || r8 > 0) ? 1u : 0u, issue);
            }
            mma_commit_p(a_empty(as), issue);
            m_a_phase ^= 1u << as;
            if (++as == p.a_stages) as = 0;
            touched |= 1u << i;
          }
          mma_commit_p(b_empty(bs), issue);
          m_b_phase ^= 1u << bs;
          bs ^= 1;
        }
        mma_commit_p(t_full, issue);
      }
      __syncwarp();
    } else if (warp >= kWgEpiWarp0 && warp < kWgMmaWarp) {
      // ============================ epilogue ============================
      const int ew = warp & 3;
      // which accumulators received at least one MMA (same rule as the step table: the offsets of the M block
      // meet the offsets present somewhere in the row range)
      uint32_t seen = 0;
      for (int rbi = lane; rbi < n_rbi; rbi += 32)
        seen |= p.tile_mask ? (p.tile_mask[((rb0 + rbi) * kRows) >> 7] & all_taps) : all_taps;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) seen |= __shfl_xor_sync(0xffffffffu, seen, d);
      uint32_t touched = 0;
      for (int mb = mb0; mb < mb1; ++mb)
        if (mb_taps(mb) & seen) touched |= 1u << (mb - mb0);
      mbar_wait_sleep(t_full, t_phase);
      tc_fence_after();
      while (touched) {
        const int i = __ffs(touched) - 1;
        touched &= touched - 1u;
        const int q = (mb0 + i) * 4 + ew;
        if (q >= p.nq) continue;  // padding chunk of the last M block (warp-uniform)
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
      mbar_arrive(t_empty);
    }
    if (warp >= kWgEpiWarp0 && warp < kWgMmaWarp) t_phase ^= 1u;
    // the A / B slot parities of a role that did nothing this item stay as they are; the MMA thread's per-slot
    // bits advanced exactly as the producers' own counters did (every filled slot was consumed)
  }
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == kWgMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

bool umma_wgrad_supported(int c_in, int c_out) {
  // (more than 256 output channels: column groups of <= 256, one launch each — ResNet14's 256->512 / 512->512)
  return c_in >= 32 && c_in % 32 == 0 && c_out >= 32 && c_out % 32 == 0 && c_out <= 1024;
}
int64_t umma_wgrad_workspace(int, int, int) { return 256; }

template <bool BF16, int RM>
static int launch_wgrad_umma(const UmmaWgradParams& p, const CUtensorMap& tmap, int grid, size_t smem,
                             cudaStream_t stream) {
  auto kern = conv_wgrad_umma_kernel<BF16, RM>;
  static int smem_set = 0;
  if ((int)smem > smem_set) {
    SPC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = (int)smem;
  }
  kern<<<grid, kWgThreads, smem, stream>>>(p, tmap);
  SPC_LAUNCHED("conv_wgrad_umma_kernel");
  return 0;
}

extern std::atomic<long long> g_conv_path_counts[4];
bool conv_wgrad_pair_eligible(int64_t m_out, int c_in, int c_grp, int K, bool bf16);
int conv_wgrad_pair_cols(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask, int64_t m_out,
                         int c_in, int c_out_full, int col0, int c_grp, int K, float* dw, cudaStream_t stream);
int conv_wgrad_umma_cols(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask,
                         int64_t m_out, int c_in, int c_out_full, int col0, int c_grp, int K, bool bf16, float* dw,
                         cudaStream_t stream);

// dw: [K, Cin, Cout] fp32.  zero_dw: clear it first (plain gradient); otherwise the partial sums are ADDED to what
// it holds (the trainer's gradient arena, zeroed once per step: no separate accumulation pass).
int conv_wgrad_umma(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask,
                    int64_t m_out, int c_in, int c_out, int K, bool bf16, float* dw, bool zero_dw,
                    cudaStream_t stream) {
  SPC_REQUIRE(umma_wgrad_supported(c_in, c_out), "shape not supported by the tcgen05 wgrad path");
  SPC_REQUIRE(K <= 32, "tcgen05 path supports kernel volume <= 32");
  SPC_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)dout % 16) == 0 && ((uintptr_t)dw % 16) == 0,
              "rows must be 16-byte aligned");
  if (zero_dw) SPC_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (m_out == 0) return 0;
  if (c_out > 256) {  // column groups: the widest multiple of 32 <= 256 that divides c_out evenly enough
    int groups = (c_out + 255) / 256;
    while (c_out % groups || (c_out / groups) % 32) ++groups;
    const int cg = c_out / groups;
    for (int g = 0; g < groups; ++g) {
      int rc = conv_wgrad_umma_cols(in, dout, nbr, tile_mask, m_out, c_in, c_out, g * cg, cg, K, bf16, dw, stream);
      if (rc) return rc;
    }
    return 0;
  }
  return conv_wgrad_umma_cols(in, dout, nbr, tile_mask, m_out, c_in, c_out, 0, c_out, K, bf16, dw, stream);
}

// one column group [col0, col0 + c_grp) of the c_out_full output channels
int conv_wgrad_umma_cols(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask,
                         int64_t m_out, int c_in, int c_out_full, int col0, int c_grp, int K, bool bf16, float* dw,
                         cudaStream_t stream) {
  // large bf16 maps: CTA pairs (conv_wgrad_umma_pair.cu) — each CTA stages half of the dout columns, twice the M
  // blocks per pass
  if (conv_wgrad_pair_eligible(m_out, c_in, c_grp, K, bf16))
    return conv_wgrad_pair_cols(in, dout, nbr, tile_mask, m_out, c_in, c_out_full, col0, c_grp, K, dw, stream);
  const int c_out = c_grp;
  const int rows = bf16 ? 128 : 64;  // out rows per pipeline step
  UmmaWgradParams p;
  p.in = in; p.nbr = nbr; p.tile_mask = tile_mask; p.dw = dw;
  p.m_out = (int)m_out; p.Cin = c_in; p.Cout = c_out; p.K = K;
  p.dw_pitch = c_out_full; p.col0 = col0;
  p.ncc = c_in / 32;
  p.nq = K * p.ncc;
  p.n_mb = (p.nq + 3) / 4;
  SPC_REQUIRE(p.n_mb <= kWgMaxMb, "too many M blocks");
  int cap = 512 / c_out;                       // accumulators that fit in TMEM
  p.n_pass = (p.n_mb + cap - 1) / cap;
  p.n_rb = (int)ceil_div(m_out, rows);
  int want_split = (2 * kNumSMs) / p.n_pass;  // <= 2 work items per CTA (static round-robin)
  if (want_split > p.n_rb) want_split = p.n_rb;
  if (want_split < 1) want_split = 1;
  p.rb_per_split = (p.n_rb + want_split - 1) / want_split;
  p.act16 = (p.n_mb + p.n_pass - 1) / p.n_pass > 8 ? 1 : 0;
  const int max_rb = p.act16 ? kWgActBytes / 2 : kWgActBytes;
  if (p.rb_per_split > max_rb) p.rb_per_split = max_rb;  // (step-table capacity: more, shorter items)
  if (!bf16 && (p.rb_per_split & 1)) ++p.rb_per_split;  // keep splits aligned to 128-row mask tiles
  p.n_split = (p.n_rb + p.rb_per_split - 1) / p.rb_per_split;
  p.n_work = p.n_split * p.n_pass;
  const int b_stage_bytes = (c_out / 32) * kWgChunkBlock;
  int a_stages = (kSmemLimit - 1024 - 256 - kWgTabBytes - kWgBStages * b_stage_bytes) / kWgAStage;
  if (a_stages > kWgMaxAStages) a_stages = kWgMaxAStages;  // one producer warp per A stage + the B warp <= 8 warps
  SPC_REQUIRE(a_stages >= 2, "wgrad tile does not fit in shared memory");
  // row-visit mode: bf16 rows are 64 B per chunk, so visits of 2 / 4 chunks touch fewer 128-byte lines; tf32 chunks are
  // whole lines already.  Two producer warps share a slot when a step has >= 2 visits: 10 A warps = 5 slots then
  int rm = 1;
  if (bf16 && g_umma_dbg[5] != 1) rm = (p.ncc % 4 == 0) ? 4 : (p.ncc % 2 == 0 ? 2 : 1);
  const int wpa = (4 / rm >= 2) ? 2 : 1;
  if (a_stages > kWgAWarps / wpa) a_stages = kWgAWarps / wpa;
  p.a_stages = a_stages;
  const size_t smem = (size_t)a_stages * kWgAStage + (size_t)kWgBStages * b_stage_bytes + 1024 + 256 + kWgTabBytes;
  const int grid = p.n_work < kNumSMs ? p.n_work : kNumSMs;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  SPC_REQUIRE(make_rows_tile_map(&tmap, dout, m_out, c_out_full, rows, bf16), "cuTensorMapEncodeTiled unavailable");
  g_conv_path_counts[bf16 ? 0 : 1].fetch_add(1, std::memory_order_relaxed);
  if (!bf16) return launch_wgrad_umma<false, 1>(p, tmap, grid, smem, stream);
  switch (rm) {
    case 4: return launch_wgrad_umma<true, 4>(p, tmap, grid, smem, stream);
    case 2: return launch_wgrad_umma<true, 2>(p, tmap, grid, smem, stream);
    default: return launch_wgrad_umma<true, 1>(p, tmap, grid, smem, stream);
  }
}

}  // namespace spc
