// coords.cu — voxel quantisation, coordinate hashing, stride maps and kernel maps.
//
// Replaces (behaviourally) MinkowskiEngine's CoordinateMapManager for the call
// sites the reference uses:
//   TensorField.sparse()            co3d_3d/src/models/mink/resnet.py:164, res16unet.py:392
//   CoordinateManager.stride()      co3d_3d/src/models/mink/modules/sparse_conv.py:403-405
//   CoordinateManager.kernel_map()  co3d_3d/src/models/mink/modules/sparse_conv.py:90-96,197-204
//
// Data layout in HBM
//   coordinate map : int32 [M,4] rows (b,x,y,z), 16 B per row (one LDG.128)
//   hash table     : 16-byte slots {u64 key, u32 first_row, u32 map_row}, two slots per
//                    32-byte bucket (= one DRAM/L2 sector), linear probing over buckets,
//                    load factor <= 0.5.  A probe costs one sector; key and value arrive
//                    together so a hit never needs a second dependent load.
//   kernel map     : int32 [K, M_out] offset-major, -1 = no neighbour.  Offset-major makes
//                    both the builder's writes and the convolution's per-offset reads
//                    fully coalesced.
//
// Row order is canonical (ME CPU semantics): a voxel's row is the rank of its first
// occurrence in the source, so results are deterministic however the atomics race.
#include "common.cuh"

namespace spc {

thread_local char g_last_error[512] = {0};
std::atomic<long long> g_launch_count{0};

constexpr int kScanThreads = 1024;

// ---------------------------------------------------------------------------
// 1. insert: quantise, pack, claim a slot, remember the smallest source row
// ---------------------------------------------------------------------------
// A slot is claimed with ONE 128-bit compare-and-swap {empty} -> {key, first = i, row = unset} (atom.cas.b128, sm_90+):
// for a source without duplicates (what the plenoxel loaders deliver) that is the only L2 transaction of the point
// besides its coalesced reads / writes — r1 / r2 spent a volatile read, a 64-bit CAS and an atomicMin on it.  A failed
// CAS returns the slot: the same key -> atomicMin on `first` (only if the snapshot is larger), another key -> next slot.
__device__ __forceinline__ void cas_slot(Slot* s, unsigned long long new_lo, unsigned long long new_hi,
                                         unsigned long long& old_lo, unsigned long long& old_hi) {
  const unsigned long long ones = kEmptyKey;
  asm volatile(
      "{\n\t"
      ".reg .b128 c, n, o;\n\t"
      "mov.b128 c, {%2, %2};\n\t"
      "mov.b128 n, {%3, %4};\n\t"
      "atom.relaxed.gpu.global.cas.b128 o, [%5], c, n;\n\t"
      "mov.b128 {%0, %1}, o;\n\t"
      "}"
      : "=l"(old_lo), "=l"(old_hi)
      : "l"(ones), "l"(new_lo), "l"(new_hi), "l"(s)
      : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// PEEK: look at the slot before trying to claim it — a stride map sees every parent voxel ~5 times, so most points
// find their key already there and an atomicMin (often not even that: `first` is read with the key) is all they need.
template <int SRC>
__global__ void __launch_bounds__(256)
insert_kernel(const void* __restrict__ src, int n, const int* __restrict__ n_dev, int3 ts, Slot* slots,
              unsigned long long bucket_mask, int* __restrict__ slot_of, int* status,
              int* __restrict__ out_count, unsigned long long* __restrict__ scan_state, int n_state) {
  constexpr bool PEEK = SRC == SPC_SRC_STRIDE;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  // this launch also clears what the row-assignment kernel behind it accumulates into: the per-voxel multiplicities
  // and the look-back words of its single-pass scan (n_state words: ticket counter + one per block)
  if (i < n) out_count[i] = 0;
  if (i < n_state) scan_state[i] = 0ull;
  if (n_dev) n = min(n, *n_dev);  // row count produced on the device by the previous level
  if (i >= n) return;
  int b, x, y, z;
  if (SRC == SPC_SRC_FLOAT) {
    float4 c = reinterpret_cast<const float4*>(src)[i];
    b = (int)floorf(c.x);
    if (ts.x == 1 && ts.y == 1 && ts.z == 1) {
      x = (int)floorf(c.y);
      y = (int)floorf(c.z);
      z = (int)floorf(c.w);
    } else {
      x = (int)(floorf(c.y / (float)ts.x) * (float)ts.x);
      y = (int)(floorf(c.z / (float)ts.y) * (float)ts.y);
      z = (int)(floorf(c.w / (float)ts.z) * (float)ts.z);
    }
    // reject NaN/inf and values whose float->int conversion saturates
    if (!(fabsf(c.x) < 1e9f && fabsf(c.y) < 1e9f && fabsf(c.z) < 1e9f && fabsf(c.w) < 1e9f)) b = -1;
  } else {
    int4 c = reinterpret_cast<const int4*>(src)[i];
    b = c.x;
    if (SRC == SPC_SRC_STRIDE) {
      x = floor_div(c.y, ts.x) * ts.x;
      y = floor_div(c.z, ts.y) * ts.y;
      z = floor_div(c.w, ts.z) * ts.z;
    } else {
      x = c.y; y = c.z; z = c.w;
    }
  }
  unsigned long long key;
  if (!pack_key(b, x, y, z, key)) {
    status[1] = 1;
    slot_of[i] = -1;
    return;
  }
  const unsigned long long mine_hi = (unsigned long long)(unsigned)i | 0xFFFFFFFF00000000ull;  // first = i, row unset
  unsigned long long bucket = hash_key(key, bucket_mask);
  for (;;) {
    Slot* s = slots + 2 * bucket;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      unsigned long long k, hi;
      bool have = false;
      if (PEEK) {
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(k), "=l"(hi) : "l"(&s[j]) : "memory");
        have = k != kEmptyKey;
      }
      if (!have) {
        cas_slot(&s[j], key, mine_hi, k, hi);
        if (k == kEmptyKey) {   // claimed (an empty key means an empty slot: the other fields change only under a key)
          slot_of[i] = (int)(2 * bucket + j);
          return;
        }
      }
      if (k == key) {
        if ((unsigned)hi > (unsigned)i) atomicMin(&s[j].first, (unsigned)i);   // (`first` only ever decreases)
        slot_of[i] = (int)(2 * bucket + j);
        return;
      }
    }
    bucket = (bucket + 1) & bucket_mask;
  }
}

// block-wide exclusive prefix of a predicate using warp ballots
__device__ __forceinline__ int block_ballot_prefix(bool flag, int* warp_sum /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned bal = __ballot_sync(0xffffffffu, flag);
  int lane_prefix = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_sum[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int w = warp_sum[lane], x = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    warp_sum[lane] = x - w;  // exclusive
  }
  __syncthreads();
  return warp_sum[warp] + lane_prefix;
}

// single-block exclusive scan (in place) of `nb` ints, total -> *total_out (pair-list export)
__global__ void __launch_bounds__(kScanThreads)
scan_blocks_kernel(int* __restrict__ data, int nb, int* __restrict__ total_out) {
  __shared__ int warp_sum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < nb; base += kScanThreads) {
    int idx = base + threadIdx.x;
    int v = idx < nb ? data[idx] : 0;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int w = warp_sum[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += y;
      }
      warp_sum[lane] = w;
    }
    __syncthreads();
    int prefix = carry + (warp > 0 ? warp_sum[warp - 1] : 0) + x - v;
    if (idx < nb) data[idx] = prefix;
    int total = warp_sum[31];
    __syncthreads();
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

// 2. rows in first-occurrence order, coordinates, unique_index, inverse map and multiplicities in ONE pass
// (r1 / r2: count_first + scan_blocks + assign_rows + inverse kernels, each with its own random read of the slots).
// A block takes a ticket t (so that lower tickets are always resident or done), owns source rows [1024 t, 1024 t + 1024),
// counts its first occurrences and gets the number of those in front of it by decoupled look-back over the per-block
// words (status << 32 | value; status 1 = the block's own count, 2 = its inclusive prefix) — the stream is read once.
// A source row that is NOT the first of its voxel finds the voxel's row in the slot: written by a thread of this
// block before the barrier, or by a block with a lower ticket (the first occurrence has the smaller index), which
// never waits for a higher one — so spinning on the slot's `row` field terminates.
__global__ void __launch_bounds__(kScanThreads)
assign_rows_kernel(Slot* slots, const int* __restrict__ slot_of, int n, const int* __restrict__ n_dev,
                   unsigned long long* scan_state, int4* __restrict__ out_coords, int* __restrict__ out_first,
                   int* __restrict__ inverse, int* __restrict__ count, int* __restrict__ status) {
  __shared__ int warp_sum[32];
  __shared__ int s_ticket, s_prefix;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (n_dev) n = min(n, *n_dev);
  if (threadIdx.x == 0) s_ticket = (int)atomicAdd(scan_state, 1ull);
  __syncthreads();
  const int t = s_ticket;
  unsigned long long* state = scan_state + 1;
  const int last = n > 0 ? (n - 1) / kScanThreads : 0;
  if (t > last) return;   // (nothing in this block's range, and no block that works looks at its word)
  const int i = t * kScanThreads + threadIdx.x;
  const int s = i < n ? slot_of[i] : -1;
  bool flag = false;
  unsigned long long key = 0;
  if (s >= 0) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(slots + s);   // key, first (final since the insert launch)
    key = v.x;
    flag = (unsigned)v.y == (unsigned)i;
  }
  const int local = block_ballot_prefix(flag, warp_sum);
  const int agg = __syncthreads_count(flag);
  if (warp == 0) {
    int excl = 0;
    if (t == 0) {
      if (lane == 0) st_relaxed_u64(state, (2ull << 32) | (unsigned)agg);
    } else {
      if (lane == 0) st_relaxed_u64(state + t, (1ull << 32) | (unsigned)agg);
      for (int j = t - 1;; j -= 32) {
        const int idx = j - lane;
        unsigned long long w = 2ull << 32;   // in front of block 0: inclusive prefix 0
        if (idx >= 0) {
          do { w = ld_relaxed_u64(state + idx); } while ((w >> 32) == 0ull);
        }
        const unsigned done = __ballot_sync(0xffffffffu, (w >> 32) == 2ull);
        int v = (int)(unsigned)w;
        if (done && lane > __ffs(done) - 1) v = 0;   // blocks behind the nearest inclusive prefix are inside it
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        excl += v;
        if (done) break;
      }
      if (lane == 0) st_relaxed_u64(state + t, (2ull << 32) | (unsigned)(excl + agg));
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();
  const int row = s_prefix + local;
  if (flag) {
    st_relaxed_u32(&slots[s].row, (unsigned)row);
    out_coords[row] = unpack_key(key);
    out_first[row] = i;
  }
  if (t == last && threadIdx.x == 0) status[0] = s_prefix + agg;
  __syncthreads();
  if (i < n) {
    int r = -1;
    if (s >= 0) {
      if (flag) r = row;
      else {
        unsigned v;
        do { v = ld_relaxed_u32(&slots[s].row); } while (v == 0xFFFFFFFFu);
        r = (int)v;
      }
      atomicAdd(&count[r], 1);
    }
    inverse[i] = r;
  }
}

// ---------------------------------------------------------------------------
// kernel map
// ---------------------------------------------------------------------------
constexpr int kMaxOffsets = 125;  // up to 5^3
struct Offsets {
  int v[kMaxOffsets * 3];
};

__global__ void __launch_bounds__(256)
kernel_map_kernel(const Slot* __restrict__ slots, unsigned long long bucket_mask,
                  const int4* __restrict__ out_coords, int m_out, Offsets offs, int K,
                  int* __restrict__ nbr, int* __restrict__ tap_count) {
  __shared__ int s_count[kMaxOffsets];
  if (threadIdx.x < K) s_count[threadIdx.x] = 0;
  __syncthreads();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = o < m_out;
  const int lane = threadIdx.x & 31;
  int4 c = valid ? out_coords[o] : make_int4(0, 0, 0, 0);
  for (int k = 0; k < K; ++k) {
    int row = -1;
    if (valid) {
      unsigned long long key;
      if (pack_key(c.x, c.y + offs.v[3 * k], c.z + offs.v[3 * k + 1], c.w + offs.v[3 * k + 2], key))
        row = table_lookup(slots, bucket_mask, key);
      nbr[(size_t)k * m_out + o] = row;
    }
    unsigned bal = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0 && bal) atomicAdd(&s_count[k], __popc(bal));
  }
  __syncthreads();
  if (threadIdx.x < K && s_count[threadIdx.x]) atomicAdd(&tap_count[threadIdx.x], s_count[threadIdx.x]);
}

// Self map with centrally symmetric offsets (every 3^3 / 5^3 stride-1 convolution: in map == out map,
// offs[K-1-k] == -offs[k]): nbr[k][o] = i  <=>  nbr[K-1-k][i] = o, and the centre offset is the identity.  Only the
// first K/2 offsets are probed; a hit also writes the mirrored entry (each (K-1-k, i) has exactly one writer; rows
// K/2+1 .. K-1 are pre-filled with -1).  Half the hash probes — what bounds this kernel — for the map that every
// stride-1 convolution of a level shares.
__global__ void __launch_bounds__(256)
kernel_map_sym_kernel(const Slot* __restrict__ slots, unsigned long long bucket_mask,
                      const int4* __restrict__ coords, int m, Offsets offs, int K,
                      int* __restrict__ nbr, int* __restrict__ tap_count) {
  __shared__ int s_count[kMaxOffsets];
  const int half = K / 2;
  if (threadIdx.x < K) s_count[threadIdx.x] = 0;
  __syncthreads();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = o < m;
  const int lane = threadIdx.x & 31;
  int4 c = valid ? coords[o] : make_int4(0, 0, 0, 0);
  for (int k = 0; k < half; ++k) {
    int row = -1;
    if (valid) {
      unsigned long long key;
      if (pack_key(c.x, c.y + offs.v[3 * k], c.z + offs.v[3 * k + 1], c.w + offs.v[3 * k + 2], key))
        row = table_lookup(slots, bucket_mask, key);
      nbr[(size_t)k * m + o] = row;
      if (row >= 0) nbr[(size_t)(K - 1 - k) * m + row] = o;
    }
    unsigned bal = __ballot_sync(0xffffffffu, row >= 0);
    if (lane == 0 && bal) atomicAdd(&s_count[k], __popc(bal));
  }
  if (valid) nbr[(size_t)half * m + o] = o;
  __syncthreads();
  if (threadIdx.x < half && s_count[threadIdx.x]) {
    atomicAdd(&tap_count[threadIdx.x], s_count[threadIdx.x]);
    atomicAdd(&tap_count[K - 1 - threadIdx.x], s_count[threadIdx.x]);
  }
  if (threadIdx.x == 0) {
    const int rows = min(256, m - (int)(blockIdx.x * blockDim.x));
    if (rows > 0) atomicAdd(&tap_count[half], rows);
  }
}

__global__ void __launch_bounds__(256)
transpose_map_kernel(const int* __restrict__ nbr, int m_out, int m_in, int K,
                     int* __restrict__ nbr_t) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  int k = blockIdx.y;
  if (o >= m_out) return;
  int i = nbr[(size_t)k * m_out + o];
  if (i >= 0) nbr_t[(size_t)k * m_in + i] = o;
}

// per-128-row-tile bitmask of offsets that have at least one neighbour in the tile
__global__ void __launch_bounds__(128)
tile_mask_kernel(const int* __restrict__ nbr, int m, int K, unsigned* __restrict__ mask) {
  const int o = blockIdx.x * 128 + threadIdx.x;
  unsigned mm = 0;
  for (int k = 0; k < K; ++k) {
    int v = (o < m) && nbr[(size_t)k * m + o] >= 0;
    if (__syncthreads_or(v)) mm |= 1u << k;
  }
  if (threadIdx.x == 0) mask[blockIdx.x] = mm;
}

// per-ROW bitmask of offsets with a neighbour (bit k; K <= 32) + the number of (128-row tile, offset) pairs a
// tensor-core kernel would execute on this row order (sum over tiles of popcount(OR of the tile's row masks)):
// what the engine-side row re-ordering (ops.reorder_rows_by_mask) sorts by and decides on
__global__ void __launch_bounds__(128)
row_masks_kernel(const int* __restrict__ nbr, int m, int K, unsigned* __restrict__ row_mask,
                 unsigned long long* __restrict__ executed) {
  const int o = blockIdx.x * 128 + threadIdx.x;
  unsigned mine = 0, tile = 0;
  for (int k = 0; k < K; ++k) {
    const int v = (o < m) && nbr[(size_t)k * m + o] >= 0;
    if (v) mine |= 1u << k;
    if (__syncthreads_or(v)) tile |= 1u << k;
  }
  if (o < m) row_mask[o] = mine;
  if (threadIdx.x == 0 && tile) atomicAdd(executed, (unsigned long long)__popc(tile));
}

// rows of a coordinate map re-numbered: row r becomes pos[r] in every occupied slot of its hash table
__global__ void __launch_bounds__(256)
table_relabel_kernel(Slot* slots, long long n_slots, const int* __restrict__ pos) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  if (slots[i].key != kEmptyKey) slots[i].row = (unsigned)pos[slots[i].row];
}

// pair-list export: count / scan / write, stable in out row
__global__ void __launch_bounds__(kScanThreads)
pairs_count_kernel(const int* __restrict__ nbr, int m_out, int nblk, int* __restrict__ block_count) {
  int o = blockIdx.x * kScanThreads + threadIdx.x;
  int k = blockIdx.y;
  bool flag = o < m_out && nbr[(size_t)k * m_out + o] >= 0;
  int c = __syncthreads_count(flag);
  if (threadIdx.x == 0) block_count[k * nblk + blockIdx.x] = c;
}

__global__ void __launch_bounds__(kScanThreads)
pairs_write_kernel(const int* __restrict__ nbr, int m_out, int nblk, int K,
                   const int* __restrict__ block_offset, const int* __restrict__ total,
                   long long pair_capacity, int* __restrict__ pairs, int* __restrict__ tap_offset) {
  __shared__ int warp_sum[32];
  int o = blockIdx.x * kScanThreads + threadIdx.x;
  int k = blockIdx.y;
  int i = o < m_out ? nbr[(size_t)k * m_out + o] : -1;
  bool flag = i >= 0;
  int base = block_offset[k * nblk + blockIdx.x];
  int pos = base + block_ballot_prefix(flag, warp_sum);
  if (flag && pos < pair_capacity) {
    pairs[pos] = i;
    pairs[pair_capacity + pos] = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tap_offset[k] = base;
    if (k == K - 1) tap_offset[K] = *total;
  }
}

// ---------------------------------------------------------------------------
// feature rows keyed by an index map
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
segment_add_kernel(const float* __restrict__ feats, const int* __restrict__ inverse,
                   const int* __restrict__ count, long long total, int C,
                   float* __restrict__ out) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long j = e / C;
  int c = (int)(e - j * C);
  int r = inverse[j];
  if (r < 0) return;
  float v = feats[e];
  float* dst = out + (long long)r * C + c;
  if (count[r] == 1) *dst = v;  // single writer: no atomic needed
  else atomicAdd(dst, v);
}

__global__ void __launch_bounds__(256)
segment_norm_kernel(float* __restrict__ out, const int* __restrict__ count, long long total, int C) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  int cnt = count[e / C];
  if (cnt > 1) out[e] = out[e] / (float)cnt;
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ index,
                   const int* __restrict__ count, long long total, int C,
                   float* __restrict__ out) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long j = e / C;
  int c = (int)(e - j * C);
  int r = index[j];
  float v = 0.f;
  if (r >= 0) {
    v = src[(long long)r * C + c];
    if (count) {
      int cnt = count[r];
      if (cnt > 1) v = v / (float)cnt;
    }
  }
  out[e] = v;
}

// float4 flavour of the gather for C % 4 == 0
__global__ void __launch_bounds__(256)
gather_rows_v4_kernel(const float4* __restrict__ src, const int* __restrict__ index,
                      const int* __restrict__ count, long long total4, int C4,
                      float4* __restrict__ out) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total4) return;
  long long j = e / C4;
  int c = (int)(e - j * C4);
  int r = index[j];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= 0) {
    v = src[(long long)r * C4 + c];
    if (count) {
      int cnt = count[r];
      if (cnt > 1) {
        float s = 1.f / (float)cnt;
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      }
    }
  }
  out[e] = v;
}

__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const float* __restrict__ src, const int* __restrict__ index,
                        long long total, int C, float* __restrict__ out) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long j = e / C;
  int c = (int)(e - j * C);
  int r = index[j];
  if (r >= 0) atomicAdd(out + (long long)r * C + c, src[e]);
}

}  // namespace spc

using namespace spc;

extern "C" {

int spc_abi_version(void) { return SPC_ABI_VERSION; }
const char* spc_last_error(void) { return g_last_error; }
int64_t spc_launch_count(void) { return (int64_t)g_launch_count.load(); }

int64_t spc_table_slots(int64_t n) {
  int64_t s = 64;
  while (s < 2 * n) s <<= 1;
  return s;
}

int64_t spc_coords_insert_workspace(int64_t n) {
  int64_t nb = ceil_div(n > 0 ? n : 1, kScanThreads);
  return align_up(n * 4, 256) + align_up((nb + 1) * 8, 256) + 256;   // slot of every source row + look-back words
}

int spc_coords_insert(const void* src, int64_t n, int src_kind, const int32_t* ts, void* slots,
                      int64_t n_slots, int32_t* out_coords, int32_t* out_first,
                      int32_t* out_inverse, int32_t* out_count, int32_t* status, void* workspace,
                      int64_t workspace_bytes, void* stream_) {
  return spc_coords_insert_dev(src, n, nullptr, src_kind, ts, slots, n_slots, out_coords, out_first, out_inverse,
                               out_count, status, workspace, workspace_bytes, stream_);
}

int spc_coords_insert_dev(const void* src, int64_t n, const int32_t* n_dev, int src_kind, const int32_t* ts,
                          void* slots, int64_t n_slots, int32_t* out_coords, int32_t* out_first,
                          int32_t* out_inverse, int32_t* out_count, int32_t* status, void* workspace,
                          int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(n >= 0 && n < (1ll << 31) - kScanThreads, "n out of range");
  SPC_REQUIRE(n_slots >= 64 && (n_slots & (n_slots - 1)) == 0 && n_slots >= 2 * n,
              "n_slots must be a power of two >= 2n (use spc_table_slots)");
  SPC_REQUIRE(workspace_bytes >= spc_coords_insert_workspace(n), "workspace too small");
  SPC_REQUIRE(ts && ts[0] > 0 && ts[1] > 0 && ts[2] > 0, "tensor stride must be positive");
  SPC_REQUIRE(src_kind >= 0 && src_kind <= 2, "bad src_kind");
  SPC_CUDA(cudaMemsetAsync(slots, 0xFF, (size_t)n_slots * sizeof(Slot), stream));
  SPC_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int), stream));
  if (n == 0) return 0;
  char* ws = (char*)workspace;
  int* slot_of = (int*)ws;
  unsigned long long* scan_state = (unsigned long long*)(ws + align_up(n * 4, 256));   // [0] ticket counter, [1 + t] block t
  const int nb = (int)ceil_div(n, kScanThreads);
  const unsigned long long bucket_mask = (unsigned long long)(n_slots / 2 - 1);
  const int3 t3 = make_int3(ts[0], ts[1], ts[2]);
  const int grid256 = (int)ceil_div(n, 256);   // (>= nb + 1 threads: the launch also clears the look-back words)
  Slot* sl = (Slot*)slots;
  if (src_kind == SPC_SRC_FLOAT)
    insert_kernel<SPC_SRC_FLOAT><<<grid256, 256, 0, stream>>>(src, (int)n, n_dev, t3, sl, bucket_mask, slot_of, status,
                                                              out_count, scan_state, nb + 1);
  else if (src_kind == SPC_SRC_INT)
    insert_kernel<SPC_SRC_INT><<<grid256, 256, 0, stream>>>(src, (int)n, n_dev, t3, sl, bucket_mask, slot_of, status,
                                                            out_count, scan_state, nb + 1);
  else
    insert_kernel<SPC_SRC_STRIDE><<<grid256, 256, 0, stream>>>(src, (int)n, n_dev, t3, sl, bucket_mask, slot_of, status,
                                                               out_count, scan_state, nb + 1);
  SPC_LAUNCHED("insert_kernel");
  assign_rows_kernel<<<nb, kScanThreads, 0, stream>>>(sl, slot_of, (int)n, n_dev, scan_state, (int4*)out_coords,
                                                      out_first, out_inverse, out_count, status);
  SPC_LAUNCHED("assign_rows_kernel");
  return 0;
}

int spc_kernel_map(const void* in_slots, int64_t in_n_slots, const int32_t* out_coords,
                   int64_t m_out, const int32_t* offsets_host, int K, int32_t* nbr,
                   int32_t* tap_count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(K >= 1 && K <= kMaxOffsets, "kernel volume must be in [1,125]");
  SPC_REQUIRE(in_n_slots >= 64 && (in_n_slots & (in_n_slots - 1)) == 0, "bad table size");
  SPC_REQUIRE(m_out >= 0 && m_out < (1ll << 31) - 1024, "m_out out of range");
  SPC_CUDA(cudaMemsetAsync(tap_count, 0, (size_t)K * sizeof(int), stream));
  if (m_out == 0) return 0;
  Offsets offs;
  memset(&offs, 0, sizeof(offs));
  memcpy(offs.v, offsets_host, (size_t)K * 3 * sizeof(int));
  kernel_map_kernel<<<(int)ceil_div(m_out, 256), 256, 0, stream>>>(
      (const Slot*)in_slots, (unsigned long long)(in_n_slots / 2 - 1), (const int4*)out_coords,
      (int)m_out, offs, K, nbr, tap_count);
  SPC_LAUNCHED("kernel_map_kernel");
  return 0;
}

int spc_kernel_map_sym(const void* slots, int64_t n_slots, const int32_t* coords, int64_t m,
                       const int32_t* offsets_host, int K, int32_t* nbr, int32_t* tap_count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(K >= 1 && K <= kMaxOffsets && (K & 1), "kernel volume must be odd and in [1,125]");
  SPC_REQUIRE(n_slots >= 64 && (n_slots & (n_slots - 1)) == 0, "bad table size");
  SPC_REQUIRE(m >= 0 && m < (1ll << 31) - 1024, "m out of range");
  for (int k = 0; k < K; ++k)
    for (int a = 0; a < 3; ++a)
      SPC_REQUIRE(offsets_host[3 * k + a] == -offsets_host[3 * (K - 1 - k) + a], "offsets are not centrally symmetric");
  SPC_CUDA(cudaMemsetAsync(tap_count, 0, (size_t)K * sizeof(int), stream));
  if (m == 0) return 0;
  const int half = K / 2;
  if (half > 0)
    SPC_CUDA(cudaMemsetAsync(nbr + (size_t)(half + 1) * m, 0xFF, (size_t)half * m * sizeof(int), stream));
  Offsets offs;
  memset(&offs, 0, sizeof(offs));
  memcpy(offs.v, offsets_host, (size_t)K * 3 * sizeof(int));
  kernel_map_sym_kernel<<<(int)ceil_div(m, 256), 256, 0, stream>>>(
      (const Slot*)slots, (unsigned long long)(n_slots / 2 - 1), (const int4*)coords, (int)m, offs, K, nbr, tap_count);
  SPC_LAUNCHED("kernel_map_sym_kernel");
  return 0;
}

int spc_kernel_map_transpose(const int32_t* nbr, int64_t m_out, int64_t m_in, int K,
                             int32_t* nbr_t, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(K >= 1 && K <= 65535, "bad K");
  if (m_in > 0) SPC_CUDA(cudaMemsetAsync(nbr_t, 0xFF, (size_t)K * m_in * sizeof(int), stream));
  if (m_out == 0 || m_in == 0) return 0;
  dim3 grid((unsigned)ceil_div(m_out, 256), (unsigned)K);
  transpose_map_kernel<<<grid, 256, 0, stream>>>(nbr, (int)m_out, (int)m_in, K, nbr_t);
  SPC_LAUNCHED("transpose_map_kernel");
  return 0;
}

int spc_tile_mask(const int32_t* nbr, int64_t m, int K, uint32_t* mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(K >= 1 && K <= 32, "tile masks support kernel volume <= 32");
  if (m == 0) return 0;
  tile_mask_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, stream>>>(nbr, (int)m, K, mask);
  SPC_LAUNCHED("tile_mask_kernel");
  return 0;
}

int spc_row_masks(const int32_t* nbr, int64_t m, int K, uint32_t* row_mask, int64_t* executed_dev, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(K >= 1 && K <= 32, "row masks support kernel volume <= 32");
  SPC_REQUIRE(m >= 0 && m < (1ll << 31) - 1024, "m out of range");
  SPC_CUDA(cudaMemsetAsync(executed_dev, 0, sizeof(int64_t), stream));
  if (m == 0) return 0;
  row_masks_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, stream>>>(nbr, (int)m, K, row_mask,
                                                                   (unsigned long long*)executed_dev);
  SPC_LAUNCHED("row_masks_kernel");
  return 0;
}

int spc_table_relabel(void* slots, int64_t n_slots, const int32_t* pos, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(n_slots >= 64 && (n_slots & (n_slots - 1)) == 0, "bad table size");
  table_relabel_kernel<<<(unsigned)ceil_div(n_slots, 256), 256, 0, stream>>>((Slot*)slots, (long long)n_slots, pos);
  SPC_LAUNCHED("table_relabel_kernel");
  return 0;
}

int64_t spc_pairs_workspace(int64_t m_out, int K) {
  int64_t nblk = ceil_div(m_out > 0 ? m_out : 1, kScanThreads);
  return align_up(nblk * K * 4, 256) + 256;
}

int spc_kernel_map_pairs(const int32_t* nbr, int64_t m_out, int K, int64_t pair_capacity,
                         int32_t* pairs, int32_t* tap_offset, void* workspace,
                         int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(K >= 1 && K <= 65535, "bad K");
  SPC_REQUIRE(workspace_bytes >= spc_pairs_workspace(m_out, K), "workspace too small");
  if (m_out == 0) {
    SPC_CUDA(cudaMemsetAsync(tap_offset, 0, (size_t)(K + 1) * sizeof(int), stream));
    return 0;
  }
  const int nblk = (int)ceil_div(m_out, kScanThreads);
  int* block_off = (int*)workspace;
  int* total = (int*)((char*)workspace + align_up((int64_t)nblk * K * 4, 256));
  dim3 grid((unsigned)nblk, (unsigned)K);
  pairs_count_kernel<<<grid, kScanThreads, 0, stream>>>(nbr, (int)m_out, nblk, block_off);
  SPC_LAUNCHED("pairs_count_kernel");
  scan_blocks_kernel<<<1, kScanThreads, 0, stream>>>(block_off, nblk * K, total);
  SPC_LAUNCHED("scan_blocks_kernel");
  pairs_write_kernel<<<grid, kScanThreads, 0, stream>>>(nbr, (int)m_out, nblk, K, block_off, total,
                                                        (long long)pair_capacity, pairs, tap_offset);
  SPC_LAUNCHED("pairs_write_kernel");
  return 0;
}

int spc_segment_reduce(const float* feats, const int32_t* inverse, const int32_t* count,
                       int64_t n, int64_t m, int C, int mode, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1, "bad C");
  if (m > 0) SPC_CUDA(cudaMemsetAsync(out, 0, (size_t)m * C * sizeof(float), stream));
  if (n == 0 || m == 0) return 0;
  long long total = (long long)n * C;
  segment_add_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(feats, inverse, count, total, C, out);
  SPC_LAUNCHED("segment_add_kernel");
  if (mode == 0) {
    long long tm = (long long)m * C;
    segment_norm_kernel<<<(unsigned)ceil_div(tm, 256), 256, 0, stream>>>(out, count, tm, C);
    SPC_LAUNCHED("segment_norm_kernel");
  }
  return 0;
}

int spc_gather_rows(const float* src, const int32_t* index, const int32_t* count, int64_t n,
                    int C, float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1, "bad C");
  if (n == 0) return 0;
  if (C % 4 == 0 && ((uintptr_t)src % 16 == 0) && ((uintptr_t)out % 16 == 0)) {
    long long total4 = (long long)n * (C / 4);
    gather_rows_v4_kernel<<<(unsigned)ceil_div(total4, 256), 256, 0, stream>>>(
        (const float4*)src, index, count, total4, C / 4, (float4*)out);
    SPC_LAUNCHED("gather_rows_v4_kernel");
  } else {
    long long total = (long long)n * C;
    gather_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, index, count, total, C, out);
    SPC_LAUNCHED("gather_rows_kernel");
  }
  return 0;
}

int spc_scatter_add_rows(const float* src, const int32_t* index, int64_t n, int64_t m, int C,
                         float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1, "bad C");
  if (m > 0) SPC_CUDA(cudaMemsetAsync(out, 0, (size_t)m * C * sizeof(float), stream));
  if (n == 0 || m == 0) return 0;
  long long total = (long long)n * C;
  scatter_add_rows_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, index, total, C, out);
  SPC_LAUNCHED("scatter_add_rows_kernel");
  return 0;
}

}  // extern "C"
