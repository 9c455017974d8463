// pipeline.cu — the two ends of the hot path that the reference runs on the CPU / with per-class torch loops
// (SURVEY.md §8f "next" rows 2 and 3):
//   plenoxel_decode : PeRFception plenoxel record -> network input.  links (flat indices of the occupied cells of the
//                     reso^3 grid) -> float coordinates (b, i, j, k) [optionally through an affine map, the form of the
//                     reference's rotation / scale / translation augmentations], uint8 spherical-harmonic coefficients ->
//                     float features  sh * scale + min                      (co3d_3d/src/data/co3d.py:164-172,196-203)
//   seg_metrics     : per-class seen / correct / predicted counts of argmax(logits) against the labels with an ignore
//                     label — IoUMeter.update                                (co3d_3d/src/metrics.py:29-41)
#include <climits>

#include "common.cuh"

namespace spc {

struct Affine {
  float m[9];  // row-major 3x3
  float t[3];
  int enabled;
};

__global__ void __launch_bounds__(256)
plenoxel_decode_kernel(const long long* __restrict__ links64, const int* __restrict__ links32, long long n,
                       int r1, int r2, float batch, Affine aff, const unsigned char* __restrict__ sh_u8, int C,
                       float sh_scale, float sh_min, float* __restrict__ coords, float* __restrict__ feats,
                       const int* __restrict__ rows) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long plane = (long long)r1 * r2;
  for (long long i = t0; i < n; i += stride) {
    const long long src = rows ? (long long)rows[i] : i;   // output row i <- record rows[i] (crop / dropout selections)
    const long long l = links64 ? links64[src] : (long long)links32[src];
    // torch.div(links, r1*r2, 'trunc'), torch.div(links % (r1*r2), r2, 'trunc'), links % r2   (co3d.py:196-203)
    const float x = (float)(l / plane), y = (float)((l % plane) / r2), z = (float)(l % r2);
    float4 c;
    c.x = batch;
    if (aff.enabled) {  // products and sums rounded one by one, left to right, like the numpy / torch reference
      c.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aff.m[0], x), __fmul_rn(aff.m[1], y)), __fmul_rn(aff.m[2], z)), aff.t[0]);
      c.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aff.m[3], x), __fmul_rn(aff.m[4], y)), __fmul_rn(aff.m[5], z)), aff.t[1]);
      c.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aff.m[6], x), __fmul_rn(aff.m[7], y)), __fmul_rn(aff.m[8], z)), aff.t[2]);
    } else {
      c.y = x; c.z = y; c.w = z;
    }
    reinterpret_cast<float4*>(coords)[i] = c;
  }
  if (rows) {   // gathered rows: one thread per (row, 4-byte group) when C % 4 == 0, else per element
    const bool v4 = C % 4 == 0 && ((uintptr_t)sh_u8 % 4 == 0) && ((uintptr_t)feats % 16 == 0);
    const int per = v4 ? C / 4 : C;
    const long long tot = n * (long long)per;
    for (long long q = t0; q < tot; q += stride) {
      const long long i = q / per;
      const int j = (int)(q - i * per);
      const long long src = (long long)rows[i];
      if (v4) {
        const uchar4 u = reinterpret_cast<const uchar4*>(sh_u8 + src * C)[j];
        float4 o;
        o.x = __fadd_rn(__fmul_rn((float)u.x, sh_scale), sh_min);
        o.y = __fadd_rn(__fmul_rn((float)u.y, sh_scale), sh_min);
        o.z = __fadd_rn(__fmul_rn((float)u.z, sh_scale), sh_min);
        o.w = __fadd_rn(__fmul_rn((float)u.w, sh_scale), sh_min);
        reinterpret_cast<float4*>(feats + i * C)[j] = o;
      } else {
        feats[i * C + j] = __fadd_rn(__fmul_rn((float)sh_u8[src * C + j], sh_scale), sh_min);
      }
    }
    return;
  }
  // features: flat over n * C bytes, 4 per thread where alignment allows;  sh.float() * scale + min  (co3d.py:169)
  const long long total = n * (long long)C;
  const bool vec = ((uintptr_t)sh_u8 % 4 == 0) && ((uintptr_t)feats % 16 == 0);
  const long long n4 = vec ? total / 4 : 0;
  for (long long q = t0; q < n4; q += stride) {
    const uchar4 u = reinterpret_cast<const uchar4*>(sh_u8)[q];
    float4 o;
    o.x = __fadd_rn(__fmul_rn((float)u.x, sh_scale), sh_min);
    o.y = __fadd_rn(__fmul_rn((float)u.y, sh_scale), sh_min);
    o.z = __fadd_rn(__fmul_rn((float)u.z, sh_scale), sh_min);
    o.w = __fadd_rn(__fmul_rn((float)u.w, sh_scale), sh_min);
    reinterpret_cast<float4*>(feats)[q] = o;
  }
  for (long long e = n4 * 4 + t0; e < total; e += stride) feats[e] = __fadd_rn(__fmul_rn((float)sh_u8[e], sh_scale), sh_min);
}

// ---- RandomCrop on the device (co3d_3d/src/data/transforms.py:195-243) ---------------------------------------------
// The crop keeps the points strictly inside a box of size `size` whose corner is u * max(extent - size, 0) above the
// minimum of the (already transformed) coordinates.  Three small kernels over the current row selection of a record:
// extent (ordered-int atomic min / max), flags + per-block counts, stable compaction into a row list — the decode
// kernel then gathers exactly the kept records (plenoxel_decode_kernel, `rows`), in the reference's order.
__device__ __forceinline__ int float_order(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;     // monotone float -> int
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

__device__ __forceinline__ void plenoxel_point(long long l, long long plane, int r2, const Affine& aff, float& px, float& py,
                                               float& pz) {
  const float x = (float)(l / plane), y = (float)((l % plane) / r2), z = (float)(l % r2);
  if (aff.enabled) {
    px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aff.m[0], x), __fmul_rn(aff.m[1], y)), __fmul_rn(aff.m[2], z)), aff.t[0]);
    py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aff.m[3], x), __fmul_rn(aff.m[4], y)), __fmul_rn(aff.m[5], z)), aff.t[1]);
    pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(aff.m[6], x), __fmul_rn(aff.m[7], y)), __fmul_rn(aff.m[8], z)), aff.t[2]);
  } else {
    px = x; py = y; pz = z;
  }
}

__global__ void __launch_bounds__(256)
crop_extent_kernel(const long long* __restrict__ links64, const int* __restrict__ links32, const int* __restrict__ rows,
                   long long n, int r1, int r2, Affine aff, int* __restrict__ ext /* [6] ordered ints: min xyz, max xyz */) {
  const long long plane = (long long)r1 * r2;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long src = rows ? (long long)rows[i] : i;
    float p[3];
    plenoxel_point(links64 ? links64[src] : (long long)links32[src], plane, r2, aff, p[0], p[1], p[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) { const int o = float_order(p[a]); lo[a] = min(lo[a], o); hi[a] = max(hi[a], o); }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int d = 16; d > 0; d >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], d));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], d));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(ext + a, lo[a]); atomicMax(ext + 3 + a, hi[a]); }
  }
}

struct CropBox { float u[3], size[3]; };

// keep flag of a point: norm = p - min; box corner = u * clip(max - min - size, 0); strict inequalities.
// (float32 arithmetic in the reference's order of operations; the reference computes in float64 when the coordinates
// are float64 — pipeline.py documents the dtype it mirrors)
__device__ __forceinline__ bool crop_keep(const float* p, const int* __restrict__ ext, const CropBox& box) {
  bool keep = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mn = order_float(ext[a]), mx = order_float(ext[3 + a]);
    const float norm = __fsub_rn(p[a], mn);
    const float range = fmaxf(__fsub_rn(__fsub_rn(mx, mn), box.size[a]), 0.f);
    const float lo = __fmul_rn(box.u[a], range), hi = __fadd_rn(lo, box.size[a]);
    keep = keep && norm > lo && norm < hi;
  }
  return keep;
}

constexpr int kCropBlock = 1024;   // records per block of the count / compaction kernels

template <bool WRITE>
__global__ void __launch_bounds__(kCropBlock)
crop_select_kernel(const long long* __restrict__ links64, const int* __restrict__ links32, const int* __restrict__ rows,
                   long long n, int r1, int r2, Affine aff, const int* __restrict__ ext, CropBox box,
                   int* __restrict__ block_counts /* WRITE: exclusive offsets */, int* __restrict__ out_rows) {
  __shared__ int s_warp[kCropBlock / 32];
  const long long plane = (long long)r1 * r2;
  const long long i = (long long)blockIdx.x * kCropBlock + threadIdx.x;
  bool keep = false;
  long long src = 0;
  if (i < n) {
    src = rows ? (long long)rows[i] : i;
    float p[3];
    plenoxel_point(links64 ? links64[src] : (long long)links32[src], plane, r2, aff, p[0], p[1], p[2]);
    keep = crop_keep(p, ext, box);
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_warp[wid] = __popc(ballot);
  __syncthreads();
  if (wid == 0) {   // exclusive scan of the 32 warp counts
    int v = s_warp[lane], incl = v;
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    s_warp[lane] = incl - v;
    if (!WRITE && lane == 31) block_counts[blockIdx.x] = incl;
  }
  if (!WRITE) return;
  __syncthreads();
  if (keep) out_rows[block_counts[blockIdx.x] + s_warp[wid] + __popc(ballot & ((1u << lane) - 1u))] = (int)src;
}

// exclusive scan of the block counts in place (one block; <= a few thousand blocks), total and "box fits" flag
__global__ void __launch_bounds__(1024)
crop_scan_kernel(int* __restrict__ block_counts, int n_blocks, const int* __restrict__ ext, CropBox box,
                 int* __restrict__ result /* [2]: rows kept, 1 if the box covers the extent on every axis */) {
  __shared__ int s_part[1024];
  const int per = (n_blocks + 1023) / 1024;
  const int b0 = threadIdx.x * per, b1 = min(b0 + per, n_blocks);
  int sum = 0;
  for (int b = b0; b < b1; ++b) sum += block_counts[b];
  s_part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < 1024; ++t) { const int v = s_part[t]; s_part[t] = run; run += v; }
    result[0] = run;
    bool fits = true;
    for (int a = 0; a < 3; ++a)
      fits = fits && !(__fsub_rn(__fsub_rn(order_float(ext[3 + a]), order_float(ext[a])), box.size[a]) > 0.f);
    result[1] = fits ? 1 : 0;
  }
  __syncthreads();
  int run = s_part[threadIdx.x];
  for (int b = b0; b < b1; ++b) { const int v = block_counts[b]; block_counts[b] = run; run += v; }
}

// counts[0][c] = #(target == c), counts[1][c] = #(target == c and argmax == c), counts[2][c] = #(argmax == c),
// over rows whose target is not the ignore label.  argmax ties -> the lowest class index (torch.argmax on CUDA
// returns one of the maxima; the reference's logits are floats where exact ties do not occur in practice).
template <int CMAX>
__global__ void __launch_bounds__(256)
seg_metrics_kernel(const float* __restrict__ logits, const long long* __restrict__ target, long long n, int C,
                   long long ignore_label, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int s[3 * CMAX];
  for (int i = threadIdx.x; i < 3 * CMAX; i += blockDim.x) s[i] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long y = target[i];
    if (y == ignore_label) continue;
    int best = 0;
    float bv = logits[i * C];
    for (int c = 1; c < C; ++c) {
      const float v = logits[i * C + c];
      if (v > bv) { bv = v; best = c; }
    }
    atomicAdd(&s[2 * CMAX + best], 1u);
    if (y >= 0 && y < C) {
      atomicAdd(&s[(int)y], 1u);
      if (best == (int)y) atomicAdd(&s[CMAX + (int)y], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
    const int which = i / C, c = i - which * C;
    const unsigned int v = s[which * CMAX + c];
    if (v) atomicAdd(&counts[which * C + c], (unsigned long long)v);
  }
}

// ---------------------------------------------------------------------------
// max pooling (ME.MinkowskiMaxPooling / MinkowskiGlobalMaxPooling of the other backbones)
// ---------------------------------------------------------------------------
// out[o, c] = max over the present neighbours nbr[k, o] of in[., c]; arg[o, c] = the input row that won (lowest k on
// ties, -1 / 0 when the output row has no neighbour).  One thread per (output row, channel).
__global__ void __launch_bounds__(256)
pool_max_fwd_kernel(const float* __restrict__ in, const int* __restrict__ nbr, long long m_out, int C, int K,
                    float* __restrict__ out, int* __restrict__ arg) {
  const long long total = m_out * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long o = e / C;
    const int c = (int)(e - o * C);
    float best = 0.f;
    int who = -1;
    for (int k = 0; k < K; ++k) {
      const int i = nbr[(long long)k * m_out + o];
      if (i < 0) continue;
      const float v = in[(long long)i * C + c];
      if (who < 0 || v > best) { best = v; who = i; }
    }
    out[e] = best;
    arg[e] = who;
  }
}

// din[arg[o, c], c] += dout[o, c]   (din zeroed by the wrapper; regions may overlap when kernel > stride)
__global__ void __launch_bounds__(256)
pool_max_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, long long total, int C,
                    float* __restrict__ din) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = arg[e];
    if (i >= 0) atomicAdd(din + (long long)i * C + (e % C), dout[e]);
  }
}

// order-preserving float <-> unsigned encoding for atomicMax
__device__ __forceinline__ unsigned enc_f(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__global__ void __launch_bounds__(256)
global_max_pass1_kernel(const float* __restrict__ in, const int4* __restrict__ coords, long long total, int C,
                        int n_batch, unsigned* __restrict__ best) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int b = coords[r].x;
    if (b >= 0 && b < n_batch) atomicMax(best + (long long)b * C + (e - r * C), enc_f(in[e]));
  }
}
__global__ void __launch_bounds__(256)
global_max_pass2_kernel(const float* __restrict__ in, const int4* __restrict__ coords, long long total, int C,
                        int n_batch, const unsigned* __restrict__ best, int* __restrict__ arg) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int b = coords[r].x;
    if (b >= 0 && b < n_batch && enc_f(in[e]) == best[(long long)b * C + (e - r * C)])
      atomicMin(arg + (long long)b * C + (e - r * C), (int)r);  // first row attaining the maximum
  }
}
__global__ void __launch_bounds__(256)
global_max_finish_kernel(const unsigned* __restrict__ best, int* __restrict__ arg, int n, float* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const bool any = arg[e] != 0x7F7F7F7F;
  out[e] = any ? dec_f(best[e]) : 0.f;
  if (!any) arg[e] = -1;
}

}  // namespace spc

using namespace spc;

extern "C" {

int spc_plenoxel_decode(const void* links, int links_is_int64, int64_t n, const int32_t* reso, int32_t batch_index,
                        const float* affine12, const uint8_t* sh_u8, int C, float sh_scale, float sh_min,
                        float* out_coords, float* out_feats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(n >= 0 && C >= 0 && reso && reso[1] > 0 && reso[2] > 0, "bad arguments");
  SPC_REQUIRE(((uintptr_t)out_coords % 16) == 0, "coordinates must be 16-byte aligned");
  if (n == 0) return 0;
  Affine aff;
  memset(&aff, 0, sizeof(aff));
  if (affine12) {
    memcpy(aff.m, affine12, 9 * sizeof(float));
    memcpy(aff.t, affine12 + 9, 3 * sizeof(float));
    aff.enabled = 1;
  }
  int64_t want = ceil_div(n, 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  plenoxel_decode_kernel<<<grid, 256, 0, stream>>>(links_is_int64 ? (const long long*)links : nullptr,
                                                    links_is_int64 ? nullptr : (const int*)links, n, reso[1], reso[2],
                                                    (float)batch_index, aff, sh_u8, C, sh_scale, sh_min, out_coords,
                                                    out_feats, nullptr);
  SPC_LAUNCHED("plenoxel_decode_kernel");
  return 0;
}

static Affine make_affine(const float* affine12) {
  Affine aff;
  memset(&aff, 0, sizeof(aff));
  if (affine12) {
    memcpy(aff.m, affine12, 9 * sizeof(float));
    memcpy(aff.t, affine12 + 9, 3 * sizeof(float));
    aff.enabled = 1;
  }
  return aff;
}

int spc_plenoxel_decode_rows(const void* links, int links_is_int64, const int32_t* rows, int64_t n_rows,
                             const int32_t* reso, int32_t batch_index, const float* affine12, const uint8_t* sh_u8, int C,
                             float sh_scale, float sh_min, float* out_coords, float* out_feats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(n_rows >= 0 && C >= 0 && rows && reso && reso[1] > 0 && reso[2] > 0, "bad arguments");
  SPC_REQUIRE(((uintptr_t)out_coords % 16) == 0, "coordinates must be 16-byte aligned");
  if (n_rows == 0) return 0;
  int64_t want = ceil_div(n_rows * (C >= 4 ? C / 4 : 1), 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  plenoxel_decode_kernel<<<grid, 256, 0, stream>>>(links_is_int64 ? (const long long*)links : nullptr,
                                                    links_is_int64 ? nullptr : (const int*)links, n_rows, reso[1], reso[2],
                                                    (float)batch_index, make_affine(affine12), sh_u8, C, sh_scale, sh_min,
                                                    out_coords, out_feats, rows);
  SPC_LAUNCHED("plenoxel_decode_kernel");
  return 0;
}

int64_t spc_plenoxel_crop_workspace(int64_t n) { return 64 + 4 * (ceil_div(n > 0 ? n : 1, kCropBlock) + 8); }

int spc_plenoxel_crop_select(const void* links, int links_is_int64, const int32_t* rows, int64_t n, const int32_t* reso,
                             const float* affine12, const float* u3, const float* size3, int32_t* out_rows,
                             int32_t* result2, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(n >= 0 && reso && reso[1] > 0 && reso[2] > 0 && u3 && size3 && out_rows && result2, "bad arguments");
  SPC_REQUIRE(workspace && workspace_bytes >= spc_plenoxel_crop_workspace(n), "workspace too small");
  SPC_REQUIRE(n < (1ll << 31), "too many rows");
  if (n == 0) { SPC_CUDA(cudaMemsetAsync(result2, 0, 8, stream)); return 0; }
  const long long* l64 = links_is_int64 ? (const long long*)links : nullptr;
  const int* l32 = links_is_int64 ? nullptr : (const int*)links;
  const Affine aff = make_affine(affine12);
  int* ext = (int*)(((uintptr_t)workspace + 15) & ~(uintptr_t)15);
  int* block_counts = ext + 8;
  const int init[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
  SPC_CUDA(cudaMemcpyAsync(ext, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  CropBox box;
  memcpy(box.u, u3, 12);
  memcpy(box.size, size3, 12);
  int64_t want = ceil_div(n, 256 * 4);
  int grid = (int)(want < kNumSMs * 8 ? (want > 0 ? want : 1) : kNumSMs * 8);
  crop_extent_kernel<<<grid, 256, 0, stream>>>(l64, l32, rows, n, reso[1], reso[2], aff, ext);
  SPC_LAUNCHED("crop_extent_kernel");
  const int n_blocks = (int)ceil_div(n, kCropBlock);
  crop_select_kernel<false><<<n_blocks, kCropBlock, 0, stream>>>(l64, l32, rows, n, reso[1], reso[2], aff, ext, box,
                                                                  block_counts, nullptr);
  SPC_LAUNCHED("crop_select_kernel");
  crop_scan_kernel<<<1, 1024, 0, stream>>>(block_counts, n_blocks, ext, box, result2);
  SPC_LAUNCHED("crop_scan_kernel");
  crop_select_kernel<true><<<n_blocks, kCropBlock, 0, stream>>>(l64, l32, rows, n, reso[1], reso[2], aff, ext, box,
                                                                 block_counts, out_rows);
  SPC_LAUNCHED("crop_select_kernel");
  return 0;
}

int spc_seg_metrics(const float* logits, const int64_t* target, int64_t n, int C, int64_t ignore_label,
                    uint64_t* counts, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && C <= 256, "seg_metrics supports 1..256 classes");
  if (n == 0) return 0;
  int64_t want = ceil_div(n, 256 * 4);
  int grid = (int)(want < kNumSMs * 8 ? (want > 0 ? want : 1) : kNumSMs * 8);
  if (C <= 64)
    seg_metrics_kernel<64><<<grid, 256, 0, stream>>>(logits, (const long long*)target, n, C, ignore_label,
                                                      (unsigned long long*)counts);
  else
    seg_metrics_kernel<256><<<grid, 256, 0, stream>>>(logits, (const long long*)target, n, C, ignore_label,
                                                       (unsigned long long*)counts);
  SPC_LAUNCHED("seg_metrics_kernel");
  return 0;
}

int spc_pool_max_fwd(const float* in, const int32_t* nbr, int64_t m_out, int C, int K, float* out, int32_t* arg,
                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && K >= 1, "bad shape");
  if (m_out == 0) return 0;
  int64_t want = ceil_div(m_out * C, 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  pool_max_fwd_kernel<<<grid, 256, 0, stream>>>(in, nbr, m_out, C, K, out, arg);
  SPC_LAUNCHED("pool_max_fwd_kernel");
  return 0;
}

int spc_pool_max_bwd(const float* dout, const int32_t* arg, int64_t m_out, int64_t m_in, int C, float* din,
                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_CUDA(cudaMemsetAsync(din, 0, (size_t)m_in * C * sizeof(float), stream));
  if (m_out == 0) return 0;
  int64_t want = ceil_div(m_out * C, 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  pool_max_bwd_kernel<<<grid, 256, 0, stream>>>(dout, arg, m_out * C, C, din);
  SPC_LAUNCHED("pool_max_bwd_kernel");
  return 0;
}

int spc_global_max_fwd(const float* in, const int32_t* coords, int64_t m, int C, int n_batch, float* out,
                       int32_t* arg, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && n_batch >= 1, "bad shape");
  SPC_REQUIRE(workspace && workspace_bytes >= (int64_t)n_batch * C * 4, "workspace too small (n_batch * C * 4 bytes)");
  unsigned* best = (unsigned*)workspace;
  SPC_CUDA(cudaMemsetAsync(best, 0, (size_t)n_batch * C * 4, stream));      // below every encoded float
  SPC_CUDA(cudaMemsetAsync(arg, 0x7F, (size_t)n_batch * C * 4, stream));    // 0x7F7F7F7F > any row index
  const int64_t total = m * C;
  int64_t want = ceil_div(total > 0 ? total : 1, 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  if (total > 0) {
    global_max_pass1_kernel<<<grid, 256, 0, stream>>>(in, (const int4*)coords, total, C, n_batch, best);
    SPC_LAUNCHED("global_max_pass1_kernel");
    global_max_pass2_kernel<<<grid, 256, 0, stream>>>(in, (const int4*)coords, total, C, n_batch, best, arg);
    SPC_LAUNCHED("global_max_pass2_kernel");
  }
  global_max_finish_kernel<<<(n_batch * C + 255) / 256, 256, 0, stream>>>(best, arg, n_batch * C, out);
  SPC_LAUNCHED("global_max_finish_kernel");
  return 0;
}

}  // extern "C"

