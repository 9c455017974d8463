// interp.cu — trilinear interpolation / splatting between a point field and a voxel map
// (ME.MinkowskiInterpolation, SparseTensor.interpolate, TensorField.splat — used by
//  co3d_3d/src/models/mink/fcnn.py:184-205 (MinkowskiSplatFCNN) and src/data/transforms.py:472,520-528).
//
// A query point (b, x, y, z) touches the 2^3 voxels of the tensor-stride lattice around it: per axis the lower voxel
// lo = floor(x / ts) * ts and the upper lo + ts, with weights (1 - f) and f, f = x / ts - floor(x / ts); a corner's
// weight is the product over the axes.  Corner k = bx + 2 by + 4 bz (first spatial axis fastest, like kernel offsets).
//   interp_corners : query -> lower corner (int32) + the 8 weights; the 8 voxel ROWS come from the ordinary kernel-map
//                    probe (spc_kernel_map with the 8 corner offsets) or, for splat, from inserting the corners
//   interp_fwd     : out[j]  = sum_k w[k, j] * feats[idx[k, j]]          (missing corners contribute nothing)
//   interp_bwd     : dfeats[idx[k, j]] += w[k, j] * dout[j]             (= the splat forward)
#include "common.cuh"

namespace spc {

__global__ void __launch_bounds__(256)
interp_corners_kernel(const float4* __restrict__ query, long long n, int ts0, int ts1, int ts2,
                      int4* __restrict__ lower, float* __restrict__ w) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const float4 q = query[j];
    const float tx = q.y / (float)ts0, ty = q.z / (float)ts1, tz = q.w / (float)ts2;
    const float fx = floorf(tx), fy = floorf(ty), fz = floorf(tz);
    int4 lo;
    lo.x = (int)floorf(q.x);
    lo.y = (int)fx * ts0;
    lo.z = (int)fy * ts1;
    lo.w = (int)fz * ts2;
    lower[j] = lo;
    const float rx = tx - fx, ry = ty - fy, rz = tz - fz;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float wx = (k & 1) ? rx : 1.f - rx;
      const float wy = (k & 2) ? ry : 1.f - ry;
      const float wz = (k & 4) ? rz : 1.f - rz;
      w[(long long)k * n + j] = wx * wy * wz;
    }
  }
}

template <typename V>
__global__ void __launch_bounds__(256)
interp_fwd_kernel(const V* __restrict__ feats, const int* __restrict__ idx, const float* __restrict__ w,
                  long long n, int CV, int K, V* __restrict__ out);

template <>
__global__ void __launch_bounds__(256)
interp_fwd_kernel<float>(const float* __restrict__ feats, const int* __restrict__ idx, const float* __restrict__ w,
                         long long n, int CV, int K, float* __restrict__ out) {
  const long long total = n * CV;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / CV;
    const int c = (int)(e - j * CV);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const int i = idx[(long long)k * n + j];
      if (i >= 0) acc += w[(long long)k * n + j] * feats[(long long)i * CV + c];
    }
    out[e] = acc;
  }
}

template <>
__global__ void __launch_bounds__(256)
interp_fwd_kernel<float4>(const float4* __restrict__ feats, const int* __restrict__ idx, const float* __restrict__ w,
                          long long n, int CV, int K, float4* __restrict__ out) {
  const long long total = n * CV;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / CV;
    const int c = (int)(e - j * CV);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const int i = idx[(long long)k * n + j];
      if (i >= 0) {
        const float a = w[(long long)k * n + j];
        const float4 v = feats[(long long)i * CV + c];
        acc.x += a * v.x; acc.y += a * v.y; acc.z += a * v.z; acc.w += a * v.w;
      }
    }
    out[e] = acc;
  }
}

__global__ void __launch_bounds__(256)
interp_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx, const float* __restrict__ w,
                  long long n, int C, int K, float* __restrict__ dfeats) {
  const long long total = n * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / C;
    const int c = (int)(e - j * C);
    const float g = dout[e];
    for (int k = 0; k < K; ++k) {
      const int i = idx[(long long)k * n + j];
      if (i >= 0) atomicAdd(dfeats + (long long)i * C + c, w[(long long)k * n + j] * g);
    }
  }
}

static int flat_grid_interp(long long total) {
  int64_t want = ceil_div(total, 256);
  return (int)(want < kNumSMs * 16 ? (want > 0 ? want : 1) : kNumSMs * 16);
}

}  // namespace spc

using namespace spc;

extern "C" {

int spc_interp_corners(const float* query, int64_t n, const int32_t* ts, int32_t* lower, float* weights,
                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(ts[0] >= 1 && ts[1] >= 1 && ts[2] >= 1, "tensor stride must be positive");
  if (n == 0) return 0;
  interp_corners_kernel<<<flat_grid_interp(n), 256, 0, stream>>>((const float4*)query, n, ts[0], ts[1], ts[2],
                                                                  (int4*)lower, weights);
  SPC_LAUNCHED("interp_corners_kernel");
  return 0;
}

int spc_interp_fwd(const float* feats, const int32_t* idx, const float* weights, int64_t n, int C, int K, float* out,
                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && K >= 1, "bad shape");
  if (n == 0) return 0;
  const bool vec = (C % 4 == 0) && ((uintptr_t)feats % 16 == 0) && ((uintptr_t)out % 16 == 0);
  if (vec)
    interp_fwd_kernel<float4><<<flat_grid_interp(n * (C / 4)), 256, 0, stream>>>((const float4*)feats, idx, weights, n,
                                                                                  C / 4, K, (float4*)out);
  else
    interp_fwd_kernel<float><<<flat_grid_interp(n * C), 256, 0, stream>>>(feats, idx, weights, n, C, K, out);
  SPC_LAUNCHED("interp_fwd_kernel");
  return 0;
}

int spc_interp_bwd(const float* dout, const int32_t* idx, const float* weights, int64_t n, int64_t m, int C, int K,
                   float* dfeats, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && K >= 1, "bad shape");
  if (m > 0) SPC_CUDA(cudaMemsetAsync(dfeats, 0, (size_t)m * C * sizeof(float), stream));
  if (n == 0 || m == 0) return 0;
  interp_bwd_kernel<<<flat_grid_interp(n * C), 256, 0, stream>>>(dout, idx, weights, n, C, K, dfeats);
  SPC_LAUNCHED("interp_bwd_kernel");
  return 0;
}

}  // extern "C"
