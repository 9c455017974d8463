// conv_simt.cu — FP32 CUDA-core sparse convolution (precision mode SPC_PREC_FP32).
//
// This is the fp32-faithful arithmetic path: the same output-stationary formulation as the
// tcgen05 kernels (conv_umma.cu) but with plain FFMA, used (a) for the strict-tolerance
// parity tests against the fp64 oracle, (b) for shapes the tensor-core path does not take
// (C not a multiple of 16), and (c) as the on-device cross-check of the TF32 kernels.
//
// Semantics (restated in-tree at co3d_3d/src/models/mink/modules/sparse_conv.py:122-143):
//   fwd   out[o,:] = sum_k in[nbr[k,o],:] @ W[k]            W: [K, Cin, Cout]
//   dgrad din[i,:] = sum_k dout[nbr_t[k,i],:] @ W[k]^T       (same kernel, TRANSPOSE_W)
//   wgrad dW[k]    = sum_o in[nbr[k,o],:]^T dout[o,:]
// Output rows are owned by exactly one CTA, so forward/dgrad need no atomics.
#include "common.cuh"

namespace spc {

constexpr int BM = 64, BN = 64, BK = 16;

// out[M, Cn] = sum_k gather(A, nbr[k])[M, Ck] * B_k[Ck, Cn]
//   TRANSPOSE_W = false: B_k[c][n] = W[k][c][n]   (ldw = Cn)
//   TRANSPOSE_W = true : B_k[c][n] = W[k][n][c]   (ldw = Ck)
template <bool TRANSPOSE_W>
__global__ void __launch_bounds__(256)
conv_gather_gemm_kernel(const float* __restrict__ A, const float* __restrict__ W,
                        const float* __restrict__ bias, const int* __restrict__ nbr, int m_out,
                        int Ck, int Cn, int K, float* __restrict__ out) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ int s_row[BM];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int o0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool vec_a = (Ck % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

  for (int k = 0; k < K; ++k) {
    int r = -1;
    if (tid < BM && o0 + tid < m_out) r = nbr[(size_t)k * m_out + o0 + tid];
    __syncthreads();  // previous tap's readers of s_row / tiles are done
    if (tid < BM) s_row[tid] = r;
    if (!__syncthreads_or(r >= 0)) continue;  // no neighbour in this tile for this tap
    const float* Wk = W + (size_t)k * Ck * Cn;
    for (int c0 = 0; c0 < Ck; c0 += BK) {
      {  // A tile: 64 rows x 16 channels, one float4 per thread
        const int row = tid / 4, c4 = (tid % 4) * 4;
        const int src = s_row[row];
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (src >= 0) {
          const float* p = A + (size_t)src * Ck + c0 + c4;
          if (vec_a && c0 + c4 + 3 < Ck) {
            float4 t = *reinterpret_cast<const float4*>(p);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c0 + c4 + j < Ck) v[j] = p[j];
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) As[c4 + j][row] = v[j];
      }
      {  // B tile: 16 x 64
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int e = tid + q * 256;
          int kk, n;
          if (TRANSPOSE_W) { kk = e % BK; n = e / BK; } else { kk = e / BN; n = e % BN; }
          float v = 0.f;
          if (c0 + kk < Ck && n0 + n < Cn)
            v = TRANSPOSE_W ? Wk[(size_t)(n0 + n) * Ck + c0 + kk] : Wk[(size_t)(c0 + kk) * Cn + n0 + n];
          Bs[kk][n] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + ty * 4 + i;
    if (o >= m_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Cn) out[(size_t)o * Cn + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// dW[k][ci][co] += sum over a chunk of out rows of in[nbr[k,o]][ci] * dout[o][co]
constexpr int WG_ROWS = 2048;  // out rows per CTA (split-K)
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dout,
                  const int* __restrict__ nbr, int m_out, int Cin, int Cout, int K, int n_co_tiles,
                  float* __restrict__ dw) {
  __shared__ __align__(16) float As[BK][BM + 4];  // [row][ci]
  __shared__ __align__(16) float Bs[BK][BN + 4];  // [row][co]
  __shared__ int s_row[BK];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int k = blockIdx.y;
  const int ci0 = (blockIdx.z / n_co_tiles) * BM, co0 = (blockIdx.z % n_co_tiles) * BN;
  const int r_begin = blockIdx.x * WG_ROWS;
  const int r_end = min(r_begin + WG_ROWS, m_out);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  bool any = false;
  const bool vec_a = (Cin % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
  const bool vec_b = (Cout % 4 == 0) && ((reinterpret_cast<uintptr_t>(dout) & 15) == 0);
  for (int o0 = r_begin; o0 < r_end; o0 += BK) {
    int r = -1;
    if (tid < BK && o0 + tid < r_end) r = nbr[(size_t)k * m_out + o0 + tid];
    __syncthreads();
    if (tid < BK) s_row[tid] = r;
    if (!__syncthreads_or(r >= 0)) continue;
    any = true;
    const int row = tid / 16, c4 = (tid % 16) * 4;
    {
      const int src = s_row[row];
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (src >= 0) {
        const float* p = in + (size_t)src * Cin + ci0 + c4;
        if (vec_a && ci0 + c4 + 3 < Cin) {
          float4 t = *reinterpret_cast<const float4*>(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ci0 + c4 + j < Cin) v[j] = p[j];
        }
      }
      *reinterpret_cast<float4*>(&As[row][c4]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (s_row[row] >= 0) {  // rows without a neighbour contribute nothing
        const float* p = dout + (size_t)(o0 + row) * Cout + co0 + c4;
        if (vec_b && co0 + c4 + 3 < Cout) {
          float4 t = *reinterpret_cast<const float4*>(p);
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (co0 + c4 + j < Cout) v[j] = p[j];
        }
      }
      *reinterpret_cast<float4*>(&Bs[row][c4]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  if (!any) return;
  float* dwk = dw + (size_t)k * Cin * Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < Cout && acc[i][j] != 0.f) atomicAdd(dwk + (size_t)ci * Cout + co, acc[i][j]);
    }
  }
}

int conv_fwd_simt(const float* in, const float* w, const float* bias, const int* nbr,
                  int64_t m_out, int c_in, int c_out, int K, bool transpose_w, float* out,
                  cudaStream_t stream) {
  if (m_out == 0) return 0;
  dim3 grid((unsigned)ceil_div(m_out, BM), (unsigned)ceil_div(c_out, BN));
  if (transpose_w)
    conv_gather_gemm_kernel<true><<<grid, 256, 0, stream>>>(in, w, bias, nbr, (int)m_out, c_in, c_out, K, out);
  else
    conv_gather_gemm_kernel<false><<<grid, 256, 0, stream>>>(in, w, bias, nbr, (int)m_out, c_in, c_out, K, out);
  SPC_LAUNCHED("conv_gather_gemm_kernel");
  return 0;
}

int conv_wgrad_simt(const float* in, const float* dout, const int* nbr, int64_t m_out, int c_in,
                    int c_out, int K, float* dw, cudaStream_t stream) {
  SPC_CUDA(cudaMemsetAsync(dw, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (m_out == 0) return 0;
  int n_ci = (int)ceil_div(c_in, BM), n_co = (int)ceil_div(c_out, BN);
  dim3 grid((unsigned)ceil_div(m_out, WG_ROWS), (unsigned)K, (unsigned)(n_ci * n_co));
  conv_wgrad_kernel<<<grid, 256, 0, stream>>>(in, dout, nbr, (int)m_out, c_in, c_out, K, n_co, dw);
  SPC_LAUNCHED("conv_wgrad_kernel");
  return 0;
}

}  // namespace spc
