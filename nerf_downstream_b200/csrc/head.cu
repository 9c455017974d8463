// head.cu — the output end of the hot path and the instance normalisation of the other in-tree backbones
// (SURVEY.md §8f "next" rows 3 and 4):
//   seg_head      : out.slice(x).F -> SegLoss -> IoUMeter.update as ONE pass over the N points: inverse-map gather of
//                   the voxel logits, weighted softmax cross-entropy with ignore index, raw gradient accumulated
//                   straight onto the voxel rows (= backward of slice), per-class seen / correct / predicted counts
//                   (co3d_3d/src/models/mink/res16unet.py:435, src/modules/segmentation_training.py:27-44,219-228,
//                   src/metrics.py:29-41).  With inverse == NULL it is a class-weighted cross-entropy over plain rows.
//   instance norm : ME.MinkowskiInstanceNorm (modules/common.py:25-26, resunet.py "IN" variants): per (batch index,
//                   channel) mean / biased variance over the rows of that instance, (x - mean) / sqrt(var + eps), then
//                   the [1, C] affine.
#include <limits.h>

#include "common.cuh"

namespace spc {

// ---------------------------------------------------------------------------------------------------------------
// One thread per point; the C <= CMAX logits of its voxel row stay in registers.
//   stats[0] += sum_i w[y_i] * (-log p_i[y_i]),  stats[1] += sum_i w[y_i]            (i over non-ignored points)
//   graw[row(i), c] (+)= w[y_i] * (p_i[c] - [c == y_i])       row(i) = inverse ? inverse[i] : i
//   counts[0][c] += #(y == c), counts[1][c] += #(y == c and argmax == c), counts[2][c] += #(argmax == c)
// With an inverse map several points may share a voxel row, so the gradient goes through red.global.add (graw zeroed by
// the wrapper); without one every row is written exactly once with plain stores (zeros for ignored rows).
// bad[0] = 1: a target outside [0, C) that is not the ignore index; 2: an inverse-map entry outside [0, m).
// ---------------------------------------------------------------------------------------------------------------
template <int CMAX>
__global__ void __launch_bounds__(256)
seg_head_kernel(const float* __restrict__ logits, long long m, const int* __restrict__ inverse,
                const long long* __restrict__ target, long long n, int C, long long ignore_index,
                const float* __restrict__ weight, float* __restrict__ graw, double* __restrict__ stats,
                unsigned long long* __restrict__ counts, int* __restrict__ bad) {
  __shared__ unsigned int s_cnt[3 * CMAX];
  __shared__ float s_loss[8], s_w[8];
  if (counts) {
    for (int i = threadIdx.x; i < 3 * CMAX; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
  }
  float loss = 0.f, wsum = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long y = target[i];
    long long row = i;
    if (inverse) {
      row = inverse[i];
      if (row < 0 || row >= m) { *bad = 2; continue; }
    }
    float* g = graw + row * C;
    if (y == ignore_index) {
      if (!inverse)
        for (int c = 0; c < C; ++c) g[c] = 0.f;
      continue;
    }
    float v[CMAX];
    float mx = -INFINITY;
    int best = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        v[c] = logits[row * C + c];
        if (v[c] > mx) { mx = v[c]; best = c; }    // ties -> lowest class index, as spc_seg_metrics
      }
    const bool in_range = y >= 0 && y < C;
    if (counts) {
      atomicAdd(&s_cnt[2 * CMAX + best], 1u);
      if (in_range) {
        atomicAdd(&s_cnt[(int)y], 1u);
        if (best == (int)y) atomicAdd(&s_cnt[CMAX + (int)y], 1u);
      }
    }
    if (!in_range) {
      *bad = 1;
      if (!inverse)
        for (int c = 0; c < C; ++c) g[c] = 0.f;
      continue;
    }
    const float w = weight ? weight[(int)y] : 1.f;
    float sum = 0.f, vy = 0.f;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        if (c == (int)y) vy = v[c];
        v[c] = __expf(v[c] - mx);
        sum += v[c];
      }
    const float inv = 1.f / sum;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) {
        const float d = w * (v[c] * inv - (c == (int)y ? 1.f : 0.f));
        if (inverse) atomicAdd(g + c, d); else g[c] = d;
      }
    loss += w * ((mx - vy) + __logf(sum));
    wsum += w;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, d);
    wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  }
  if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss; s_w[threadIdx.x >> 5] = wsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0, k = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { l += (double)s_loss[w]; k += (double)s_w[w]; }
    if (k != 0.0 || l != 0.0) { atomicAdd(stats, l); atomicAdd(stats + 1, k); }
  }
  if (counts)
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
      const int which = i / C, c = i - which * C;
      const unsigned int v = s_cnt[which * CMAX + c];
      if (v) atomicAdd(&counts[which * C + c], (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// instance norm.  256 threads = 32 channel lanes x 8 row slots; blockIdx.y picks the 32-channel tile, blockIdx.x a
// block of rows.  Rows of one batch index are (almost always) contiguous, so a thread keeps running sums and flushes
// them with double atomics when the batch index changes.
//   FWD:  sums[b][0][c] += x          sums[b][1][c] += x^2          cnt[b] += 1 (channel 0 only)
//   BWD:  sums[b][0][c] += dy         sums[b][1][c] += dy * xhat    xhat = (x - mean[b,c]) * rstd[b,c]
// ---------------------------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(256)
inst_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const int4* __restrict__ coords, long long m, int C, int n_batch,
                 int rows_per_block, double* __restrict__ sums, int* __restrict__ cnt) {
  const int c = blockIdx.y * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  if (c >= C) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < m ? r0 + rows_per_block : m;
  float a0 = 0.f, a1 = 0.f, mu = 0.f, rs = 0.f;
  int cur = -1, k = 0;
  for (long long r = r0 + rl; r < r1; r += 8) {
    const int b = coords[r].x;
    if (b != cur) {
      if (cur >= 0 && cur < n_batch && k) {
        atomicAdd(sums + ((long long)cur * 2 + 0) * C + c, (double)a0);
        atomicAdd(sums + ((long long)cur * 2 + 1) * C + c, (double)a1);
        if (!BWD && c == 0) atomicAdd(cnt + cur, k);
      }
      cur = b; a0 = a1 = 0.f; k = 0;
      if (BWD && b >= 0 && b < n_batch) { mu = mean[(long long)b * C + c]; rs = rstd[(long long)b * C + c]; }
    }
    const float v = x[r * C + c];
    if (BWD) {
      const float g = dy[r * C + c];
      a0 += g;
      a1 += g * ((v - mu) * rs);
    } else {
      a0 += v;
      a1 += v * v;
    }
    ++k;
  }
  if (cur >= 0 && cur < n_batch && k) {
    atomicAdd(sums + ((long long)cur * 2 + 0) * C + c, (double)a0);
    atomicAdd(sums + ((long long)cur * 2 + 1) * C + c, (double)a1);
    if (!BWD && c == 0) atomicAdd(cnt + cur, k);
  }
}

// mean = s0 / n, var = s1 / n - mean^2 (biased, clamped at 0), rstd = 1 / sqrt(var + eps)
__global__ void inst_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ cnt, int n_batch, int C,
                                     float eps, float* __restrict__ mean, float* __restrict__ rstd) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_batch * C) return;
  const int b = e / C, c = e - b * C;
  const int n = cnt[b];
  double mu = 0.0, var = 0.0;
  if (n > 0) {
    mu = sums[((long long)b * 2 + 0) * C + c] / n;
    var = sums[((long long)b * 2 + 1) * C + c] / n - mu * mu;
    if (var < 0.0) var = 0.0;
  }
  mean[e] = (float)mu;
  rstd[e] = (float)(1.0 / sqrt(var + (double)eps));
}

// y = (x - mean[b]) * rstd[b] * gamma + beta
__global__ void __launch_bounds__(256)
inst_apply_kernel(const float* __restrict__ x, const int4* __restrict__ coords, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                  long long total, int C, int n_batch, float* __restrict__ y) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    const int b = coords[r].x;
    float o = 0.f;
    if (b >= 0 && b < n_batch) {
      const long long s = (long long)b * C + c;
      o = (x[e] - mean[s]) * rstd[s];
      o = o * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f);
    }
    y[e] = o;
  }
}

// dx = gamma * rstd * (dy - S0 / n - xhat * S1 / n),  S0 = sum dy, S1 = sum dy * xhat over the instance
__global__ void __launch_bounds__(256)
inst_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const int4* __restrict__ coords,
                      const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                      const double* __restrict__ sums, const int* __restrict__ cnt, long long total, int C, int n_batch,
                      float* __restrict__ dx) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / C;
    const int c = (int)(e - r * C);
    const int b = coords[r].x;
    float o = 0.f;
    if (b >= 0 && b < n_batch) {
      const long long s = (long long)b * C + c;
      const float inv_n = 1.f / (float)cnt[b];
      const float s0 = (float)sums[((long long)b * 2 + 0) * C + c] * inv_n;
      const float s1 = (float)sums[((long long)b * 2 + 1) * C + c] * inv_n;
      const float xh = (x[e] - mean[s]) * rstd[s];
      o = (gamma ? gamma[c] : 1.f) * rstd[s] * (dy[e] - s0 - xh * s1);
    }
    dx[e] = o;
  }
}

// ---- 16-byte versions (C % 4 == 0): a thread owns a channel QUAD, issues four rows' loads before it consumes them, and
// the per-thread partial sums of a block that lies inside one instance meet in shared memory, so that only `lanes`
// threads per block touch the double atomics.  lanes = quads per block row (<= 256), rows = 256 / lanes row slots.
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

__device__ __forceinline__ void inst_flush_direct(double* sums, int* cnt, int cur, int n_batch, int C, int cq,
                                                  const float4& a0, const float4& a1, int k, bool count) {
  if (cur < 0 || cur >= n_batch || k == 0) return;
  double* p0 = sums + ((long long)cur * 2 + 0) * C + 4 * cq;
  double* p1 = sums + ((long long)cur * 2 + 1) * C + 4 * cq;
  atomicAdd(p0 + 0, (double)a0.x); atomicAdd(p0 + 1, (double)a0.y); atomicAdd(p0 + 2, (double)a0.z); atomicAdd(p0 + 3, (double)a0.w);
  atomicAdd(p1 + 0, (double)a1.x); atomicAdd(p1 + 1, (double)a1.y); atomicAdd(p1 + 2, (double)a1.z); atomicAdd(p1 + 3, (double)a1.w);
  if (count && cq == 0) atomicAdd(cnt + cur, k);
}

template <bool BWD>
__global__ void __launch_bounds__(256)
inst_sums4_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float* __restrict__ mean,
                  const float* __restrict__ rstd, const int4* __restrict__ coords, long long m, int C, int C4,
                  int n_batch, int lanes, int rows, int rows_per_block, double* __restrict__ sums,
                  int* __restrict__ cnt) {
  __shared__ float4 s0[256], s1[256];
  __shared__ int s_k[256];
  __shared__ int s_cur;
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const int cq = blockIdx.y * lanes + lane;
  const bool active = rl < rows && cq < C4;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < m ? r0 + rows_per_block : m;
  float4 a0 = f4_zero(), a1 = f4_zero(), mu = f4_zero(), rs = f4_zero();
  int cur = -1, k = 0;
  if (active) {
    for (long long r = r0 + rl; r < r1; r += (long long)rows * 4) {
      int b[4];
      float4 v[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long rr = r + (long long)u * rows;
        b[u] = INT_MIN;
        if (rr < r1) {
          b[u] = coords[rr].x;
          v[u] = x[rr * C4 + cq];
          if (BWD) g[u] = dy[rr * C4 + cq];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b[u] == INT_MIN) continue;
        if (b[u] != cur) {
          inst_flush_direct(sums, cnt, cur, n_batch, C, cq, a0, a1, k, !BWD);
          cur = b[u]; a0 = f4_zero(); a1 = f4_zero(); k = 0;
          if (BWD && cur >= 0 && cur < n_batch) {
            mu = *reinterpret_cast<const float4*>(mean + (long long)cur * C + 4 * cq);
            rs = *reinterpret_cast<const float4*>(rstd + (long long)cur * C + 4 * cq);
          }
        }
        if (BWD) {
          f4_add(a0, g[u]);
          a1.x += g[u].x * ((v[u].x - mu.x) * rs.x); a1.y += g[u].y * ((v[u].y - mu.y) * rs.y);
          a1.z += g[u].z * ((v[u].z - mu.z) * rs.z); a1.w += g[u].w * ((v[u].w - mu.w) * rs.w);
        } else {
          f4_add(a0, v[u]);
          a1.x += v[u].x * v[u].x; a1.y += v[u].y * v[u].y; a1.z += v[u].z * v[u].z; a1.w += v[u].w * v[u].w;
        }
        ++k;
      }
    }
  }
  // final flush: through shared memory when every partial sum of the block belongs to the same instance
  if (threadIdx.x == 0) s_cur = cur;                 // thread 0 always owns row r0 of a launched block
  __syncthreads();
  const int bcur = s_cur;
  const bool mine = active && k > 0;
  const int uniform = __syncthreads_and(!mine || cur == bcur);
  if (!uniform || bcur < 0 || bcur >= n_batch) {
    if (mine) inst_flush_direct(sums, cnt, cur, n_batch, C, cq, a0, a1, k, !BWD);
    return;
  }
  s0[threadIdx.x] = mine ? a0 : f4_zero();
  s1[threadIdx.x] = mine ? a1 : f4_zero();
  s_k[threadIdx.x] = mine ? k : 0;
  __syncthreads();
  if (rl == 0 && cq < C4) {
    float4 t0 = f4_zero(), t1 = f4_zero();
    int kk = 0;
    for (int sl = 0; sl < rows; ++sl) {
      f4_add(t0, s0[sl * lanes + lane]);
      f4_add(t1, s1[sl * lanes + lane]);
      kk += s_k[sl * lanes + lane];
    }
    inst_flush_direct(sums, cnt, bcur, n_batch, C, cq, t0, t1, kk, !BWD);
  }
}

__global__ void __launch_bounds__(256)
inst_apply4_kernel(const float4* __restrict__ x, const int4* __restrict__ coords, const float* __restrict__ mean,
                   const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                   long long total4, int C, int C4, int n_batch, float4* __restrict__ y) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += stride) {
    const long long r = e / C4;
    const int cq = (int)(e - r * C4);
    const int b = coords[r].x;
    const float4 v = x[e];
    float4 o = f4_zero();
    if (b >= 0 && b < n_batch) {
      const float4 mu = *reinterpret_cast<const float4*>(mean + (long long)b * C + 4 * cq);
      const float4 rs = *reinterpret_cast<const float4*>(rstd + (long long)b * C + 4 * cq);
      const float4 ga = gamma ? *reinterpret_cast<const float4*>(gamma + 4 * cq) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 be = beta ? *reinterpret_cast<const float4*>(beta + 4 * cq) : f4_zero();
      o.x = (v.x - mu.x) * rs.x * ga.x + be.x; o.y = (v.y - mu.y) * rs.y * ga.y + be.y;
      o.z = (v.z - mu.z) * rs.z * ga.z + be.z; o.w = (v.w - mu.w) * rs.w * ga.w + be.w;
    }
    y[e] = o;
  }
}

__global__ void __launch_bounds__(256)
inst_bwd_apply4_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const int4* __restrict__ coords,
                       const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                       const double* __restrict__ sums, const int* __restrict__ cnt, long long total4, int C, int C4,
                       int n_batch, float4* __restrict__ dx) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += stride) {
    const long long r = e / C4;
    const int cq = (int)(e - r * C4);
    const int b = coords[r].x;
    const float4 v = x[e], g = dy[e];
    float4 o = f4_zero();
    if (b >= 0 && b < n_batch) {
      const float4 mu = *reinterpret_cast<const float4*>(mean + (long long)b * C + 4 * cq);
      const float4 rs = *reinterpret_cast<const float4*>(rstd + (long long)b * C + 4 * cq);
      const float4 ga = gamma ? *reinterpret_cast<const float4*>(gamma + 4 * cq) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float inv_n = 1.f / (float)cnt[b];
      const double* p0 = sums + ((long long)b * 2 + 0) * C + 4 * cq;
      const double* p1 = sums + ((long long)b * 2 + 1) * C + 4 * cq;
      o.x = ga.x * rs.x * (g.x - (float)p0[0] * inv_n - (v.x - mu.x) * rs.x * ((float)p1[0] * inv_n));
      o.y = ga.y * rs.y * (g.y - (float)p0[1] * inv_n - (v.y - mu.y) * rs.y * ((float)p1[1] * inv_n));
      o.z = ga.z * rs.z * (g.z - (float)p0[2] * inv_n - (v.z - mu.z) * rs.z * ((float)p1[2] * inv_n));
      o.w = ga.w * rs.w * (g.w - (float)p0[3] * inv_n - (v.w - mu.w) * rs.w * ((float)p1[3] * inv_n));
    }
    dx[e] = o;
  }
}

struct QuadMap { int lanes, rows, ytiles; };
static QuadMap quad_map(int C4) {
  QuadMap q;
  q.lanes = C4 < 256 ? C4 : 256;
  q.rows = 256 / q.lanes;
  q.ytiles = (int)ceil_div(C4, q.lanes);
  return q;
}
static bool aligned16(const void* p) { return ((uintptr_t)p % 16) == 0; }
static int g_inst_scalar_only = 0;   // debug knob: force the scalar kernels (cross-check of the 16-byte path)

static int flat_grid(long long total) {
  int64_t want = ceil_div(total, 256 * 4);
  return (int)(want < kNumSMs * 16 ? (want > 0 ? want : 1) : kNumSMs * 16);
}

}  // namespace spc

using namespace spc;

extern "C" {

int spc_seg_head_fwd(const float* logits, int64_t m, const int32_t* inverse, const int64_t* target, int64_t n, int C,
                     int64_t ignore_index, const float* class_weight, float* grad_raw, double* stats,
                     uint64_t* counts, int32_t* bad_target, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && C <= 64, "the segmentation head supports 1..64 classes");
  SPC_REQUIRE(inverse != nullptr || m == n, "without an inverse map there is one target per logits row");
  SPC_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double), stream));
  SPC_CUDA(cudaMemsetAsync(bad_target, 0, sizeof(int32_t), stream));
  if (inverse && m > 0) SPC_CUDA(cudaMemsetAsync(grad_raw, 0, (size_t)m * C * sizeof(float), stream));
  if (n == 0) return 0;
  int64_t want = ceil_div(n, 256);
  int grid = (int)(want < kNumSMs * 8 ? want : kNumSMs * 8);
  if (C <= 32)
    seg_head_kernel<32><<<grid, 256, 0, stream>>>(logits, m, inverse, (const long long*)target, n, C, ignore_index,
                                                   class_weight, grad_raw, stats, (unsigned long long*)counts, bad_target);
  else
    seg_head_kernel<64><<<grid, 256, 0, stream>>>(logits, m, inverse, (const long long*)target, n, C, ignore_index,
                                                   class_weight, grad_raw, stats, (unsigned long long*)counts, bad_target);
  SPC_LAUNCHED("seg_head_kernel");
  return 0;
}

void spc_inst_norm_force_scalar(int on) { g_inst_scalar_only = on; }

int spc_inst_norm_fwd(const float* x, const int32_t* coords, int64_t m, int C, int n_batch, const float* gamma,
                      const float* beta, float eps, float* y, float* mean, float* rstd, int32_t* cnt, double* ws,
                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && n_batch >= 1, "bad shape");
  SPC_CUDA(cudaMemsetAsync(ws, 0, (size_t)n_batch * 2 * C * sizeof(double), stream));
  SPC_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n_batch * sizeof(int32_t), stream));
  const bool vec = !g_inst_scalar_only && C % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(mean) &&
                   aligned16(rstd) && (!gamma || aligned16(gamma)) && (!beta || aligned16(beta));
  const int rows_per_block = vec ? 1024 : 512;
  if (m > 0 && vec) {
    const QuadMap qm = quad_map(C / 4);
    dim3 grid((unsigned)ceil_div(m, rows_per_block), (unsigned)qm.ytiles);
    inst_sums4_kernel<false><<<grid, 256, 0, stream>>>((const float4*)x, nullptr, nullptr, nullptr, (const int4*)coords,
                                                        m, C, C / 4, n_batch, qm.lanes, qm.rows, rows_per_block, ws, cnt);
    SPC_LAUNCHED("inst_sums4_kernel");
  } else if (m > 0) {
    dim3 grid((unsigned)ceil_div(m, rows_per_block), (unsigned)ceil_div(C, 32));
    inst_sums_kernel<false><<<grid, 256, 0, stream>>>(x, nullptr, nullptr, nullptr, (const int4*)coords, m, C, n_batch,
                                                       rows_per_block, ws, cnt);
    SPC_LAUNCHED("inst_sums_kernel");
  }
  inst_finalize_kernel<<<(int)ceil_div((int64_t)n_batch * C, 256), 256, 0, stream>>>(ws, cnt, n_batch, C, eps, mean, rstd);
  SPC_LAUNCHED("inst_finalize_kernel");
  if (m == 0) return 0;
  const long long total = (long long)m * C;
  if (vec) {
    inst_apply4_kernel<<<flat_grid(total), 256, 0, stream>>>((const float4*)x, (const int4*)coords, mean, rstd, gamma,
                                                              beta, total / 4, C, C / 4, n_batch, (float4*)y);
    SPC_LAUNCHED("inst_apply4_kernel");
    return 0;
  }
  inst_apply_kernel<<<flat_grid(total), 256, 0, stream>>>(x, (const int4*)coords, mean, rstd, gamma, beta, total, C,
                                                           n_batch, y);
  SPC_LAUNCHED("inst_apply_kernel");
  return 0;
}

int spc_inst_norm_bwd(const float* x, const float* dy, const int32_t* coords, int64_t m, int C, int n_batch,
                      const float* gamma, const float* mean, const float* rstd, const int32_t* cnt, float* dx,
                      double* sums, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && n_batch >= 1, "bad shape");
  SPC_CUDA(cudaMemsetAsync(sums, 0, (size_t)n_batch * 2 * C * sizeof(double), stream));
  if (m == 0) return 0;
  const long long total = (long long)m * C;
  const bool vec = !g_inst_scalar_only && C % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) &&
                   aligned16(mean) && aligned16(rstd) && (!gamma || aligned16(gamma));
  if (vec) {
    const int rpb = 1024;
    const QuadMap qm = quad_map(C / 4);
    dim3 grid4((unsigned)ceil_div(m, rpb), (unsigned)qm.ytiles);
    inst_sums4_kernel<true><<<grid4, 256, 0, stream>>>((const float4*)x, (const float4*)dy, mean, rstd, (const int4*)coords,
                                                        m, C, C / 4, n_batch, qm.lanes, qm.rows, rpb, sums, nullptr);
    SPC_LAUNCHED("inst_sums4_kernel");
    inst_bwd_apply4_kernel<<<flat_grid(total), 256, 0, stream>>>((const float4*)x, (const float4*)dy, (const int4*)coords,
                                                                  mean, rstd, gamma, sums, cnt, total / 4, C, C / 4,
                                                                  n_batch, (float4*)dx);
    SPC_LAUNCHED("inst_bwd_apply4_kernel");
    return 0;
  }
  const int rows_per_block = 512;
  dim3 grid((unsigned)ceil_div(m, rows_per_block), (unsigned)ceil_div(C, 32));
  inst_sums_kernel<true><<<grid, 256, 0, stream>>>(x, dy, mean, rstd, (const int4*)coords, m, C, n_batch, rows_per_block,
                                                    sums, nullptr);
  SPC_LAUNCHED("inst_sums_kernel");
  inst_bwd_apply_kernel<<<flat_grid(total), 256, 0, stream>>>(x, dy, (const int4*)coords, mean, rstd, gamma, sums, cnt,
                                                               total, C, n_batch, dx);
  SPC_LAUNCHED("inst_bwd_apply_kernel");
  return 0;
}

}  // extern "C"
