// elementwise.cu — bandwidth-bound kernels on voxel feature rows [M, C] (fp32, row-major):
// BatchNorm statistics / apply / backward with fused ReLU and residual add, ReLU, add,
// local and global pooling, fused SGD.  SURVEY.md §8a rows a7-a10.
//
// Thread mapping for all [M,C] kernels: a block owns R consecutive rows per iteration and
// every thread a fixed group of VEC channels (VEC = 4 when C % 4 == 0, else 1), so
//   - per-channel parameters (scale/shift) live in registers,
//   - a warp always touches one contiguous span of memory (rows are contiguous),
//   - the grid is a multiple of the SM count and strides over row groups.
#include <cuda_bf16.h>

#include "common.cuh"

namespace spc {

struct RowMap {
  int lanes;    // threads per row = C / VEC
  int rows;     // rows per block iteration
  int threads;  // lanes * rows
};
static inline RowMap make_row_map(int C, int vec) {
  RowMap r;
  r.lanes = C / vec;
  r.rows = r.lanes >= 256 ? 1 : 256 / r.lanes;
  r.threads = r.lanes * r.rows;
  return r;
}
static inline int pick_vec(int C, const void* a, const void* b = nullptr, const void* c = nullptr,
                           const void* d = nullptr) {
  auto al = [](const void* p) { return p == nullptr || ((uintptr_t)p % 16) == 0; };
  return (C % 4 == 0 && al(a) && al(b) && al(c) && al(d)) ? 4 : 1;
}
static inline int pick_grid(int64_t m, int rows_per_iter, int blocks_per_sm) {
  int64_t need = ceil_div(m, rows_per_iter);
  int64_t cap = (int64_t)kNumSMs * blocks_per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

template <int VEC> struct Vec;
// round-to-nearest-even bf16 copy of VEC values (the operand format of the SPC_PREC_BF16 convolutions)
template <int VEC>
__device__ __forceinline__ void store_bf16(__nv_bfloat16* p, const float* v) {
  if (VEC == 4) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = o;
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) p[j] = __float2bfloat16_rn(v[j]);
  }
}
template <> struct Vec<1> {
  float v[1];
  __device__ static Vec load(const float* p) { Vec r; r.v[0] = *p; return r; }
  __device__ void store(float* p) const { *p = v[0]; }
};
template <> struct Vec<4> {
  float v[4];
  __device__ static Vec load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    Vec r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
  }
  __device__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// ReLU mask source: the fp32 output rows, or their bf16 copy (half the bytes; the sign / zero-ness of a
// value survives round-to-nearest bf16 for everything above 1e-40)
template <int VEC>
__device__ __forceinline__ Vec<VEC> load_mask_rows(const float* y, const __nv_bfloat16* yb, long long off) {
  if (yb == nullptr) return Vec<VEC>::load(y + off);
  Vec<VEC> r;
  if (VEC == 4) {
    uint2 t = *reinterpret_cast<const uint2*>(yb + off);
    r.v[0] = __uint_as_float(t.x << 16);
    r.v[1] = __uint_as_float(t.x & 0xFFFF0000u);
    r.v[2] = __uint_as_float(t.y << 16);
    r.v[3] = __uint_as_float(t.y & 0xFFFF0000u);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) r.v[j] = __bfloat162float(yb[off + j]);
  }
  return r;
}

// The affine form of BatchNorm as the apply kernel evaluates it: y = fma(x, sc, sh).  Backward re-evaluates exactly
// this expression to recover the ReLU mask of a BN+ReLU WITHOUT a residual (relu mode 2: y > 0 <=> fma(x, sc, sh) > 0)
// instead of reading the output rows again, so all three kernels share the one definition.
__device__ __forceinline__ void bn_scale_shift(float gamma, float beta, float mean, float var, float eps, float& sc,
                                               float& sh) {
  sc = gamma * rsqrtf(var + eps);
  sh = __fmaf_rn(-mean, sc, beta);
}

// ---------------------------------------------------------------------------
// column reductions: up to two sums per channel, partials per block, double finalise
// ---------------------------------------------------------------------------
// MODE 0: s0 = sum x, s1 = sum x^2                     (BN statistics)
// MODE 1: s0 = sum dy', s1 = sum dy' * (x - mean)      (BN backward; dy' = relu-masked dy)
// Per-channel double sums live in `gsum[2C]` (zeroed by the host wrapper); the block that takes
// the last ticket of `counter` finalises, so statistics cost ONE launch.
struct ColFinal {
  double* gsum;          // [2C]
  unsigned* counter;
  float* out0;           // MODE 0: mean    | MODE 1: dbeta
  float* out1;           // MODE 0: var     | MODE 1: dgamma
  float* raw;            // MODE 1: float sums [2C] for the apply kernel
  float* run_mean;       // MODE 0, optional
  float* run_var;
  const float* var;      // MODE 1
  float momentum, eps;
  int accumulate;        // MODE 1: out0 / out1 += (slices of a gradient arena zeroed once per step) instead of =
  long long* tracked;    // MODE 0, optional: nn.BatchNorm1d.num_batches_tracked, incremented by the last block
};

template <int VEC, int MODE>
__global__ void __launch_bounds__(1024)
col_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y,
                  const __nv_bfloat16* __restrict__ yb, const float* __restrict__ dy, long long dy_pitch,
                  const float* __restrict__ mean, const float* __restrict__ gamma, const float* __restrict__ beta,
                  long long m, int C, int lanes, int rows, int relu, ColFinal fin) {
  extern __shared__ float sm[];  // [rows][2*C]
  __shared__ bool s_last;
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const int c0 = lane * VEC;
  float s0[VEC], s1[VEC], mu[VEC];
  // MODE 0 accumulates around a pivot (row 0) so that E[x^2] - E[x]^2 does not cancel
#pragma unroll
  for (int j = 0; j < VEC; ++j) { s0[j] = 0.f; s1[j] = 0.f; mu[j] = MODE == 1 ? mean[c0 + j] : x[c0 + j]; }
  float sc[VEC], sh[VEC];  // relu mode 2 (MODE 1): the forward affine, to recompute the ReLU mask from x
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sc[j] = 0.f; sh[j] = 0.f;
    if (MODE == 1 && relu == 2)
      bn_scale_shift(gamma ? gamma[c0 + j] : 1.f, beta ? beta[c0 + j] : 0.f, mu[j], fin.var[c0 + j], fin.eps, sc[j], sh[j]);
  }
  // U row groups per iteration: all their loads are issued before the first use, so a thread keeps U (MODE 0) or
  // up to 3 U (MODE 1) 16-byte loads in flight instead of one per stream (r1: 0.78-0.84 of the HBM peak)
  constexpr int U = MODE == 0 ? 4 : 2;
  const long long stride = (long long)gridDim.x * rows;
  for (long long r = (long long)blockIdx.x * rows + rl; r < m; r += U * stride) {
    Vec<VEC> xv[U], g[U], yv[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * stride;
      ok[u] = ru < m;
      if (ok[u]) {
        const long long off = ru * C + c0;
        xv[u] = Vec<VEC>::load(x + off);
        if (MODE == 1) {
          g[u] = Vec<VEC>::load(dy + ru * dy_pitch + c0);
          if (relu == 1) yv[u] = load_mask_rows<VEC>(y, yb, off);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { float d = xv[u].v[j] - mu[j]; s0[j] += d; s1[j] += d * d; }
      } else {
        if (relu == 1) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) if (!(yv[u].v[j] > 0.f)) g[u].v[j] = 0.f;
        } else if (relu == 2) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) if (!(fmaf(xv[u].v[j], sc[j], sh[j]) > 0.f)) g[u].v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s0[j] += g[u].v[j]; s1[j] += g[u].v[j] * (xv[u].v[j] - mu[j]); }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sm[rl * 2 * C + c0 + j] = s0[j];
    sm[rl * 2 * C + C + c0 + j] = s1[j];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 2 * C; idx += blockDim.x) {
    float acc = 0.f;
    for (int q = 0; q < rows; ++q) acc += sm[q * 2 * C + idx];
    atomicAdd(fin.gsum + idx, (double)acc);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(fin.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double a = __ldcg(fin.gsum + c), b = __ldcg(fin.gsum + C + c);
    if (MODE == 0) {
      const double dm = a / (double)m;
      const double mu0 = (double)x[c] + dm;            // pivot + mean of the deviations
      double v = b / (double)m - dm * dm;              // biased variance
      v = v > 0.0 ? v : 0.0;
      fin.out0[c] = (float)mu0;
      fin.out1[c] = (float)v;
      if (fin.run_mean) {  // nn.BatchNorm1d: running_var tracks the UNBIASED variance
        const double unb = m > 1 ? v * (double)m / (double)(m - 1) : v;
        fin.run_mean[c] = (float)((1.0 - fin.momentum) * fin.run_mean[c] + fin.momentum * mu0);
        fin.run_var[c] = (float)((1.0 - fin.momentum) * fin.run_var[c] + fin.momentum * unb);
      }
    } else {
      fin.raw[c] = (float)a;
      fin.raw[C + c] = (float)b;
      const float db = (float)a, dg = (float)(b / sqrt((double)fin.var[c] + (double)fin.eps));
      if (fin.out0) fin.out0[c] = fin.accumulate ? fin.out0[c] + db : db;    // dbeta
      if (fin.out1) fin.out1[c] = fin.accumulate ? fin.out1[c] + dg : dg;    // dgamma
    }
  }
  if (MODE == 0 && fin.tracked && threadIdx.x == 0) *fin.tracked += 1;
}

template <int VEC>
__global__ void __launch_bounds__(1024)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                const float* __restrict__ var, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ res, long long m, int C,
                int lanes, int rows, float eps, int relu, float* __restrict__ y,
                __nv_bfloat16* __restrict__ yb) {
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const int c0 = lane * VEC;
  float sc[VEC], sh[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    bn_scale_shift(gamma ? gamma[c0 + j] : 1.f, beta ? beta[c0 + j] : 0.f, mean[c0 + j], var[c0 + j], eps, sc[j], sh[j]);
  }
  constexpr int U = 4;  // row groups per iteration, loads first (see col_reduce_kernel)
  const long long stride = (long long)gridDim.x * rows;
  for (long long r = (long long)blockIdx.x * rows + rl; r < m; r += U * stride) {
    Vec<VEC> xv[U], rv[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * stride;
      ok[u] = ru < m;
      if (ok[u]) {
        xv[u] = Vec<VEC>::load(x + ru * C + c0);
        if (res) rv[u] = Vec<VEC>::load(res + ru * C + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const long long off = (r + u * stride) * C + c0;
      Vec<VEC> o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) o.v[j] = fmaf(xv[u].v[j], sc[j], sh[j]);
      if (res) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o.v[j] += rv[u].v[j];
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) o.v[j] = fmaxf(o.v[j], 0.f);
      }
      if (y) o.store(y + off);   // (null: only the bf16 operand copy is wanted, see ops.py "hollow" rows)
      if (yb) store_bf16<VEC>(yb + off, o.v);
    }
  }
}

// dx = scale * (dy' - sum_dy/m - (x-mean) * invstd^2 * sum_dy_xmu/m), dres = dy'
template <int VEC>
__global__ void __launch_bounds__(1024)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y,
                    const __nv_bfloat16* __restrict__ yb, const float* __restrict__ dy, long long dy_pitch,
                    const float* __restrict__ mean,
                    const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ sums /* [2C]: sum dy', sum dy'(x-mean) */,
                    long long m, int C, int lanes, int rows, float eps, int relu, int training,
                    float* __restrict__ dx, float* __restrict__ dres, __nv_bfloat16* __restrict__ dxb) {
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const int c0 = lane * VEC;
  float sc[VEC], mu[VEC], k0[VEC], k1[VEC], fsc[VEC], fsh[VEC];
  const float inv_m = 1.f / (float)m;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    float istd = rsqrtf(var[c0 + j] + eps);
    sc[j] = (gamma ? gamma[c0 + j] : 1.f) * istd;
    mu[j] = mean[c0 + j];
    fsc[j] = 0.f; fsh[j] = 0.f;
    if (relu == 2)
      bn_scale_shift(gamma ? gamma[c0 + j] : 1.f, beta ? beta[c0 + j] : 0.f, mu[j], var[c0 + j], eps, fsc[j], fsh[j]);
    k0[j] = training ? sums[c0 + j] * inv_m : 0.f;
    k1[j] = training ? sums[C + c0 + j] * inv_m * istd * istd : 0.f;
  }
  constexpr int U = 2;  // row groups per iteration, loads first (see col_reduce_kernel)
  const long long stride = (long long)gridDim.x * rows;
  for (long long r = (long long)blockIdx.x * rows + rl; r < m; r += U * stride) {
    Vec<VEC> g[U], yv[U], xv[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * stride;
      ok[u] = ru < m;
      if (ok[u]) {
        const long long off = ru * C + c0;
        g[u] = Vec<VEC>::load(dy + ru * dy_pitch + c0);
        if (relu == 1) yv[u] = load_mask_rows<VEC>(y, yb, off);
        xv[u] = Vec<VEC>::load(x + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const long long off = (r + u * stride) * C + c0;
      if (relu == 1) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) if (!(yv[u].v[j] > 0.f)) g[u].v[j] = 0.f;
      } else if (relu == 2) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) if (!(fmaf(xv[u].v[j], fsc[j], fsh[j]) > 0.f)) g[u].v[j] = 0.f;
      }
      if (dres) g[u].store(dres + off);
      Vec<VEC> o;
#pragma unroll
      for (int j = 0; j < VEC; ++j) o.v[j] = sc[j] * (g[u].v[j] - k0[j] - (xv[u].v[j] - mu[j]) * k1[j]);
      if (dx) o.store(dx + off);   // (null: the only consumer is a bf16 convolution reading dxb)
      if (dxb) store_bf16<VEC>(dxb + off, o.v);
    }
  }
}

// ---------------------------------------------------------------------------
// flat elementwise
// ---------------------------------------------------------------------------
// OP 0: y = max(a,0)   OP 1: y = b * (a > 0)   OP 2: y = a + b
template <int OP>
__global__ void __launch_bounds__(256)
flat_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
            float* __restrict__ y, int vec_ok) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long n4 = vec_ok ? n / 4 : 0;
  for (long long q = i; q < n4; q += stride) {
    float4 av = reinterpret_cast<const float4*>(a)[q], o;
    if (OP == 0) {
      o = make_float4(fmaxf(av.x, 0.f), fmaxf(av.y, 0.f), fmaxf(av.z, 0.f), fmaxf(av.w, 0.f));
    } else {
      float4 bv = reinterpret_cast<const float4*>(b)[q];
      if (OP == 1) o = make_float4(av.x > 0.f ? bv.x : 0.f, av.y > 0.f ? bv.y : 0.f,
                                   av.z > 0.f ? bv.z : 0.f, av.w > 0.f ? bv.w : 0.f);
      else o = make_float4(av.x + bv.x, av.y + bv.y, av.z + bv.z, av.w + bv.w);
    }
    reinterpret_cast<float4*>(y)[q] = o;
  }
  for (long long q = n4 * 4 + i; q < n; q += stride) {
    float av = a[q];
    if (OP == 0) y[q] = fmaxf(av, 0.f);
    else if (OP == 1) y[q] = av > 0.f ? b[q] : 0.f;
    else y[q] = av + b[q];
  }
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
           float lr, float mom, float wd, float gscale, int first) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = p[i];
    float d = fmaf(wd, w, g[i] * gscale);
    float b = first ? d : fmaf(mom, buf[i], d);
    buf[i] = b;
    p[i] = w - lr * b;
  }
}

// ---------------------------------------------------------------------------
// pooling
// ---------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(1024)
pool_fwd_kernel(const float* __restrict__ in, const int* __restrict__ nbr, long long m_out, int C,
                int K, int lanes, int rows, int avg, float* __restrict__ out) {
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const int c0 = lane * VEC;
  for (long long o = (long long)blockIdx.x * rows + rl; o < m_out; o += (long long)gridDim.x * rows) {
    Vec<VEC> acc;
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc.v[j] = 0.f;
    int cnt = 0;
    for (int k = 0; k < K; ++k) {
      int i = nbr[(long long)k * m_out + o];
      if (i >= 0) {
        Vec<VEC> v = Vec<VEC>::load(in + (long long)i * C + c0);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc.v[j] += v.v[j];
        ++cnt;
      }
    }
    if (avg && cnt > 1) {
      float s = 1.f / (float)cnt;
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc.v[j] *= s;
    }
    acc.store(out + o * C + c0);
  }
}

// global pooling: rows of one batch index are (almost always) contiguous, so each thread
// keeps a running sum and flushes with one atomic per (batch change, channel).
__global__ void __launch_bounds__(1024)
global_pool_fwd_kernel(const float* __restrict__ in, const int4* __restrict__ coords, long long m,
                       int C, int n_batch, int lanes, int rows, int rows_per_block,
                       float* __restrict__ out, int* __restrict__ cnt) {
  const int c = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block < m ? r0 + rows_per_block : m;
  float acc = 0.f;
  int cur = -1, n = 0;
  for (long long r = r0 + rl; r < r1; r += rows) {
    int b = coords[r].x;
    if (b != cur) {
      if (cur >= 0 && cur < n_batch) {
        atomicAdd(out + (long long)cur * C + c, acc);
        if (c == 0) atomicAdd(cnt + cur, n);
      }
      cur = b; acc = 0.f; n = 0;
    }
    acc += in[r * C + c];
    ++n;
  }
  if (cur >= 0 && cur < n_batch) {
    atomicAdd(out + (long long)cur * C + c, acc);
    if (c == 0) atomicAdd(cnt + cur, n);
  }
}

__global__ void global_pool_norm_kernel(float* __restrict__ out, const int* __restrict__ cnt,
                                        int n_batch, int C) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_batch * C) return;
  int n = cnt[e / C];
  if (n > 1) out[e] = out[e] / (float)n;
}

__global__ void __launch_bounds__(256)
global_pool_bwd_kernel(const float* __restrict__ dout, const int4* __restrict__ coords,
                       const int* __restrict__ cnt, long long total, int C, int n_batch, int avg,
                       float* __restrict__ din) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long r = e / C;
  int c = (int)(e - r * C);
  int b = coords[r].x;
  float v = 0.f;
  if (b >= 0 && b < n_batch) {
    v = dout[(long long)b * C + c];
    if (avg) v = v / (float)cnt[b];
  }
  din[e] = v;
}

// ---------------------------------------------------------------------------
// softmax cross-entropy over rows [n, C] with ignore_index (segmentation_training.py:27-44,
// classification_training.py:33): one thread per row, rows stay in registers (C <= 64).
// Writes stats[0] += sum_i -log p_i[y_i], stats[1] += #non-ignored rows, and the raw gradient
// graw[i, c] = p_i[c] - [c == y_i] (zero for ignored rows); backward scales by gout / count.
// ---------------------------------------------------------------------------
template <int CMAX>
__global__ void __launch_bounds__(256)
ce_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target, long long n,
              int C, long long ignore_index, float* __restrict__ graw, double* __restrict__ stats,
              int* __restrict__ bad_target) {
  float loss = 0.f;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long y = target[i];
    float* g = graw + i * C;
    if (y == ignore_index || y < 0 || y >= C) {
      // ignored rows and rows with an out-of-range label (flagged: the host raises) contribute no gradient;
      // graw comes from an uninitialised allocation, so the row must be written either way
      if (y != ignore_index) *bad_target = 1;
      for (int c = 0; c < C; ++c) g[c] = 0.f;
      continue;
    }
    float v[CMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) { v[c] = logits[i * C + c]; mx = fmaxf(mx, v[c]); }
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
      if (c < C) g[c] = v[c] * inv - (c == (int)y ? 1.f : 0.f);
    loss += (mx - vy) + __logf(sum);
    ++cnt;
  }
  // block reduction -> one double atomic per block
  __shared__ float s_loss[8];
  __shared__ int s_cnt[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if ((threadIdx.x & 31) == 0) { s_loss[threadIdx.x >> 5] = loss; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    long long k = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { l += (double)s_loss[w]; k += s_cnt[w]; }
    atomicAdd(stats, l);
    atomicAdd(stats + 1, (double)k);
  }
}

// dlogits = graw * (gout / count)
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ graw, const double* __restrict__ stats,
              const float* __restrict__ gout, long long total, float* __restrict__ dlogits) {
  const double cnt = stats[1];
  const float sc = cnt > 0.0 ? (float)((double)gout[0] / cnt) : 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x)
    dlogits[e] = graw[e] * sc;
}

// mean / biased variance (+ running statistics) from per-channel sums gsum[0..C) = sum x, gsum[C..2C) = sum x^2
// accumulated in double by the convolution epilogue (spc_conv_fwd_stats)
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const double* __restrict__ gsum, long long m, int C, float* __restrict__ mean,
                   float* __restrict__ var, float* run_mean, float* run_var, float momentum, long long* tracked) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && tracked) *tracked += 1;
  if (c >= C) return;
  const double mu = gsum[c] / (double)m;
  double v = gsum[C + c] / (double)m - mu * mu;
  v = v > 0.0 ? v : 0.0;
  mean[c] = (float)mu;
  var[c] = (float)v;
  if (run_mean) {  // nn.BatchNorm1d: running_var tracks the UNBIASED variance
    const double unb = m > 1 ? v * (double)m / (double)(m - 1) : v;
    run_mean[c] = (float)((1.0 - momentum) * run_mean[c] + momentum * mu);
    run_var[c] = (float)((1.0 - momentum) * run_var[c] + momentum * unb);
  }
}

}  // namespace spc

using namespace spc;

// rows of `row_bytes` bytes from a pitched source to a pitched destination (channel concatenation of operand
// copies: dst = a column slice of the wider row).  T = the widest type every pitch / pointer / width is a multiple of.
template <typename T>
__global__ void __launch_bounds__(256)
copy_rows_kernel(const char* __restrict__ src, long long src_pitch, char* __restrict__ dst, long long dst_pitch,
                 int per_row, long long rows) {
  const long long total = rows * per_row;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / per_row;
    const int j = (int)(e - r * per_row);
    reinterpret_cast<T*>(dst + r * dst_pitch)[j] = reinterpret_cast<const T*>(src + r * src_pitch)[j];
  }
}

#define DISPATCH_VEC(vec, ...)        \
  if ((vec) == 4) { constexpr int VEC = 4; __VA_ARGS__; } \
  else { constexpr int VEC = 1; __VA_ARGS__; }

extern "C" {

int64_t spc_bn_workspace(int64_t m, int C) {
  (void)m;
  return (int64_t)2 * C * 8 + 2 * C * 4 + 256;  // double sums [2C], counter, float sums [2C]
}

struct BnWs {
  double* gsum;
  unsigned* counter;
  float* raw;
};
static BnWs bn_ws(void* workspace, int C) {
  BnWs w;
  char* base = (char*)(((uintptr_t)workspace + 15) & ~(uintptr_t)15);
  w.gsum = (double*)base;
  w.counter = (unsigned*)(base + (size_t)2 * C * 8);
  w.raw = (float*)(base + (size_t)2 * C * 8 + 64);
  return w;
}

static int col_reduce_launch(int mode, const float* x, const float* y, const void* y_bf16, const float* dy,
                             int64_t dy_pitch, const float* mean, const float* gamma, const float* beta, int64_t m, int C,
                             int relu, ColFinal fin, cudaStream_t stream) {
  int vec = pick_vec(C, x, y, dy);
  if (y_bf16 && ((uintptr_t)y_bf16 % 8)) vec = 1;
  if (dy && dy_pitch % 4) vec = 1;
  const __nv_bfloat16* yb = (const __nv_bfloat16*)y_bf16;
  RowMap rm = make_row_map(C, vec);
  if (rm.threads > 1024) return fail("col_reduce", "C too large (max 1024 scalar / 4096 vec4)");
  int grid = pick_grid(m, rm.rows, 4);
  size_t smem = (size_t)rm.rows * 2 * C * sizeof(float);
  if (smem > 40 * 1024) return fail("col_reduce", "shared memory");
  SPC_CUDA(cudaMemsetAsync(fin.gsum, 0, (size_t)2 * C * 8 + 64, stream));  // sums + ticket counter
  DISPATCH_VEC(vec,
    if (mode == 0) col_reduce_kernel<VEC, 0><<<grid, rm.threads, smem, stream>>>(x, y, yb, dy, dy_pitch, mean, gamma, beta, m, C, rm.lanes, rm.rows, relu, fin);
    else col_reduce_kernel<VEC, 1><<<grid, rm.threads, smem, stream>>>(x, y, yb, dy, dy_pitch, mean, gamma, beta, m, C, rm.lanes, rm.rows, relu, fin));
  SPC_LAUNCHED("col_reduce_kernel");
  return 0;
}

int spc_bn_stats(const float* x, int64_t m, int C, float* mean, float* var, float* running_mean,
                 float* running_var, float momentum, void* workspace, int64_t workspace_bytes,
                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(m >= 1 && C >= 1, "empty input");
  SPC_REQUIRE(workspace_bytes >= spc_bn_workspace(m, C), "workspace too small");
  BnWs w = bn_ws(workspace, C);
  ColFinal fin;
  fin.gsum = w.gsum; fin.counter = w.counter; fin.out0 = mean; fin.out1 = var; fin.raw = nullptr;
  fin.run_mean = running_mean; fin.run_var = running_var; fin.var = nullptr;
  fin.momentum = momentum; fin.eps = 0.f; fin.accumulate = 0; fin.tracked = nullptr;
  return col_reduce_launch(0, x, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, m, C, 0, fin, stream);
}

int spc_bn_stats_tracked(const float* x, int64_t m, int C, float* mean, float* var, float* running_mean,
                         float* running_var, float momentum, int64_t* num_batches_tracked, void* workspace,
                         int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(m >= 1 && C >= 1, "empty input");
  SPC_REQUIRE(workspace_bytes >= spc_bn_workspace(m, C), "workspace too small");
  BnWs w = bn_ws(workspace, C);
  ColFinal fin;
  fin.gsum = w.gsum; fin.counter = w.counter; fin.out0 = mean; fin.out1 = var; fin.raw = nullptr;
  fin.run_mean = running_mean; fin.run_var = running_var; fin.var = nullptr;
  fin.momentum = momentum; fin.eps = 0.f; fin.accumulate = 0; fin.tracked = (long long*)num_batches_tracked;
  return col_reduce_launch(0, x, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, m, C, 0, fin, stream);
}

int spc_bn_finalize(const double* sums, int64_t m, int C, float* mean, float* var, float* running_mean,
                    float* running_var, float momentum, int64_t* num_batches_tracked, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(m >= 1 && C >= 1 && sums, "empty input");
  bn_finalize_kernel<<<(C + 255) / 256, 256, 0, stream>>>(sums, m, C, mean, var, running_mean, running_var, momentum,
                                                          (long long*)num_batches_tracked);
  SPC_LAUNCHED("bn_finalize_kernel");
  return 0;
}

int spc_bn_apply(const float* x, const float* mean, const float* var, const float* gamma,
                 const float* beta, const float* residual, int64_t m, int C, float eps, int relu,
                 float* y, void* y_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1, "bad C");
  SPC_REQUIRE(y || y_bf16, "no output (y and y_bf16 are both null)");
  if (m == 0) return 0;
  int vec = pick_vec(C, x, residual, y);
  if (y_bf16 && ((uintptr_t)y_bf16 % 8)) vec = 1;
  RowMap rm = make_row_map(C, vec);
  SPC_REQUIRE(rm.threads <= 1024, "C too large");
  int grid = pick_grid(m, rm.rows, 8);
  DISPATCH_VEC(vec, bn_apply_kernel<VEC><<<grid, rm.threads, 0, stream>>>(
      x, mean, var, gamma, beta, residual, m, C, rm.lanes, rm.rows, eps, relu, y, (__nv_bfloat16*)y_bf16));
  SPC_LAUNCHED("bn_apply_kernel");
  return 0;
}

int spc_bn_bwd(const float* x, const float* y, const void* y_bf16, const float* dy, int64_t dy_pitch, const float* mean,
               const float* var, const float* gamma, int64_t m, int C, float eps, int relu,
               int training, float* dx, void* dx_bf16, float* dresidual, float* dgamma, float* dbeta,
               void* workspace, int64_t workspace_bytes, void* stream_) {
  return spc_bn_bwd_acc(x, y, y_bf16, dy, dy_pitch, mean, var, gamma, nullptr, m, C, eps, relu, training, dx, dx_bf16,
                        dresidual, dgamma, dbeta, 0, workspace, workspace_bytes, stream_);
}

int spc_bn_bwd_acc(const float* x, const float* y, const void* y_bf16, const float* dy, int64_t dy_pitch,
                   const float* mean, const float* var, const float* gamma, const float* beta, int64_t m, int C,
                   float eps, int relu,
                   int training, float* dx, void* dx_bf16, float* dresidual, float* dgamma, float* dbeta,
                   int accumulate_param_grads, void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(m >= 1 && C >= 1, "empty input");
  SPC_REQUIRE(workspace_bytes >= spc_bn_workspace(m, C), "workspace too small");
  SPC_REQUIRE(relu >= 0 && relu <= 2, "relu: 0 none, 1 mask from y, 2 mask recomputed from x");
  SPC_REQUIRE(relu != 1 || y || y_bf16, "relu backward (mode 1) needs y (fp32 rows or their bf16 copy)");
  SPC_REQUIRE(dx || dx_bf16, "no output (dx and dx_bf16 are both null)");
  SPC_REQUIRE(dy_pitch >= C, "dy pitch smaller than C");
  BnWs w = bn_ws(workspace, C);
  float* sums = w.raw;
  ColFinal fin;
  fin.gsum = w.gsum; fin.counter = w.counter; fin.out0 = dbeta; fin.out1 = dgamma; fin.raw = sums;
  fin.run_mean = nullptr; fin.run_var = nullptr; fin.var = var; fin.momentum = 0.f; fin.eps = eps;
  fin.accumulate = accumulate_param_grads; fin.tracked = nullptr;
  if (relu != 1) { y = nullptr; y_bf16 = nullptr; }
  int rc = col_reduce_launch(1, x, y, y_bf16, dy, dy_pitch, mean, gamma, beta, m, C, relu, fin, stream);
  if (rc) return rc;
  int vec = pick_vec(C, x, y, dy, dx);
  if (dresidual && ((uintptr_t)dresidual % 16)) vec = 1;
  if (dx_bf16 && ((uintptr_t)dx_bf16 % 8)) vec = 1;
  if (y_bf16 && ((uintptr_t)y_bf16 % 8)) vec = 1;
  if (dy_pitch % 4) vec = 1;
  RowMap rm = make_row_map(C, vec);
  SPC_REQUIRE(rm.threads <= 1024, "C too large");
  int grid = pick_grid(m, rm.rows, 8);
  DISPATCH_VEC(vec, bn_bwd_apply_kernel<VEC><<<grid, rm.threads, 0, stream>>>(
      x, y, (const __nv_bfloat16*)y_bf16, dy, dy_pitch, mean, var, gamma, beta, sums, m, C, rm.lanes, rm.rows, eps, relu, training, dx,
      dresidual, (__nv_bfloat16*)dx_bf16));
  SPC_LAUNCHED("bn_bwd_apply_kernel");
  return 0;
}

int spc_copy_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t row_bytes, int64_t rows,
                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(row_bytes >= 0 && rows >= 0 && src_pitch >= row_bytes && dst_pitch >= row_bytes, "bad shape");
  if (rows == 0 || row_bytes == 0) return 0;
  const uintptr_t all = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)src_pitch | (uintptr_t)dst_pitch | (uintptr_t)row_bytes;
  const int w = (all % 16 == 0) ? 16 : (all % 4 == 0 ? 4 : (all % 2 == 0 ? 2 : 1));
  const int per_row = (int)(row_bytes / w);
  const int64_t want = ceil_div(rows * per_row, 256);
  const int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  const char* s = (const char*)src;
  char* d = (char*)dst;
  if (w == 16) copy_rows_kernel<uint4><<<grid, 256, 0, stream>>>(s, src_pitch, d, dst_pitch, per_row, rows);
  else if (w == 4) copy_rows_kernel<uint32_t><<<grid, 256, 0, stream>>>(s, src_pitch, d, dst_pitch, per_row, rows);
  else if (w == 2) copy_rows_kernel<uint16_t><<<grid, 256, 0, stream>>>(s, src_pitch, d, dst_pitch, per_row, rows);
  else copy_rows_kernel<uint8_t><<<grid, 256, 0, stream>>>(s, src_pitch, d, dst_pitch, per_row, rows);
  SPC_LAUNCHED("copy_rows_kernel");
  return 0;
}

static int flat_launch(int op, const float* a, const float* b, int64_t n, float* y, cudaStream_t stream) {
  if (n == 0) return 0;
  int vec_ok = ((uintptr_t)a % 16 == 0) && ((uintptr_t)y % 16 == 0) && (!b || (uintptr_t)b % 16 == 0);
  int64_t want = ceil_div(ceil_div(n, 4), 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  if (op == 0) flat_kernel<0><<<grid, 256, 0, stream>>>(a, b, n, y, vec_ok);
  else if (op == 1) flat_kernel<1><<<grid, 256, 0, stream>>>(a, b, n, y, vec_ok);
  else flat_kernel<2><<<grid, 256, 0, stream>>>(a, b, n, y, vec_ok);
  SPC_LAUNCHED("flat_kernel");
  return 0;
}
int spc_relu_fwd(const float* x, int64_t n, float* y, void* stream) { return flat_launch(0, x, nullptr, n, y, (cudaStream_t)stream); }
int spc_relu_bwd(const float* y, const float* dy, int64_t n, float* dx, void* stream) { return flat_launch(1, y, dy, n, dx, (cudaStream_t)stream); }
int spc_add(const float* a, const float* b, int64_t n, float* y, void* stream) { return flat_launch(2, a, b, n, y, (cudaStream_t)stream); }

int spc_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr,
                 float momentum, float weight_decay, float grad_scale, int first_step, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return 0;
  int64_t want = ceil_div(n, 256);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  sgd_kernel<<<grid, 256, 0, stream>>>(param, grad, momentum_buf, n, lr, momentum, weight_decay, grad_scale, first_step);
  SPC_LAUNCHED("sgd_kernel");
  return 0;
}

int spc_pool_fwd(const float* in, const int32_t* nbr, int64_t m_out, int C, int K, int avg,
                 float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && K >= 1, "bad shape");
  if (m_out == 0) return 0;
  int vec = pick_vec(C, in, out);
  RowMap rm = make_row_map(C, vec);
  SPC_REQUIRE(rm.threads <= 1024, "C too large");
  int grid = pick_grid(m_out, rm.rows, 8);
  DISPATCH_VEC(vec, pool_fwd_kernel<VEC><<<grid, rm.threads, 0, stream>>>(in, nbr, m_out, C, K, rm.lanes, rm.rows, avg, out));
  SPC_LAUNCHED("pool_fwd_kernel");
  return 0;
}

int spc_global_pool_fwd(const float* in, const int32_t* coords, int64_t m, int C, int n_batch,
                        int avg, float* out, int32_t* cnt, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && C <= 1024 && n_batch >= 1, "bad shape");
  SPC_CUDA(cudaMemsetAsync(out, 0, (size_t)n_batch * C * sizeof(float), stream));
  SPC_CUDA(cudaMemsetAsync(cnt, 0, (size_t)n_batch * sizeof(int), stream));
  if (m == 0) return 0;
  RowMap rm = make_row_map(C, 1);
  const int rows_per_block = rm.rows * 64;
  int grid = (int)ceil_div(m, rows_per_block);
  global_pool_fwd_kernel<<<grid, rm.threads, 0, stream>>>(in, (const int4*)coords, m, C, n_batch, rm.lanes, rm.rows, rows_per_block, out, cnt);
  SPC_LAUNCHED("global_pool_fwd_kernel");
  if (avg) {
    global_pool_norm_kernel<<<(int)ceil_div((int64_t)n_batch * C, 256), 256, 0, stream>>>(out, cnt, n_batch, C);
    SPC_LAUNCHED("global_pool_norm_kernel");
  }
  return 0;
}

int spc_global_pool_bwd(const float* dout, const int32_t* coords, const int32_t* cnt, int64_t m,
                        int C, int n_batch, int avg, float* din, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (m == 0) return 0;
  long long total = (long long)m * C;
  global_pool_bwd_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(dout, (const int4*)coords, cnt, total, C, n_batch, avg, din);
  SPC_LAUNCHED("global_pool_bwd_kernel");
  return 0;
}

int spc_ce_fwd(const float* logits, const int64_t* target, int64_t n, int C, int64_t ignore_index,
               float* grad_raw, double* stats, int32_t* bad_target, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SPC_REQUIRE(C >= 1 && C <= 64, "cross entropy supports 1..64 classes");
  SPC_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(double), stream));
  SPC_CUDA(cudaMemsetAsync(bad_target, 0, sizeof(int32_t), stream));
  if (n == 0) return 0;
  int64_t want = ceil_div(n, 256);
  int grid = (int)(want < kNumSMs * 8 ? want : kNumSMs * 8);
  if (C <= 32)
    ce_fwd_kernel<32><<<grid, 256, 0, stream>>>(logits, (const long long*)target, n, C, ignore_index, grad_raw, stats, bad_target);
  else
    ce_fwd_kernel<64><<<grid, 256, 0, stream>>>(logits, (const long long*)target, n, C, ignore_index, grad_raw, stats, bad_target);
  SPC_LAUNCHED("ce_fwd_kernel");
  return 0;
}

int spc_ce_bwd(const float* grad_raw, const double* stats, const float* grad_out, int64_t n, int C,
               float* dlogits, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n == 0) return 0;
  long long total = (long long)n * C;
  int64_t want = ceil_div(total, 256 * 4);
  int grid = (int)(want < kNumSMs * 16 ? want : kNumSMs * 16);
  ce_bwd_kernel<<<grid, 256, 0, stream>>>(grad_raw, stats, grad_out, total, dlogits);
  SPC_LAUNCHED("ce_bwd_kernel");
  return 0;
}

}  // extern "C"

