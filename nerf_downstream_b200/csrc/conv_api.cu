// conv_api.cu — C-ABI entry points of the sparse convolution; dispatches between the
// tcgen05 TF32 kernels (conv_umma.cu) and the FP32 CUDA-core kernels (conv_simt.cu).
// Reference call sites: MinkowskiConvolution / MinkowskiConvolutionTranspose constructed
// at co3d_3d/src/models/mink/modules/common.py:117-125,172-180; forward semantics
// restated at modules/sparse_conv.py:122-143.
#include "common.cuh"

namespace spc {
int conv_fwd_simt(const float* in, const float* w, const float* bias, const int* nbr,
                  int64_t m_out, int c_in, int c_out, int K, bool transpose_w, float* out,
                  cudaStream_t stream);
int conv_wgrad_simt(const float* in, const float* dout, const int* nbr, int64_t m_out, int c_in,
                    int c_out, int K, float* dw, cudaStream_t stream);
// conv_umma.cu
bool umma_fwd_supported(int c_in, int c_out);
int64_t umma_fwd_workspace(int K, int c_in, int c_out);
int conv_fwd_umma(const void* in, const float* w, const float* bias, const int* nbr,
                  const uint32_t* tile_mask, int64_t m_out, int c_in, int c_out, int K,
                  bool transpose_w, bool bf16, float* out, double* stats, int* stats_fused, void* workspace,
                  int64_t workspace_bytes, cudaStream_t stream);
bool umma_wgrad_supported(int c_in, int c_out);
int64_t umma_wgrad_workspace(int K, int c_in, int c_out);
int conv_wgrad_umma(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask,
                    int64_t m_out, int c_in, int c_out, int K, bool bf16, float* dw, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream);
int to_bf16(const float* src, int64_t rows, int c_src, int64_t src_pitch, int c_dst, void* dst, cudaStream_t stream);
void umma_set_force_mt(int mt);
void umma_debug_set(int idx, int val);
int umma_debug_read(long long* host, int n);
}  // namespace spc

using namespace spc;

extern "C" {

/* test hook: force the number of 128-row sub-tiles per CTA of the tcgen05 kernel (0 = auto) */
void spc_debug_force_mt(int mt) { umma_set_force_mt(mt); }
/* test hook: operand-layout knobs of the tcgen05 wgrad kernel (0 = built-in default) */
void spc_debug_set(int idx, int val) { umma_debug_set(idx, val); }
/* test hook: cycle counters of the last tcgen05 wgrad launch (8 x int64 per CTA) into a HOST buffer */
int spc_debug_read(long long* host, int n) { return umma_debug_read(host, n); }

int64_t spc_conv_workspace(int K, int c_in, int c_out, int precision) {
  if (precision == SPC_PREC_FP32) return 256;
  int64_t a = umma_fwd_supported(c_in, c_out) ? umma_fwd_workspace(K, c_in, c_out) : 0;
  int64_t b = umma_fwd_supported(c_out, c_in) ? umma_fwd_workspace(K, c_out, c_in) : 0;
  int64_t c = umma_wgrad_supported(c_in, c_out) ? umma_wgrad_workspace(K, c_in, c_out) : 0;
  int64_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

static bool prec_ok(int p) { return p == SPC_PREC_FP32 || p == SPC_PREC_TF32 || p == SPC_PREC_BF16; }

int spc_to_bf16(const float* src, int64_t rows, int c_src, int64_t src_pitch, int c_dst, void* dst, void* stream) {
  SPC_REQUIRE(rows >= 0 && c_src >= 0 && c_dst >= c_src && src_pitch >= c_src, "bad shape");
  return to_bf16(src, rows, c_src, src_pitch, c_dst, dst, (cudaStream_t)stream);
}

int spc_conv_fwd(const void* in, const float* w, const float* bias, const int32_t* nbr,
                 const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision,
                 float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  return spc_conv_fwd_stats(in, w, bias, nbr, tile_mask, m_in, m_out, c_in, c_out, K, precision, out, nullptr, nullptr,
                            workspace, workspace_bytes, stream);
}

int spc_conv_fwd_stats(const void* in, const float* w, const float* bias, const int32_t* nbr,
                       const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                       int precision, float* out, double* bn_stats, int32_t* stats_fused, void* workspace,
                       int64_t workspace_bytes, void* stream) {
  if (stats_fused) *stats_fused = 0;
  (void)m_in;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(prec_ok(precision), "bad precision mode");
  const bool umma = K <= 32 && umma_fwd_supported(c_in, c_out);
  if (precision == SPC_PREC_BF16) {
    SPC_REQUIRE(umma, "bf16 mode needs a tensor-core shape (Cin % 32 == 0, Cout % 16 == 0, K <= 32)");
    return conv_fwd_umma(in, w, bias, nbr, tile_mask, m_out, c_in, c_out, K, false, true, out, bn_stats, stats_fused, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  }
  if (precision == SPC_PREC_TF32 && umma)
    return conv_fwd_umma(in, w, bias, nbr, tile_mask, m_out, c_in, c_out, K, false, false, out, bn_stats, stats_fused, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  return conv_fwd_simt((const float*)in, w, bias, nbr, m_out, c_in, c_out, K, false, out, (cudaStream_t)stream);
}

int spc_conv_dgrad(const void* dout, const float* w, const int32_t* nbr_t,
                   const uint32_t* tile_mask_t, int64_t m_in,
                   int64_t m_out, int c_in, int c_out, int K, int precision, float* din,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  (void)m_out;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(prec_ok(precision), "bad precision mode");
  // din[M_in, Cin] = sum_k gather(dout, nbr_t[k]) [M_in, Cout] x W[k]^T [Cout, Cin]
  const bool umma = K <= 32 && umma_fwd_supported(c_out, c_in);
  if (precision == SPC_PREC_BF16) {
    SPC_REQUIRE(umma, "bf16 mode needs a tensor-core shape (Cout % 32 == 0, Cin % 16 == 0, K <= 32)");
    return conv_fwd_umma(dout, w, nullptr, nbr_t, tile_mask_t, m_in, c_out, c_in, K, true, true, din, nullptr, nullptr, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  }
  if (precision == SPC_PREC_TF32 && umma)
    return conv_fwd_umma(dout, w, nullptr, nbr_t, tile_mask_t, m_in, c_out, c_in, K, true, false, din, nullptr, nullptr, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  return conv_fwd_simt((const float*)dout, w, nullptr, nbr_t, m_in, c_out, c_in, K, true, din, (cudaStream_t)stream);
}

int spc_conv_wgrad(const void* in, const void* dout, const int32_t* nbr,
                   const uint32_t* tile_mask, int64_t m_in,
                   int64_t m_out, int c_in, int c_out, int K, int precision, float* dw,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  (void)m_in;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(prec_ok(precision), "bad precision mode");
  const bool umma = K <= 32 && (int64_t)K * c_in <= 128 * 128 && umma_wgrad_supported(c_in, c_out);
  if (precision == SPC_PREC_BF16) {
    SPC_REQUIRE(umma, "bf16 mode needs a tensor-core shape (Cin, Cout % 32 == 0, Cout <= 256, K <= 32)");
    return conv_wgrad_umma(in, dout, nbr, tile_mask, m_out, c_in, c_out, K, true, dw, workspace, workspace_bytes,
                           (cudaStream_t)stream);
  }
  if (precision == SPC_PREC_TF32 && umma)
    return conv_wgrad_umma(in, dout, nbr, tile_mask, m_out, c_in, c_out, K, false, dw, workspace, workspace_bytes,
                           (cudaStream_t)stream);
  return conv_wgrad_simt((const float*)in, (const float*)dout, nbr, m_out, c_in, c_out, K, dw, (cudaStream_t)stream);
}

}  // extern "C"
