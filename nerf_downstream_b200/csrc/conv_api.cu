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
int conv_fwd_umma(const float* in, const float* w, const float* bias, const int* nbr,
                  const uint32_t* tile_mask, int64_t m_out, int c_in, int c_out, int K, bool transpose_w, float* out,
                  void* workspace, int64_t workspace_bytes, cudaStream_t stream);
bool umma_wgrad_supported(int c_in, int c_out);
int64_t umma_wgrad_workspace(int K, int c_in, int c_out);
int conv_wgrad_umma(const float* in, const float* dout, const int* nbr, const uint32_t* tile_mask,
                    int64_t m_out, int c_in, int c_out, int K, float* dw, void* workspace,
                    int64_t workspace_bytes, cudaStream_t stream);
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
  if (precision != SPC_PREC_TF32) return 256;
  int64_t a = umma_fwd_supported(c_in, c_out) ? umma_fwd_workspace(K, c_in, c_out) : 0;
  int64_t b = umma_fwd_supported(c_out, c_in) ? umma_fwd_workspace(K, c_out, c_in) : 0;
  int64_t c = umma_wgrad_supported(c_in, c_out) ? umma_wgrad_workspace(K, c_in, c_out) : 0;
  int64_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

int spc_conv_fwd(const float* in, const float* w, const float* bias, const int32_t* nbr,
                 const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision,
                 float* out, void* workspace, int64_t workspace_bytes, void* stream) {
  (void)m_in;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(precision == SPC_PREC_FP32 || precision == SPC_PREC_TF32, "bad precision mode");
  if (precision == SPC_PREC_TF32 && K <= 32 && umma_fwd_supported(c_in, c_out))
    return conv_fwd_umma(in, w, bias, nbr, tile_mask, m_out, c_in, c_out, K, false, out, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  return conv_fwd_simt(in, w, bias, nbr, m_out, c_in, c_out, K, false, out, (cudaStream_t)stream);
}

int spc_conv_dgrad(const float* dout, const float* w, const int32_t* nbr_t,
                   const uint32_t* tile_mask_t, int64_t m_in,
                   int64_t m_out, int c_in, int c_out, int K, int precision, float* din,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  (void)m_out;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(precision == SPC_PREC_FP32 || precision == SPC_PREC_TF32, "bad precision mode");
  // din[M_in, Cin] = sum_k gather(dout, nbr_t[k]) [M_in, Cout] x W[k]^T [Cout, Cin]
  if (precision == SPC_PREC_TF32 && K <= 32 && umma_fwd_supported(c_out, c_in))
    return conv_fwd_umma(dout, w, nullptr, nbr_t, tile_mask_t, m_in, c_out, c_in, K, true, din, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  return conv_fwd_simt(dout, w, nullptr, nbr_t, m_in, c_out, c_in, K, true, din, (cudaStream_t)stream);
}

int spc_conv_wgrad(const float* in, const float* dout, const int32_t* nbr,
                   const uint32_t* tile_mask, int64_t m_in,
                   int64_t m_out, int c_in, int c_out, int K, int precision, float* dw,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  (void)m_in;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(precision == SPC_PREC_FP32 || precision == SPC_PREC_TF32, "bad precision mode");
  if (precision == SPC_PREC_TF32 && K <= 32 && (int64_t)K * c_in <= 128 * 128 && umma_wgrad_supported(c_in, c_out))
    return conv_wgrad_umma(in, dout, nbr, tile_mask, m_out, c_in, c_out, K, dw, workspace, workspace_bytes,
                           (cudaStream_t)stream);
  return conv_wgrad_simt(in, dout, nbr, m_out, c_in, c_out, K, dw, (cudaStream_t)stream);
}

}  // extern "C"
