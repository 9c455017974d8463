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
int64_t umma_packed_bytes(int K, int c_in, int c_out);
int conv_pack_weights(const float* w, int K, int Ck, int Cn, int flags, bool bf16, void* packed,
                      cudaStream_t stream);
int conv_pack_weights_batch(const long long* desc_dev, int n_layers, cudaStream_t stream);
int conv_fwd_umma(const void* in, const float* w, const void* packed, const float* bias, const int* nbr,
                  const uint32_t* tile_mask, int64_t m_out, int c_in, int c_out, int K,
                  bool transpose_w, bool bf16, float* out, void* workspace,
                  int64_t workspace_bytes, cudaStream_t stream, bool accumulate = false, double* stats = nullptr,
                  int* stats_fused = nullptr);
// conv_wgrad_umma.cu
bool umma_wgrad_supported(int c_in, int c_out);
int64_t umma_wgrad_workspace(int K, int c_in, int c_out);
int conv_wgrad_umma(const void* in, const void* dout, const int* nbr, const uint32_t* tile_mask,
                    int64_t m_out, int c_in, int c_out, int K, bool bf16, float* dw, bool zero_dw,
                    cudaStream_t stream);
extern std::atomic<long long> g_conv_path_counts[4];
int to_bf16(const float* src, int64_t rows, int c_src, int64_t src_pitch, int c_dst, void* dst, cudaStream_t stream);
void umma_set_force_mt(int mt);
void umma_debug_set(int idx, int val);
}  // namespace spc

using namespace spc;

extern "C" {

/* test hook: force the number of 128-row sub-tiles per CTA of the tcgen05 kernel (0 = auto) */
void spc_debug_force_mt(int mt) { umma_set_force_mt(mt); }
/* test hook: operand-layout knobs of the tcgen05 wgrad kernel (0 = built-in default) */
void spc_debug_set(int idx, int val) { umma_debug_set(idx, val); }
/* launches per route since the last reset: [0] tcgen05 bf16, [1] tcgen05 tf32, [2] CUDA-core fp32 */
void spc_conv_path_counts(long long* out3, int reset) {
  for (int i = 0; i < 3; ++i) {
    if (out3) out3[i] = g_conv_path_counts[i].load(std::memory_order_relaxed);
    if (reset) g_conv_path_counts[i].store(0, std::memory_order_relaxed);
  }
}

static bool tc_fwd_ok(int K, int c_in, int c_out) { return K <= 32 && umma_fwd_supported(c_in, c_out); }
static bool tc_wgrad_ok(int K, int c_in, int c_out) {
  return K <= 32 && (int64_t)K * c_in <= 128 * 128 && umma_wgrad_supported(c_in, c_out);
}

/* 1 if spc_conv_fwd (what = 0), spc_conv_dgrad (1) or spc_conv_wgrad (2) runs this shape on the tensor cores */
int spc_conv_tensor_core(int what, int K, int c_in, int c_out, int precision) {
  if (precision == SPC_PREC_FP32) return 0;
  if (what == 0) return tc_fwd_ok(K, c_in, c_out) ? 1 : 0;
  if (what == 1) return tc_fwd_ok(K, c_out, c_in) ? 1 : 0;
  return tc_wgrad_ok(K, c_in, c_out) ? 1 : 0;
}

int64_t spc_conv_packed_bytes(int K, int c_in, int c_out) { return umma_packed_bytes(K, c_in, c_out); }

int spc_conv_pack_weights(const float* w, int K, int c_in, int c_out, int dgrad, int precision, void* packed,
                          void* stream) {
  SPC_REQUIRE(precision == SPC_PREC_TF32 || precision == SPC_PREC_BF16, "packed weights are a tensor-core format");
  SPC_REQUIRE(dgrad ? tc_fwd_ok(K, c_out, c_in) : tc_fwd_ok(K, c_in, c_out), "shape not supported by the tcgen05 path");
  SPC_REQUIRE(dgrad >= 0 && dgrad <= 2, "dgrad: 0 forward image, 1 W^T, 2 W^T with the offsets reversed");
  // forward contracts over Cin (rows of W[k]); dgrad contracts over Cout with W[k]^T; dgrad == 2 additionally stores
  // W[K-1-k]^T in slab k: the dgrad of a centrally symmetric self map then reads the FORWARD map (nbr_t[k] == nbr[K-1-k])
  return dgrad ? conv_pack_weights(w, K, c_out, c_in, dgrad == 2 ? 3 : 1, precision == SPC_PREC_BF16, packed, (cudaStream_t)stream)
               : conv_pack_weights(w, K, c_in, c_out, 0, precision == SPC_PREC_BF16, packed, (cudaStream_t)stream);
}

int spc_conv_pack_weights_batch(const int64_t* desc_dev, int n_layers, void* stream) {
  return conv_pack_weights_batch((const long long*)desc_dev, n_layers, (cudaStream_t)stream);
}

int spc_conv_fwd_packed(const void* in, const void* w_packed, const float* bias, const int32_t* nbr,
                        const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                        int precision, float* out, void* stream) {
  return spc_conv_fwd_packed_stats(in, w_packed, bias, nbr, tile_mask, m_in, m_out, c_in, c_out, K, precision, out,
                                   nullptr, nullptr, stream);
}

int spc_conv_fwd_packed_stats(const void* in, const void* w_packed, const float* bias, const int32_t* nbr,
                              const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                              int precision, float* out, double* bn_sums, int32_t* stats_fused, void* stream) {
  (void)m_in;
  SPC_REQUIRE(precision == SPC_PREC_TF32 || precision == SPC_PREC_BF16, "packed weights are a tensor-core format");
  SPC_REQUIRE(w_packed && tc_fwd_ok(K, c_in, c_out), "shape not supported by the tcgen05 path");
  SPC_REQUIRE(bn_sums == nullptr || stats_fused != nullptr, "bn_sums needs stats_fused");
  int fused = 0;
  int rc = conv_fwd_umma(in, nullptr, w_packed, bias, nbr, tile_mask, m_out, c_in, c_out, K, false,
                         precision == SPC_PREC_BF16, out, nullptr, 0, (cudaStream_t)stream, false, bn_sums, &fused);
  if (stats_fused) *stats_fused = fused;
  return rc;
}

int spc_conv_dgrad_packed(const void* dout, const void* w_packed_t, const int32_t* nbr_t, const uint32_t* tile_mask_t,
                          int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision, float* din,
                          void* stream) {
  return spc_conv_dgrad_packed_acc(dout, w_packed_t, nbr_t, tile_mask_t, m_in, m_out, c_in, c_out, K, precision, din, 0,
                                   stream);
}

int spc_conv_dgrad_packed_acc(const void* dout, const void* w_packed_t, const int32_t* nbr_t,
                              const uint32_t* tile_mask_t, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                              int precision, float* din, int accumulate, void* stream) {
  (void)m_out;
  SPC_REQUIRE(precision == SPC_PREC_TF32 || precision == SPC_PREC_BF16, "packed weights are a tensor-core format");
  SPC_REQUIRE(w_packed_t && tc_fwd_ok(K, c_out, c_in), "shape not supported by the tcgen05 path");
  return conv_fwd_umma(dout, nullptr, w_packed_t, nullptr, nbr_t, tile_mask_t, m_in, c_out, c_in, K, true,
                       precision == SPC_PREC_BF16, din, nullptr, 0, (cudaStream_t)stream, accumulate != 0);
}

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
  // (the epilogue statistics of round 1 cost the forward kernel more than the BatchNorm pass they saved —
  // profiles/r1_gather4_experiment.md — and are gone: *stats_fused is always 0, bn_stats is not written)
  if (stats_fused) *stats_fused = 0;
  (void)m_in; (void)bn_stats;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(prec_ok(precision), "bad precision mode");
  const bool umma = tc_fwd_ok(K, c_in, c_out);
  if (precision == SPC_PREC_BF16) {
    SPC_REQUIRE(umma, "bf16 mode needs a tensor-core shape (Cin % 32 == 0, Cout % 16 == 0, K <= 32)");
    return conv_fwd_umma(in, w, nullptr, bias, nbr, tile_mask, m_out, c_in, c_out, K, false, true, out, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  }
  if (precision == SPC_PREC_TF32 && umma)
    return conv_fwd_umma(in, w, nullptr, bias, nbr, tile_mask, m_out, c_in, c_out, K, false, false, out, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  g_conv_path_counts[2].fetch_add(1, std::memory_order_relaxed);
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
  const bool umma = tc_fwd_ok(K, c_out, c_in);
  if (precision == SPC_PREC_BF16) {
    SPC_REQUIRE(umma, "bf16 mode needs a tensor-core shape (Cout % 32 == 0, Cin % 16 == 0, K <= 32)");
    return conv_fwd_umma(dout, w, nullptr, nullptr, nbr_t, tile_mask_t, m_in, c_out, c_in, K, true, true, din, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  }
  if (precision == SPC_PREC_TF32 && umma)
    return conv_fwd_umma(dout, w, nullptr, nullptr, nbr_t, tile_mask_t, m_in, c_out, c_in, K, true, false, din, workspace,
                         workspace_bytes, (cudaStream_t)stream);
  g_conv_path_counts[2].fetch_add(1, std::memory_order_relaxed);
  return conv_fwd_simt((const float*)dout, w, nullptr, nbr_t, m_in, c_out, c_in, K, true, din, (cudaStream_t)stream);
}

int spc_conv_wgrad(const void* in, const void* dout, const int32_t* nbr,
                   const uint32_t* tile_mask, int64_t m_in,
                   int64_t m_out, int c_in, int c_out, int K, int precision, float* dw,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  (void)workspace; (void)workspace_bytes;
  return spc_conv_wgrad_acc(in, dout, nbr, tile_mask, m_in, m_out, c_in, c_out, K, precision, dw, 0, stream);
}

int spc_conv_wgrad_acc(const void* in, const void* dout, const int32_t* nbr, const uint32_t* tile_mask, int64_t m_in,
                       int64_t m_out, int c_in, int c_out, int K, int precision, float* dw, int accumulate,
                       void* stream) {
  (void)m_in;
  SPC_REQUIRE(c_in >= 1 && c_out >= 1 && K >= 1, "bad shape");
  SPC_REQUIRE(prec_ok(precision), "bad precision mode");
  const bool umma = tc_wgrad_ok(K, c_in, c_out);
  if (precision == SPC_PREC_BF16) {
    SPC_REQUIRE(umma, "bf16 mode needs a tensor-core shape (Cin, Cout % 32 == 0, Cout <= 256, K <= 32)");
    return conv_wgrad_umma(in, dout, nbr, tile_mask, m_out, c_in, c_out, K, true, dw, !accumulate, (cudaStream_t)stream);
  }
  if (precision == SPC_PREC_TF32 && umma)
    return conv_wgrad_umma(in, dout, nbr, tile_mask, m_out, c_in, c_out, K, false, dw, !accumulate, (cudaStream_t)stream);
  SPC_REQUIRE(!accumulate, "the CUDA-core wgrad kernel overwrites dw (accumulate needs a tensor-core shape)");
  g_conv_path_counts[2].fetch_add(1, std::memory_order_relaxed);
  return conv_wgrad_simt((const float*)in, (const float*)dout, nbr, m_out, c_in, c_out, K, dw, (cudaStream_t)stream);
}

}  // extern "C"
