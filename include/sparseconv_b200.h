/*
 * sparseconv_b200.h — C ABI of the B200-native sparse-convolution hot path.
 *
 * This is the drop-in boundary.  The reference
 * (POSTECH-CVLab/NeRF-Downstream) reaches this path only through the Python
 * package `MinkowskiEngine`, whose native backend (`MinkowskiEngineBackend._C`,
 * third-party, not vendored) it imports at
 *   co3d_3d/src/models/mink/modules/sparse_conv.py:7-12.
 * Every entry point below cites the reference call site whose behaviour it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; the text of the last
 *     error on the calling thread is returned by spc_last_error();
 *   - all data pointers are DEVICE pointers (torch `tensor.data_ptr()`), sizes
 *     are int64_t, `stream` is a cudaStream_t passed as void*;
 *   - nothing is allocated inside: callers pass outputs and workspaces;
 *   - no function synchronises the stream; sizes that are only known on the
 *     device (number of unique voxels) are written to a device int32 the caller
 *     reads back when it needs them.
 *
 * Coordinate rows are int32 [M,4] = (batch, x, y, z)  (co3d_3d/src/data/co3d.py:121
 * strips column 0 as the batch index).  Supported range: batch in [0,1022],
 * x,y,z in [-131072,131071]; anything else sets the error flag (`status[0]`).
 */
#ifndef SPARSECONV_B200_H_
#define SPARSECONV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPC_ABI_VERSION 1

/* bytes per hash-table slot: {u64 key, u32 first_row, u32 map_row} */
#define SPC_SLOT_BYTES 16

/* coordinate source kinds for spc_coords_insert */
#define SPC_SRC_FLOAT 0   /* float32 [N,4]; quantised with floor(x/ts)*ts  */
#define SPC_SRC_INT 1     /* int32   [N,4]; inserted as is                  */
#define SPC_SRC_STRIDE 2  /* int32   [N,4]; floor_div(c, ts)*ts per axis    */

/* precision modes of the convolution kernels */
#define SPC_PREC_FP32 0   /* CUDA-core FP32 FMA, fp32-faithful             */
#define SPC_PREC_TF32 1   /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM */
#define SPC_PREC_BF16 2   /* tcgen05.mma kind::f16 on bf16 copies of the rows (spc_to_bf16), fp32 accumulate */

int spc_abi_version(void);
const char* spc_last_error(void);

/* Number of table slots to allocate for n keys (power of two, load <= 0.5). */
int64_t spc_table_slots(int64_t n);
/* Workspace bytes needed by spc_coords_insert for n source rows. */
int64_t spc_coords_insert_workspace(int64_t n);

/*
 * Voxel quantisation + coordinate hashing + unique / inverse map.
 * Replaces ME.TensorField(...).sparse() (co3d_3d/src/models/mink/resnet.py:164,
 * res16unet.py:392, base_model.py:10-13) and CoordinateManager.stride()
 * (modules/sparse_conv.py:403-405).
 *   src        [n,4] float32 or int32 (see SPC_SRC_*)
 *   ts[3]      tensor stride applied when quantising (1,1,1 for plain floor)
 *   slots      n_slots*16 bytes, n_slots = spc_table_slots(n); (re)initialised here
 *   out_coords [n,4] int32; rows [0,M) valid, first-occurrence order
 *   out_first  [n] int32;  out_first[r] = first source row of voxel r (ME unique_index)
 *   out_inverse[n] int32;  voxel row of every source row (ME inverse_mapping)
 *   out_count  [n] int32;  rows [0,M): number of source rows per voxel
 *   status     [2] int32;  status[0]=M, status[1]=error flag (1 = coordinate out of range)
 */
int spc_coords_insert(const void* src, int64_t n, int src_kind, const int32_t* ts,
                      void* slots, int64_t n_slots,
                      int32_t* out_coords, int32_t* out_first, int32_t* out_inverse,
                      int32_t* out_count, int32_t* status,
                      void* workspace, int64_t workspace_bytes, void* stream);
/* Same, with the number of valid source rows read on the DEVICE: rows = min(n, *n_dev) when n_dev is
 * not NULL (`n` is then an upper bound that sizes the launch and every buffer).  Lets a pyramid of
 * stride maps (status[0] of one level = n_dev of the next) be enqueued without a host round trip
 * per level. */
int spc_coords_insert_dev(const void* src, int64_t n, const int32_t* n_dev, int src_kind, const int32_t* ts,
                          void* slots, int64_t n_slots, int32_t* out_coords, int32_t* out_first,
                          int32_t* out_inverse, int32_t* out_count, int32_t* status, void* workspace,
                          int64_t workspace_bytes, void* stream);

/*
 * Kernel-map construction (CoordinateManager.kernel_map, sparse_conv.py:90-96,
 * :197-204).  For every out row o and offset k: nbr[k*M_out+o] = in row whose
 * coordinate is coord_out[o] + offsets[k], or -1.
 *   offsets  HOST pointer, [K,3] int32, already scaled by in tensor stride x dilation;
 *            index order: first spatial axis fastest (sparse_conv.py:375-379)
 *   nbr      [K, M_out] int32 (offset-major)
 *   tap_count[K] int32 device: number of pairs per offset (zeroed here)
 */
int spc_kernel_map(const void* in_slots, int64_t in_n_slots,
                   const int32_t* out_coords, int64_t m_out,
                   const int32_t* offsets_host, int K,
                   int32_t* nbr, int32_t* tap_count, void* stream);

/*
 * Per-tile offset mask: bit k of mask[t] is set iff some row of the 128-row tile t has a
 * neighbour at offset k.  The tcgen05 convolution skips offsets whose bit is clear.
 *   mask [ceil(m/128)] uint32, K <= 32.
 */
int spc_tile_mask(const int32_t* nbr, int64_t m, int K, uint32_t* mask, void* stream);

/* spc_kernel_map for a SELF map (out map == the map the table indexes) with centrally symmetric offsets
 * (offsets[K-1-k] == -offsets[k], K odd — every odd-kernel stride-1 convolution): nbr[k][o] = i <=> nbr[K-1-k][i] = o,
 * so only K/2 offsets are probed and the mirrored entries are written by the thread that found them.  Same result
 * as spc_kernel_map, half the hash probes. */
int spc_kernel_map_sym(const void* slots, int64_t n_slots, const int32_t* coords, int64_t m,
                       const int32_t* offsets_host, int K, int32_t* nbr, int32_t* tap_count, void* stream);
/* Transposed dense map: nbr_t[k*M_in + i] = o  iff  nbr[k*M_out + o] = i. */
int spc_kernel_map_transpose(const int32_t* nbr, int64_t m_out, int64_t m_in, int K,
                             int32_t* nbr_t, void* stream);

/*
 * ME-style pair lists (sparse_conv.py:122-143): for each offset k the pairs
 * (in,out) ascending in out row, concatenated k-major.
 *   pairs [2, P] int32 (row 0 = in rows, row 1 = out rows), P = sum tap_count
 *   tap_offset [K+1] int32 device: start of each offset's segment
 *   workspace: spc_pairs_workspace(m_out, K) bytes
 */
int64_t spc_pairs_workspace(int64_t m_out, int K);
int spc_kernel_map_pairs(const int32_t* nbr, int64_t m_out, int K, int64_t pair_capacity,
                         int32_t* pairs, int32_t* tap_offset,
                         void* workspace, int64_t workspace_bytes, void* stream);

/*
 * Per-voxel feature reduction of TensorField.sparse() (UNWEIGHTED_AVERAGE,
 * res16unet.py:665-667): out[r] = mean/sum of feats[j] over inverse[j]==r.
 *   mode 0 = average, 1 = sum.  `out` [m, C] is zeroed here.
 */
int spc_segment_reduce(const float* feats, const int32_t* inverse, const int32_t* count,
                       int64_t n, int64_t m, int C, int mode, float* out, void* stream);
/* Backward of the above and SparseTensor.slice (res16unet.py:435):
 * out[j] = src[index[j]] * (count ? 1/count[index[j]] : 1). */
int spc_gather_rows(const float* src, const int32_t* index, const int32_t* count,
                    int64_t n, int C, float* out, void* stream);
/* out[index[j]] += src[j]  (backward of slice); out zeroed here. m rows. */
int spc_scatter_add_rows(const float* src, const int32_t* index, int64_t n, int64_t m, int C,
                         float* out, void* stream);

/*
 * Sparse convolution (MinkowskiConvolution / MinkowskiConvolutionTranspose,
 * modules/common.py:117-125,172-180; arithmetic restated at sparse_conv.py:122-143).
 *   fwd  : out[o,:]  = sum_k  in[nbr[k,o],:] @ W[k]            (+ bias)
 *   dgrad: din[i,:]  = sum_k  dout[nbr_t[k,i],:] @ W[k]^T
 *   wgrad: dW[k]     = sum_o  in[nbr[k,o],:]^T  dout[o,:]
 *   W is [K, Cin, Cout] row-major fp32.
 *   tile_mask / tile_mask_t: spc_tile_mask of nbr / nbr_t, or NULL (= all offsets active).
 *   workspace: spc_conv_workspace(...) bytes (packed weights for the TF32 path).
 */
void spc_debug_force_mt(int mt);
void spc_debug_set(int idx, int val); /* test / measurement hook, every knob 0 = default: 0 forward: number of producer
                                        * groups (1, 2, 4, 8); 2 forward: 1 = st.global epilogue instead of TMA stores;
                                        * 3 forward: epilogue writes nothing (timing only, wrong results); 4 forward: 1 =
                                        * one 32-channel chunk per stage; 5 wgrad: 1 = one row visit per chunk;
                                        * 6 forward: producer warps (8 / 16); 7 forward: 1 = no shared ring slots;
                                        * 8 forward / dgrad: 1 = never the CTA-pair kernel (cta_group::2), 2 = on every
                                        * eligible shape whatever the map size; 9 wgrad: 2 = the CTA-pair wgrad kernel;
                                        * 10 pair kernel: ring slots; 11 pair kernel: 1 = one staging block.  Results do
                                        * not depend on knobs 0, 2, 4-11. */
/* launches per convolution route since the last reset: out3[0] tcgen05 bf16, [1] tcgen05 tf32, [2] CUDA-core fp32
 * (out3 may be NULL); 1 if spc_conv_fwd (what = 0) / dgrad (1) / wgrad (2) runs the shape on the tensor cores */
void spc_conv_path_counts(long long* out3, int reset);
int spc_conv_tensor_core(int what, int K, int c_in, int c_out, int precision);
int64_t spc_conv_workspace(int K, int c_in, int c_out, int precision);
/* fp32 rows -> dense bf16 rows (round to nearest even) for the SPC_PREC_BF16 convolutions: src has `rows` rows of
 * c_src valid columns at a pitch of src_pitch elements (a column slice of a wider tensor is read in place); dst
 * is [rows, c_dst] with c_dst >= c_src, columns past c_src zero (channel padding to the tensor-core widths). */
int spc_to_bf16(const float* src, int64_t rows, int c_src, int64_t src_pitch, int c_dst, void* dst_bf16, void* stream);
/* `in` / `dout` point to fp32 rows, or to bf16 rows when precision == SPC_PREC_BF16. */
int spc_conv_fwd(const void* in, const float* w, const float* bias, const int32_t* nbr,
                 const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision,
                 float* out, void* workspace, int64_t workspace_bytes, void* stream);
/* spc_conv_fwd that also accumulates, in its epilogue, the per-channel sum and sum of squares of the output
 * rows into bn_stats[2 * c_out] (doubles, zeroed here) for the BatchNorm that follows (spc_bn_finalize), saving
 * that layer's statistics pass over the rows.  Done only when the launch has one output-channel tile and no
 * offset split (large maps, c_out <= 256): *stats_fused (HOST int, may be NULL) says whether bn_stats is valid. */
int spc_conv_fwd_stats(const void* in, const float* w, const float* bias, const int32_t* nbr,
                       const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                       int precision, float* out, double* bn_stats, int32_t* stats_fused, void* workspace,
                       int64_t workspace_bytes, void* stream);
int spc_conv_dgrad(const void* dout, const float* w, const int32_t* nbr_t,
                   const uint32_t* tile_mask_t, int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision,
                   float* din, void* workspace, int64_t workspace_bytes, void* stream);
int spc_conv_wgrad(const void* in, const void* dout, const int32_t* nbr,
                   const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision,
                   float* dw, void* workspace, int64_t workspace_bytes, void* stream);
/* spc_conv_wgrad with `accumulate` != 0: dw is not cleared, the gradient is ADDED to what it holds (a slice of
 * the trainer's gradient arena, zeroed once per step) — tensor-core shapes only (spc_conv_tensor_core(2, ...)). */
int spc_conv_wgrad_acc(const void* in, const void* dout, const int32_t* nbr, const uint32_t* tile_mask, int64_t m_in,
                       int64_t m_out, int c_in, int c_out, int K, int precision, float* dw, int accumulate,
                       void* stream);
/* Weights in the tensor-core kernels' shared-memory image ([K][C/32][C'][32] swizzled rows, independent of the
 * tile shape), so that a layer packs ONCE per optimiser step instead of once per launch: `dgrad` = 0 packs W for
 * spc_conv_fwd_packed, 1 packs W^T for spc_conv_dgrad_packed, 2 packs W^T with the kernel offsets REVERSED (slab k =
 * W[K-1-k]^T): on a centrally symmetric self map (odd kernel, stride 1, spc_kernel_map_sym) nbr_t[k] == nbr[K-1-k],
 * so spc_conv_dgrad_packed is then called with the FORWARD map and tile mask and no transposed map is ever built.
 * `packed`: spc_conv_packed_bytes() bytes, 1024-byte aligned.  precision: SPC_PREC_TF32 or SPC_PREC_BF16; shapes: spc_conv_tensor_core(0 / 1, ...). */
int64_t spc_conv_packed_bytes(int K, int c_in, int c_out);
int spc_conv_pack_weights(const float* w, int K, int c_in, int c_out, int dgrad, int precision, void* packed,
                          void* stream);
/* spc_conv_pack_weights for n_layers (layer, direction) pairs in ONE launch.  desc_dev: DEVICE int64 [n_layers][8] =
 * { w (fp32 [K, c_in, c_out]), packed (1024-byte aligned), K, Ck, Cn, flags, bf16, 0 } with (Ck, Cn, flags) =
 * (c_in, c_out, 0) for the forward image, (c_out, c_in, 1) for the dgrad image and (c_out, c_in, 3) for the dgrad
 * image with reversed offsets. */
int spc_conv_pack_weights_batch(const int64_t* desc_dev, int n_layers, void* stream);
int spc_conv_fwd_packed(const void* in, const void* w_packed, const float* bias, const int32_t* nbr,
                        const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                        int precision, float* out, void* stream);
/* spc_conv_fwd_packed that also leaves the per-column sum and sum of squares of `out` in bn_sums[2 * c_out] (DEVICE
 * doubles, cleared here) for the BatchNorm that follows (spc_bn_finalize) — the statistics pass over the rows (4 of the
 * 10-18 bytes per element BatchNorm forward moves) is then not needed.  The epilogue reads the sums back from the
 * staging blocks of its TMA stores.  *stats_fused (HOST int) = 1 if the sums were produced: large maps (one owner per
 * output row), c_out <= 128, no bias; otherwise 0 and bn_sums is untouched (run spc_bn_stats). */
int spc_conv_fwd_packed_stats(const void* in, const void* w_packed, const float* bias, const int32_t* nbr,
                              const uint32_t* tile_mask, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                              int precision, float* out, double* bn_sums, int32_t* stats_fused, void* stream);
int spc_conv_dgrad_packed(const void* dout, const void* w_packed_t, const int32_t* nbr_t, const uint32_t* tile_mask_t,
                          int64_t m_in, int64_t m_out, int c_in, int c_out, int K, int precision, float* din,
                          void* stream);
/* spc_conv_dgrad_packed with accumulate != 0: din += the input gradient (TMA reduce-add epilogue).  `din` then holds
 * the gradient that reached the same rows through the block's residual connection (resnet_block.py:53-69: x feeds
 * conv1 AND `out += residual`), so the sum autograd would form with a separate pass over both is written once. */
int spc_conv_dgrad_packed_acc(const void* dout, const void* w_packed_t, const int32_t* nbr_t,
                              const uint32_t* tile_mask_t, int64_t m_in, int64_t m_out, int c_in, int c_out, int K,
                              int precision, float* din, int accumulate, void* stream);

/*
 * BatchNorm over voxel rows (MinkowskiBatchNorm == nn.BatchNorm1d on .F,
 * modules/common.py:24) with optional fused ReLU and residual add
 * (resnet_block.py:66-67).
 *   stats  : sum[c], sumsq[c] over m rows -> mean, biased var (double-accumulated)
 *   apply  : y = (x-mean)*rsqrt(var+eps)*gamma+beta (+res) ; relu optional
 *   bwd    : dx, dgamma, dbeta (and dres = masked dy when res was fused)
 */
int64_t spc_bn_workspace(int64_t m, int C);
int spc_bn_stats(const float* x, int64_t m, int C, float* mean, float* var,
                 float* running_mean, float* running_var, float momentum, /* nullable */
                 void* workspace, int64_t workspace_bytes, void* stream);
/* spc_bn_stats that also increments nn.BatchNorm1d's num_batches_tracked (device int64, nullable) in the
 * reduction's last block — no separate one-element launch per BatchNorm layer and step. */
int spc_bn_stats_tracked(const float* x, int64_t m, int C, float* mean, float* var, float* running_mean,
                         float* running_var, float momentum, int64_t* num_batches_tracked, void* workspace,
                         int64_t workspace_bytes, void* stream);
/* mean / biased variance (+ running statistics) from the sums written by spc_conv_fwd_stats. */
int spc_bn_finalize(const double* sums, int64_t m, int C, float* mean, float* var, float* running_mean,
                    float* running_var, float momentum, int64_t* num_batches_tracked /* nullable, incremented */,
                    void* stream);
int spc_bn_apply(const float* x, const float* mean, const float* var, const float* gamma,
                 const float* beta, const float* residual, int64_t m, int C, float eps,
                 int relu, float* y /* NULL: only the bf16 copy is produced */,
                 void* y_bf16 /* optional bf16 copy of y, or NULL */, void* stream);
int spc_bn_bwd(const float* x, const float* y, const void* y_bf16 /* ReLU mask from the bf16 copy of y instead of y, or NULL */,
               const float* dy, int64_t dy_pitch /* elements between rows of dy (>= C; a column slice is read in place) */,
               const float* mean, const float* var, const float* gamma, int64_t m, int C, float eps, int relu,
               int training, float* dx, void* dx_bf16 /* optional bf16 copy of dx, or NULL */,
               float* dresidual, float* dgamma, float* dbeta,
               void* workspace, int64_t workspace_bytes, void* stream);

/* spc_bn_bwd whose dgamma / dbeta (either may be NULL) are ADDED to what the buffers hold when
 * accumulate_param_grads != 0 (slices of a gradient arena that is zeroed once per step).
 *   relu = 2 : the ReLU mask is RE-COMPUTED from x with the forward affine (gamma, beta, mean, var: y > 0 <=>
 *              fma(x, sc, sh) > 0, the expression spc_bn_apply evaluates) — valid when no residual was added before
 *              the ReLU; y / y_bf16 are then not read at all.  relu = 1 reads y (or y_bf16) as spc_bn_bwd does.
 *   dx may be NULL when dx_bf16 is given (the consumer is a bf16 convolution's dgrad / wgrad only). */
int spc_bn_bwd_acc(const float* x, const float* y, const void* y_bf16, const float* dy, int64_t dy_pitch,
                   const float* mean, const float* var, const float* gamma, const float* beta, int64_t m, int C,
                   float eps, int relu, int training, float* dx, void* dx_bf16, float* dresidual, float* dgamma, float* dbeta,
                   int accumulate_param_grads, void* workspace, int64_t workspace_bytes, void* stream);

/* dst[r, 0:row_bytes) = src[r, 0:row_bytes) for pitched rows (pitches in BYTES): ME.cat (res16unet.py:410-425) of the
 * bf16 operand copies straight into column slices of the concatenated operand. */
int spc_copy_rows(const void* src, int64_t src_pitch, void* dst, int64_t dst_pitch, int64_t row_bytes, int64_t rows,
                  void* stream);

/* y = relu(x) ; dx = dy * (y > 0) ; y = a + b  (MinkowskiReLU, SparseTensor +=). */
int spc_relu_fwd(const float* x, int64_t n, float* y, void* stream);
int spc_relu_bwd(const float* y, const float* dy, int64_t n, float* dx, void* stream);
int spc_add(const float* a, const float* b, int64_t n, float* y, void* stream);

/*
 * Local pooling with kernel_size == stride (MinkowskiSumPooling / AvgPooling,
 * resnet.py:62-64, co3d.py:107-111) over the stride map `parent` (in row ->
 * out row, the inverse map of spc_coords_insert with SPC_SRC_STRIDE).
 *   fwd: out[o] = sum_k in[nbr[k,o]]  (avg: / number of present inputs), a gather over the
 *        kernel map `nbr` [K, m_out] of the pooling region, so no atomics are needed.
 *   bwd: din[j] = dout[parent[j]] (avg: / count[parent[j]]) == spc_gather_rows(dout, parent, count).
 */
int spc_pool_fwd(const float* in, const int32_t* nbr, int64_t m_out, int C, int K, int avg,
                 float* out, void* stream);

/*
 * Global average / sum pooling (MinkowskiGlobalAvgPooling, resnet.py:18,175):
 * out[b,:] = mean over rows with batch index b; rows ordered by batch index.
 *   coords [m,4] int32; n_batch rows of output; cnt [n_batch] int32 (written).
 */
int spc_global_pool_fwd(const float* in, const int32_t* coords, int64_t m, int C, int n_batch,
                        int avg, float* out, int32_t* cnt, void* stream);
int spc_global_pool_bwd(const float* dout, const int32_t* coords, const int32_t* cnt, int64_t m,
                        int C, int n_batch, int avg, float* din, void* stream);

/* Softmax cross-entropy with ignore_index over rows [n, C] (C <= 64), mean over the non-ignored rows
 * (classification_training.py:33, segmentation_training.py:27-44: nn.CrossEntropyLoss(ignore_index=)).
 * fwd: stats[0] = sum of -log p[target], stats[1] = number of non-ignored rows (doubles, device),
 *      grad_raw[i, c] = p[i, c] - [c == target[i]] (0 for ignored rows); *bad_target = 1 if a target
 *      is outside [0, C) and is not ignore_index.  loss = stats[0] / stats[1].
 * bwd: dlogits = grad_raw * grad_out[0] / stats[1]. */
int spc_ce_fwd(const float* logits, const int64_t* target, int64_t n, int C, int64_t ignore_index,
               float* grad_raw, double* stats, int32_t* bad_target, void* stream);
int spc_ce_bwd(const float* grad_raw, const double* stats, const float* grad_out, int64_t n, int C,
               float* dlogits, void* stream);


/* ---- max pooling (ME.MinkowskiMaxPooling / MinkowskiGlobalMaxPooling) --------------------
 * local: out[o, c] = max over present neighbours of the kernel map, arg[o, c] = winning input row (-1: none);
 * backward routes dout to the winning rows (din zeroed inside).  global: per batch index, rows found through
 * coords[:, 0]; arg[b, c] = first row attaining the maximum; backward = spc_pool_max_bwd with m_out = n_batch. */
int spc_pool_max_fwd(const float* in, const int32_t* nbr, int64_t m_out, int C, int K, float* out, int32_t* arg,
                     void* stream);
int spc_pool_max_bwd(const float* dout, const int32_t* arg, int64_t m_out, int64_t m_in, int C, float* din,
                     void* stream);
int spc_global_max_fwd(const float* in, const int32_t* coords, int64_t m, int C, int n_batch, float* out,
                       int32_t* arg, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- the two ends of the path (SURVEY.md §8f "next" rows 2 and 3) ------------------------------------------
 * spc_plenoxel_decode: one PeRFception plenoxel record -> network input (co3d_3d/src/data/co3d.py:164-172,196-203).
 *   links[n] (int32 or int64 flat indices of the occupied cells of a reso[0] x reso[1] x reso[2] grid) ->
 *   out_coords[n,4] float (batch_index, i, j, k) with i = links / (r1 r2), j = (links % (r1 r2)) / r2, k = links % r2,
 *   optionally mapped through affine12 = HOST float[12] (row-major 3x3 then translation; NULL = identity), and
 *   sh_u8[n,C] uint8 -> out_feats[n,C] float = sh * sh_scale + sh_min (two roundings, as numpy / torch compute it).
 * spc_seg_metrics: IoUMeter.update (co3d_3d/src/metrics.py:29-41) in one pass: counts[3][C] (uint64, ACCUMULATED,
 *   caller zeroes) += per-class #seen, #correct, #predicted of argmax(logits[n,C]) over rows with target != ignore. */
int spc_plenoxel_decode(const void* links, int links_is_int64, int64_t n, const int32_t* reso, int32_t batch_index,
                        const float* affine12, const uint8_t* sh_u8, int C, float sh_scale, float sh_min,
                        float* out_coords, float* out_feats, void* stream);
int spc_seg_metrics(const float* logits, const int64_t* target, int64_t n, int C, int64_t ignore_label,
                    uint64_t* counts, void* stream);

/* Row selections of a plenoxel record on the device: the reference's RandomCrop and CoordinateDropout
 * (co3d_3d/src/data/transforms.py:195-265) drop points BEFORE the network sees them; here they become a row list and
 * the record is decoded once, for the kept rows only, in the reference's row order.
 * spc_plenoxel_decode_rows: spc_plenoxel_decode for output row j <- record rows[j] (rows: DEVICE int32 [n_rows]; a
 *   CoordinateDropout's `np.random.choice` index list, or the list spc_plenoxel_crop_select wrote).
 * spc_plenoxel_crop_select: RandomCrop.__call__ for one draw of the box position.  Over the records rows[0..n) (NULL:
 *   all n records) with lattice coordinates mapped through affine12 (HOST float[12] or NULL: the affine chain applied
 *   before the crop): extent, then keep the points with  u * clip(max - min - size, 0) < p - min < ... + size  on every
 *   axis (u3, size3: HOST float[3]; float32 arithmetic), stable compaction of the kept record numbers into out_rows
 *   (DEVICE int32 [n]).  result2 (DEVICE int32 [2]) = { rows kept, 1 if the box covers the extent on every axis (the
 *   reference then returns its input unchanged) }.  The caller reads result2 once (RandomCrop retries an empty box). */
int spc_plenoxel_decode_rows(const void* links, int links_is_int64, const int32_t* rows, int64_t n_rows,
                             const int32_t* reso, int32_t batch_index, const float* affine12, const uint8_t* sh_u8, int C,
                             float sh_scale, float sh_min, float* out_coords, float* out_feats, void* stream);
int64_t spc_plenoxel_crop_workspace(int64_t n);
int spc_plenoxel_crop_select(const void* links, int links_is_int64, const int32_t* rows, int64_t n, const int32_t* reso,
                             const float* affine12, const float* u3, const float* size3, int32_t* out_rows,
                             int32_t* result2, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- segmentation head --------------------------------------------------------------------
 * spc_seg_head_fwd: `out.slice(x).F` (res16unet.py:435) -> `SegLoss` (segmentation_training.py:27-44) ->
 *   `IoUMeter.update` (metrics.py:29-41) in ONE pass over the n points.
 *   logits[m, C]      voxel rows of the network output
 *   inverse[n]        point -> voxel row (ME's inverse_mapping, int32); NULL = identity (then m == n): a plain
 *                     class-weighted F.cross_entropy over rows
 *   target[n]         int64 labels; == ignore_index rows contribute nothing
 *   class_weight[C]   float or NULL (all one) — F.cross_entropy's `weight` (SegLoss: ones, void_weight on the last)
 *   grad_raw[m, C]    out: sum over the points of a voxel of w[y] * (softmax - onehot)  (= backward of the slice
 *                     gather already applied); spc_ce_bwd(grad_raw, stats, gout, m, C) turns it into dlogits
 *   stats[2]          out (double): sum w[y] * nll, sum w[y]  -> loss = stats[0] / stats[1]
 *   counts[3, C]      uint64, ACCUMULATED (caller zeroes), or NULL: #seen, #correct, #predicted per class
 *   bad_target[1]     out: 1 = a label outside [0, C) other than ignore_index, 2 = inverse entry outside [0, m) */
int spc_seg_head_fwd(const float* logits, int64_t m, const int32_t* inverse, const int64_t* target, int64_t n, int C,
                     int64_t ignore_index, const float* class_weight, float* grad_raw, double* stats,
                     uint64_t* counts, int32_t* bad_target, void* stream);

/* ---- instance normalisation (ME.MinkowskiInstanceNorm, modules/common.py:25-26) ------------
 * Per (batch index = coords[r, 0], channel): mean and biased variance over the rows of that instance,
 *   y = (x - mean) / sqrt(var + eps) * gamma[c] + beta[c]      (gamma / beta may be NULL)
 * fwd writes mean[n_batch, C], rstd[n_batch, C], cnt[n_batch] for the backward; ws = n_batch * 2 * C doubles.
 * bwd: dx = gamma * rstd * (dy - mean_b(dy) - xhat * mean_b(dy * xhat)); sums[n_batch, 2, C] (double) returns
 *   sum dy and sum dy * xhat per instance, from which dgamma = sum_b sums[b, 1], dbeta = sum_b sums[b, 0]. */
int spc_inst_norm_fwd(const float* x, const int32_t* coords, int64_t m, int C, int n_batch, const float* gamma,
                      const float* beta, float eps, float* y, float* mean, float* rstd, int32_t* cnt, double* ws,
                      void* stream);
int spc_inst_norm_bwd(const float* x, const float* dy, const int32_t* coords, int64_t m, int C, int n_batch,
                      const float* gamma, const float* mean, const float* rstd, const int32_t* cnt, float* dx,
                      double* sums, void* stream);
/* Debug knob: 1 = always use the scalar kernels (C % 4 == 0 inputs normally take the 16-byte kernels); tests use it
 * to cross-check the two paths. */
void spc_inst_norm_force_scalar(int on);

/* ---- trilinear interpolation / splat (ME.MinkowskiInterpolation, SparseTensor.interpolate,
 * TensorField.splat; fcnn.py:184-205, transforms.py:472,520-528) ----------------------------------------------------
 * spc_interp_corners: query[n,4] float (b,x,y,z) -> lower[n,4] int32 = (floor(b), floor(x/ts)*ts, ...) and
 *   weights[8,n]: corner k = bx + 2 by + 4 bz sits at lower + (bx,by,bz)*ts with weight prod(axis: b ? f : 1 - f),
 *   f = x/ts - floor(x/ts).  ts = HOST int32[3].  The rows idx[8,n] of the corners in a voxel map are
 *   spc_kernel_map(table, lower, n, the 8 corner offsets) (interpolate) or the inverse map of inserting them (splat).
 * spc_interp_fwd: out[j,:] = sum_k weights[k,j] * feats[idx[k,j],:]   (idx < 0: corner not in the map, skipped)
 * spc_interp_bwd: dfeats[idx[k,j],:] += weights[k,j] * dout[j,:]      (dfeats [m,C] zeroed here) = the splat forward */
int spc_interp_corners(const float* query, int64_t n, const int32_t* ts, int32_t* lower, float* weights, void* stream);
int spc_interp_fwd(const float* feats, const int32_t* idx, const float* weights, int64_t n, int C, int K, float* out,
                   void* stream);
int spc_interp_bwd(const float* dout, const int32_t* idx, const float* weights, int64_t n, int64_t m, int C, int K,
                   float* dfeats, void* stream);

/* Fused SGD step on a flat arena (co3d_cls.gin:33-39; optim.py:60-69):
 * g = grad*grad_scale + wd*p ; buf = mom*buf + g ; p -= lr*buf. */
int spc_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr,
                 float momentum, float weight_decay, float grad_scale, int first_step,
                 void* stream);

/*
 * Engine-side row order (no ME counterpart: MinkowskiEngine's GPU maps leave the row order unspecified).  A tensor-core
 * convolution executes an offset for a whole 128-row tile if ANY row of the tile has a neighbour there; on thin
 * geometry (the faithful ScanNet-plenoxel lattices: 5-8 neighbours per voxel, but 24-27 offsets per raster-ordered
 * tile) grouping rows with equal neighbourhoods inside windows lets the tile masks skip most of that.
 *   spc_row_masks     row_mask[o] = bit set of offsets k with nbr[k, o] >= 0 (K <= 32); *executed_dev (DEVICE int64) =
 *                     number of (tile, offset) pairs executed on the CURRENT row order
 *   spc_table_relabel every occupied slot's map row r becomes pos[r] (pos: DEVICE int32 [M], a permutation): after the
 *                     caller has permuted the coordinate rows, look-ups and kernel maps come out in the new order
 */
int spc_row_masks(const int32_t* nbr, int64_t m, int K, uint32_t* row_mask, int64_t* executed_dev, void* stream);
int spc_table_relabel(void* slots, int64_t n_slots, const int32_t* pos, void* stream);

/* Number of kernels launched through this library since load (bench `gpu_launches`). */
int64_t spc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SPARSECONV_B200_H_ */
