/*
 * densefusion_b200.h -- C ABI of libdensefusion_b200.so (hand-written sm_100a CUDA kernels for the
 * DenseFusion per-pixel pose-hypothesis path).
 *
 * Conventions (mirroring the reference's only native interface, lib/knn/src/knn_cuda_kernel.h:14-16
 * and lib/knn/src/knn_pytorch.h:1):
 *   - plain pointers and sizes, no torch types; every pointer is a DEVICE pointer;
 *   - the caller owns all buffers including outputs and scratch; nothing is retained between calls;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), no host synchronisation, no
 *     allocation -> every entry point is CUDA-graph capturable and re-entrant across streams;
 *   - return 0 on success, DF_ERR_ARG (-1) / DF_ERR_UNSUPPORTED (-2) for rejected arguments, or the
 *     positive cudaError_t of a failed launch.  Nothing throws across this boundary.
 *   - fp32 data, int64 indices, row-major, "point-major" activations: (rows = crops*points, channels).
 */
#ifndef DENSEFUSION_B200_H
#define DENSEFUSION_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------- */
/* ABI version and a bit mask of compiled features (bit 0: tcgen05 GEMM path). */
int df_abi_version(void);
int df_features(void);
/* Measured-peak probe for the roofline report: launches blocks x 256 threads x 8 independent FFMA chains of `iters` steps
 * (no memory traffic) and returns the number of flops launched (negative: error).  Not part of the pose path. */
long long df_probe_ffma(float* sink, int blocks, int iters, void* stream);

/* ---- K4: k nearest neighbours -----------------------------------------------------------------
 * Replaces   int knn(THCudaTensor *ref, THCudaTensor *query, THCudaLongTensor *idx)
 *            (lib/knn/src/knn_pytorch.h:1, lib/knn/src/knn_pytorch.c:6-48)   and
 *            void knn_device(float*, int, float*, int, int, int, float*, long*, cudaStream_t)
 *            (lib/knn/src/knn_cuda_kernel.h:14-16).
 * ref (batch, dim, R), query (batch, dim, Q) dim-major fp32 -> ind (batch, k, Q) int64, 1-BASED.
 * Bit-identical indices to the reference kernels (FMA-chain distance, strict '<', lowest index on
 * ties, NaN never inserted).  No R x Q scratch is needed (the reference's dist_dev argument is gone).
 * dim == 3 && k == 1 is the tiled fast path; other (dim, k <= 64) use the general kernel. */
int df_knn(const float* ref, const float* query, int64_t* ind, int batch, int dim, int R, int Q, int k,
           void* stream);

/* ---- K3: fused ADD / ADD-S loss + hypothesis scoring + argmax selection ------------------------
 * Replaces lib/loss.py:13-70 (pred_c != NULL, hyp_points = the cloud) and lib/loss_refiner.py:12-62
 * (pred_c == NULL, hyp_points == NULL, P == 1), batched over B crops.
 *   pred_r (B,P,4) un-normalised (w,x,y,z); pred_t (B,P,3); pred_c (B,P); target, model_points (B,M,3);
 *   hyp_points (B,P,3) per-hypothesis anchor (points + pred_t); points (B,N,3) cloud to re-express;
 *   idx (B) int64 object id; bit o of sym_mask set <=> object o is symmetric; allow_sym = 0 reproduces
 *   Loss(..., refine=True).
 * Outputs: loss (B), dis_sel (B) distance of the selected hypothesis, which (B) its index,
 *   new_points (B,N,3), new_target (B,M,3); dis_all (B,P), sum_u (B,P,3), sum_um (B,P,9) are the
 *   per-hypothesis state df_loss_backward consumes.  tickets: (B) uint32, zero before the first call
 *   (the kernel leaves it zero).  dbg_pred (B,P,M,3) / dbg_nn (B,P,M) optional (NULL in production):
 *   the transformed model points and the 0-based nearest-target index, for parity tests. */
int df_loss_forward(const float* pred_r, const float* pred_t, const float* pred_c, const float* target,
                    const float* model_points, const float* hyp_points, const float* points,
                    const int64_t* idx, unsigned long long sym_mask, int allow_sym, float w,
                    int B, int P, int M, int N,
                    float* dis_all, float* sum_u, float* sum_um, float* loss, float* dis_sel,
                    int64_t* which, float* new_points, float* new_target, unsigned int* tickets,
                    float* dbg_pred, int* dbg_nn, void* stream);

/* Gradients of  sum_b g_loss[b]*loss[b] + g_dis[b]*dis_sel[b]  w.r.t. pred_r / pred_t / pred_c
 * (what autograd derives from lib/loss.py; the kNN indices are constants).  g_loss / g_dis may be
 * NULL (zero).  pred_c == NULL selects the refiner-loss form. */
int df_loss_backward(const float* pred_r, const float* pred_c, const float* dis_all, const float* sum_u,
                     const float* sum_um, const int64_t* which, const float* g_loss, const float* g_dis,
                     float w, int B, int P, float* g_pred_r, float* g_pred_t, float* g_pred_c, void* stream);

/* ---- K1 / K2: dense-fusion head and refiner building blocks -------------------------------------
 * C[m,n] = act(sum_k A[m,k] W[n,k] + bias[n])  (Conv1d k=1 / Linear of lib/network.py:42-49, :77-91,
 * :139-146, :176-183).  W is the torch (out,in) weight.  K % 16 == 0, N % 4 == 0.
 *   bias_crop_stride != 0 : bias row = bias + (m / rows_per_crop) * bias_crop_stride (folded global feature)
 *   groups > 1            : blockIdx.z batches block-diagonal layers (r/t/c towers); *_group_stride are
 *                           element offsets between groups for A (columns), W, bias and C (columns)
 *   pool_partial != NULL  : C is not stored; (crops, tiles, N) column sums of act(.) over crop-aligned
 *                           128-row tiles are written instead (tiles = ceil(rows_per_crop / 128))
 * df_gemm_fp32: exact fp32 FFMA arithmetic.  df_gemm_tc: tcgen05/TMEM tensor-core arithmetic,
 * precision 1 = 3xTF32 (error-compensated, fp32-parity), 2 = single-pass TF32 (looser bound). */
int df_gemm_fp32(const float* A, int lda, const float* W, int ldw, const float* bias, int bias_crop_stride,
                 float* C, int ldc, int M, int N, int K, int relu, int rows_per_crop, int groups,
                 long long a_group_stride, long long w_group_stride, long long bias_group_stride,
                 long long c_group_stride, float* pool_partial, void* stream);
int df_gemm_rows_per_pool_tile(void);

/* Tensor-core form of the same contract.  W_hi / W_lo: the weight split once by df_split_tf32
 * (hi = w & 0xffffe000, lo = w - hi), group g's rows at g*N (stacked, pitch ldw).  precision 1 = 3xTF32
 * (needs W_lo), 2 = single-pass TF32 (W_lo may be NULL).  K % 32 == 0.  Activations A and outputs stay plain fp32:
 * the hi / lo split of A happens in registers on the way into TMEM.
 * variant 0 = auto (6, falling back to 4 for layouts a 2-D TMA tensor cannot address; env DF_TC_VARIANT overrides);
 *   1/2 = one tile per CTA, N tile 128/256, A through TMEM;  3 = A through shared memory (bring-up);
 *   4 = persistent warp-specialised kernel, stagers read A from global memory;
 *   5 = persistent, A by TMA (128B-swizzled tiles), 128x128 tiles;
 *   6/7 = 5 on CTA pairs: tcgen05.mma.cta_group::2, 256-row tiles, each CTA stages half of the weight tile. */
int df_gemm_tc(const float* A, int lda, const float* W_hi, const float* W_lo, int ldw, const float* bias,
               int bias_crop_stride, float* C, int ldc, int M, int N, int K, int relu, int rows_per_crop,
               int groups, long long a_group_stride, long long bias_group_stride, long long c_group_stride,
               float* pool_partial, int precision, int variant, void* stream);
int df_split_tf32(const float* x, float* hi, float* lo, long long n, void* stream);
/* precision 3 ("hybrid", generation-2 kernels / df_conv_tc only): D += A_hi W_hi in TF32 plus the two correction terms
 * A_lo W and A W_lo in bf16 (operands ~2^-11 of the main term, so 8 mantissa bits keep them to 2^-20 of the result) --
 * 8 MMA instructions per 32-wide k-block instead of 12, same fp32-parity bound.  W_lo then points to the packed pair
 * tensor made by df_pack_bf16_pairs: per row and k-block 64 bf16 = [bf16(w) x32 | bf16(w - tf32(w)) x32]; needs ldw == K. */
int df_pack_bf16_pairs(const float* w, void* out, long long rows, int K, void* stream);
/* precision 4 ("hybrid16", same kernels): the main term on fp16 operands (11 significant bits like TF32, but K = 16 per
 * instruction: half the tensor time) plus the same two bf16 correction terms -- 6 instructions per k-block.  fp16 saturates at
 * +-65504 and flushes below 6e-8; what it drops is carried exactly by the correction terms (a - fp16(a) in bf16), so the bound
 * stays 2^-20 per product.  W_hi / W_lo then point to the two packed tensors made by df_pack_f16_pairs: per row and k-block
 * t1 = [fp16(w) x32 | bf16(w) x32] (the byte size of the weight), t2 = bf16(w - fp16(w)) row-major (half of it); needs ldw == K. */
int df_pack_f16_pairs(const float* w, void* t1, void* t2, long long rows, int K, void* stream);
/* precision 6 ("hybrid16s", CTA-pair kernel): EVERY term on fp16 operands -- x = fp16(x s) + fp16(x s - fp16(x s)) for both operands,
 * D += A_hi W_hi + A_lo W_hi + A_hi W_lo, 6 instructions per k-block like precision 4 but only TWO 16-bit planes per operand: 4 instead
 * of 6 bytes per weight element through the SM's fabric port (what bounds this kernel) and half the TMEM per A stage.
 * The power-of-two scales s keep the remainder planes out of fp16's subnormals: the weight's is chosen by df_pack_f16s from the tensor's
 * own maximum (planes = per row and k-block [fp16 hi x32 | fp16 lo x32], the byte size of the weight; scale = 4 floats: 1/s, s, and
 * scratch), the activation's is passed as log2 in bits 16..23 of `precision` (signed; parity needs the entries that carry the dot
 * product inside [2^-8, 2^15] after scaling: 22 significant bits above 0.25, an absolute error of 2^-25 below).  W_hi = planes, W_lo = scale;
 * needs ldw == K. */
int df_pack_f16s(const float* w, void* planes, float* scale, long long rows, int K, void* stream);
/* Debug aid of the tensor-core kernel (env DF_TC_DBG bit 256): clock64() timeline of cluster 0's first 96 k-blocks of the last launch,
 * [17 events][96]: per k-block producer slot free / TMA issued, issuer stage landed / A handed over / MMAs committed, stager bytes landed /
 * split done / TMEM slot free / A handed over; per accumulation run epilogue waits for / has / has drained the accumulator, issuer waits
 * for / has a free accumulator; per chunk of epilogue warp 10 accumulator in registers / transposed / stored.  count <= 1632 uint64
 * copied to host memory. */
int df_tc_trace_read(unsigned long long* host_out, int count);
/* Accumulation runs: long k loops are cut into runs on fresh accumulators, summed in fp32 through C (the tensor core
 * truncates while accumulating; the bias grows with the number of chained instructions).  Default 216 MMA instructions per
 * run; bits 8..15 of `precision` (df_gemm_tc and df_conv_tc) select another length in units of 12 instructions -- the
 * training path passes 9 (108): weight gradients sum that bias over all pixels. */
/* torch convolution weight (Cout,Cin,kh,kw) -> (rows, taps*cols) tap-major GEMM operand split for the tensor-core modes in
 * one pass: hi always, lo (3xTF32) and / or pairs (hybrid).  rotate = 1: the data-gradient kernel (rows = Cin, taps reversed). */
int df_pack_conv_weight(const float* w, float* hi, float* lo, void* pairs, int Cout, int Cin, int taps, int rotate, void* stream);
/* ... and for precision 4: the two packed tensors of df_pack_f16_pairs, from the torch convolution weight in one pass. */
int df_pack_conv_weight16(const float* w, void* t1, void* t2, int Cout, int Cin, int taps, int rotate, void* stream);
/* Weight gradient of a stride-1 3x3 (padding == dilation) / 1x1 convolution (training; replaces cuDNN's wgrad):
 * dW (Cout, taps*Cin) tap-major = sum over pixels of dY (B,H,W,Cout; pitch ldy) x shifted X (B,H,W,Cin; pitch ldx), one 3xTF32
 * GEMM over the zero-padded, flattened pixel axis.  Cin % 64 == 0; `scratch` holds df_conv_wgrad_scratch_floats() floats. */
long long df_conv_wgrad_scratch_floats(int B, int H, int W, int Cin, int Cout, int taps, int dilation);
int df_conv_wgrad_tc(const float* X, int ldx, const float* dY, int ldy, int B, int H, int W, int Cin, int Cout, int taps,
                     int dilation, float* scratch, float* dW, void* stream);

/* emb[b,c,n] = feat[b,c,choose[b,n]]  (lib/network.py:98-102).  feat is addressed with explicit element
 * strides so NCHW and channels-last encoders both work.  emb_pm (B*N,32) point-major and/or emb_cm
 * (B,32,N) reference layout; either may be NULL. */
int df_gather_embedding(const float* feat, const int64_t* choose, float* emb_pm, float* emb_cm,
                        long long stride_b, long long stride_c, long long stride_pix, int B, int N, int HW,
                        void* stream);

/* out[row, 0:64] = relu(W (64,3) . x[row] + bias)   -- conv1 of PoseNetFeat / PoseRefineNetFeat. */
int df_xyz_conv(const float* x, const float* W, const float* bias, float* out, int ldo, long long rows,
                void* stream);

/* g[crop, c] = (sum_tiles partial[crop, tile, c]) / rows_per_crop  -- AvgPool1d(num_points). */
int df_pool_finish(const float* partial, float* g, int crops, int tiles, int channels, int rows_per_crop,
                   void* stream);

/* Last tower layer for the selected object only (lib/network.py:118-130 / :198-204):
 * h (rows, ldh) holds the 128-channel r | t | c branch activations at columns 0 / 128 / 256.
 * out_r (rows,4), out_t (rows,3), out_c (rows) = sigmoid(.).  Wc == NULL: refiner (no confidence). */
int df_select_out(const float* h, int ldh, const float* Wr, const float* br, const float* Wt, const float* bt,
                  const float* Wc, const float* bc, const int64_t* obj, int rows_per_crop, int num_obj,
                  long long rows, float* out_r, float* out_t, float* out_c, void* stream);

/* ---- backward of K1 / K2 and the optimiser step (training, config C4; tools/train.py:152-169) ------------
 * What autograd derives from lib/network.py, as explicit kernels (exact fp32):
 *   df_gemm_dgrad_fp32 : dX (+)= (dY . W) (*) [relu_mask > 0]; Wt is the TRANSPOSED layer weight (K_out, N_red) so the
 *                        forward kernel is reused; relu_mask has dX's leading dimension / group stride.
 *   df_gemm_wgrad_fp32 : partial[s][g][n,k] = sum over row slice s of dY[m, g*dy_gs+n] X[m, g*x_gs+k]
 *   df_reduce_partials : out (+)= sum_s partial[s]  (fixed order -> deterministic)
 *   df_colsum_rows     : out[group, c] (+)= sum_{r < rows_per_group} X[group*rows_per_group + r, c]
 *   df_relu_mask_inplace, df_pool_backward, df_select_out_backward, df_gather_embedding_backward: see backward.cu
 *   df_adam_step       : torch.optim.Adam (no weight decay / amsgrad) on a flat parameter arena */
int df_gemm_dgrad_fp32(const float* dY, int ldy, const float* Wt, int ldw, float* dX, int ldx, int M, int N, int K,
                       int groups, long long dy_group_stride, long long w_group_stride, long long dx_group_stride,
                       const float* relu_mask, int accumulate, void* stream);
int df_gemm_wgrad_fp32(const float* dY, int ldy, const float* X, int ldx, float* partial, int M, int N, int K,
                       int groups, int splits, long long dy_group_stride, long long x_group_stride, void* stream);
int df_reduce_partials(const float* partial, int splits, long long count, float* out, int accumulate, void* stream);
int df_colsum_rows(const float* X, int ldx, int rows_per_group, int groups, int C, float* out, int accumulate,
                   void* stream);
int df_relu_mask_inplace(float* d, const float* act, int ld, int cols, long long rows, void* stream);
int df_pool_backward(const float* dg, const float* h, float* dh, int rows_per_crop, int C, long long rows, void* stream);
#define DF_SELECT_SPLITS 8      /* row slices per crop: blk is (crops, DF_SELECT_SPLITS, 8, 128), bsum (crops, DF_SELECT_SPLITS, 8) */
int df_select_out_backward(const float* g_r, const float* g_t, const float* g_c, const float* out_c, const float* h,
                           int ldh, const float* Wr, const float* Wt, const float* Wc, const int64_t* obj,
                           int rows_per_crop, int num_obj, long long rows, float* dh, float* gz, float* blk,
                           float* bsum, float* dWr, float* dbr, float* dWt, float* dbt, float* dWc, float* dbc,
                           void* stream);
int df_gather_embedding_backward(const float* demb, const int64_t* choose, float* dfeat, long long stride_b,
                                 long long stride_c, long long stride_pix, int B, int N, int HW, void* stream);
int df_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                 float beta1, float beta2, float eps, int step, void* stream);
/* Same update with the step number kept on the device: uses *step_counter + 1 and then increments it, so the launch
 * can be captured in a CUDA graph and replayed. */
int df_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                     float beta1, float beta2, float eps, int* step_counter, void* stream);

/* Multiply-adds df_conv_tc executes on real output pixels for this geometry (dense count minus the all-padding taps its pixel
 * patches skip); host-side arithmetic only, for the time-weighted roofline of bench.py. */
long long df_conv_tc_macs(int B, int H, int W, int Cin, int Cout, int taps, int dilation);

/* The tile plan df_gemm_tc takes for a hybrid16s GEMM (M x K) . (groups*N x K)^T on `clusters` CTA pairs (74 on a B200).  Host-side
 * arithmetic only -- the launcher's own width / operand-placement rules, exposed so that tests can pin them (tests/test_conv_schedule.py).
 *   pooled != 0: the column-sum epilogue of conv6 (lib/network.py:66-68), tiles aligned to rows_per_crop.
 *   out (6 ints): [0] tile width in accumulator columns, [1] 1 = activation planes staged in shared memory (two 256-column
 *   accumulators) / 0 = in TMEM, [2] accumulators a tile alternates between, [3] tiles, [4] rounds of the persistent kernel,
 *   [5] accumulation runs per tile.  Returns 0, DF_ERR_ARG or DF_ERR_UNSUPPORTED. */
int df_gemm_tc_plan(int M, int N, int K, int groups, int pooled, int rows_per_crop, int clusters, int* out);

/* The work schedule df_conv_tc uses for a hybrid16s 3x3 convolution of this geometry on `clusters` CTA pairs (256-wide tiles): long-K
 * convolutions whose tile count does not fill the last round of the persistent kernel are cut into contiguous per-cluster ranges
 * of (tile, accumulation run) units of equal weight instead of whole tiles dealt round robin.  Host-side arithmetic only.
 *   out (5 + 4*clusters ints): [0] k-blocks of the busiest cluster under round robin, [1] under the balanced schedule, [2] tiles,
 *   [3] k-blocks per accumulation run, [4] k-blocks of a full tile, then per cluster (first tile, first run, last tile, end run or 0x7fff).
 * Returns 1: balanced schedule taken; 0: round robin kept; < 0: DF_ERR_ARG. */
int df_conv_tc_schedule(int B, int H, int W, int Cin, int Cout, int dilation, int clusters, int* out);

/* ---- colour encoder on the tensor cores (lib/extractors.py:78-124, lib/pspnet.py:7-77) ----
 * Activations are NHWC.  df_conv_tc: 3x3 (stride 1, padding == dilation) or 1x1 convolution as an implicit GEMM on the
 * CTA-pair tcgen05 kernel -- the A operand is a 4-D TMA box shifted by the tap, its out-of-image part zero-filled.
 *   X (B,H,W,Cin) pixel pitch ldx; W_hi/W_lo (Cout, taps*Cin) tap-major repack of the torch (Cout,Cin,kh,kw) weight, split by
 *   df_split_tf32; act 0 none / 1 ReLU / 2 PReLU(prelu[0]); residual (pixel pitch ldr) is added before the activation;
 *   Y pixel pitch ldy (>= Cout: the output may be a channel slice of a wider buffer).  precision as df_gemm_tc.
 * df_enc_*: the HBM-bound glue (im2col for the three stride-2 layers, pooling, resizing into a channel slice, log-softmax). */
int df_conv_tc(const float* X, int B, int H, int W, int Cin, int ldx, const float* W_hi, const float* W_lo, int taps,
               int dilation, const float* bias, const float* residual, int ldr, const float* prelu, int act, float* Y,
               int ldy, int Cout, int precision, void* stream);
int df_enc_im2col_conv1(const float* img, float* A, int B, int H, int W, int ldk, void* stream);
/* conv1 (7x7 / stride 2 / pad 3, 3 -> Cout <= 64 channels, lib/extractors.py:82) WITHOUT the patch matrix: a GEMM on the tcgen05 kernel whose A
   operand the kernel gathers from the NCHW image itself (hybrid16s arithmetic).  planes / scale = df_pack_f16s of the (Cout, 160) weight
   matrix, column c*49 + ky*7 + kx, zero-padded from 147.  Y (B*Ho*Wo, Cout) NHWC with pixel pitch ldy; relu != 0 fuses the ReLU. */
int df_enc_conv1_tc(const float* img, int B, int H, int W, const void* planes, const float* scale, float* Y, int ldy, int Cout,
                    int relu, void* stream);
int df_enc_maxpool(const float* in, float* out, int B, int H, int W, int C, void* stream);
int df_enc_im2col_s2(const float* in, float* A, int B, int H, int W, int C, void* stream);
int df_enc_col2im_s2(const float* dA, float* dx, int B, int H, int W, int C, void* stream);   /* transpose of df_enc_im2col_s2 (training) */
int df_enc_adaptive_avgpool(const float* in, int ldi, float* out, int B, int H, int W, int C, int S, void* stream);
/* Folded pyramid (lib/pspnet.py:17-24): the four adaptive average pools (1,2,3,6) of `in` (B,H,W,C; pixel pitch ldi) in one
   pass -> out (50 B, C) stage-major (rows [B x 1 | B x 4 | B x 9 | B x 36]); and out[b,y,x,:] = Y[b,:] + sum_s bilinear(Y cells of s) resized to
   H x W (align_corners = False) -- the pooled branches after their (bottleneck . stage) products, summed at full resolution. */
int df_enc_pyramid_pool(const float* in, int ldi, float* out, int B, int H, int W, int C, void* stream);
int df_enc_pyramid_sum(const float* Y, float* out, int ldo, int B, int H, int W, int C, void* stream);
/* Decoder stage (lib/pspnet.py:27-37: x2 bilinear resize, align_corners; 3x3 convolution; PReLU) evaluated at the low
   resolution: Z (B,h,w,>=9*C; pixel pitch ldz) = x . [W_tap0 .. W_tap8] (one GEMM, tap-major columns); this call sums the nine
   shifted bilinear samples of Z per output pixel, adds `bias` (may be NULL) and applies PReLU(prelu[0]) -> out (B,2h,2w,C). */
int df_enc_upconv_finish(const float* Z, int ldz, const float* bias, const float* prelu, float* out, int ldo, int B, int h, int w,
                         int C, void* stream);
int df_enc_upsample(const float* in, int ldi, float* out, int ldo, int B, int hin, int win, int hout, int wout, int C,
                    int align_corners, void* stream);
int df_enc_upsample_backward(const float* gout, int ldo, float* gin, int ldi, int B, int hin, int win, int hout, int wout,
                             int C, int align_corners, void* stream);      /* gather form of the resize's transpose */
int df_enc_log_softmax32(float* x, long long pixels, void* stream);

/* ---- element-wise / pooling kernels of the encoder's TRAINING graph (autograd of lib/extractors.py:78-124, lib/pspnet.py:7-77) ----
 * NHWC fp32 (torch channels_last storage), C % 4 == 0, gather form (deterministic).
 *   df_ew_relu_mask             out = d * [act > 0] (pixel pitch ld, `cols` channels)           -- backward of a fused ReLU epilogue
 *   df_ew_maxpool_backward      3x3 / stride 2 / pad 1; window maxima recomputed from x, first maximum wins (ATen's rule)
 *   df_ew_pyramid_pool_backward the four adaptive average pools (1,2,3,6) at once; dpool (50 B, C) stage-major as df_enc_pyramid_pool
 *   df_ew_log_softmax32_backward dx = dy - exp(y) sum_c dy
 *   df_ew_prelu / _backward     one slope (nn.PReLU()); dslope = sum dy x [x <= 0], fixed-order two-stage reduction
 *                               (scratch: df_ew_prelu_scratch_floats() floats)
 *   df_ew_dropout_mask          Dropout2d decisions, one per (sample, channel): mask[i] in {0, 1/(1-p)} from a counter-based hash of
 *                               state = {seed, counter} (device, two uint64); the launch advances the counter (graph replays differ)
 *   df_ew_scale_bc              y[b,pixel,c] = x[b,pixel,c] * mask[b,c]
 *   df_ew_copy2d                pitched row copy (a channel slice of an NHWC buffer: the pyramid concat is written slice by slice)
 *   df_ew_add                   out = a + b */
int df_ew_relu_mask(const float* d, const float* act, float* out, int ld, int cols, long long rows, void* stream);
int df_ew_maxpool_backward(const float* x, const float* gout, float* gin, int B, int H, int W, int C, void* stream);
int df_ew_pyramid_pool_backward(const float* dpool, float* dx, int ldo, int B, int H, int W, int C, void* stream);
int df_ew_log_softmax32_backward(const float* y, const float* dy, float* dx, long long pixels, void* stream);
int df_ew_prelu(const float* x, const float* slope, float* y, long long n, void* stream);
int df_ew_prelu_scratch_floats(void);
int df_ew_prelu_backward(const float* x, const float* slope, const float* dy, float* dx, float* dslope, float* scratch, long long n,
                         void* stream);
int df_ew_dropout_mask(float* mask, int n, float p, unsigned long long* state, void* stream);
int df_ew_scale_bc(const float* x, const float* mask, float* y, int B, long long HW, int C, void* stream);
int df_ew_copy2d(const float* src, int lds, float* dst, int ldd, long long rows, int cols, void* stream);
int df_ew_add(const float* a, const float* b, float* out, long long n, void* stream);
/* Sparse last decoder stage: the 3x3 patches of the x2-upsampled (align_corners) map `in` (B,h,w,C) around the N chosen
 * pixels of every crop (choose (B,N), indices into the (2h x 2w) image), A (B*N, 9*C) tap-major; zero outside the image. */
int df_enc_gather_up_patches(const float* in, const int64_t* choose, float* A, int B, int N, int h, int w, int C,
                             void* stream);

/* ---- input preparation on the device (tools/eval_ycb.py:147-190) ---------------------
 * One bucket of b objects whose snapped boxes share the size (h,w).  rgb (F,H,W,3) uint8, depth (F,H,W) fp32 (raw sensor
 * units), label (F,H,W) int32, meta (b,6) int32 = frame, item id, rmin, rmax, cmin, cmax (host arrays; cam / mean_std are
 * HOST pointers: cx cy fx fy scale / mean[3] std[3]).  Outputs: img (b,3,h,w) normalised, choose (b,N) int64 indices into
 * the box, cloud (b,N,3), count (b) = masked pixels found (0: object lost, outputs zero). */
int df_build_crops(const uint8_t* rgb, const float* depth, const int* label, const int* meta, int b, int H, int W, int h, int w,
                   int N, const float* cam, const float* mean_std, unsigned seed, float* out_img, int64_t* out_choose,
                   float* out_cloud, int* out_count, void* stream);

/* ---- encoder helper -----------------------------------------------------------------------------
 * NCHW bilinear up-sampling (lib/pspnet.py:20-23 F.upsample(size=...), :30-34 nn.Upsample(scale_factor=2,
 * align_corners=True)): in (planes, hin, win) -> out (planes, hout, wout), planes = batch*channels. */
int df_upsample_bilinear(const float* in, float* out, long long planes, int hin, int win, int hout, int wout,
                         int align_corners, void* stream);

/* ---- K5: on-device pose state for the refinement loop (tools/eval_ycb.py:193-233) ----------------
 * pose (B,7) float64 [qw qx qy qz tx ty tz]. */
int df_select_pose(const float* pred_r, const float* pred_t, const float* pred_c, const float* points,
                   int B, int N, double* pose, int64_t* which, void* stream);
int df_cloud_transform(const float* cloud, const double* pose, float* out, int B, int N, void* stream);
int df_pose_compose(double* pose, const float* r2, const float* t2, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DENSEFUSION_B200_H */
