/* denet_b200 C-ABI: B200 (sm_100a) kernels for the DeNet training hot path.
 *
 * This is the drop-in boundary.  Every entry point takes raw DEVICE pointers, sizes and a
 * cudaStream_t, enqueues work on that stream and returns without synchronising.  Nothing here allocates:
 * outputs and workspaces are owned by the caller.  Return value: 0 on success, <0 on error; the message of
 * the last error on the calling host thread is returned by denet_last_error().  Calls on distinct streams
 * are thread-safe.  No entry point has a CPU fallback.
 *
 * Layout convention: activations are NHWC ("pixel-major"): element (n, h, w, c) of a tensor with pixel pitch
 * `ld` lives at ((n*H + h)*W + w)*ld + c.  Dtypes are DENET_F32 or DENET_BF16.  Filters keep the reference's
 * layout (Cout, Cin, R, S) fp32 *for the true (flipped) convolution* that Theano's conv2d computes
 * (reference denet/layer/convolution.py:83), so reference checkpoints load unchanged.  Host-facing tensors
 * that the reference exposes in NCHW (corner_pr, targets, images) keep that layout here.
 *
 * Each function names the reference interface it replaces (paths relative to the reference repository).
 */
#ifndef DENET_B200_H
#define DENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DENET_ABI_VERSION 4

#define DENET_F32 0
#define DENET_BF16 1

#define DENET_ERR_ARG (-1)
#define DENET_ERR_CUDA (-2)

const char* denet_last_error(void);
int denet_abi_version(void);
/* number of CUDA kernels this library has enqueued in this process so far (all streams, all threads) */
long long denet_launch_count(void);
/* A/B switch: 1 launches the CTA-pair convolution and the batch-norm kernels with programmatic stream serialisation (a
 * kernel's prologue overlaps its predecessor's tail; griddepcontrol.wait orders the data), 0 (default: measured faster
 * inside CUDA graphs) fully serialised.  Results are identical. */
int denet_set_pdl(int on);

/* ------------------------------------------------------------------------------------------------ convolution
 * Replaces tensor.nnet.conv2d / its autodiff gradients, i.e. the cuDNN fprop / bwd-data / bwd-filter calls
 * (denet/layer/convolution.py:76-92).  tcgen05 tensor-core implicit GEMM, bf16 operands,
 * fp32 accumulate.  Passing the *_lo pointers selects the error-compensated bf16x3 split (fp32-parity mode);
 * passing NULL selects plain bf16 (throughput mode).
 */

/* Reference filters -> GEMM operand.  mode 0: fprop operand [Cout][R*S][pad64(Cin)];
 * mode 1: dgrad operand [Cin][R*S][pad64(Cout)].  b_lo may be NULL. */
int denet_conv_weight_prep(const float* w, int Cout, int Cin, int R, int S, int mode, void* b_hi, void* b_lo,
                           cudaStream_t stream);

/* The operands of ALL conv layers in one launch.  `entries` is a device array of denet_weight_prep_entry_bytes()-sized
 * records {const float* w; bf16* hi; bf16* lo (or NULL); long long total; int Cout, Cin, R, S, mode, Cp} with mode
 * 0 / 1 as above, 2 = row-folded stem operand (Cp = padded channels) and 3 = dgrad operand of one parity class of a
 * strided convolution, [Cin][Rc*Sc][pad64(Cout)] with tap (tr, ts) = filter element (r0 + sh*tr, s0 + sw*ts) and
 * Cp = r0 | s0<<4 | sh<<8 | sw<<12 | Rc<<16 | Sc<<20; `total` counts work items = operand rows x
 * padded K columns (an item covers that column for all filter taps); block i prepares items
 * [block_offset[i], +denet_weight_prep_chunk()) of operand block_entry[i]. */
int denet_weight_prep_entry_bytes(void);
int denet_weight_prep_chunk(void);
int denet_conv_weight_prep_multi(const void* entries, const int* block_entry, const long long* block_offset,
                                 int nblocks, cudaStream_t stream);

/* fp32 -> bf16 hi (+ lo = bf16(x - hi)) operand split.  lo may be NULL. */
int denet_split_bf16(const float* x, void* hi, void* lo, long long n, cudaStream_t stream);

/* R x S correlation (stride stride_h x stride_w) of an NHWC bf16 tensor with a prepared operand:
 *   y[n,h,w,co] = sum_{r,s,ci} x[n, h*stride_h+r-pad_h, w*stride_w+s-pad_w, ci] * B[co][r*S+s][ci]  (zero outside)
 * followed by the fused epilogue  (+bias[co]) (+residual) (relu)  and optional per-channel sum / sum-of-squares
 * accumulation of the conv output (before residual/relu) for batch-norm statistics.
 * With mode-0 operands this is the reference fprop; with mode-1 operands, Cin/Cout swapped, stride 1 and
 * pad = R-1-pad it is the reference dgrad (of a stride-1 convolution, or of a strided one after denet_dilate). */
int denet_conv2d_fprop(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin, long long ldx,
                       const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w, int stride_h,
                       int stride_w, void* y, int y_dtype, long long ldy, int Ho, int Wo, const float* bias,
                       const void* residual, int relu, float* stat_sum, float* stat_sqsum, cudaStream_t stream);

/* denet_conv2d_fprop (stride 1, no bias / relu / statistics) whose output pixel (n, h, w) is written to pixel
 * (n, h*osh + ooh, w*osw + oow) of a tensor with Hf x Wf pixels per image (`residual`, if given, is read at the same
 * place).  The data gradient of a convolution with stride s is s*s such launches on the UNDILATED dy, one per parity
 * class (a, b) of the output pixel, each a small stride-1 correlation over the filter taps r = r0 + s*tr that reach that
 * class (operand from denet_conv_weight_prep_multi mode 3): no zero-insertion buffer, no multiplications by zero. */
int denet_conv2d_fprop_scatter(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin, long long ldx,
                               const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w, void* y,
                               int y_dtype, long long ldy, int Ho, int Wo, int Hf, int Wf, int osh, int osw, int ooh,
                               int oow, const void* residual, cudaStream_t stream);

/* Data gradient of a stride-1 convolution whose INPUT was produced by a batch-norm(+ReLU) layer, with the first pass of
 * that layer's backward fused into the epilogue (replaces cuDNN bwd-data followed by the first half of cuDNN BN grad,
 * denet/layer/convolution.py:83 + batch_norm.py:47-53 / batch_norm_relu.py:50-54).  Arguments up to `residual` as
 * denet_conv2d_fprop with mode-1 operands (dz = dgrad (+ residual)).  The epilogue masks dz by the layer's ReLU mask
 * (bn_relu: bn_yout > 0 when given, else recomputed from bn_x with the forward's expression), stores the MASKED
 * gradient dz' and accumulates sum_dz[c] += sum dz', sum_dz_xhat[c] += sum dz' * (bn_x - mean) * invstd with fp32
 * atomics (both buffers zeroed by the caller).  denet_bn_backward_sums finishes the layer's backward from them. */
int denet_conv2d_dgrad_bnbwd(const void* dy_hi, const void* dy_lo, int N, int Hi, int Wi, int Cin, long long lddy,
                             const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w, void* dz,
                             int dz_dtype, long long lddz, int Ho, int Wo, const void* residual, const void* bn_x,
                             const void* bn_yout, const float* bn_mean, const float* bn_invstd, const float* bn_gamma,
                             const float* bn_beta, int bn_relu, float* sum_dz, float* sum_dz_xhat, cudaStream_t stream);

/* Kernel selection of conv2d_fprop / conv2d_rowfold_fprop (profiling / A-B tests): bit 0 (default on) lets stride-1
 * multi-tap filters and the row-folded stem use the tap-group kernel (one halo'd input box per channel chunk serves
 * all filter taps), bit 1 (default on) keeps filters that fit resident in shared memory; 0 = one box per tap. */
int denet_conv2d_fprop_set_mode(int mode);
/* Profiling aid: device buffer of 3*64*4 int64 in which CTA 0 of the tap-group kernel records per-tile clock64()
 * stamps of its producer / MMA / epilogue roles (NULL switches it off). */
int denet_conv2d_fprop_set_timeline(void* buf);

/* Filter gradient in the reference layout:
 *   dw[co][ci][R-1-r][S-1-s] (+)= sum_pixels dy[p,co] * x[p*stride+(r,s)-pad, ci].
 * Split-K partial sums go through `workspace` (size from denet_conv2d_wgrad_workspace) and are reduced in a
 * fixed order (deterministic). */
size_t denet_conv2d_wgrad_workspace(int N, int Ho, int Wo, int Cout, int Cin, int R, int S);
int denet_conv2d_wgrad(const void* dy_hi, const void* dy_lo, int N, int Ho, int Wo, int Cout, long long lddy,
                       const void* x_hi, const void* x_lo, int Hi, int Wi, int Cin, long long ldx, int R, int S,
                       int pad_h, int pad_w, int stride_h, int stride_w, float* dw, int accumulate, float* workspace,
                       size_t workspace_bytes, cudaStream_t stream);

/* Deferred reduction: conv2d_wgrad / conv2d_rowfold_wgrad called with dw = NULL leave their split-K partial sums in
 * `workspace` ([splits][Cout][R*S][pad4(Cin)] fp32, resp. [splits][Cout][R][64] for the row-folded stem; the split
 * count comes from denet_conv2d_wgrad_splits / denet_conv2d_rowfold_wgrad_splits) and denet_wgrad_reduce_multi
 * reduces the partials of MANY layers in one launch, in the same fixed order and into the same reference layout.
 * `entries` is a device array of denet_wgrad_reduce_entry_bytes()-sized records {const float* ws; float* dw;
 * long long total (= Cout*Cin*R*S); int splits, Cout, Cin, R, S, ldws, mode (0 generic, 1 row-folded), Cp,
 * accumulate, pad}.  Work of block i on entry block_entry[i]: for 1x1 filters and the row-folded stem, the elements
 * [block_offset[i], +denet_wgrad_reduce_chunk()); for multi-tap filters (mode 0, R*S > 1) the items
 * [block_offset[i], +denet_wgrad_reduce_items()) out of Cout * ceil(Cin/G), G = denet_wgrad_reduce_group(), an item being
 * one output channel times G consecutive input channels times all taps (transposed through shared memory so that both
 * sides are coalesced; the 4-float pad columns of a workspace row may be read, never written out). */
int denet_conv2d_wgrad_splits(int N, int Ho, int Wo, int Cout, int Cin, int R, int S, int stride_h, int stride_w);
int denet_conv2d_rowfold_wgrad_splits(int N, int Ho, int Wo, int Cout, int R, int stride_h);
int denet_wgrad_reduce_entry_bytes(void);
int denet_wgrad_reduce_chunk(void);
int denet_wgrad_reduce_items(void);
int denet_wgrad_reduce_group(void);   /* input channels per work item of the tiled (multi-tap) path */
int denet_wgrad_reduce_multi(const void* entries, const int* block_entry, const long long* block_offset, int nblocks,
                             cudaStream_t stream);

/* Kernel selection of conv2d_wgrad (profiling / A-B tests): row_shared = 1 (default) lets stride-1 S>1 layers use the
 * row-shared kernel (one halo'd X box serves the S taps of a filter row), 0 forces one tap per tile. */
int denet_conv2d_wgrad_set_mode(int row_shared);

/* Zero-insertion upsampling y[n, h*sh, w*sw, :] = x[n,h,w,:] (zero elsewhere), (Hd, Wd) = extent of y: turns the
 * data gradient of a strided convolution into a stride-1 correlation (cuDNN bwd-data of convolution.py:83). */
int denet_dilate(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int sh, int sw, void* y, int Hd,
                 int Wd, long long ldy, cudaStream_t stream);
/* the same plus `add` (a tensor shaped and pitched like y, or NULL): y = dilate(x) + add in one pass - the data gradient
 * of a strided 1x1 projection meeting the gradient of the block's other branch (denet/layer/resnet.py:96-113). */
int denet_dilate_add(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int sh, int sw,
                     const void* add, void* y, int Hd, int Wd, long long ldy, cudaStream_t stream);

/* Explicit im2col / col2im (column order k = (r*S+s)*C + c) for the 3-channel stem, whose 147-element patch rows
 * are too ragged for the TMA implicit-GEMM path; the GEMM then runs as a 1x1 convolution over the column matrix.
 * weight_to_im2col / weight_grad_from_im2col permute between (Cout,Cin,R,S) [true convolution] and (Cout, R*S*Cin). */
int denet_im2col(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int R, int S, int sh, int sw,
                 int ph, int pw, int Ho, int Wo, void* col, long long ldc, cudaStream_t stream);
int denet_col2im(const void* dcol, int dtype, long long ldc, int N, int H, int W, int C, long long ldx, int R, int S,
                 int sh, int sw, int ph, int pw, int Ho, int Wo, void* dx, cudaStream_t stream);
int denet_weight_to_im2col(const float* w, int Cout, int Cin, int R, int S, float* w2, cudaStream_t stream);
int denet_weight_grad_from_im2col(const float* dw2, int Cout, int Cin, int R, int S, float* dw, int accumulate,
                                  cudaStream_t stream);

/* Row-folded convolution for inputs with very few channels (the 3-channel image stem C.B[64,7,2] of
 * examples/resnet34-imagenet.sh:7; same reference op as above, denet/layer/convolution.py:76-92).  The image lives
 * zero-padded in an NHWC buffer (N, Hp, Wp, Cp) bf16 with Cp = 4 or 8 channels (stride_w*Cp and Wp*Cp multiples of 8,
 * S*Cp <= 64), image pixel (h, w) at (h + pad_h, w + pad_w); TMA reads overlapping windows of it, one filter row per
 * K chunk, so no im2col matrix is ever materialised.  nchw_to_padded_nhwc fills the interior (borders must already be
 * zero); weight_prep_rowfold lays the filters out as [Cout][R][64]; rowfold_wgrad returns dw in the reference layout. */
int denet_nchw_to_padded_nhwc(const float* x, int N, int C, int H, int W, int Cp, int pad_h, int pad_w, int Hp, int Wp,
                              void* y_hi, void* y_lo, cudaStream_t stream);
int denet_conv_weight_prep_rowfold(const float* w, int Cout, int Cin, int R, int S, int Cp, void* b_hi, void* b_lo,
                                   cudaStream_t stream);
int denet_conv2d_rowfold_fprop(const void* x_hi, const void* x_lo, int N, int Hp, int Wp, int Cp, int Cin,
                               const void* b_hi, const void* b_lo, int Cout, int R, int S, int stride_h, int stride_w,
                               void* y, int y_dtype, long long ldy, int Ho, int Wo, const float* bias, int relu,
                               float* stat_sum, float* stat_sqsum, cudaStream_t stream);
size_t denet_conv2d_rowfold_wgrad_workspace(int N, int Ho, int Wo, int Cout, int R, int stride_h);
int denet_conv2d_rowfold_wgrad(const void* dy_hi, const void* dy_lo, int N, int Ho, int Wo, int Cout, long long lddy,
                               const void* x_hi, const void* x_lo, int Hp, int Wp, int Cp, int Cin, int R, int S,
                               int stride_h, int stride_w, float* dw, int accumulate, float* workspace,
                               size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ batch norm
 * Replaces dnn_batch_normalization_train / _test and BatchNormReluOp + k_relu (denet/layer/batch_norm.py:47-53,
 * 75-76; denet/layer/batch_norm_relu.py:31-57).  x is (M = N*H*W, C) with pitch ld.
 * bn_stats: batch mean / inverse std (biased variance, 1/sqrt(var+eps)) by a deterministic two-stage shifted
 * reduction, plus the reference's EMA of mean and INVERSE STD (run_* may be NULL).
 * bn_finalize_sums: same outputs from the sum / sum-of-squares accumulated by denet_conv2d_fprop's epilogue.
 * bn_apply: y = [relu]((x-mean)*gamma*invstd + beta [+ residual]).
 * bn_inference_invstd: the reference's test-time quirk var = (1/stdinv)^2, eps added again (batch_norm.py:50-52).
 * bn_backward: dy' = dy*[y>0] if relu (yout = the forward output; when the forward had no residual input, yout may be
 *   NULL and the mask is recomputed from x, gamma, beta with the forward's exact expression - one tensor read less);
 *   dx = gamma*invstd*(dy' - mean(dy') - xhat*mean(dy'*xhat));
 *   dgamma/dbeta (+)=; dres (optional) receives dy' (gradient of the residual input). */
size_t denet_bn_workspace_bytes(long long M, int C);
int denet_bn_stats(const void* x, int dtype, long long M, int C, long long ld, float eps, float* mean, float* invstd,
                   float* run_mean, float* run_stdinv, float momentum, float* workspace, size_t workspace_bytes,
                   cudaStream_t stream);
int denet_bn_finalize_sums(const float* sum, const float* sqsum, long long M, int C, float eps, float* mean,
                           float* invstd, float* run_mean, float* run_stdinv, float momentum, cudaStream_t stream);
int denet_bn_apply(const void* x, int dtype, long long M, int C, long long ld, const float* mean, const float* invstd,
                   const float* gamma, const float* beta, const void* residual, int relu, void* y,
                   cudaStream_t stream);
/* bn_apply with the statistics given as the raw per-channel sums a convolution epilogue accumulated (sum, sqsum over M
 * rows): derives mean / invstd like denet_bn_finalize_sums, writes them to `mean` / `invstd` (the backward pass reads
 * them) and updates the running statistics - one launch instead of two per batch-norm layer. */
int denet_bn_apply_sums(const void* x, int dtype, long long M, int C, long long ld, const float* sum,
                        const float* sqsum, float eps, const float* gamma, const float* beta, const void* residual,
                        int relu, void* y, float* mean, float* invstd, float* run_mean, float* run_stdinv,
                        float momentum, cudaStream_t stream);
int denet_bn_inference_invstd(const float* run_stdinv, float eps, float* out, int C, cudaStream_t stream);
/* A/B switch for measurements: bit0 = denet_bn_backward as ONE launch (two grid-wide barriers between the reductions and
 * the apply pass; default) instead of three kernels; bit1 = apply passes launched as ONE wave of blocks and the fused
 * backward's slab totals taken 8 channels x 32 slab lanes per block (default).  Every mode sums in a fixed order
 * (deterministic); the default is 3. */
int denet_bn_set_mode(int mode);
int denet_bn_backward(const void* dy, const void* yout, const void* x, int dtype, long long M, int C, long long ld,
                      const float* mean, const float* invstd, const float* gamma, const float* beta, int relu,
                      void* dx, void* dres,
                      float* dgamma, float* dbeta, int accumulate, float* workspace, size_t workspace_bytes,
                      cudaStream_t stream);
/* Second half of the batch-norm backward when the producer of dy already masked it and accumulated the two
 * per-channel sums (denet_conv2d_dgrad_bnbwd): dx = gamma*invstd * (dy - sum_dy/M - xhat * sum_dy_xhat/M);
 * dgamma (+)= sum_dy_xhat, dbeta (+)= sum_dy (may both be NULL).  One launch. */
int denet_bn_backward_sums(const void* dy, const void* x, int dtype, long long M, int C, long long ld, const float* mean,
                           const float* invstd, const float* gamma, const float* sum_dy, const float* sum_dy_xhat,
                           void* dx, float* dgamma, float* dbeta, int accumulate, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ elementwise
 * relu: tensor.nnet.relu (denet/layer/activation.py:32-34).  add: skip / residual sums (denet/layer/skip.py:78-86,
 * denet/layer/resnet.py:113).  convert: dtype / pitch conversion.  nchw<->nhwc: host-facing layout conversion. */
int denet_relu_fwd(const void* x, int dtype, long long M, int C, long long ld, void* y, cudaStream_t stream);
int denet_relu_bwd(const void* dy, const void* y, int dtype, long long M, int C, long long ld, void* dx,
                   cudaStream_t stream);
int denet_add(const void* a, const void* b, int dtype, long long M, int C, long long ld, int relu, void* out,
              cudaStream_t stream);
/* colsum: out[c] (+)= sum over rows of x[:, c] - the bias gradient of a convolution (autodiff of
 * `y += beta[None,:,None,None]`, denet/layer/convolution.py:88-89); workspace as for denet_bn_stats. */
int denet_colsum(const void* x, int dtype, long long M, int C, long long ld, float* out, int accumulate,
                 float* workspace, size_t workspace_bytes, cudaStream_t stream);
int denet_convert(const void* x, int src_dtype, long long M, int C, long long ldx, void* y, int dst_dtype,
                  long long ldy, cudaStream_t stream);
int denet_nchw_to_nhwc(const float* x, int N, int C, int H, int W, void* y, int dtype, long long ld,
                       cudaStream_t stream);
int denet_nhwc_to_nchw(const void* x, int dtype, long long ld, int N, int C, int H, int W, float* y,
                       cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ pooling
 * pool: dnn_pool max / average_inc_pad (denet/layer/pool.py:36-38); mode 0 = max (argmax buffer
 * N*Ho*Wo*C bytes), 1 = average_inc_pad.  pool_inv: PoolInvOp / PoolInvGradOp, kernels k_pool_inv_WxH and
 * k_pool_inv_grad_WxH (denet/layer/pool_inv_op.py:38-63, 144-169); x is the small (H, W) tensor. */
int denet_pool_fwd(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int mode, int kh, int kw,
                   int sh, int sw, int ph, int pw, void* y, int Ho, int Wo, long long ldy, uint8_t* argmax,
                   cudaStream_t stream);
int denet_pool_bwd(const void* dy, int dtype, int N, int H, int W, int C, long long ldx, int mode, int kh, int kw,
                   int sh, int sw, int ph, int pw, int Ho, int Wo, long long ldy, const uint8_t* argmax, void* dx,
                   cudaStream_t stream);
int denet_pool_inv_fwd(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int sw, int sh, void* y,
                       long long ldy, cudaStream_t stream);
int denet_pool_inv_bwd(const void* dy, int dtype, int N, int H, int W, int C, long long ldx, int sw, int sh, void* dx,
                       long long ldy, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ DSS head
 * sparse_sample_fwd/bwd: DeNetSparseOp / DeNetSparseGradOp, kernels k_sparse_sample<gs> / k_sparse_sample_grad<gs>
 * (denet/layer/denet_sparse_op.py:42-85, 171-212).  fmap (B,H,W,F) pitch ldf; bbox
 * (B*rois_per_image, 4) fp32 (x0,y0,x1,y1); out rows (B*rois_per_image, gs*gs*F+2) pitch ldo in the reference's
 * channel order (yi*gs+xi)*F+f, then (box height, box width).  bwd zeroes dfmap (B,H,W,F fp32) and scatter-adds.
 * sparse_sample_index: the integer grid coordinates alone (ys, xs: (nroi, gs) int32), for parity tests. */
int denet_sparse_sample_fwd(const void* fmap, int dtype, int B, int H, int W, int F, long long ldf, const float* bbox,
                            int rois_per_image, int gs, void* out, int out_dtype, long long ldo, cudaStream_t stream);
int denet_sparse_sample_bwd(const void* dy, int dtype, long long ldo, const float* bbox, int B, int H, int W, int F,
                            int rois_per_image, int gs, float* dfmap, cudaStream_t stream);
int denet_sparse_sample_index(const float* bbox, long long nroi, int gs, int H, int W, int* ys, int* xs,
                              cudaStream_t stream);

/* build_samples: the reference's C++ extension entry point build_samples(thread_num, corner_pr, corner_threshold,
 * sample_num, max_corners, local_max, cluster_threshold) (denet/layer/denet_sparse.cc:559-668, run_build_samples
 * :489-557, search_corners :321-374, get_sample :271-308) for 4 corner types without clustering
 * (cluster_threshold >= 1, the DNS default).  corner_pr: (B,2,4,H,W) fp32 log-probabilities on the device.
 * Outputs per image, sorted by probability descending: out_pr (B,K) fp32, out_bbox (B,K,4) fp32 normalised
 * (x0/W, y0/H, (x1+1)/W, (y1+1)/H), out_ibox (B,K,4) int32 corner positions, out_count (B) int32 (<= K = sample_num^2),
 * out_ncand (B) int32 number of unique candidate boxes before the top-K (may be NULL).
 * workspace: denet_build_samples_workspace(B, H, W, max_corners) bytes. */
size_t denet_build_samples_workspace(int B, int H, int W, int max_corners);
int denet_build_samples(const float* corner_pr, int B, int H, int W, float corner_threshold, int sample_num,
                        int max_corners, int local_max, float* out_pr, float* out_bbox, int* out_ibox, int* out_count,
                        int* out_ncand, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* the same with `corner_num` maps per image: 4 corner types, or 5 when the corner layer also predicts box centres
 * (model token DNC.C, denet/layer/denet_corner.py:34-37): centres pair with every corner type
 * (denet_sparse.cc:377-468) and the centre probability joins every box score (:296-303).
 * corner_pr is (B, 2, corner_num, H, W). */
/* the same followed by the reference's corner clustering (apply_cluster, denet_sparse.cc:165-242) for every image whose
 * corner search found more than sample_num^2 boxes, when cluster_threshold < 1 (the sparse layer's nmsThreshold;
 * constructor default 0.7, denet/layer/denet_sparse.py:29): the clusters are the connected components of
 * "IoU > cluster_threshold" over the (at most 10 * sample_num^2 best) boxes; the largest sample_num^2 clusters survive
 * and each contributes its 1 + floor(size * ratio) best boxes.  Needs the larger workspace. */
size_t denet_build_samples_cluster_workspace(int B, int H, int W, int max_corners, int sample_num);
int denet_build_samples_cluster(const float* corner_pr, int B, int corner_num, int H, int W, float corner_threshold,
                                int sample_num, int max_corners, int local_max, float cluster_threshold, float* out_pr,
                                float* out_bbox, int* out_ibox, int* out_count, int* out_ncand, void* workspace,
                                size_t workspace_bytes, cudaStream_t stream);
int denet_build_samples_cn(const float* corner_pr, int B, int corner_num, int H, int W, float corner_threshold,
                           int sample_num, int max_corners, int local_max, float* out_pr, float* out_bbox,
                           int* out_ibox, int* out_count, int* out_ncand, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ costs
 * corner_logprob: DeNetCornerLayer.corner_pr (denet/layer/denet_corner.py:50-53 + common/theano_util.py:27-29):
 *   z (B*H*W rows, channels [0,cn)) -> (B,2,cn,H,W) fp32 log_softmax([z,-z]).
 * corner_cost: DeNetCornerLayer.cost (:126-134) value + gradient wrt z (written over channels [0,cn) of dz).
 * detect_cost: DeNetDetectLayer.get_errors/.cost (denet/layer/denet_detect.py:238-313), Fast R-CNN box loss variant;
 *   cost2[0] = detection cost, cost2[1] = box cost (factors included); gradient over [0, ncols_grad) of dout rows.
 * softmax_nll: RegressionLayer log-softmax + cost (denet/layer/regression.py:65-68, 97-98).
 * grad_factor multiplies the gradient only (ModelCNN cost_factors, model/model_cnn.py:229-235). */
size_t denet_loss_workspace_bytes(void);
int denet_corner_logprob(const void* z, int dtype, long long ldz, int B, int cn, int H, int W, float* corner_pr,
                         cudaStream_t stream);
int denet_corner_cost(const void* z, int dtype, long long ldz, int B, int cn, int H, int W, const float* target,
                      float cost_factor, float grad_factor, void* dz, float* cost, float* workspace,
                      cudaStream_t stream);
int denet_detect_cost(const void* o, int dtype, long long ld, int B, int sn, int s0, int use_bbox,
                      const float* target_det, const float* target_valid, const float* target_reg, float cost_factor,
                      float bbox_factor, float grad_factor, void* dout, int ncols_grad, float* cost2, float* workspace,
                      cudaStream_t stream);
/* detect_cost_v2: the same with the v2 variants of the layer: box_mode 0 none / 1 Fast R-CNN (:288-295) / 2 bounded IoU
 *   (:266-286, on the box decoded against sample_bbox (B*sn*sn, 4) fp32 as :80-97), nfit = 6 independent-fitness logits
 *   after the box outputs (:100-104, 297-299) with target_fit (B, nfit, sn, sn); joint fitness (:58-61) only changes s0
 *   to classNum*5+1.  cost3 = {detection, box, fitness} cost, factors included (:308-312). */
int denet_detect_cost_v2(const void* o, int dtype, long long ld, int B, int sn, int s0, int box_mode, int nfit,
                         const float* sample_bbox, const float* target_det, const float* target_valid,
                         const float* target_reg, const float* target_fit, float cost_factor, float bbox_factor,
                         float fit_factor, float grad_factor, void* dout, int ncols_grad, float* cost3, float* workspace,
                         cudaStream_t stream);
int denet_softmax_nll(const void* o, int dtype, long long ld, int B, int classes, const int* label, float grad_factor,
                      void* dout, float* logp_out, float* cost, float* workspace, cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ RoI post-processing
 * HOST functions (no device work).  sparse_postprocess is the python-`random` post-processing of the ranked RoIs in
 * DeNetSparseLayer.get_target (denet/layer/denet_sparse.py:184-201) for a whole batch: sub-sample to n_keep with
 * random.sample, pad to K with random boxes.  mt_state (624 words) / mt_pos are the interpreter's Mersenne-Twister
 * state (random.getstate()), advanced in place exactly as the reference's python loops would advance it.
 * pr32 (B,K) / bbox32 (B,K,4) fp32 ranked samples, count (B) int64; outputs pr (B,K) / bbox (B,K,4) doubles.
 * pyrandom_sample = random.sample(range(n), k); pyrandom_random = n x random.random(). */
int denet_sparse_postprocess(uint32_t* mt_state, int* mt_pos, const float* pr32, const float* bbox32,
                             const long long* count, int B, int K, int n_keep, double* pr, double* bbox);
int denet_pyrandom_sample(uint32_t* mt_state, int* mt_pos, int n, int k, int* out_index);
/* The same post-processing with the generator stepped AHEAD of time (while the GPU still runs the forward trunk):
 * denet_pyrandom_ahead copies the state, produces the tempered words of the next nblocks regenerations
 * (out_words, (624 - mt_pos) + 624 * nblocks of them, stream order) and the raw state after each regeneration
 * (out_states [nblocks][624]); denet_sparse_postprocess_ahead consumes words from that buffer, reports how many it used
 * and returns 1 if the buffer ran out (repeat on the live generator).  Host-only, no device work. */
int denet_pyrandom_ahead(const uint32_t* mt_state, int mt_pos, int nblocks, uint32_t* out_states, uint32_t* out_words,
                         long long* out_nwords);
int denet_sparse_postprocess_ahead(const uint32_t* words, long long nwords, long long* used, const float* pr32,
                                   const float* bbox32, const long long* count, int B, int K, int n_keep, double* pr,
                                   double* bbox);
int denet_pyrandom_random(uint32_t* mt_state, int* mt_pos, long long n, double* out);

/* ------------------------------------------------------------------------------------------------ targets
 * The host target builders of the reference, evaluated on the device from the ground-truth boxes:
 * corner_target: DeNetCornerLayer.get_target (denet/layer/denet_corner.py:81-123, dropout 0) -> (B,2,cn,H,W) fp32.
 * detect_target: DeNetDetectLayer.get_target (denet/layer/denet_detect.py:147-235, IoU of common/theano_util.py:38-59)
 *   -> target_det (B,classNum+1,sn,sn), target_valid (B,sn,sn), target_reg (B,8,sn,sn) fp32 (the three pieces of the
 *   reference's flattened yt_value).  gt_bbox (B,G,4) and sample_bbox (B,sn*sn,4) are DOUBLES (x0,y0,x1,y1) - python
 *   floats in the reference; gt_class (B,G) int32; gt_count (B) int32 (<= G <= 64).  Results equal the host
 *   builders' bit for bit. */
int denet_corner_target(const double* gt_bbox, const int* gt_count, int B, int G, int cn, int H, int W, float* target,
                        cudaStream_t stream);
int denet_detect_target(const double* gt_bbox, const int* gt_class, const int* gt_count, const double* sample_bbox,
                        int B, int G, int sn, int class_num, float thr0, float thr1, int use_bbox, float* target_det,
                        float* target_valid, float* target_reg, cudaStream_t stream);
/* detect_target_v2: fit_mode bit0 = joint fitness (:58-61,179-182: target_det (B, classNum*5+1, sn, sn), a positive RoI
 *   marks class*5 + fitness bin), bit1 = independent fitness (:187-191: target_fit (B, 6, sn, sn)); thr0 / thr1 as
 *   doubles: the IoU comparison uses them rounded to fp32 like the reference's array comparison, the fitness bin
 *   (:177) is computed in double as numpy 1.x promotes the reference's scalar arithmetic. */
int denet_detect_target_v2(const double* gt_bbox, const int* gt_class, const int* gt_count, const double* sample_bbox,
                           int B, int G, int sn, int class_num, double thr0, double thr1, int use_bbox, int fit_mode,
                           float* target_det, float* target_valid, float* target_reg, float* target_fit,
                           cudaStream_t stream);

/* ------------------------------------------------------------------------------------------------ solver
 * ModelCNN.build_train_func update rules (model/model_cnn.py:282-305, 320-324) for ALL parameter
 * tensors in one launch.  `entries` is a device array of denet_solver_entry_bytes()-sized records
 * {float* p; const float* g; float* m; float* v; long long n; int is_weight; int pad}; block i updates elements
 * [block_offset[i], +denet_solver_chunk()) of tensor block_tensor[i].  solver: 0 sgd, 1 nesterov/"torch", 2 adam.
 * grad_scale multiplies g first (1/world_size after a gradient all-reduce-sum). */
int denet_solver_entry_bytes(void);
int denet_solver_chunk(void);
int denet_solver_update(const void* entries, const int* block_tensor, const long long* block_offset, int nblocks,
                        int solver, float lr, float momentum0, float momentum1, float decay, int iteration,
                        int bias_decay, float grad_scale, cudaStream_t stream);
/* same update with the per-step scalars read from DEVICE memory, hp = {lr, momentum0, momentum1, decay, iteration,
 * grad_scale} (6 floats), so that a CUDA graph captured once replays with the values of the current step. */
int denet_solver_update_dev(const void* entries, const int* block_tensor, const long long* block_offset, int nblocks,
                            int solver, const float* hp, int bias_decay, cudaStream_t stream);

/* [total, cost_0, ...]: cost_i = sum of the lens[i] floats at srcs[i] (device array of n device pointers), total =
 * sum factors[i] * cost_i - what the reference's train function returns (`[cost] + costs`, model/model_cnn.py:229-235). */
int denet_pack_costs(const void* srcs, const int* lens, const float* factors, int n, float* out, cudaStream_t stream);

/* ---- inference tail: detect-layer outputs and per-class NMS --------------------------------------
 * detect_outputs: what DeNetDetectLayer.get_detections' compiled Theano function returns
 *   (denet/layer/denet_detect.py:60-107,330-362): det_pr (B, s0, sn, sn) = log_softmax over the s0 class channels of
 *   the layer's logits (rows = RoIs in (b, j, i) order, pitch ld floats) and, when bbox_out != NULL, the boxes
 *   (B, sn, sn, 4): the Fast R-CNN decode of channels [s0, s0+4) against sample_bbox (use_bbox = 1) or the sample boxes
 *   themselves (use_bbox = 0).
 * detections_nms: replaces the reference's CPython extension function build_detections_nms
 *   (denet/layer/denet_detect.cc:101-173; hard NMS :74-99, Gaussian soft-NMS :35-72).  det_pr / fitness element
 *   (b, cls, k = j*sn + i) lives at b*stride_b + cls*stride_c + k*stride_k (elements), so both the reference's
 *   (B, classes+1, sn, sn) arrays and NHWC log-probabilities are accepted; bbox (B, K, 4) fp32; bbox_num (B) = samples
 *   of each image that are real (the rest is padding).  Per (image, class) the surviving instances are written in the
 *   reference's output order (sample order for hard NMS, pick order for soft-NMS): out_score = exp(fitness) with libm
 *   expf's bits, out_index = sample index k; out_count (B, class_num).  Results are bit-identical to the reference. */
int denet_detect_outputs(const float* logits, long long ld, int B, int sn, int s0, int use_bbox,
                         const float* sample_bbox, float* det_pr, float* bbox_out, cudaStream_t stream);
/* detect_outputs_v2: fit_mode bit0 = joint fitness (s0 = classNum*5+1 channels folded to det_pr (B, classNum+1, sn, sn)
 *   and fitness (B, classNum+1, sn, sn), :332-348), bit1 = independent fitness (fitness = det_pr + log E[fitness],
 *   :392-397); fit_mode 0 with fitness != NULL copies det_pr (:386).  thr0 = overlapThreshold[0]. */
int denet_detect_outputs_v2(const float* logits, long long ld, int B, int sn, int s0, int use_bbox, int class_num,
                            int fit_mode, float thr0, const float* sample_bbox, float* det_pr, float* fitness,
                            float* bbox_out, cudaStream_t stream);
int denet_detections_nms(const float* det_pr, const float* fitness, long long stride_b, long long stride_c,
                         long long stride_k, const float* bbox, const int* bbox_num, int B, int class_num, int K,
                         float pr_threshold, float nms_threshold, int use_soft_nms, float* out_score, int* out_index,
                         int* out_count, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DENET_B200_H */
