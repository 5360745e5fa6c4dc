/* denet_b200 C-ABI: B200 (sm_100a) kernels for the DeNet training hot path.
 *
 * This is the drop-in boundary.  Every entry point takes raw DEVICE pointers, sizes and a
 * cudaStream_t, enqueues work on that stream and returns without synchronising.  Nothing here allocates:
 * outputs and workspaces are owned by the caller.  Return value: 0 on success, <0 on error; the message of
 * the last error on the calling host thread is returned by denet_last_error().  Calls on distinct streams
 * are thread-safe.
 *
 * Layout convention: activations are NHWC ("pixel-major"): element (n, h, w, c) of a tensor with pixel pitch
 * `ld` lives at ((n*H + h)*W + w)*ld + c.  Dtypes are DENET_F32 or DENET_BF16.  Filters keep the reference's
 * layout (Cout, Cin, R, S) fp32 *for the true (flipped) convolution* that Theano's conv2d computes
 * (reference denet/layer/convolution.py:83), so reference checkpoints load unchanged.
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

#define DENET_ABI_VERSION 1

#define DENET_F32 0
#define DENET_BF16 1

#define DENET_ERR_ARG (-1)
#define DENET_ERR_CUDA (-2)

const char* denet_last_error(void);
int denet_abi_version(void);

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

/* fp32 -> bf16 hi (+ lo = bf16(x - hi)) operand split.  lo may be NULL. */
int denet_split_bf16(const float* x, void* hi, void* lo, long long n, cudaStream_t stream);

/* Stride-1 R x S correlation of an NHWC bf16 tensor with a prepared operand:
 *   y[n,h,w,co] = sum_{r,s,ci} x[n, h+r-pad_h, w+s-pad_w, ci] * B[co][r*S+s][ci]   (zero outside the image)
 * followed by the fused epilogue  (+bias[co]) (+residual) (relu)  and optional per-channel sum / sum-of-squares
 * accumulation of the conv output (before residual/relu) for batch-norm statistics.
 * With mode-0 operands this is the reference fprop; with mode-1 operands, Cin/Cout swapped and
 * pad = R-1-pad it is the reference dgrad. */
int denet_conv2d_fprop(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin, long long ldx,
                       const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w, void* y,
                       int y_dtype, long long ldy, int Ho, int Wo, const float* bias, const void* residual, int relu,
                       float* stat_sum, float* stat_sqsum, cudaStream_t stream);

/* Filter gradient in the reference layout: dw[co][ci][R-1-r][S-1-s] (+)= sum_pixels dy[p,co] * x[p+(r,s)-pad, ci].
 * Split-K partial sums go through `workspace` (size from denet_conv2d_wgrad_workspace) and are reduced in a
 * fixed order (deterministic). */
size_t denet_conv2d_wgrad_workspace(int N, int Ho, int Wo, int Cout, int Cin, int R, int S);
int denet_conv2d_wgrad(const void* dy_hi, const void* dy_lo, int N, int Ho, int Wo, int Cout, long long lddy,
                       const void* x_hi, const void* x_lo, int Hi, int Wi, int Cin, long long ldx, int R, int S,
                       int pad_h, int pad_w, float* dw, int accumulate, float* workspace, size_t workspace_bytes,
                       cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DENET_B200_H */
