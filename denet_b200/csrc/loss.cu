// Cost layers of the DeNet hot path: forward value and gradient w.r.t. the producing conv output in one pass.
//
// Reference semantics (paths relative to the reference repository):
//   corner map     denet/layer/denet_corner.py:50-53 (logits [+z,-z] -> 2-way log-softmax, theano_util.py:27-29),
//                  cost :126-131   -sum(t*logp) per image, mean over batch, / ln 2, * cost_factor
//   detection      denet/layer/denet_detect.py:76-78 (log-softmax over classNum+1), :257 (-sum(t*logp)/ln(s0)),
//                  :289-295 (Fast R-CNN smooth-L1 box loss), :304-313 (sums / batch, factors)
//   classification denet/layer/regression.py:41,65-68,97-98 (log-softmax, -mean(logp[target]))
// Costs are reduced in two stages with a fixed order (deterministic).  Gradients are written in the activation
// dtype directly into the buffer that becomes the dY operand of the layer's 1x1 convolution.
#include <algorithm>

#include "common.cuh"
#include "pack.cuh"

namespace dn {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
    __syncthreads();
    return r;  // valid in thread 0
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, int ncols, float scale,
                                    float* __restrict__ out) {
    // out[col] = scale * sum_i partial[i*ncols + col], summed in order (double accumulator)
    const int col = threadIdx.x;
    if (col >= ncols) return;
    double a = 0.0;
    for (int i = 0; i < n; ++i) a += partial[(long long)i * ncols + col];
    out[col] = (float)(a * scale);
}

__global__ void sum_partials2_kernel(const float* __restrict__ partial, int n, float scale0, float scale1,
                                     float* __restrict__ out) {
    if (threadIdx.x >= 2) return;
    double a = 0.0;
    for (int i = 0; i < n; ++i) a += partial[(long long)i * 2 + threadIdx.x];
    out[threadIdx.x] = (float)(a * (threadIdx.x == 0 ? scale0 : scale1));
}

// one warp: lane l adds partials l, l+32, ... in order (double), then the 32 lane sums are added in lane order
__global__ void sum_partials3_kernel(const float* __restrict__ partial, int n, float scale0, float scale1, float scale2,
                                     float* __restrict__ out) {
    const int lane = threadIdx.x;
    double a[3] = {0.0, 0.0, 0.0};
    for (int i = lane; i < n; i += 32)
#pragma unroll
        for (int q = 0; q < 3; ++q) a[q] += partial[(long long)i * 3 + q];
    __shared__ double sh[3][32];
#pragma unroll
    for (int q = 0; q < 3; ++q) sh[q][lane] = a[q];
    __syncwarp();
    if (lane < 3) {
        double t = 0.0;
        for (int l = 0; l < 32; ++l) t += sh[lane][l];
        out[lane] = (float)(t * (lane == 0 ? scale0 : (lane == 1 ? scale1 : scale2)));
    }
}

// z: conv output rows = B*H*W pixels (NHWC), corner channels [0, cn).  corner_pr: (B, 2, cn, H, W) fp32.
template <typename T>
__global__ void corner_logprob_kernel(const T* __restrict__ z, long long ldz, int B, int cn, int H, int W,
                                      float* __restrict__ corner_pr) {
    const long long HW = (long long)H * W;
    const long long total = (long long)B * cn * HW;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long pix = idx % HW;
        long long t = idx / HW;
        const int c = (int)(t % cn);
        const int b = (int)(t / cn);
        const float v = to_f<T>(z[((long long)b * HW + pix) * ldz + c]);
        // log_softmax([v, -v]) exactly as theano_util.log_softmax: xdev = x - max; xdev - log(sum(exp(xdev)))
        const float m = fabsf(v);
        const float d0 = v - m, d1 = -v - m;
        const float lse = logf(expf(d0) + expf(d1));
        corner_pr[(((long long)b * 2 + 0) * cn + c) * HW + pix] = d0 - lse;
        corner_pr[(((long long)b * 2 + 1) * cn + c) * HW + pix] = d1 - lse;
    }
}

// cost partial sums + gradient w.r.t. z.  target: (B,2,cn,H,W) fp32.  grad_scale = total factor / (B * ln 2).
template <typename T>
__global__ void __launch_bounds__(kLossThreads) corner_cost_kernel(const T* __restrict__ z, long long ldz, int B, int cn,
                                                                    int H, int W, const float* __restrict__ target,
                                                                    float grad_scale, T* __restrict__ dz,
                                                                    float* __restrict__ partial) {
    __shared__ float sh[kLossThreads / 32];
    const long long HW = (long long)H * W;
    const long long total = (long long)B * cn * HW;
    float acc = 0.f;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long pix = idx % HW;
        long long t = idx / HW;
        const int c = (int)(t % cn);
        const int b = (int)(t / cn);
        const long long zoff = ((long long)b * HW + pix) * ldz + c;
        const float v = to_f<T>(z[zoff]);
        const float m = fabsf(v);
        const float d0 = v - m, d1 = -v - m;
        const float e0 = expf(d0), e1 = expf(d1);
        const float lse = logf(e0 + e1);
        const float t0 = target[(((long long)b * 2 + 0) * cn + c) * HW + pix];
        const float t1 = target[(((long long)b * 2 + 1) * cn + c) * HW + pix];
        acc += t0 * (d0 - lse) + t1 * (d1 - lse);
        // d/dv [t0*lp0 + t1*lp1] with lp0 = v - lse(v,-v), lp1 = -v - lse(v,-v), dlse/dv = (e0 - e1)/(e0 + e1)
        const float th = (e0 - e1) / (e0 + e1);
        const float g = t0 * (1.0f - th) + t1 * (-1.0f - th);
        dz[zoff] = from_f<T>(-g * grad_scale);
    }
    const float s = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// One warp per RoI row. o: conv output rows (R, ld) with [0,s0) class logits (classNum+1, or classNum*5+1 with joint
// fitness), [s0,s0+4) box regression when box_mode != 0, then nfit independent-fitness logits when nfit != 0.
// target_det (B,s0,sn,sn), target_valid (B,sn,sn), target_reg (B,8,sn,sn), target_fit (B,nfit,sn,sn) in the reference's
// NCHW packing.  box_mode 1: Fast R-CNN smooth-L1 on (tx,ty,tw,th) (denet_detect.py:288-295); 2: bounded IoU (:266-286)
// on the decoded box (:80-97), sample_bbox (R,4) fp32.
// partial[blk] = {sum t*logp, sum box term, sum t_fit*logp_fit}; gradients for channels [0, s0+4+nfit), zero beyond.
template <typename T>
__global__ void __launch_bounds__(kLossThreads) detect_cost_kernel(const T* __restrict__ o, long long ld, int R, int s0,
                                                                    int sn2, int box_mode, int nfit,
                                                                    const float* __restrict__ sample_bbox,
                                                                    const float* __restrict__ target_det,
                                                                    const float* __restrict__ target_valid,
                                                                    const float* __restrict__ target_reg,
                                                                    const float* __restrict__ target_fit,
                                                                    float det_grad_scale, float bbox_factor,
                                                                    float bbox_grad_scale, float fit_grad_scale,
                                                                    T* __restrict__ dout, int ncols_grad,
                                                                    float* __restrict__ partial) {
    __shared__ float sh[kLossThreads / 32];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int s1 = box_mode ? 4 : 0;
    float acc_det = 0.f, acc_box = 0.f, acc_fit = 0.f;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < R;
         row += (long long)gridDim.x * warps_per_block) {
        const int b = (int)(row / sn2);
        const int ji = (int)(row % sn2);
        const T* orow = o + row * ld;
        T* grow = dout + row * ld;
        float mx = -INFINITY;
        for (int k = lane; k < s0; k += 32) mx = fmaxf(mx, to_f<T>(orow[k]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float se = 0.f, st = 0.f;
        for (int k = lane; k < s0; k += 32) {
            se += expf(to_f<T>(orow[k]) - mx);
            st += target_det[((long long)b * s0 + k) * sn2 + ji];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, off);
            st += __shfl_xor_sync(0xffffffffu, st, off);
        }
        const float lse = logf(se);
        for (int k = lane; k < s0; k += 32) {
            const float xd = to_f<T>(orow[k]) - mx;
            const float lp = xd - lse;
            const float t = target_det[((long long)b * s0 + k) * sn2 + ji];
            acc_det += t * lp;
            const float p = expf(lp);
            grow[k] = from_f<T>(-(t - p * st) * det_grad_scale);
        }
        if (box_mode && lane < 4) {
            const int xy = lane & 1;
            const float valid = target_valid[(long long)b * sn2 + ji];
            const float tg_c = target_reg[((long long)b * 8 + xy) * sn2 + ji];              // target centre x|y
            const float tg_s = target_reg[((long long)b * 8 + 2 + xy) * sn2 + ji];          // target w|h
            float g;                                                                         // d(smooth-L1 term)/d(output)
            float l;
            if (box_mode == 1) {
                const float sm_c = target_reg[((long long)b * 8 + 4 + xy) * sn2 + ji];      // sample centre
                const float sm_s = target_reg[((long long)b * 8 + 6 + xy) * sn2 + ji];      // sample w|h
                const float t = lane < 2 ? (tg_c - sm_c) / sm_s : logf(tg_s / sm_s);
                const float d = t - to_f<T>(orow[s0 + lane]);
                const float ad = fabsf(d);
                l = ad < 1.0f ? 0.5f * d * d : ad - 0.5f;
                g = -(ad < 1.0f ? d : (d > 0.f ? 1.0f : -1.0f));
            } else {
                const float* sb = sample_bbox + row * 4;
                const float b0 = sb[xy], b1 = sb[2 + xy];
                const float s_c = 0.5f * (b0 + b1), s_s = b1 - b0;                           // :84-87
                const float pc = to_f<T>(orow[s0 + xy]) * s_s + s_c;                         // :89-92
                const float ps = expf(to_f<T>(orow[s0 + 2 + xy])) * s_s;
                const float p0 = pc - ps * 0.5f, p1 = pc + ps * 0.5f;                        // :93-96
                const float predict_c = 0.5f * (p0 + p1), predict_s = p1 - p0;               // :271-274
                const float eps = 0.001f;
                float c, dc;                                                                 // cost and d cost / d output
                if (lane < 2) {
                    const float d = tg_c - predict_c;                                        // :276-283
                    if (d >= 0.0f) {
                        const float den = tg_s + d + eps;
                        c = 2.0f * d / den;
                        dc = 2.0f * (tg_s + eps) / (den * den);
                    } else {
                        const float den = tg_s - d + eps;
                        c = -2.0f * d / den;
                        dc = -2.0f * (tg_s + eps) / (den * den);
                    }
                    dc *= -s_s;                                                              // d(d)/d(output) = -sample size
                } else {
                    const float q0 = tg_s / (predict_s + eps), q1 = predict_s / (tg_s + eps);   // :284-285
                    if (q0 < q1) {
                        c = 1.0f - q0;
                        dc = tg_s / ((predict_s + eps) * (predict_s + eps));
                    } else {
                        c = 1.0f - q1;
                        dc = -1.0f / (tg_s + eps);
                    }
                    dc *= ps;                                                                // d(size)/d(output) = size
                }
                const float ac = fabsf(c);
                l = ac < 1.0f ? 0.5f * c * c : ac - 0.5f;
                g = (ac < 1.0f ? c : (c > 0.f ? 1.0f : -1.0f)) * dc;
            }
            acc_box += bbox_factor * valid * l;
            grow[s0 + lane] = from_f<T>(bbox_factor * valid * g * bbox_grad_scale);
        }
        if (nfit) {                                                                          // :100-104, 297-299
            const int base = s0 + s1;
            const bool in = lane < nfit;
            const float v = in ? to_f<T>(orow[base + lane]) : -INFINITY;
            float fm = v;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) fm = fmaxf(fm, __shfl_xor_sync(0xffffffffu, fm, off));
            float fe = in ? expf(v - fm) : 0.f;
            float ft = in ? target_fit[((long long)b * nfit + lane) * sn2 + ji] : 0.f;
            const float t = ft;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                fe += __shfl_xor_sync(0xffffffffu, fe, off);
                ft += __shfl_xor_sync(0xffffffffu, ft, off);
            }
            if (in) {
                const float lp = (v - fm) - logf(fe);
                acc_fit += t * lp;
                grow[base + lane] = from_f<T>(-(t - expf(lp) * ft) * fit_grad_scale);
            }
        }
        for (int k = s0 + s1 + nfit + lane; k < ncols_grad; k += 32) grow[k] = from_f<T>(0.f);
    }
    const float s_det = block_sum(acc_det, sh);
    const float s_box = block_sum(acc_box, sh);
    const float s_fit = block_sum(acc_fit, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x * 3 + 0] = s_det;
        partial[blockIdx.x * 3 + 1] = s_box;
        partial[blockIdx.x * 3 + 2] = s_fit;
    }
}

// classification head: rows (B, ld) logits over `classes`; label[b] int32.  cost = -mean(logp[label]).
template <typename T>
__global__ void __launch_bounds__(kLossThreads) softmax_nll_kernel(const T* __restrict__ o, long long ld, int B,
                                                                    int classes, const int* __restrict__ label,
                                                                    float grad_scale, T* __restrict__ dout,
                                                                    float* __restrict__ logp_out,
                                                                    float* __restrict__ partial) {
    __shared__ float sh[kLossThreads / 32];
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    float acc = 0.f;
    for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < B;
         row += (long long)gridDim.x * warps_per_block) {
        const T* orow = o + row * ld;
        float mx = -INFINITY;
        for (int k = lane; k < classes; k += 32) mx = fmaxf(mx, to_f<T>(orow[k]));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float se = 0.f;
        for (int k = lane; k < classes; k += 32) se += expf(to_f<T>(orow[k]) - mx);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
        const float lse = logf(se);
        const int lab = label[row];
        for (int k = lane; k < classes; k += 32) {
            const float lp = to_f<T>(orow[k]) - mx - lse;
            if (logp_out) logp_out[row * classes + k] = lp;
            if (k == lab) acc += lp;
            if (dout) dout[row * ld + k] = from_f<T>((expf(lp) - (k == lab ? 1.0f : 0.0f)) * grad_scale);
        }
    }
    const float s = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

}  // namespace dn

using namespace dn;

static const int kLossBlocks = 148;
static const int kDetectBlocks = 148 * 8;   // detect_cost: one or two RoIs per warp (each RoI is a chain of dependent,
                                            // strided target reads: 16 RoIs per warp in sequence took 78 us for 13 MB)

extern "C" size_t denet_loss_workspace_bytes(void) { return sizeof(float) * (3 * kDetectBlocks + 4); }

extern "C" int denet_corner_logprob(const void* z, int dtype, long long ldz, int B, int cn, int H, int W,
                                    float* corner_pr, cudaStream_t stream) {
    DN_REQUIRE(z && corner_pr, "corner_logprob: null pointer");
    const long long total = (long long)B * cn * H * W;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 8);
    if (dtype == DENET_F32)
        corner_logprob_kernel<float><<<DN_G(grid), 256, 0, stream>>>((const float*)z, ldz, B, cn, H, W, corner_pr);
    else
        corner_logprob_kernel<__nv_bfloat16><<<DN_G(grid), 256, 0, stream>>>((const __nv_bfloat16*)z, ldz, B, cn, H, W,
                                                                        corner_pr);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_corner_cost(const void* z, int dtype, long long ldz, int B, int cn, int H, int W,
                                 const float* target, float cost_factor, float grad_factor, void* dz, float* cost,
                                 float* workspace, cudaStream_t stream) {
    DN_REQUIRE(z && target && dz && cost && workspace, "corner_cost: null pointer");
    const float inv = 1.0f / ((float)B * 0.6931471805599453f);
    if (dtype == DENET_F32)
        corner_cost_kernel<float><<<DN_G(kLossBlocks), kLossThreads, 0, stream>>>(
            (const float*)z, ldz, B, cn, H, W, target, cost_factor * grad_factor * inv, (float*)dz, workspace);
    else
        corner_cost_kernel<__nv_bfloat16><<<DN_G(kLossBlocks), kLossThreads, 0, stream>>>(
            (const __nv_bfloat16*)z, ldz, B, cn, H, W, target, cost_factor * grad_factor * inv, (__nv_bfloat16*)dz,
            workspace);
    sum_partials_kernel<<<DN_G(1), 32, 0, stream>>>(workspace, kLossBlocks, 1, -cost_factor * inv, cost);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_detect_cost_v2(const void* o, int dtype, long long ld, int B, int sn, int s0, int box_mode, int nfit,
                                    const float* sample_bbox, const float* target_det, const float* target_valid,
                                    const float* target_reg, const float* target_fit, float cost_factor,
                                    float bbox_factor, float fit_factor, float grad_factor, void* dout, int ncols_grad,
                                    float* cost3, float* workspace, cudaStream_t stream) {
    DN_REQUIRE(o && target_det && dout && cost3 && workspace, "detect_cost: null pointer");
    DN_REQUIRE(box_mode >= 0 && box_mode <= 2, "detect_cost: box_mode must be 0 (none), 1 (Fast R-CNN) or 2 (bounded IoU)");
    DN_REQUIRE(!box_mode || (target_valid && target_reg), "detect_cost: bbox targets missing");
    DN_REQUIRE(box_mode != 2 || sample_bbox, "detect_cost: bounded IoU needs the sample boxes");
    DN_REQUIRE(nfit >= 0 && nfit <= 32 && (!nfit || target_fit), "detect_cost: fitness target missing (or nfit > 32)");
    DN_REQUIRE(s0 + (box_mode ? 4 : 0) + nfit <= ncols_grad, "detect_cost: gradient row narrower than the outputs");
    const int sn2 = sn * sn;
    const int R = B * sn2;
    const float det_scale = cost_factor / ((float)B * logf((float)s0));
    const float box_scale = bbox_factor / (float)B;
    const float fit_scale = nfit ? fit_factor / ((float)B * logf((float)nfit)) : 0.f;
    const int grid = std::min(kDetectBlocks, ceil_div(R, kLossThreads / 32));
    if (dtype == DENET_F32)
        detect_cost_kernel<float><<<DN_G(grid), kLossThreads, 0, stream>>>(
            (const float*)o, ld, R, s0, sn2, box_mode, nfit, sample_bbox, target_det, target_valid, target_reg, target_fit,
            det_scale * grad_factor, bbox_factor, box_scale * grad_factor, fit_scale * grad_factor, (float*)dout,
            ncols_grad, workspace);
    else
        detect_cost_kernel<__nv_bfloat16><<<DN_G(grid), kLossThreads, 0, stream>>>(
            (const __nv_bfloat16*)o, ld, R, s0, sn2, box_mode, nfit, sample_bbox, target_det, target_valid, target_reg,
            target_fit, det_scale * grad_factor, bbox_factor, box_scale * grad_factor, fit_scale * grad_factor,
            (__nv_bfloat16*)dout, ncols_grad, workspace);
    // cost3 = {detection, box, independent fitness} cost, each including its factors (reference :308-312)
    sum_partials3_kernel<<<DN_G(1), 32, 0, stream>>>(workspace, grid, -det_scale, box_scale, -fit_scale, cost3);
    DN_CHECK_LAUNCH();
    return 0;
}

// the round-1 entry point: Fast R-CNN box loss, no fitness head; cost2 = {detection, box}
extern "C" int denet_detect_cost(const void* o, int dtype, long long ld, int B, int sn, int s0, int use_bbox,
                                 const float* target_det, const float* target_valid, const float* target_reg,
                                 float cost_factor, float bbox_factor, float grad_factor, void* dout, int ncols_grad,
                                 float* cost2, float* workspace, cudaStream_t stream) {
    DN_REQUIRE(cost2 && workspace, "detect_cost: null pointer");
    // the third sum lands in the workspace tail (the kernel's partials use 3 * kDetectBlocks floats of 3 * kDetectBlocks + 4)
    float* cost3 = workspace + 3 * kDetectBlocks;
    const int rc = denet_detect_cost_v2(o, dtype, ld, B, sn, s0, use_bbox ? 1 : 0, 0, nullptr, target_det, target_valid,
                                        target_reg, nullptr, cost_factor, bbox_factor, 0.f, grad_factor, dout, ncols_grad,
                                        cost3, workspace, stream);
    if (rc) return rc;
    DN_CHECK_CUDA(cudaMemcpyAsync(cost2, cost3, 2 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    return 0;
}

extern "C" int denet_softmax_nll(const void* o, int dtype, long long ld, int B, int classes, const int* label,
                                 float grad_factor, void* dout, float* logp_out, float* cost, float* workspace,
                                 cudaStream_t stream) {
    DN_REQUIRE(o && label && cost && workspace, "softmax_nll: null pointer");
    const int grid = std::min(kLossBlocks, ceil_div(B, kLossThreads / 32));
    if (dtype == DENET_F32)
        softmax_nll_kernel<float><<<DN_G(grid), kLossThreads, 0, stream>>>((const float*)o, ld, B, classes, label,
                                                                     grad_factor / (float)B, (float*)dout, logp_out,
                                                                     workspace);
    else
        softmax_nll_kernel<__nv_bfloat16><<<DN_G(grid), kLossThreads, 0, stream>>>(
            (const __nv_bfloat16*)o, ld, B, classes, label, grad_factor / (float)B, (__nv_bfloat16*)dout, logp_out,
            workspace);
    sum_partials_kernel<<<DN_G(1), 32, 0, stream>>>(workspace, grid, 1, -1.0f / (float)B, cost);
    DN_CHECK_LAUNCH();
    return 0;
}


// [total, cost_0, cost_1, ...] of the cost layers in one tiny launch (ModelCNN._pack_costs; the reference returns
// `[cost] + costs` from its compiled train function, model/model_cnn.py:229-235,445): cost_i = sum of the lens[i] floats
// at srcs[i] (a layer may keep its cost as several terms, e.g. detection + box cost), total = sum factors[i] * cost_i.
namespace dn {
__global__ void pack_costs_kernel(const float* const* __restrict__ srcs, const int* __restrict__ lens,
                                  const float* __restrict__ factors, int n, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float total = 0.f;
    for (int i = 0; i < n; ++i) {
        float c = 0.f;
        for (int j = 0; j < lens[i]; ++j) c += srcs[i][j];
        out[1 + i] = c;
        total += factors[i] * c;
    }
    out[0] = total;
}
}  // namespace dn

extern "C" int denet_pack_costs(const void* srcs, const int* lens, const float* factors, int n, float* out,
                                cudaStream_t stream) {
    DN_REQUIRE(srcs && lens && factors && out && n > 0, "pack_costs: null pointer");
    dn::pack_costs_kernel<<<DN_G(1), 32, 0, stream>>>(reinterpret_cast<const float* const*>(srcs), lens, factors, n, out);
    DN_CHECK_LAUNCH();
    return 0;
}
