// Inference tail of the DSS detector on the device: class log-probabilities + box decode of the detect layer, and the
// per-class non-maximum suppression of the reference's CPython extension `denet_detect`.
//
// Reference semantics followed (paths relative to the reference repository):
//   detect outputs   denet/layer/denet_detect.py:60-107  (det_pr = log_softmax over the s0 class channels,
//                    bbox_predict = Fast R-CNN style decode of the 4 regression outputs against the sample box)
//   NMS              denet/layer/denet_detect.cc:12-31 (IoU, fp32), :35-72 (Gaussian soft-NMS, scores kept as logs),
//                    :74-99 (hard NMS: an instance is dropped iff ANY strictly higher-scoring instance of its class
//                    overlaps it by more than the threshold - not the greedy variant), :101-173 (instance selection
//                    `log_pr >= log(pr_threshold)` among the first bbox_num samples, output score exp(fitness))
// The reference walks B x classes x K^2 on one CPU thread; here one CTA owns one (image, class) pair.  Every fp32
// operation is written with explicit round-to-nearest intrinsics in the reference's order (no FMA contraction), and
// exp() is the bit-exact glibc expf restatement, so the detection lists equal the reference's bit for bit.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "expf_glibc.cuh"

namespace dn {

// ------------------------------------------------------------------------------------------------ detect outputs
// logits: rows = B*sn*sn RoIs in (b, j, i) order, `ld` floats apart, channels [0,s0) classes, [s0,s0+4) box regression,
// then 6 independent-fitness logits (fit_mode bit1).  fit_mode bit0 (joint fitness, denet_detect.py:332-348): the s0 =
// classNum*5+1 log-probabilities are folded into det_pr (B, classNum+1, sn, sn) = logsumexp over the 5 fitness bins of
// a class (+ the null class) and fitness (B, classNum+1, sn, sn; null channel unused) = log(sum_f p(c,f) * val_f),
// val_f = thr0 + f (1 - thr0) / 5.  bit1 (:392-397): fitness = det_pr + log(sum_f q_f * val'_f), q = softmax of the 6
// fitness logits, val' = {0, thr0 + i (1 - thr0) / 5}.
__global__ void detect_outputs_kernel(const float* __restrict__ logits, long long ld, int B, int sn, int s0, int use_bbox,
                                      int class_num, int fit_mode, float thr0,
                                      const float* __restrict__ sample_bbox, float* __restrict__ det_pr,
                                      float* __restrict__ fitness, float* __restrict__ bbox_out) {
    const int roi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // one warp per RoI
    const int lane = threadIdx.x & 31;
    const int nroi = B * sn * sn;
    if (roi >= nroi) return;
    const float* z = logits + (long long)roi * ld;
    float m = -FLT_MAX;
    for (int c = lane; c < s0; c += 32) m = fmaxf(m, z[c]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int c = lane; c < s0; c += 32) s += expf(z[c] - m);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float ls = logf(s);
    const int b = roi / (sn * sn), ji = roi % (sn * sn);
    const long long plane = (long long)sn * sn;
    float fit_add = 0.f;
    if (fit_mode & 2) {
        const int base = s0 + (use_bbox ? 4 : 0);
        const bool in = lane < 6;
        const float v = in ? z[base + lane] : -FLT_MAX;
        float fm = v;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) fm = fmaxf(fm, __shfl_xor_sync(0xffffffffu, fm, o));
        float fe = in ? expf(v - fm) : 0.f;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) fe += __shfl_xor_sync(0xffffffffu, fe, o);
        const float q = in ? expf((v - fm) - logf(fe)) : 0.f;
        // (the reference sums in float64 and rounds once; 6 terms in [0,1]: the double sum below does the same)
        double term = (in && lane > 0) ? (double)q * ((double)thr0 + (lane - 1) * (1.0 - (double)thr0) / 5.0) : 0.0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
        fit_add = logf((float)term);
    }
    if (fit_mode & 1) {
        const int s_out = class_num + 1;
        for (int c = lane; c < class_num; c += 32) {
            float lp[5], mm = -FLT_MAX;
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                lp[f] = (z[c * 5 + f] - m) - ls;
                mm = fmaxf(mm, lp[f]);
            }
            float se = 0.f, sv = 0.f;
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                se += expf(lp[f] - mm);
                sv += expf(lp[f]) * (thr0 + (float)f * (1.0f - thr0) / 5.0f);
            }
            det_pr[((long long)b * s_out + c) * plane + ji] = mm + logf(se);
            fitness[((long long)b * s_out + c) * plane + ji] = logf(sv);
        }
        if (lane == 0) {
            const float lp = (z[class_num * 5] - m) - ls;
            det_pr[((long long)b * s_out + class_num) * plane + ji] = lp;
            fitness[((long long)b * s_out + class_num) * plane + ji] = lp;
        }
    } else {
        // reference layout (B, s0, sn, sn); theano_util.log_softmax: (x - max) - log(sum(exp(x - max)))
        for (int c = lane; c < s0; c += 32) {
            const float lp = (z[c] - m) - ls;
            det_pr[((long long)b * s0 + c) * plane + ji] = lp;
            if (fitness) fitness[((long long)b * s0 + c) * plane + ji] = lp + fit_add;
        }
    }
    if (lane == 0 && bbox_out) {
        const float* sb = sample_bbox + (long long)roi * 4;
        float x0 = sb[0], y0 = sb[1], x1 = sb[2], y1 = sb[3];
        if (use_bbox) {
            const float cx = 0.5f * (x0 + x1), cy = 0.5f * (y0 + y1), w = x1 - x0, h = y1 - y0;
            const float pcx = z[s0 + 0] * w + cx, pcy = z[s0 + 1] * h + cy;
            const float pw = expf(z[s0 + 2]) * w, ph = expf(z[s0 + 3]) * h;
            x0 = pcx - pw * 0.5f; y0 = pcy - ph * 0.5f; x1 = pcx + pw * 0.5f; y1 = pcy + ph * 0.5f;
        }
        float* o = bbox_out + (long long)roi * 4;
        o[0] = x0; o[1] = y0; o[2] = x1; o[3] = y1;
    }
}

// ------------------------------------------------------------------------------------------------ NMS
__device__ __forceinline__ float iou_ref(const float4 a, const float4 b) {
    // denet_detect.cc:12-31, fp32, operation for operation
    const float dx = fmaxf(0.0f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float dy = fmaxf(0.0f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float ai = __fmul_rn(dx, dy);
    const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float au = __fsub_rn(__fadd_rn(aa, ab), ai);
    return __fdiv_rn(ai, au);
}

constexpr int kNmsThreads = 256;
constexpr int kNmsMaxK = 4096;      // samples per image the shared-memory arrays hold (reference: sn <= 48 -> 2304)

// det_pr / fitness element (b, cls, k) at b*sb + cls*sc + k*sk (k = j*sn + i): the reference's (B, C, sn, sn) arrays
// (sc = K, sk = 1) or the detect layer's NHWC log-probabilities (sc = 1, sk = pitch) without a transpose.
__global__ void __launch_bounds__(kNmsThreads) detections_nms_kernel(
    const float* __restrict__ det_pr, const float* __restrict__ fitness, long long sb, long long sc, long long sk,
    const float* __restrict__ bbox, const int* __restrict__ bbox_num, int class_num, int K, float log_pr_threshold,
    float nms_threshold, int use_soft_nms, float* __restrict__ out_score, int* __restrict__ out_index,
    int* __restrict__ out_count) {
    extern __shared__ uint8_t nms_smem[];
    float4* box = reinterpret_cast<float4*>(nms_smem);                  // [n] instance boxes
    float* score = reinterpret_cast<float*>(box + K);                   // [n] fitness (log basis)
    int* idx = reinterpret_cast<int*>(score + K);                       // [n] sample index j*sn + i
    int* flag = idx + K;                                                // [K] scratch
    __shared__ int s_n, s_scan[kNmsThreads], s_best;
    __shared__ float s_red[kNmsThreads];
    __shared__ int s_redi[kNmsThreads];

    const int b = blockIdx.x / class_num, cls = blockIdx.x % class_num;
    const int tid = threadIdx.x;
    int nb = bbox_num[b];
    nb = nb < K ? nb : K;
    const float* pr = det_pr + (long long)b * sb + (long long)cls * sc;
    const float* fit = fitness + (long long)b * sb + (long long)cls * sc;
    const float* bb = bbox + (long long)b * K * 4;
    float* o_score = out_score + ((long long)b * class_num + cls) * K;
    int* o_index = out_index + ((long long)b * class_num + cls) * K;

    // ---- ordered compaction of the instances of this class: log_pr >= log(pr_threshold), first nb samples
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kNmsThreads) {
        const int k = base + tid;
        const int take = (k < nb) && (pr[(long long)k * sk] >= log_pr_threshold);
        s_scan[tid] = take;
        __syncthreads();
        for (int o = 1; o < kNmsThreads; o <<= 1) {                     // inclusive scan
            const int v = tid >= o ? s_scan[tid - o] : 0;
            __syncthreads();
            s_scan[tid] += v;
            __syncthreads();
        }
        const int pos = s_n + s_scan[tid] - take;
        if (take) {
            box[pos] = *reinterpret_cast<const float4*>(bb + (long long)k * 4);
            score[pos] = fit[(long long)k * sk];
            idx[pos] = k;
        }
        __syncthreads();
        if (tid == kNmsThreads - 1) s_n += s_scan[tid];
        __syncthreads();
    }
    const int n = s_n;
    const bool nms_on = (nms_threshold > 0.0f) && (nms_threshold < 1.0f) && n > 0;     // perform_nms :77-78
    if (!nms_on) {
        for (int i = tid; i < n; i += kNmsThreads) {
            o_score[i] = expf_glibc(score[i]);
            o_index[i] = idx[i];
        }
        if (tid == 0) out_count[blockIdx.x] = n;
        return;
    }
    if (!use_soft_nms) {
        // ---- hard NMS: unique iff no strictly better instance overlaps by more than the threshold
        for (int a = tid; a < n; a += kNmsThreads) {
            const float4 ba = box[a];
            const float sa = score[a];
            int unique = 1;
            for (int j = 0; j < n; ++j) {
                if (sa < score[j] && iou_ref(ba, box[j]) > nms_threshold) {
                    unique = 0;
                    break;
                }
            }
            flag[a] = unique;
        }
        __syncthreads();
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int base = 0; base < n; base += kNmsThreads) {
            const int a = base + tid;
            const int take = (a < n) && flag[a];
            s_scan[tid] = take;
            __syncthreads();
            for (int o = 1; o < kNmsThreads; o <<= 1) {
                const int v = tid >= o ? s_scan[tid - o] : 0;
                __syncthreads();
                s_scan[tid] += v;
                __syncthreads();
            }
            const int pos = s_n + s_scan[tid] - take;
            if (take) {
                o_score[pos] = expf_glibc(score[a]);
                o_index[pos] = idx[a];
            }
            __syncthreads();
            if (tid == kNmsThreads - 1) s_n += s_scan[tid];
            __syncthreads();
        }
        if (tid == 0) out_count[blockIdx.x] = s_n;
        return;
    }
    // ---- Gaussian soft-NMS (:35-72): pick the best (first of equals in list order), rescore the rest, drop < -6.9
    const float discard = -6.9f;           // `const float& discard_threshold = -6.9`
    for (int i = tid; i < n; i += kNmsThreads) flag[i] = 1;             // alive
    __syncthreads();
    int nout = 0;
    while (true) {
        float best = -FLT_MAX;
        int besti = 0x7fffffff;
        for (int i = tid; i < n; i += kNmsThreads) {
            if (flag[i] && (score[i] > best)) {          // strict: the earliest of equal scores wins within a thread
                best = score[i];
                besti = i;
            }
        }
        s_red[tid] = best;
        s_redi[tid] = besti;
        __syncthreads();
        for (int o = kNmsThreads / 2; o >= 1; o >>= 1) {
            if (tid < o) {
                const float ob = s_red[tid + o];
                const int oi = s_redi[tid + o];
                // an alive instance always beats "none" (index 0x7fffffff); equal scores: the smaller list position
                if (oi != 0x7fffffff && (s_redi[tid] == 0x7fffffff || ob > s_red[tid] ||
                                         (ob == s_red[tid] && oi < s_redi[tid]))) {
                    s_red[tid] = ob;
                    s_redi[tid] = oi;
                }
            }
            __syncthreads();
        }
        if (tid == 0) s_best = s_redi[0];
        __syncthreads();
        const int m = s_best;
        if (m == 0x7fffffff) break;
        if (tid == 0) {
            o_score[nout] = expf_glibc(score[m]);
            o_index[nout] = idx[m];
            flag[m] = 0;
        }
        ++nout;
        const float4 bm = box[m];
        __syncthreads();
        for (int i = tid; i < n; i += kNmsThreads) {
            if (flag[i]) {
                const float iou = iou_ref(bm, box[i]);
                const float sc2 = __fsub_rn(score[i], __fdiv_rn(__fmul_rn(iou, iou), nms_threshold));
                score[i] = sc2;
                if (sc2 < discard) flag[i] = 0;
            }
        }
        __syncthreads();
    }
    if (tid == 0) out_count[blockIdx.x] = nout;
}

}  // namespace dn

using namespace dn;

extern "C" int denet_detect_outputs_v2(const float* logits, long long ld, int B, int sn, int s0, int use_bbox,
                                       int class_num, int fit_mode, float thr0, const float* sample_bbox, float* det_pr,
                                       float* fitness, float* bbox_out, cudaStream_t stream) {
    DN_REQUIRE(logits && det_pr, "detect_outputs: null pointer");
    DN_REQUIRE(!bbox_out || sample_bbox, "detect_outputs: the box output needs the sample boxes");
    DN_REQUIRE(B > 0 && sn > 0 && s0 > 0, "detect_outputs: empty tensor");
    DN_REQUIRE((fit_mode & 3) != 3, "detect_outputs: joint and independent fitness exclude each other");
    DN_REQUIRE(!fit_mode || fitness, "detect_outputs: the fitness modes need the fitness output");
    DN_REQUIRE(!(fit_mode & 1) || s0 == class_num * 5 + 1, "detect_outputs: joint fitness expects classNum*5+1 channels");
    const int nroi = B * sn * sn;
    const int warps = 8;
    detect_outputs_kernel<<<DN_G(ceil_div(nroi, warps)), warps * 32, 0, stream>>>(
        logits, ld, B, sn, s0, use_bbox, class_num, fit_mode, thr0, sample_bbox, det_pr, fitness, bbox_out);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_detect_outputs(const float* logits, long long ld, int B, int sn, int s0, int use_bbox,
                                    const float* sample_bbox, float* det_pr, float* bbox_out, cudaStream_t stream) {
    return denet_detect_outputs_v2(logits, ld, B, sn, s0, use_bbox, s0 - 1, 0, 0.f, sample_bbox, det_pr, nullptr, bbox_out,
                                   stream);
}

extern "C" int denet_detections_nms(const float* det_pr, const float* fitness, long long stride_b, long long stride_c,
                                    long long stride_k, const float* bbox, const int* bbox_num, int B, int class_num,
                                    int K, float pr_threshold, float nms_threshold, int use_soft_nms, float* out_score,
                                    int* out_index, int* out_count, cudaStream_t stream) {
    DN_REQUIRE(det_pr && fitness && bbox && bbox_num && out_score && out_index && out_count, "detections_nms: null pointer");
    DN_REQUIRE(B > 0 && class_num > 0 && K > 0 && K <= kNmsMaxK, "detections_nms: need 0 < K <= %d samples per image", kNmsMaxK);
    const size_t smem = (size_t)K * (sizeof(float4) + sizeof(float) + 2 * sizeof(int));
    DN_CHECK_CUDA(cudaFuncSetAttribute(detections_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float log_thr = logf(pr_threshold);      // std::log(float) of the reference (:126) = the same libm logf
    detections_nms_kernel<<<DN_G(B * class_num), kNmsThreads, smem, stream>>>(
        det_pr, fitness, stride_b, stride_c, stride_k, bbox, bbox_num, class_num, K, log_thr, nms_threshold, use_soft_nms,
        out_score, out_index, out_count);
    DN_CHECK_LAUNCH();
    return 0;
}
