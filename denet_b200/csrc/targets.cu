// Training targets of the DSS head built on the device from the ground-truth boxes (a few hundred bytes per image)
// instead of on the host (the reference fills dense numpy arrays in python loops and uploads ~11 MB per step):
//
//   corner_target  DeNetCornerLayer.get_target   denet/layer/denet_corner.py:81-123 (dropout = 0)
//   detect_target  DeNetDetectLayer.get_target   denet/layer/denet_detect.py:147-235 with the IoU matrix of
//                  common/theano_util.py:38-59 (float32, like the compiled Theano function)
//
// The arithmetic is the reference's, type for type: python `round()` on the double product (rint = half to even),
// float32 IoU with every operation rounded separately (no FMA contraction), box-regression targets computed in
// double and then stored as float32, normalisers applied as float32 divisions.  Ground-truth boxes and the RoI boxes
// therefore come in as DOUBLES (python floats in the reference).  Outputs use the reference's flattened NCHW layouts.
#include <algorithm>

#include "common.cuh"

namespace dn {

// one thread per ground-truth box; the dense map is pre-filled by corner_target_fill_kernel
__global__ void corner_target_fill_kernel(float* __restrict__ target, int cn, long long HW, long long total,
                                          float v_not) {
    // (B, 2, cn, H, W): plane 0 ("not a corner") = 1/(W*H*cn), plane 1 = 0
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long plane = (idx / (cn * HW)) & 1;
        target[idx] = plane == 0 ? v_not : 0.f;
    }
}

__global__ void corner_target_scatter_kernel(const double* __restrict__ gt, const int* __restrict__ gt_count, int B,
                                             int G, int cn, int H, int W, float v, float* __restrict__ target) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * G) return;
    const int b = idx / G, g = idx % G;
    if (g >= gt_count[b]) return;
    const double* bb = gt + (long long)idx * 4;
    // denet_corner.py:96-99  x0 = int(round(bbox[0]*W)); x1 = max(x0, int(round(bbox[2]*W)) - 1)
    const int x0 = (int)rint(__dmul_rn(bb[0], (double)W));
    const int y0 = (int)rint(__dmul_rn(bb[1], (double)H));
    int x1 = (int)rint(__dmul_rn(bb[2], (double)W)) - 1;
    int y1 = (int)rint(__dmul_rn(bb[3], (double)H)) - 1;
    x1 = x1 > x0 ? x1 : x0;
    y1 = y1 > y0 ? y1 : y0;
    const bool x0v = x0 >= 0 && x0 < W, y0v = y0 >= 0 && y0 < H, x1v = x1 >= 0 && x1 < W, y1v = y1 >= 0 && y1 < H;
    const long long HW = (long long)H * W;
    float* t0 = target + (long long)b * 2 * cn * HW;   // plane 0
    float* t1 = t0 + (long long)cn * HW;               // plane 1
    auto mark = [&](int c, int y, int x) {
        const long long o = (long long)c * HW + (long long)y * W + x;
        t1[o] = v;       // 1/(W*H*cn)
        t0[o] = 0.f;     // (1 - 1)/(W*H*cn)
    };
    if (x0v && y0v) mark(0, y0, x0);
    if (x1v && y0v) mark(1, y0, x1);
    if (x0v && y1v) mark(2, y1, x0);
    if (x1v && y1v) mark(3, y1, x1);
    if (cn == 5) {
        // :111-114  centre = round((x0+x2)*0.5*W)
        const int cx = (int)rint(__dmul_rn(__dmul_rn(__dadd_rn(bb[0], bb[2]), 0.5), (double)W));
        const int cy = (int)rint(__dmul_rn(__dmul_rn(__dadd_rn(bb[1], bb[3]), 0.5), (double)H));
        if (cx >= 0 && cx < W && cy >= 0 && cy < H) mark(4, cy, cx);
    }
}

constexpr int kMaxGt = 64;   // ground-truth boxes per image handled on the device

// one thread per RoI.  fit_mode bit0: joint fitness (denet_detect.py:58-61,179-182: classNum x 5 fitness bins + null),
// bit1: independent fitness (:100-104,187-191: a second 6-way target, bin 0 = "no object").
__global__ void detect_target_kernel(const double* __restrict__ gt, const int* __restrict__ gt_class,
                                     const int* __restrict__ gt_count, const double* __restrict__ samples, int B, int G,
                                     int sn, int class_num, float thr0, float thr1, int use_bbox, int fit_mode,
                                     double thr0_d, float* __restrict__ det, float* __restrict__ valid,
                                     float* __restrict__ reg, float* __restrict__ fit) {
    const int K = sn * sn;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)B * K) return;
    const int b = (int)(idx / K), k = (int)(idx % K);
    const bool joint = (fit_mode & 1) != 0, indfit = (fit_mode & 2) != 0;
    const int fitness_num = joint ? 5 : 6;
    const int null_class = joint ? class_num * fitness_num : class_num;
    const int s0 = null_class + 1;
    const float nfactor = (float)K;
    const double* sb = samples + idx * 4;
    const float y0 = (float)sb[0], y1 = (float)sb[1], y2 = (float)sb[2], y3 = (float)sb[3];
    const float y_area = __fmul_rn(__fsub_rn(y2, y0), __fsub_rn(y3, y1));
    int pos_cls[kMaxGt];
    int npos = 0;
    unsigned int fit_bins = 0;
    float best = 0.f;
    int best_g = -1;
    const int ng = gt_count[b];
    for (int g = 0; g < ng; ++g) {
        const double* gb = gt + ((long long)b * G + g) * 4;
        const float x0 = (float)gb[0], x1 = (float)gb[1], x2 = (float)gb[2], x3 = (float)gb[3];
        // theano_util.py:38-59 in float32, one rounding per operation
        const float x_area = __fmul_rn(__fsub_rn(x2, x0), __fsub_rn(x3, x1));
        const float dx = fmaxf(__fsub_rn(fminf(x2, y2), fmaxf(x0, y0)), 0.f);
        const float dy = fmaxf(__fsub_rn(fminf(x3, y3), fmaxf(x1, y1)), 0.f);
        const float inter = __fmul_rn(dx, dy);
        const float uni = __fsub_rn(__fadd_rn(x_area, y_area), inter);
        const float iou = __fdiv_rn(inter, uni);
        if (iou > thr0) {                       // :172-177 positives (NaN compares false like numpy)
            int c = gt_class[b * G + g];
            if (fit_mode) {
                // :177 sample_f in double (numpy float32 scalar with python floats under numpy 1.x promotion)
                const double sample_f = __ddiv_rn(__dsub_rn((double)iou, thr0_d), __dsub_rn(1.0, thr0_d));
                if (joint) {                    // :180 int() truncates toward zero
                    int f = (int)__dmul_rn((double)fitness_num, sample_f);
                    f = f < 0 ? 0 : (f > fitness_num - 1 ? fitness_num - 1 : f);
                    c = c * fitness_num + f;
                }
                if (indfit) {                   // :188-189
                    int f = 1 + (int)floor(__dmul_rn((double)(fitness_num - 1), sample_f));
                    f = f < 1 ? 1 : (f > fitness_num - 1 ? fitness_num - 1 : f);
                    fit_bins |= 1u << f;
                }
            }
            bool seen = false;
            for (int i = 0; i < npos; ++i) seen |= (pos_cls[i] == c);
            if (!seen) pos_cls[npos++] = c;
        }
        if (best_g < 0 || iou > best) {         // numpy argmax: first maximum (NaN handling differs; boxes are finite)
            best = iou;
            best_g = g;
        }
    }
    // det_pr: one-hot (possibly several classes) or null class, normalised per RoI, then / sn^2   (:229-232)
    const long long plane = (long long)K;
    float* d = det + (long long)b * s0 * plane + k;
    const float v_null = __fdiv_rn(__fdiv_rn(1.f, 1.f), nfactor);
    for (int c = 0; c < s0; ++c) d[(long long)c * plane] = 0.f;
    if (npos == 0) {
        d[(long long)null_class * plane] = v_null;
    } else {
        const float v = __fdiv_rn(__fdiv_rn(1.f, (float)npos), nfactor);
        for (int i = 0; i < npos; ++i) d[(long long)pos_cls[i] * plane] = v;
    }
    if (indfit) {
        float* f = fit + (long long)b * fitness_num * plane + k;
        const int nb = __popc(fit_bins);
        const float v = __fdiv_rn(__fdiv_rn(1.f, (float)(nb ? nb : 1)), nfactor);
        f[0] = nb ? 0.f : v;
        for (int i = 1; i < fitness_num; ++i) f[(long long)i * plane] = ((fit_bins >> i) & 1u) ? v : 0.f;
    }
    if (use_bbox) {
        float r[8] = {0.f, 0.f, 1.f, 1.f, 0.f, 0.f, 1.f, 1.f};
        float vv = 0.f;
        if (best_g >= 0 && best > thr1) {       // :198-212
            const double* t = gt + ((long long)b * G + best_g) * 4;
            vv = __fdiv_rn(1.f, nfactor);
            r[0] = (float)__dmul_rn(0.5, __dadd_rn(t[0], t[2]));
            r[1] = (float)__dmul_rn(0.5, __dadd_rn(t[1], t[3]));
            r[2] = (float)__dsub_rn(t[2], t[0]);
            r[3] = (float)__dsub_rn(t[3], t[1]);
            r[4] = (float)__dmul_rn(0.5, __dadd_rn(sb[0], sb[2]));
            r[5] = (float)__dmul_rn(0.5, __dadd_rn(sb[1], sb[3]));
            r[6] = (float)__dsub_rn(sb[2], sb[0]);
            r[7] = (float)__dsub_rn(sb[3], sb[1]);
        }
        valid[idx] = vv;
        float* rr = reg + (long long)b * 8 * plane + k;
#pragma unroll
        for (int i = 0; i < 8; ++i) rr[(long long)i * plane] = r[i];
    }
}

}  // namespace dn

using namespace dn;

extern "C" int denet_corner_target(const double* gt_bbox, const int* gt_count, int B, int G, int cn, int H, int W,
                                   float* target, cudaStream_t stream) {
    DN_REQUIRE(gt_bbox && gt_count && target, "corner_target: null pointer");
    DN_REQUIRE(cn == 4 || cn == 5, "corner_target: corner_num must be 4 or 5");
    DN_REQUIRE(B > 0 && G > 0 && H > 0 && W > 0, "corner_target: empty problem");
    const long long HW = (long long)H * W;
    const long long total = (long long)B * 2 * cn * HW;
    const float v = 1.0f / (float)((long long)W * H * cn);   // numpy: float32 array /= python int
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 16);
    corner_target_fill_kernel<<<DN_G(grid), 256, 0, stream>>>(target, cn, HW, total, v);
    corner_target_scatter_kernel<<<DN_G(ceil_div(B * G, 128)), 128, 0, stream>>>(gt_bbox, gt_count, B, G, cn, H, W, v,
                                                                                  target);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_detect_target_v2(const double* gt_bbox, const int* gt_class, const int* gt_count,
                                      const double* sample_bbox, int B, int G, int sn, int class_num, double thr0,
                                      double thr1, int use_bbox, int fit_mode, float* target_det, float* target_valid,
                                      float* target_reg, float* target_fit, cudaStream_t stream) {
    DN_REQUIRE(gt_bbox && gt_class && gt_count && sample_bbox && target_det, "detect_target: null pointer");
    DN_REQUIRE(!use_bbox || (target_valid && target_reg), "detect_target: box targets requested without buffers");
    DN_REQUIRE(!(fit_mode & 2) || target_fit, "detect_target: fitness target requested without a buffer");
    DN_REQUIRE((fit_mode & 3) != 3, "detect_target: joint and independent fitness exclude each other");
    DN_REQUIRE(G > 0 && G <= kMaxGt, "detect_target: at most %d ground-truth boxes per image (got %d)", kMaxGt, G);
    const long long total = (long long)B * sn * sn;
    detect_target_kernel<<<DN_G((int)ceil_div_ll(total, 128)), 128, 0, stream>>>(
        gt_bbox, gt_class, gt_count, sample_bbox, B, G, sn, class_num, (float)thr0, (float)thr1, use_bbox, fit_mode, thr0,
        target_det, target_valid, target_reg, target_fit);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_detect_target(const double* gt_bbox, const int* gt_class, const int* gt_count,
                                   const double* sample_bbox, int B, int G, int sn, int class_num, float thr0,
                                   float thr1, int use_bbox, float* target_det, float* target_valid, float* target_reg,
                                   cudaStream_t stream) {
    return denet_detect_target_v2(gt_bbox, gt_class, gt_count, sample_bbox, B, G, sn, class_num, thr0, thr1, use_bbox, 0,
                                  target_det, target_valid, target_reg, nullptr, stream);
}
