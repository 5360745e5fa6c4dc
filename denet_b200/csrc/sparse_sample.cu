// Directed-sparse-sampling RoI feature gather / scatter on NHWC feature maps.
//
// Replaces the reference GpuOps DeNetSparseOp / DeNetSparseGradOp, i.e. the inline CUDA kernels k_sparse_sample<gs>
// (denet/layer/denet_sparse_op.py:42-85) and k_sparse_sample_grad<gs> (:171-212); SURVEY.md §8 rows a12/a13.
//
// Index contract (bit-exact with the reference kernel): for grid step g in [0, gs)
//     t  = float(g) * extent            (fp32 product)
//     y  = fma(t, 1/(gs-1), y0)         (nvcc contracts the reference's  y0 + t*k  into one FMA)
//     ys = lroundf(max(0, min(H-1, y*H)))   (round half away from zero)
// Output row layout = the reference channel order: [(yi*gs + xi)*F + f] for the gs*gs grid points, then bbox_h, bbox_w.
// In NHWC this makes every grid point one contiguous F-channel copy: the reference's one-thread-per-RoI loop with
// stride-HW scalar accesses becomes coalesced vector copies (8- or 16-byte), one CTA per RoI.  The feature map and
// the gathered rows may differ in dtype (fp32 map -> bf16 GEMM operand) so no separate conversion pass is needed.
#include "common.cuh"
#include "pack.cuh"

namespace dn {

__device__ __forceinline__ int grid_index(float b0, float extent, int g, float k, int size) {
    const float t = __fmul_rn((float)g, extent);
    const float y = __fmaf_rn(t, k, b0);
    const float v = fmaxf(0.0f, fminf((float)size - 1.0f, __fmul_rn(y, (float)size)));
    return (int)lroundf(v);
}

constexpr int kMaxGrid = 16;
constexpr int kSsThreads = 256;

template <typename TI, typename TO, int VEC>
__global__ void __launch_bounds__(kSsThreads) sparse_sample_fwd_kernel(const TI* __restrict__ fmap, long long ldf,
                                                                        const float* __restrict__ bbox, int B, int F,
                                                                        int H, int W, int rois_per_image, int gs,
                                                                        TO* __restrict__ out, long long ldo) {
    const long long roi = blockIdx.x;
    const int b = (int)(roi / rois_per_image);
    __shared__ int s_off[kMaxGrid * kMaxGrid];
    __shared__ float s_hw[2];
    const float x0 = bbox[roi * 4 + 0], y0 = bbox[roi * 4 + 1], x1 = bbox[roi * 4 + 2], y1 = bbox[roi * 4 + 3];
    const float bh = y1 - y0, bw = x1 - x0;
    const float k = 1.0f / (float)(gs - 1);
    for (int gp = threadIdx.x; gp < gs * gs; gp += blockDim.x) {
        const int ys = grid_index(y0, bh, gp / gs, k, H);
        const int xs = grid_index(x0, bw, gp % gs, k, W);
        s_off[gp] = ys * W + xs;
    }
    if (threadIdx.x == 0) {
        s_hw[0] = bh;
        s_hw[1] = bw;
    }
    __syncthreads();
    const int FV = F / VEC;
    const int total = gs * gs * FV;
    const TI* src = fmap + (long long)b * H * W * ldf;
    TO* dst = out + roi * ldo;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int gp = i / FV;
        const int fv = i - gp * FV;
        Pack<TI, VEC> p;
        p.load(src + (long long)s_off[gp] * ldf + fv * VEC);
        Pack<TO, VEC> q;
#pragma unroll
        for (int j = 0; j < VEC; ++j) q.v[j] = p.v[j];
        q.store(dst + (long long)gp * F + fv * VEC);
    }
    if (threadIdx.x < 2) dst[(long long)gs * gs * F + threadIdx.x] = from_f<TO>(s_hw[threadIdx.x]);
}

// scatter-add into an fp32 accumulation map (B,H,W,F); like the reference the summation order is not fixed
template <typename T, int VEC>
__global__ void __launch_bounds__(kSsThreads) sparse_sample_bwd_kernel(const T* __restrict__ dy, long long ldo,
                                                                        const float* __restrict__ bbox, int B, int F,
                                                                        int H, int W, int rois_per_image, int gs,
                                                                        float* __restrict__ dfmap) {
    const long long roi = blockIdx.x;
    const int b = (int)(roi / rois_per_image);
    __shared__ int s_off[kMaxGrid * kMaxGrid];
    const float x0 = bbox[roi * 4 + 0], y0 = bbox[roi * 4 + 1], x1 = bbox[roi * 4 + 2], y1 = bbox[roi * 4 + 3];
    const float bh = y1 - y0, bw = x1 - x0;
    const float k = 1.0f / (float)(gs - 1);
    for (int gp = threadIdx.x; gp < gs * gs; gp += blockDim.x) {
        const int ys = grid_index(y0, bh, gp / gs, k, H);
        const int xs = grid_index(x0, bw, gp % gs, k, W);
        s_off[gp] = ys * W + xs;
    }
    __syncthreads();
    const int FV = F / VEC;
    const int total = gs * gs * FV;
    const T* src = dy + roi * ldo;
    float* dst = dfmap + (long long)b * H * W * F;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int gp = i / FV;
        const int fv = i - gp * FV;
        Pack<T, VEC> p;
        p.load(src + (long long)gp * F + fv * VEC);
        float* d = dst + (long long)s_off[gp] * F + fv * VEC;
        if constexpr (VEC == 8) {
            atomicAdd(reinterpret_cast<float4*>(d), make_float4(p.v[0], p.v[1], p.v[2], p.v[3]));
            atomicAdd(reinterpret_cast<float4*>(d) + 1, make_float4(p.v[4], p.v[5], p.v[6], p.v[7]));
        } else if constexpr (VEC == 4) {
            atomicAdd(reinterpret_cast<float4*>(d), make_float4(p.v[0], p.v[1], p.v[2], p.v[3]));
        } else {
#pragma unroll
            for (int j = 0; j < VEC; ++j) atomicAdd(d + j, p.v[j]);
        }
    }
}

__global__ void sparse_sample_index_kernel(const float* __restrict__ bbox, long long nroi, int gs, int H, int W,
                                           int* __restrict__ ys, int* __restrict__ xs) {
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= nroi * gs) return;
    const long long roi = idx / gs;
    const int g = (int)(idx % gs);
    const float x0 = bbox[roi * 4 + 0], y0 = bbox[roi * 4 + 1], x1 = bbox[roi * 4 + 2], y1 = bbox[roi * 4 + 3];
    const float k = 1.0f / (float)(gs - 1);
    ys[idx] = grid_index(y0, y1 - y0, g, k, H);
    xs[idx] = grid_index(x0, x1 - x0, g, k, W);
}

// widest legal vector: all pointers aligned to VEC elements of their type, channel count and pitches divisible
static int pick_vec(int F, long long ld_a, const void* a, int esz_a, long long ld_b, const void* b, int esz_b) {
    for (int vec = 8; vec > 1; vec >>= 1) {
        const bool ok = (F % vec == 0) && (ld_a % vec == 0) && (ld_b % vec == 0) &&
                        (reinterpret_cast<uintptr_t>(a) % (size_t)(vec * esz_a) == 0) &&
                        (reinterpret_cast<uintptr_t>(b) % (size_t)(vec * esz_b) == 0);
        if (ok && vec != 2) return vec;
    }
    return 1;
}

template <typename TI, typename TO>
static void launch_fwd(int vec, const void* fmap, long long ldf, const float* bbox, int B, int F, int H, int W, int rpi,
                       int gs, void* out, long long ldo, cudaStream_t stream) {
    const int grid = B * rpi;
    if (vec == 8)
        sparse_sample_fwd_kernel<TI, TO, 8><<<DN_G(grid), kSsThreads, 0, stream>>>((const TI*)fmap, ldf, bbox, B, F, H, W, rpi,
                                                                             gs, (TO*)out, ldo);
    else if (vec == 4)
        sparse_sample_fwd_kernel<TI, TO, 4><<<DN_G(grid), kSsThreads, 0, stream>>>((const TI*)fmap, ldf, bbox, B, F, H, W, rpi,
                                                                             gs, (TO*)out, ldo);
    else
        sparse_sample_fwd_kernel<TI, TO, 1><<<DN_G(grid), kSsThreads, 0, stream>>>((const TI*)fmap, ldf, bbox, B, F, H, W, rpi,
                                                                             gs, (TO*)out, ldo);
}

}  // namespace dn

using namespace dn;

extern "C" int denet_sparse_sample_fwd(const void* fmap, int dtype, int B, int H, int W, int F, long long ldf,
                                       const float* bbox, int rois_per_image, int gs, void* out, int out_dtype,
                                       long long ldo, cudaStream_t stream) {
    DN_REQUIRE(fmap && bbox && out, "sparse_sample_fwd: null pointer");
    DN_REQUIRE(gs >= 2 && gs <= kMaxGrid, "sparse_sample_fwd: grid size must be in [2,%d]", kMaxGrid);
    DN_REQUIRE(ldo >= (long long)gs * gs * F + 2, "sparse_sample_fwd: output pitch too small");
    if (B * rois_per_image == 0) return 0;
    const int ei = dtype == DENET_F32 ? 4 : 2, eo = out_dtype == DENET_F32 ? 4 : 2;
    const int vec = pick_vec(F, ldf, fmap, ei, ldo, out, eo);
    if (dtype == DENET_F32 && out_dtype == DENET_F32)
        launch_fwd<float, float>(vec, fmap, ldf, bbox, B, F, H, W, rois_per_image, gs, out, ldo, stream);
    else if (dtype == DENET_F32)
        launch_fwd<float, __nv_bfloat16>(vec, fmap, ldf, bbox, B, F, H, W, rois_per_image, gs, out, ldo, stream);
    else if (out_dtype == DENET_F32)
        launch_fwd<__nv_bfloat16, float>(vec, fmap, ldf, bbox, B, F, H, W, rois_per_image, gs, out, ldo, stream);
    else
        launch_fwd<__nv_bfloat16, __nv_bfloat16>(vec, fmap, ldf, bbox, B, F, H, W, rois_per_image, gs, out, ldo, stream);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_sparse_sample_bwd(const void* dy, int dtype, long long ldo, const float* bbox, int B, int H, int W,
                                       int F, int rois_per_image, int gs, float* dfmap, cudaStream_t stream) {
    DN_REQUIRE(dy && bbox && dfmap, "sparse_sample_bwd: null pointer");
    DN_REQUIRE(gs >= 2 && gs <= kMaxGrid, "sparse_sample_bwd: grid size must be in [2,%d]", kMaxGrid);
    DN_CHECK_CUDA(cudaMemsetAsync(dfmap, 0, sizeof(float) * (size_t)B * H * W * F, stream));
    if (B * rois_per_image == 0) return 0;
    const int es = dtype == DENET_F32 ? 4 : 2;
    const int vec = pick_vec(F, ldo, dy, es, F, dfmap, 4);
    const int grid = B * rois_per_image;
#define DN_SS_BWD(T, V)                                                                                              \
    sparse_sample_bwd_kernel<T, V><<<DN_G(grid), kSsThreads, 0, stream>>>((const T*)dy, ldo, bbox, B, F, H, W,             \
                                                                    rois_per_image, gs, dfmap)
    if (dtype == DENET_F32) {
        if (vec == 8) DN_SS_BWD(float, 8); else if (vec == 4) DN_SS_BWD(float, 4); else DN_SS_BWD(float, 1);
    } else {
        if (vec == 8) DN_SS_BWD(__nv_bfloat16, 8); else if (vec == 4) DN_SS_BWD(__nv_bfloat16, 4);
        else DN_SS_BWD(__nv_bfloat16, 1);
    }
#undef DN_SS_BWD
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_sparse_sample_index(const float* bbox, long long nroi, int gs, int H, int W, int* ys, int* xs,
                                         cudaStream_t stream) {
    DN_REQUIRE(bbox && ys && xs, "sparse_sample_index: null pointer");
    if (nroi == 0) return 0;
    const long long total = nroi * gs;
    sparse_sample_index_kernel<<<DN_G((unsigned)ceil_div_ll(total, 256)), 256, 0, stream>>>(bbox, nroi, gs, H, W, ys, xs);
    DN_CHECK_LAUNCH();
    return 0;
}
