// Explicit im2col / col2im for the few convolutions the TMA implicit-GEMM path does not cover (stride > 1 and the
// 3-channel stem): they run as 1x1 GEMMs over the column matrix.  Column order k = (r*S + s)*C + c.
// Weight permutations between the reference filter layout (Cout, Cin, R, S) [true convolution, reference
// denet/layer/convolution.py:83] and the (Cout, R*S*Cin) matrix that multiplies the column matrix.
#include <algorithm>

#include "common.cuh"
#include "pack.cuh"

namespace dn {

template <typename T, int VEC>
__global__ void im2col_kernel(const T* __restrict__ x, int N, int H, int W, int C, long long ldx, int R, int S, int sh,
                              int sw, int ph, int pw, int Ho, int Wo, T* __restrict__ col, long long ldc) {
    const int CV = C / VEC;
    const int taps = R * S;
    const long long total = (long long)N * Ho * Wo * taps * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int tap = (int)(t % taps);
        const long long pix = t / taps;
        const int wo = (int)(pix % Wo);
        const int ho = (int)((pix / Wo) % Ho);
        const int n = (int)(pix / ((long long)Wo * Ho));
        const int h = ho * sh - ph + tap / S;
        const int w = wo * sw - pw + tap % S;
        Pack<T, VEC> p;
        if (h >= 0 && h < H && w >= 0 && w < W) {
            p.load(x + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) p.v[i] = 0.f;
        }
        p.store(col + pix * ldc + (long long)tap * C + (long long)cv * VEC);
    }
}

template <typename T>
__global__ void zero_tail_kernel(T* __restrict__ col, long long rows, int k0, long long ldc) {
    const int tail = (int)(ldc - k0);
    const long long total = rows * tail;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
        col[(idx / tail) * ldc + k0 + idx % tail] = from_f<T>(0.f);
}

// gather form of col2im (deterministic): each input pixel sums the column entries that were copied from it
template <typename T, int VEC>
__global__ void col2im_kernel(const T* __restrict__ dcol, long long ldc, int N, int H, int W, int C, long long ldx,
                              int R, int S, int sh, int sw, int ph, int pw, int Ho, int Wo, T* __restrict__ dx) {
    const int CV = C / VEC;
    const long long total = (long long)N * H * W * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int w = (int)(t % W);
        t /= W;
        const int h = (int)(t % H);
        const int n = (int)(t / H);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int r = 0; r < R; ++r) {
            const int hn = h + ph - r;
            if (hn < 0 || hn % sh != 0) continue;
            const int ho = hn / sh;
            if (ho >= Ho) continue;
            for (int s = 0; s < S; ++s) {
                const int wn = w + pw - s;
                if (wn < 0 || wn % sw != 0) continue;
                const int wo = wn / sw;
                if (wo >= Wo) continue;
                Pack<T, VEC> g;
                g.load(dcol + (((long long)n * Ho + ho) * Wo + wo) * ldc + (long long)(r * S + s) * C +
                       (long long)cv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += g.v[i];
            }
        }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
        o.store(dx + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
    }
}

// w2[co][(r*S+s)*Cin + c] = w[co][c][R-1-r][S-1-s]   (dir 0);  inverse with optional accumulation (dir 1)
__global__ void weight_im2col_perm_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin,
                                          int R, int S, int dir, int accumulate) {
    const long long total = (long long)Cout * Cin * R * S;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % Cin);
        long long t = idx / Cin;
        const int tap = (int)(t % (R * S));
        const int co = (int)(t / (R * S));
        const int r = tap / S, s = tap % S;
        const long long iw = (((long long)co * Cin + c) * R + (R - 1 - r)) * S + (S - 1 - s);
        if (dir == 0)
            dst[idx] = src[iw];
        else
            dst[iw] = accumulate ? dst[iw] + src[idx] : src[idx];
    }
}

static inline int grid_for(long long items) {
    return (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(items, 256), (long long)num_sms() * 16));
}

}  // namespace dn

using namespace dn;

extern "C" int denet_im2col(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int R, int S, int sh,
                            int sw, int ph, int pw, int Ho, int Wo, void* col, long long ldc, cudaStream_t stream) {
    DN_REQUIRE(x && col, "im2col: null pointer");
    DN_REQUIRE(ldc >= (long long)R * S * C, "im2col: column pitch too small");
    const bool v = vec8_ok(C, ldx, x) && vec8_ok(C, ldc, col);
    const long long rows = (long long)N * Ho * Wo;
    DN_DISPATCH(dtype, v, {
        im2col_kernel<T, VEC><<<DN_G(grid_for(rows * R * S * (C / VEC))), 256, 0, stream>>>(
            (const T*)x, N, H, W, C, ldx, R, S, sh, sw, ph, pw, Ho, Wo, (T*)col, ldc);
        if (ldc > (long long)R * S * C)
            zero_tail_kernel<T><<<DN_G(grid_for(rows * (ldc - (long long)R * S * C))), 256, 0, stream>>>((T*)col, rows,
                                                                                                    R * S * C, ldc);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_col2im(const void* dcol, int dtype, long long ldc, int N, int H, int W, int C, long long ldx, int R,
                            int S, int sh, int sw, int ph, int pw, int Ho, int Wo, void* dx, cudaStream_t stream) {
    DN_REQUIRE(dcol && dx, "col2im: null pointer");
    const bool v = vec8_ok(C, ldx, dx) && vec8_ok(C, ldc, dcol);
    DN_DISPATCH(dtype, v, {
        col2im_kernel<T, VEC><<<DN_G(grid_for((long long)N * H * W * (C / VEC))), 256, 0, stream>>>(
            (const T*)dcol, ldc, N, H, W, C, ldx, R, S, sh, sw, ph, pw, Ho, Wo, (T*)dx);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_weight_to_im2col(const float* w, int Cout, int Cin, int R, int S, float* w2, cudaStream_t stream) {
    DN_REQUIRE(w && w2, "weight_to_im2col: null pointer");
    weight_im2col_perm_kernel<<<DN_G(grid_for((long long)Cout * Cin * R * S)), 256, 0, stream>>>(w, w2, Cout, Cin, R, S, 0, 0);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_weight_grad_from_im2col(const float* dw2, int Cout, int Cin, int R, int S, float* dw,
                                             int accumulate, cudaStream_t stream) {
    DN_REQUIRE(dw2 && dw, "weight_grad_from_im2col: null pointer");
    weight_im2col_perm_kernel<<<DN_G(grid_for((long long)Cout * Cin * R * S)), 256, 0, stream>>>(dw2, dw, Cout, Cin, R, S, 1,
                                                                                          accumulate);
    DN_CHECK_LAUNCH();
    return 0;
}
