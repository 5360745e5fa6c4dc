// HBM-bound layers of the DeNet hot path on NHWC tensors: batch-norm (+ReLU, +residual) forward/backward, ReLU,
// add, max / average pooling, pool-inv (nearest-neighbour upsampling) and NCHW<->NHWC layout conversion.
//
// Reference semantics followed (paths relative to the reference repository):
//   batch-norm        denet/layer/batch_norm.py:47-53,75-76 (cuDNN spatial BN; EMA of mean and INVERSE STD)
//   batch-norm + relu denet/layer/batch_norm_relu.py:31-57 (k_relu, grad masked by y > 0)
//   pooling           denet/layer/pool.py:28-40 (cuDNN max / average_inc_pad)
//   pool-inv          denet/layer/pool_inv_op.py:38-63 (fwd), :144-169 (bwd: fp32 running sum in (ry, rx) order)
//   relu              denet/layer/activation.py:32-34
// All kernels are grid-stride over 8-channel packs (16 B bf16 / 32 B fp32 per access, channels contiguous) and are
// sized in multiples of the SM count; reductions are two-stage with a fixed order (deterministic).
#include <string.h>
#include <algorithm>

#include "common.cuh"
#include "pack.cuh"

namespace dn {

static inline int ew_grid(long long work_items, int block) {
    long long g = ceil_div_ll(work_items, block);
    long long cap = (long long)num_sms() * 16;
    return (int)std::max<long long>(1, std::min(g, cap));
}

// ------------------------------------------------------------------------------------------------ batch norm
// stage 1: per row-slab partial sums of (x - K) and (x - K)^2 with K = x[row 0] (shifted sums: no cancellation)
constexpr int kBnThreads = 256;

template <typename T, int VEC>
__global__ void __launch_bounds__(kBnThreads) bn_stats_partial_kernel(const T* __restrict__ x, long long M, int C,
                                                                        long long ld, int rows_per_block,
                                                                        float* __restrict__ partial) {
    // thread -> (channel pack cv, row lane rl); block covers rows [r0, r1) and channel packs [cv0, cv0 + cvt)
    const int CV = C / VEC;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    const int rlanes = kBnThreads / cvt;
    const int cv = blockIdx.y * cvt + threadIdx.x % cvt;
    const int rl = threadIdx.x / cvt;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    __shared__ float red[2 * kBnThreads * (VEC == 8 ? 8 : 1)];

    float s[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = s2[i] = 0.f;
    const bool active = (cv < CV) && (rl < rlanes);
    if (active) {
        Pack<T, VEC> k;
        k.load(x + (long long)cv * VEC);
        for (long long r = r0 + rl; r < r1; r += rlanes) {
            Pack<T, VEC> p;
            p.load(x + r * ld + (long long)cv * VEC);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float d = p.v[i] - k.v[i];
                s[i] += d;
                s2[i] += d * d;
            }
        }
    }
    // reduce over row lanes (fixed order)
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        red[(threadIdx.x * VEC + i) * 2 + 0] = s[i];
        red[(threadIdx.x * VEC + i) * 2 + 1] = s2[i];
    }
    __syncthreads();
    if (active && rl == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float a = 0.f, b = 0.f;
            for (int l = 0; l < rlanes; ++l) {
                const int t = l * cvt + threadIdx.x;
                a += red[(t * VEC + i) * 2 + 0];
                b += red[(t * VEC + i) * 2 + 1];
            }
            const int c = cv * VEC + i;
            partial[((long long)blockIdx.x * 2 + 0) * C + c] = a;
            partial[((long long)blockIdx.x * 2 + 1) * C + c] = b;
        }
    }
}

// Second stage of the two-stage reductions: partial[(slab*NQ + q)*C + c] summed over slabs in a FIXED order
// (deterministic).  Block = 32 channels x 32 slab lanes: lane j adds slabs j, j+32, ... in double, then the 32 lane
// sums are combined in lane order by the thread with slab lane 0, which returns true and holds the totals.
constexpr int kFinThreads = 1024;
template <int NQ>
__device__ __forceinline__ bool reduce_slabs(const float* __restrict__ partial, int nslabs, int C, int* c_out,
                                             double (&tot)[NQ]) {
    __shared__ double red[NQ][32][33];
    const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
    if (c < C)
        for (int s = sl; s < nslabs; s += 32)
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] += (double)partial[((long long)s * NQ + q) * C + c];
#pragma unroll
    for (int q = 0; q < NQ; ++q) red[q][sl][cl] = acc[q];
    __syncthreads();
    *c_out = c;
    if (sl != 0 || c >= C) return false;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double t = 0.0;
        for (int j = 0; j < 32; ++j) t += red[q][j][cl];
        tot[q] = t;
    }
    return true;
}

// stage 2: combine slabs in order, produce mean / inverse std and the running-statistics EMA
template <typename T>
__global__ void bn_stats_finalize_kernel(const T* __restrict__ x, const float* __restrict__ partial, int nslabs,
                                         long long M, int C, float eps, float* __restrict__ mean,
                                         float* __restrict__ invstd, float* __restrict__ run_mean,
                                         float* __restrict__ run_stdinv, float momentum) {
    int c;
    double tot[2];
    if (!reduce_slabs<2>(partial, nslabs, C, &c, tot)) return;
    const double a = tot[0], b = tot[1];
    const double k = to_f<T>(x[c]);
    const double dm = a / (double)M;
    double var = b / (double)M - dm * dm;
    if (var < 0.0) var = 0.0;
    const float m = (float)(k + dm);
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = m;
    invstd[c] = is;
    if (run_mean) run_mean[c] = momentum * run_mean[c] + (1.0f - momentum) * m;
    if (run_stdinv) run_stdinv[c] = momentum * run_stdinv[c] + (1.0f - momentum) * is;
}

// Thread layout shared by the batch-norm passes: a thread owns ONE channel pack (8 channels = 16 B) for the whole
// kernel, so the per-channel constants live in registers, and walks the rows of its slab `rlanes` apart with the
// loads of kBnUnroll rows in flight.  A warp covers 32 consecutive packs = 512 contiguous bytes of a pixel row (or of
// several rows when C < 256).
constexpr int kBnUnroll = 4;      // forward apply: 2 streams x 4 rows in flight
constexpr int kBnUnrollBwd = 4;   // backward passes: 2-3 streams x 4 rows, held as raw 16-byte packs

// Statistics handed over as raw per-channel sums (the producing convolution's epilogue accumulated them): the apply
// kernel derives mean / inverse std itself - same arithmetic as bn_finalize_sums_kernel, one thread per channel of the
// block's channel range, shared through shared memory - and the blocks of row slab 0 publish them (backward needs
// them) and update the running statistics.  Saves one tiny launch per batch-norm layer.
struct BnSums {
    const float* sum;        // null: mean / invstd are given
    const float* sqsum;
    long long M;
    float eps;
    float* mean_out;
    float* invstd_out;
    float* run_mean;
    float* run_stdinv;
    float momentum;
};

// y = [relu]( (x - mean) * (gamma * invstd) + beta [+ residual] )
template <typename T, int VEC>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const T* __restrict__ x, long long M, int C, long long ld,
                                                              int rows_per_block, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const T* __restrict__ residual, int relu,
                                                              T* __restrict__ y, const BnSums sums) {
    const int CV = C / VEC;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    const int rlanes = kBnThreads / cvt;
    const int cv = blockIdx.y * cvt + threadIdx.x % cvt;
    const int rl = threadIdx.x / cvt;
    __shared__ float s_mu[kBnThreads * (VEC == 8 ? 8 : 1)], s_is[kBnThreads * (VEC == 8 ? 8 : 1)];
    // gamma / beta do not depend on the statistics: their loads overlap the prologue's
    float a[VEC], b[VEC];
    if (cv < CV) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            a[i] = gamma[cv * VEC + i];
            b[i] = beta[cv * VEC + i];
        }
    }
    griddep_wait();                    // x, the statistics and the residual come from the kernels just before
    if (sums.sum) {
        const int c0 = blockIdx.y * cvt * VEC, nc = cvt * VEC;
        for (int j = threadIdx.x; j < nc; j += blockDim.x) {
            const int c = c0 + j;
            if (c < C) {
                const double m = (double)sums.sum[c] / (double)sums.M;
                double var = (double)sums.sqsum[c] / (double)sums.M - m * m;
                if (var < 0.0) var = 0.0;
                const float mf = (float)m;
                const float is = (float)(1.0 / sqrt(var + (double)sums.eps));
                s_mu[j] = mf;
                s_is[j] = is;
                if (blockIdx.x == 0) {
                    sums.mean_out[c] = mf;
                    sums.invstd_out[c] = is;
                    if (sums.run_mean) sums.run_mean[c] = sums.momentum * sums.run_mean[c] + (1.0f - sums.momentum) * mf;
                    if (sums.run_stdinv)
                        sums.run_stdinv[c] = sums.momentum * sums.run_stdinv[c] + (1.0f - sums.momentum) * is;
                }
            }
        }
        __syncthreads();
    }
    if (cv >= CV || rl >= rlanes) return;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    float mu[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = cv * VEC + i;
        const int j = (threadIdx.x % cvt) * VEC + i;
        mu[i] = sums.sum ? s_mu[j] : mean[c];
        a[i] = a[i] * (sums.sum ? s_is[j] : invstd[c]);
    }
    const long long coff = (long long)cv * VEC;
    for (long long r = r0 + rl; r < r1; r += (long long)rlanes * kBnUnroll) {
        Raw<T, VEC> p[kBnUnroll], q[kBnUnroll];
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const long long rr = r + (long long)u * rlanes;
            if (rr < r1) {
                p[u].load(x + rr * ld + coff);
                if (residual) q[u].load(residual + rr * ld + coff);
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const long long rr = r + (long long)u * rlanes;
            if (rr < r1) {
                Pack<T, VEC> o;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float v = (p[u].get(i) - mu[i]) * a[i] + b[i];
                    if (residual) v += q[u].get(i);
                    if (relu) v = fmaxf(v, 0.f);
                    o.v[i] = v;
                }
                o.store(y + rr * ld + coff);
            }
        }
    }
}

// backward stage 1: per-slab partial sums of dy' and dy' * xhat, dy' = dy * [y > 0] when relu
template <typename T, int VEC>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_partial_kernel(const T* __restrict__ dy, const T* __restrict__ yout,
                                                                      const T* __restrict__ x, long long M, int C,
                                                                      long long ld, int rows_per_block,
                                                                      const float* __restrict__ mean,
                                                                      const float* __restrict__ invstd, int relu,
                                                                      float* __restrict__ partial,
                                                                      const float* __restrict__ gamma,
                                                                      const float* __restrict__ beta) {
    const int CV = C / VEC;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    const int rlanes = kBnThreads / cvt;
    const int cv = blockIdx.y * cvt + threadIdx.x % cvt;
    const int rl = threadIdx.x / cvt;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    __shared__ float red[2 * kBnThreads * (VEC == 8 ? 8 : 1)];
    float s[VEC], s2[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = s2[i] = 0.f;
    const bool active = (cv < CV) && (rl < rlanes);
    if (active) {
        // relu mask: y > 0.  Without a residual input y = relu((x - mean) * (gamma*invstd) + beta), so the mask is
        // recomputed from x with the forward pass's exact expression (yout == nullptr) instead of reading y back.
        const bool mask_from_x = relu && (yout == nullptr);
        float mu[VEC], is[VEC], fa[VEC], fb[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            mu[i] = mean[cv * VEC + i];
            is[i] = invstd[cv * VEC + i];
            fa[i] = mask_from_x ? gamma[cv * VEC + i] * is[i] : 0.f;
            fb[i] = mask_from_x ? beta[cv * VEC + i] : 0.f;
        }
        const long long coff = (long long)cv * VEC;
        for (long long r = r0 + rl; r < r1; r += (long long)rlanes * kBnUnrollBwd) {
            Raw<T, VEC> g[kBnUnrollBwd], xv[kBnUnrollBwd], yo[kBnUnrollBwd];
#pragma unroll
            for (int u = 0; u < kBnUnrollBwd; ++u) {
                const long long rr = r + (long long)u * rlanes;
                if (rr < r1) {
                    g[u].load(dy + rr * ld + coff);
                    xv[u].load(x + rr * ld + coff);
                    if (relu && !mask_from_x) yo[u].load(yout + rr * ld + coff);
                }
            }
#pragma unroll
            for (int u = 0; u < kBnUnrollBwd; ++u) {
                const long long rr = r + (long long)u * rlanes;
                if (rr < r1) {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float d = g[u].get(i);
                        const float xc = xv[u].get(i) - mu[i];
                        if (mask_from_x) {
                            if (!(xc * fa[i] + fb[i] > 0.f)) d = 0.f;
                        } else if (relu && !(yo[u].get(i) > 0.f)) {
                            d = 0.f;
                        }
                        s[i] += d;
                        s2[i] += d * (xc * is[i]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        red[(threadIdx.x * VEC + i) * 2 + 0] = s[i];
        red[(threadIdx.x * VEC + i) * 2 + 1] = s2[i];
    }
    __syncthreads();
    if (active && rl == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float a = 0.f, b = 0.f;
            for (int l = 0; l < rlanes; ++l) {
                const int t = l * cvt + threadIdx.x;
                a += red[(t * VEC + i) * 2 + 0];
                b += red[(t * VEC + i) * 2 + 1];
            }
            const int c = cv * VEC + i;
            partial[((long long)blockIdx.x * 2 + 0) * C + c] = a;
            partial[((long long)blockIdx.x * 2 + 1) * C + c] = b;
        }
    }
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nslabs, int C,
                                       float* __restrict__ sum_dy, float* __restrict__ sum_dy_xhat,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    int c;
    double tot[2];
    if (!reduce_slabs<2>(partial, nslabs, C, &c, tot)) return;
    const double a = tot[0], b = tot[1];
    sum_dy[c] = (float)a;
    sum_dy_xhat[c] = (float)b;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)a : (float)a;
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)b : (float)b;
}

// backward stage 2: dx = gamma*invstd * (dy' - sum_dy/M - xhat * sum_dy_xhat/M); optionally export dy' (residual grad)
template <typename T, int VEC>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ yout,
                                                                  const T* __restrict__ x, long long M, int C,
                                                                  long long ld, int rows_per_block,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ sum_dy,
                                                                  const float* __restrict__ sum_dy_xhat, int relu,
                                                                  T* __restrict__ dx, T* __restrict__ dres,
                                                                  const float* __restrict__ beta,
                                                                  float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                  int accumulate) {
    const int CV = C / VEC;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    const int rlanes = kBnThreads / cvt;
    const int cv = blockIdx.y * cvt + threadIdx.x % cvt;
    const int rl = threadIdx.x / cvt;
    if (cv >= CV || rl >= rlanes) return;
    if (dgamma && blockIdx.x == 0 && rl == 0) {
        // the sums arrived complete (dgrad epilogue): they ARE the parameter gradients (no finalize launch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = cv * VEC + i;
            dgamma[c] = accumulate ? dgamma[c] + sum_dy_xhat[c] : sum_dy_xhat[c];
            dbeta[c] = accumulate ? dbeta[c] + sum_dy[c] : sum_dy[c];
        }
    }
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    const float inv_m = 1.0f / (float)M;
    const bool mask_from_x = relu && (yout == nullptr);
    float mu[VEC], is[VEC], k1[VEC], c1[VEC], c2[VEC], fb[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = cv * VEC + i;
        mu[i] = mean[c];
        is[i] = invstd[c];
        k1[i] = gamma[c] * is[i];
        c1[i] = sum_dy[c] * inv_m;
        c2[i] = sum_dy_xhat[c] * inv_m;
        fb[i] = mask_from_x ? beta[c] : 0.f;
    }
    const long long coff = (long long)cv * VEC;
    for (long long r = r0 + rl; r < r1; r += (long long)rlanes * kBnUnrollBwd) {
        Raw<T, VEC> g[kBnUnrollBwd], xv[kBnUnrollBwd], yo[kBnUnrollBwd];
#pragma unroll
        for (int u = 0; u < kBnUnrollBwd; ++u) {
            const long long rr = r + (long long)u * rlanes;
            if (rr < r1) {
                g[u].load(dy + rr * ld + coff);
                xv[u].load(x + rr * ld + coff);
                if (relu && !mask_from_x) yo[u].load(yout + rr * ld + coff);
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnrollBwd; ++u) {
            const long long rr = r + (long long)u * rlanes;
            if (rr < r1) {
                Pack<T, VEC> o, gm;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    float d = g[u].get(i);
                    const float xc = xv[u].get(i) - mu[i];
                    if (mask_from_x) {   // k1 = gamma * invstd: the forward pass's (x - mean) * (gamma*invstd) + beta
                        if (!(xc * k1[i] + fb[i] > 0.f)) d = 0.f;
                    } else if (relu && !(yo[u].get(i) > 0.f)) {
                        d = 0.f;
                    }
                    gm.v[i] = d;
                    const float xh = xc * is[i];
                    o.v[i] = k1[i] * (d - c1[i] - xh * c2[i]);
                }
                o.store(dx + rr * ld + coff);
                if (dres) gm.store(dres + rr * ld + coff);
            }
        }
    }
}

// ---- batch-norm backward as ONE launch (the three kernels above cost two launch gaps and a 10 us finalize per layer:
// 84 extra launches per DeNet-34 step).  All blocks are co-resident (grid sized from the occupancy of this kernel), so
// the two reductions are separated by grid-wide barriers on module-scope counters instead of kernel boundaries:
//   phase 1  per-slab partial sums of dy' and dy' * xhat            (same arithmetic as bn_bwd_partial_kernel)
//   barrier  -> the first ceil(C/32) blocks add the slabs in a FIXED order, in double (reduce_slabs' order: 8 slab
//            lanes, then lanes in order): deterministic, independent of which block runs when
//   barrier  -> phase 2: dx (and the residual gradient) over the SAME rows the block has just read (L2 / L1 hits)
// The counters live in a module-scope array (zero at load); the last block to finish resets its slot, so a slot is zero
// again before the stream can launch the next kernel that uses it.
constexpr int kBnSyncSlots = 64;
__device__ unsigned int g_bn_sync[kBnSyncSlots][4];

// Release / acquire at GPU scope around a relaxed counter (the block's writes reach thread 0 through bar.sync; the data the
// other blocks publish - slab partials, totals - is read with L1-bypassing loads).  __threadfence() + a volatile poll
// compiled to MEMBAR.SC.GPU + CGAERRBAR + CCTL.IVALL twice per barrier and system-scope polling loads: ~4 us per barrier.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int v;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < expected);
    }
    __syncthreads();
}

template <typename T, int VEC, int U>
__global__ void __launch_bounds__(kBnThreads, 2)
    bn_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ yout, const T* __restrict__ x, long long M, int C,
                        long long ld, int rows_per_block, const float* __restrict__ mean,
                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                        const float* __restrict__ beta, int relu, T* __restrict__ dx, T* __restrict__ dres,
                        float* __restrict__ partial, float* __restrict__ sums, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, int accumulate, int sync_slot, int wide_totals, int debug) {
    const int CV = C / VEC;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    const int rlanes = kBnThreads / cvt;
    const int cv = blockIdx.y * cvt + threadIdx.x % cvt;
    const int rl = threadIdx.x / cvt;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    const int nslabs = gridDim.x;
    const unsigned int nblocks = gridDim.x * gridDim.y;
    const unsigned int bid = blockIdx.y * gridDim.x + blockIdx.x;
    unsigned int* sync = g_bn_sync[sync_slot];
    __shared__ float red[2 * kBnThreads * (VEC == 8 ? 8 : 1)];
    __shared__ double fin[2][8][33];
    __shared__ double fin8[2][8][8];
    const bool active = (cv < CV) && (rl < rlanes);
    const bool mask_from_x = relu && (yout == nullptr);
    const long long coff = (long long)cv * VEC;
    griddep_wait();
    float mu[VEC], is[VEC], k1[VEC], fb[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = cv * VEC + i;
        mu[i] = active ? mean[c] : 0.f;
        is[i] = active ? invstd[c] : 0.f;
        k1[i] = active ? gamma[c] * is[i] : 0.f;                      // the forward pass's gamma * invstd
        fb[i] = (active && mask_from_x) ? beta[c] : 0.f;
    }
    // ---------------------------------------------------------------- phase 1
    {
        float s[VEC], s2[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) s[i] = s2[i] = 0.f;
        if (active) {
            for (long long r = r0 + rl; r < r1; r += (long long)rlanes * U) {
                Raw<T, VEC> g[U], xv[U], yo[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long rr = r + (long long)u * rlanes;
                    if (rr < r1) {
                        g[u].load(dy + rr * ld + coff);
                        xv[u].load(x + rr * ld + coff);
                        if (relu && !mask_from_x) yo[u].load(yout + rr * ld + coff);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const long long rr = r + (long long)u * rlanes;
                    if (rr < r1) {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) {
                            float d = g[u].get(i);
                            const float xc = xv[u].get(i) - mu[i];
                            if (mask_from_x) {
                                if (!(xc * k1[i] + fb[i] > 0.f)) d = 0.f;
                            } else if (relu && !(yo[u].get(i) > 0.f)) {
                                d = 0.f;
                            }
                            s[i] += d;
                            s2[i] += d * (xc * is[i]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            red[(threadIdx.x * VEC + i) * 2 + 0] = s[i];
            red[(threadIdx.x * VEC + i) * 2 + 1] = s2[i];
        }
        __syncthreads();
        if (active && rl == 0) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float a = 0.f, b = 0.f;
                for (int l = 0; l < rlanes; ++l) {
                    const int t = l * cvt + threadIdx.x;
                    a += red[(t * VEC + i) * 2 + 0];
                    b += red[(t * VEC + i) * 2 + 1];
                }
                const int c = cv * VEC + i;
                partial[((long long)blockIdx.x * 2 + 0) * C + c] = a;
                partial[((long long)blockIdx.x * 2 + 1) * C + c] = b;
            }
        }
    }
    if (!(debug & 1)) grid_barrier(&sync[0], nblocks);
    // ---------------------------------------------------------------- slab totals (fixed order, double)
    // wide form: a block totals 8 channels with 32 slab lanes, every lane's <= 10 loads issued before the first add
    // (the 8-lane form below walks 37 slabs per thread: a chain of L2 latencies while every other block waits)
    if (wide_totals) {
        constexpr int KU = 5;
        for (unsigned int cb = bid; cb * 8u < (unsigned int)C; cb += nblocks) {
            const int cl = threadIdx.x & 7, sl = threadIdx.x >> 3;
            const int c = cb * 8 + cl;
            double a0 = 0.0, a1 = 0.0;
            if (c < C) {
                for (int sb0 = sl; sb0 < nslabs; sb0 += 32 * KU) {
                    float v0[KU], v1[KU];
#pragma unroll
                    for (int k = 0; k < KU; ++k) {
                        const int sb = sb0 + 32 * k;
                        v0[k] = sb < nslabs ? __ldcg(partial + ((long long)sb * 2 + 0) * C + c) : 0.f;
                        v1[k] = sb < nslabs ? __ldcg(partial + ((long long)sb * 2 + 1) * C + c) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < KU; ++k) {
                        a0 += (double)v0[k];
                        a1 += (double)v1[k];
                    }
                }
            }
            // lanes of a warp: 4 slab lanes x 8 channels -> fixed tree over the slab lanes, then the 8 warps in order
            a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
            if ((threadIdx.x & 31) < 8) {
                fin8[0][threadIdx.x >> 5][cl] = a0;
                fin8[1][threadIdx.x >> 5][cl] = a1;
            }
            __syncthreads();
            if (threadIdx.x < 8 && c < C) {
                double t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    t0 += fin8[0][j][cl];
                    t1 += fin8[1][j][cl];
                }
                sums[c] = (float)t0;
                sums[C + c] = (float)t1;
                if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)t0 : (float)t0;
                if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)t1 : (float)t1;
            }
            __syncthreads();
        }
    } else
    for (unsigned int cb = bid; cb * 32u < (unsigned int)C; cb += nblocks) {
        const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
        const int c = cb * 32 + cl;
        double a0 = 0.0, a1 = 0.0;
        if (c < C)
            for (int sb = sl; sb < nslabs; sb += 8) {
                a0 += (double)partial[((long long)sb * 2 + 0) * C + c];
                a1 += (double)partial[((long long)sb * 2 + 1) * C + c];
            }
        fin[0][sl][cl] = a0;
        fin[1][sl][cl] = a1;
        __syncthreads();
        if (sl == 0 && c < C) {
            double t0 = 0.0, t1 = 0.0;
            for (int j = 0; j < 8; ++j) {
                t0 += fin[0][j][cl];
                t1 += fin[1][j][cl];
            }
            sums[c] = (float)t0;
            sums[C + c] = (float)t1;
            if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)t0 : (float)t0;
            if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)t1 : (float)t1;
        }
        __syncthreads();
    }
    if (!(debug & 1)) grid_barrier(&sync[1], nblocks);
    // ---------------------------------------------------------------- phase 2
    if (active && !(debug & 2)) {
        const float inv_m = 1.0f / (float)M;
        float c1[VEC], c2[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = cv * VEC + i;
            c1[i] = __ldcg(sums + c) * inv_m;            // written by another block after kernel start: bypass L1
            c2[i] = __ldcg(sums + C + c) * inv_m;
        }
        for (long long r = r0 + rl; r < r1; r += (long long)rlanes * U) {
            Raw<T, VEC> g[U], xv[U], yo[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long rr = r + (long long)u * rlanes;
                if (rr < r1) {
                    g[u].load(dy + rr * ld + coff);
                    xv[u].load(x + rr * ld + coff);
                    if (relu && !mask_from_x) yo[u].load(yout + rr * ld + coff);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long rr = r + (long long)u * rlanes;
                if (rr < r1) {
                    Pack<T, VEC> o, gm;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        float d = g[u].get(i);
                        const float xc = xv[u].get(i) - mu[i];
                        if (mask_from_x) {
                            if (!(xc * k1[i] + fb[i] > 0.f)) d = 0.f;
                        } else if (relu && !(yo[u].get(i) > 0.f)) {
                            d = 0.f;
                        }
                        gm.v[i] = d;
                        const float xh = xc * is[i];
                        o.v[i] = k1[i] * (d - c1[i] - xh * c2[i]);
                    }
                    o.store(dx + rr * ld + coff);
                    if (dres) gm.store(dres + rr * ld + coff);
                }
            }
        }
    }
    // the last block out re-arms the slot
    __syncthreads();
    if (threadIdx.x == 0 && !(debug & 1)) {
        const unsigned int ticket = atomicAdd(&sync[2], 1u);
        if (ticket == nblocks - 1) {
            sync[0] = 0;
            sync[1] = 0;
            sync[2] = 0;
            __threadfence();
        }
    }
}

__global__ void bn_inference_invstd_kernel(const float* __restrict__ run_stdinv, float eps, float* __restrict__ out,
                                           int C) {
    // batch_norm.py:50-52: var = (1/stdinv)^2 is handed to cuDNN inference, which adds eps again
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sd = 1.0f / run_stdinv[c];
    out[c] = 1.0f / sqrtf(sd * sd + eps);
}

// ------------------------------------------------------------------------------------------------ relu / add
template <typename T, int VEC>
__global__ void relu_fwd_kernel(const T* __restrict__ x, long long M, int C, long long ld, T* __restrict__ y) {
    const int CV = C / VEC;
    const long long total = M * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long off = (idx / CV) * ld + (idx % CV) * VEC;
        Pack<T, VEC> p;
        p.load(x + off);
#pragma unroll
        for (int i = 0; i < VEC; ++i) p.v[i] = fmaxf(p.v[i], 0.f);
        p.store(y + off);
    }
}

template <typename T, int VEC>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, long long M, int C, long long ld,
                                T* __restrict__ dx) {
    const int CV = C / VEC;
    const long long total = M * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long off = (idx / CV) * ld + (idx % CV) * VEC;
        Pack<T, VEC> g, o;
        g.load(dy + off);
        o.load(y + off);
#pragma unroll
        for (int i = 0; i < VEC; ++i) g.v[i] = o.v[i] > 0.f ? g.v[i] : 0.f;
        g.store(dx + off);
    }
}

template <typename T, int VEC>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, long long M, int C, long long ld, int relu,
                           T* __restrict__ out) {
    const int CV = C / VEC;
    const long long total = M * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long off = (idx / CV) * ld + (idx % CV) * VEC;
        Pack<T, VEC> p, q;
        p.load(a + off);
        q.load(b + off);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            p.v[i] += q.v[i];
            if (relu) p.v[i] = fmaxf(p.v[i], 0.f);
        }
        p.store(out + off);
    }
}

// ------------------------------------------------------------------------------------------------ layout
// NCHW fp32 (host-facing layout of the reference) -> NHWC T with pitch ld; tiled transpose through shared memory
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int C, long long HW, long long ld, T* __restrict__ y) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const float* xs = x + (long long)n * C * HW;
    T* ys = y + (long long)n * HW * ld;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && p < HW) ? xs[(long long)c * HW + p] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long p = p0 + i;
        const int c = c0 + threadIdx.x;
        if (p < HW && c < C) ys[p * ld + c] = from_f<T>(tile[threadIdx.x][i]);
    }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int C, long long HW, long long ld, float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long long p0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const T* xs = x + (long long)n * HW * ld;
    float* ys = y + (long long)n * C * HW;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long p = p0 + i;
        const int c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (p < HW && c < C) ? to_f<T>(xs[p * ld + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long p = p0 + threadIdx.x;
        if (c < C && p < HW) ys[(long long)c * HW + p] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------ pooling
template <typename T, int VEC>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, long long ldx, int kh, int kw,
                                   int sh, int sw, int ph, int pw, int Ho, int Wo, long long ldy, T* __restrict__ y,
                                   uint8_t* __restrict__ argmax) {
    const int CV = C / VEC;
    const long long total = (long long)N * Ho * Wo * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float best[VEC];
        int bi[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            best[i] = -INFINITY;
            bi[i] = 0;
        }
        for (int r = 0; r < kh; ++r) {
            const int h = ho * sh - ph + r;
            if (h < 0 || h >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int w = wo * sw - pw + s;
                if (w < 0 || w >= W) continue;
                Pack<T, VEC> p;
                p.load(x + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    if (p.v[i] > best[i]) {
                        best[i] = p.v[i];
                        bi[i] = r * kw + s;
                    }
            }
        }
        const long long opix = ((long long)n * Ho + ho) * Wo + wo;
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            o.v[i] = best[i];
            argmax[opix * C + cv * VEC + i] = (uint8_t)bi[i];
        }
        o.store(y + opix * ldy + (long long)cv * VEC);
    }
}

// gather form (deterministic): every input pixel sums dy of the windows whose recorded argmax points at it.
// The windows covering input row h are ho in [ceil((h+ph-kh+1)/sh), floor((h+ph)/sh)] (at most ceil(kh/sh) of them,
// 2 x 2 for the 3x3 stride-2 stem pool); they are enumerated directly - no scan over the kh*kw taps - and the tap
// bytes and dy packs of all of them are requested before the first use, so that the loads overlap.
template <typename T, int VEC>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ argmax, int N, int H, int W,
                                   int C, long long ldx, int kh, int kw, int sh, int sw, int ph, int pw, int Ho, int Wo,
                                   long long ldy, T* __restrict__ dx) {
    const int CV = C / VEC;
    const long long total = (long long)N * H * W * CV;
    const bool small = total < 0x7fffffffLL;
    const bool quad = (kh + sh - 1) / sh <= 2 && (kw + sw - 1) / sw <= 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int cv, w, h, n;
        if (small) {
            unsigned u = (unsigned)idx;
            cv = (int)(u % (unsigned)CV); u /= (unsigned)CV;
            w = (int)(u % (unsigned)W); u /= (unsigned)W;
            h = (int)(u % (unsigned)H);
            n = (int)(u / (unsigned)H);
        } else {
            cv = (int)(idx % CV);
            long long t = idx / CV;
            w = (int)(t % W);
            t /= W;
            h = (int)(t % H);
            n = (int)(t / H);
        }
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        const int hp = h + ph, wp = w + pw;
        int ho_hi = hp / sh, wo_hi = wp / sw;
        if (ho_hi > Ho - 1) ho_hi = Ho - 1;
        if (wo_hi > Wo - 1) wo_hi = Wo - 1;
        const int hlo = hp - kh + 1, wlo = wp - kw + 1;
        const int ho_lo = hlo <= 0 ? 0 : (hlo + sh - 1) / sh;
        const int wo_lo = wlo <= 0 ? 0 : (wlo + sw - 1) / sw;
        if (VEC == 8 && quad) {
            uint2 am[4];
            Raw<T, VEC> g[4];
            uint32_t tapw[4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int ho = ho_hi - a, wo = wo_hi - b;
                    const bool ok = ho >= ho_lo && wo >= wo_lo;
                    const int q = 2 * a + b;
                    if (ok) {
                        const long long opix = ((long long)n * Ho + ho) * Wo + wo;
                        tapw[q] = (uint32_t)((hp - ho * sh) * kw + (wp - wo * sw)) * 0x01010101u;
                        am[q] = *reinterpret_cast<const uint2*>(argmax + opix * C + cv * 8);
                        g[q].load(dy + opix * ldy + (long long)cv * VEC);
                    } else {
                        tapw[q] = 0;
                        am[q] = make_uint2(0xffffffffu, 0xffffffffu);      // matches no tap
                        g[q].load(dy);                                      // never used
                    }
                }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // byte-wise compare of the 8 recorded taps with this pixel's tap: 0xff per matching channel
                const uint32_t m0 = __vcmpeq4(am[q].x, tapw[q]), m1 = __vcmpeq4(am[q].y, tapw[q]);
                if ((m0 | m1) == 0) continue;
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    if (((i < 4 ? m0 : m1) >> (8 * (i & 3))) & 1u) acc[i] += g[q].get(i);
            }
        } else {
            for (int ho = ho_hi; ho >= ho_lo; --ho)
                for (int wo = wo_hi; wo >= wo_lo; --wo) {
                    const long long opix = ((long long)n * Ho + ho) * Wo + wo;
                    const int tap = (hp - ho * sh) * kw + (wp - wo * sw);
                    Pack<T, VEC> g;
                    g.load(dy + opix * ldy + (long long)cv * VEC);
#pragma unroll
                    for (int i = 0; i < VEC; ++i)
                        if (argmax[opix * C + cv * VEC + i] == tap) acc[i] += g.v[i];
                }
        }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
        o.store(dx + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
    }
}

// 3x3 / stride 2 / pad 1 max-pool backward (the ResNet stem pool), bf16 x 8 channels: one thread per 2x2 quad of input
// pixels.  The quad (2i..2i+1, 2j..2j+1) is covered by exactly the windows (i..i+1, j..j+1): 4 tap words + 4 dy packs
// are loaded once and serve the 9 (window, pixel) pairs of the quad, and the index arithmetic is done once per quad
// (the per-pixel kernel was instruction-bound: ~250 instructions per 16-byte store).  Same sums in the same order as
// maxpool_bwd_kernel (windows in decreasing ho, then decreasing wo).
__global__ void __launch_bounds__(256) maxpool_bwd_3x3s2_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                 const uint8_t* __restrict__ argmax, int N, int H, int W,
                                                                 int C, long long ldx, int Ho, int Wo, long long ldy,
                                                                 __nv_bfloat16* __restrict__ dx) {
    const int CV = C / 8;
    const int QH = (H + 1) / 2, QW = (W + 1) / 2;
    const unsigned total = (unsigned)N * QH * QW * CV;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        unsigned u = idx;
        const int cv = (int)(u % (unsigned)CV); u /= (unsigned)CV;
        const int j = (int)(u % (unsigned)QW); u /= (unsigned)QW;
        const int i = (int)(u % (unsigned)QH);
        const int n = (int)(u / (unsigned)QH);
        uint2 am[4];
        Raw<__nv_bfloat16, 8> g[4];
        bool ok[4];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int q = 2 * a + b;
                const int ho = i + a, wo = j + b;
                ok[q] = ho < Ho && wo < Wo;
                const long long opix = ((long long)n * Ho + (ok[q] ? ho : 0)) * Wo + (ok[q] ? wo : 0);
                am[q] = ok[q] ? *reinterpret_cast<const uint2*>(argmax + opix * C + cv * 8) : make_uint2(~0u, ~0u);
                g[q].load(dy + opix * ldy + (long long)cv * 8);
            }
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                const int h = 2 * i + py, w = 2 * j + px;
                if (h >= H || w >= W) continue;
                float acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = 0.f;
                // windows of pixel (py, px): a in {0} (py == 0) or {1, 0} (py == 1), likewise b; tap = r*3 + s with
                // r = py + 1 - 2a, s = px + 1 - 2b
#pragma unroll
                for (int a = 1; a >= 0; --a)
#pragma unroll
                    for (int b = 1; b >= 0; --b) {
                        const int r = py + 1 - 2 * a, sx = px + 1 - 2 * b;
                        if (r < 0 || sx < 0) continue;                 // compile-time after unrolling
                        const int q = 2 * a + b;
                        const uint32_t tapw = (uint32_t)(r * 3 + sx) * 0x01010101u;
                        const uint32_t m0 = __vcmpeq4(am[q].x, tapw), m1 = __vcmpeq4(am[q].y, tapw);
                        if ((m0 | m1) == 0) continue;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (((c < 4 ? m0 : m1) >> (8 * (c & 3))) & 1u) acc[c] += g[q].get(c);
                    }
                Pack<__nv_bfloat16, 8> o;
#pragma unroll
                for (int c = 0; c < 8; ++c) o.v[c] = acc[c];
                o.store(dx + (((long long)n * H + h) * W + w) * ldx + (long long)cv * 8);
            }
    }
}

// average_inc_pad: divisor is always kh*kw
template <typename T, int VEC>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, long long ldx, int kh, int kw,
                                   int sh, int sw, int ph, int pw, int Ho, int Wo, long long ldy, T* __restrict__ y) {
    const int CV = C / VEC;
    const long long total = (long long)N * Ho * Wo * CV;
    const float inv = 1.0f / (float)(kh * kw);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int wo = (int)(t % Wo);
        t /= Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        float acc[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
        for (int r = 0; r < kh; ++r) {
            const int h = ho * sh - ph + r;
            if (h < 0 || h >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int w = wo * sw - pw + s;
                if (w < 0 || w >= W) continue;
                Pack<T, VEC> p;
                p.load(x + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += p.v[i];
            }
        }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = acc[i] * inv;
        o.store(y + (((long long)n * Ho + ho) * Wo + wo) * ldy + (long long)cv * VEC);
    }
}

template <typename T, int VEC>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, int N, int H, int W, int C, long long ldx, int kh, int kw,
                                   int sh, int sw, int ph, int pw, int Ho, int Wo, long long ldy, T* __restrict__ dx) {
    const int CV = C / VEC;
    const long long total = (long long)N * H * W * CV;
    const float inv = 1.0f / (float)(kh * kw);
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
        for (int r = 0; r < kh; ++r) {
            const int hn = h + ph - r;
            if (hn < 0 || hn % sh != 0) continue;
            const int ho = hn / sh;
            if (ho >= Ho) continue;
            for (int s = 0; s < kw; ++s) {
                const int wn = w + pw - s;
                if (wn < 0 || wn % sw != 0) continue;
                const int wo = wn / sw;
                if (wo >= Wo) continue;
                Pack<T, VEC> g;
                g.load(dy + (((long long)n * Ho + ho) * Wo + wo) * ldy + (long long)cv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += g.v[i];
            }
        }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = acc[i] * inv;
        o.store(dx + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
    }
}

// pool-inv forward: one thread per OUTPUT pixel pack (coalesced writes; the 4 reads of an input pack hit L1/L2)
template <typename T, int VEC>
__global__ void pool_inv_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, long long ldx, int sw, int sh,
                                    long long ldy, T* __restrict__ y) {
    const int CV = C / VEC;
    const int RH = H * sh, RW = W * sw;
    const long long total = (long long)N * RH * RW * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int rx = (int)(t % RW);
        t /= RW;
        const int ry = (int)(t % RH);
        const int n = (int)(t / RH);
        Pack<T, VEC> p;
        p.load(x + (((long long)n * H + ry / sh) * W + rx / sw) * ldx + (long long)cv * VEC);
        p.store(y + (((long long)n * RH + ry) * RW + rx) * ldy + (long long)cv * VEC);
    }
}

// pool-inv backward: fp32 running sum over the sh x sw window in (ry, rx) order starting from 0 (reference order)
template <typename T, int VEC>
__global__ void pool_inv_bwd_kernel(const T* __restrict__ dy, int N, int H, int W, int C, long long ldx, int sw, int sh,
                                    long long ldy, T* __restrict__ dx) {
    const int CV = C / VEC;
    const int RH = H * sh, RW = W * sw;
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
        for (int ry = h * sh; ry < h * sh + sh; ++ry)
            for (int rx = w * sw; rx < w * sw + sw; ++rx) {
                Pack<T, VEC> g;
                g.load(dy + (((long long)n * RH + ry) * RW + rx) * ldy + (long long)cv * VEC);
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[i] += g.v[i];
            }
        Pack<T, VEC> o;
#pragma unroll
        for (int i = 0; i < VEC; ++i) o.v[i] = acc[i];
        o.store(dx + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
    }
}

// zero-insertion upsampling used by the strided dgrad: out[n, h*sh, w*sw, :] = x[n, h, w, :], zero elsewhere.
// One thread per OUTPUT pack so every byte of `out` is written exactly once (no separate memset pass).
template <typename T, int VEC>
__global__ void dilate_kernel(const T* __restrict__ x, int N, int H, int W, int C, long long ldx, int sh, int sw, int Hd,
                              int Wd, long long ldy, T* __restrict__ y, const T* __restrict__ add) {
    const int CV = C / VEC;
    const long long total = (long long)N * Hd * Wd * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long t = idx / CV;
        const int wd = (int)(t % Wd);
        t /= Wd;
        const int hd = (int)(t % Hd);
        const int n = (int)(t / Hd);
        Pack<T, VEC> p;
        const int h = hd / sh, w = wd / sw;
        if (hd % sh == 0 && wd % sw == 0 && h < H && w < W) {
            p.load(x + (((long long)n * H + h) * W + w) * ldx + (long long)cv * VEC);
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) p.v[i] = 0.f;
        }
        const long long o = (((long long)n * Hd + hd) * Wd + wd) * ldy + (long long)cv * VEC;
        if (add) {      // (the other branch's gradient: saves the separate add pass over the full tensor)
            Pack<T, VEC> q;
            q.load(add + o);
#pragma unroll
            for (int i = 0; i < VEC; ++i) p.v[i] += q.v[i];
        }
        p.store(y + o);
    }
}

// batch statistics from the per-channel sum / sum-of-squares accumulated by the conv epilogue (throughput mode)
__global__ void bn_finalize_sums_kernel(const float* __restrict__ sum, const float* __restrict__ sqsum, long long M,
                                        int C, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                                        float* __restrict__ run_mean, float* __restrict__ run_stdinv, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = (double)sum[c] / (double)M;
    double var = (double)sqsum[c] / (double)M - m * m;
    if (var < 0.0) var = 0.0;
    const float mf = (float)m;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = mf;
    invstd[c] = is;
    if (run_mean) run_mean[c] = momentum * run_mean[c] + (1.0f - momentum) * mf;
    if (run_stdinv) run_stdinv[c] = momentum * run_stdinv[c] + (1.0f - momentum) * is;
}

// dtype conversion of a pitched NHWC tensor (fp32 <-> bf16), e.g. fp32 master gradient -> bf16 dY operand
template <typename TI, typename TO, int VEC>
__global__ void convert_kernel(const TI* __restrict__ x, long long M, int C, long long ldx, long long ldy,
                               TO* __restrict__ y) {
    const int CV = C / VEC;
    const long long total = M * CV;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / CV;
        const int cv = (int)(idx % CV);
        Pack<TI, VEC> p;
        p.load(x + r * ldx + (long long)cv * VEC);
        Pack<TO, VEC> q;
#pragma unroll
        for (int i = 0; i < VEC; ++i) q.v[i] = p.v[i];
        q.store(y + r * ldy + (long long)cv * VEC);
    }
}

// per-channel sum over rows (bias gradient of a convolution: sum over pixels of dy), two-stage, fixed order
template <typename T, int VEC>
__global__ void __launch_bounds__(kBnThreads) colsum_partial_kernel(const T* __restrict__ x, long long M, int C,
                                                                      long long ld, int rows_per_block,
                                                                      float* __restrict__ partial) {
    const int CV = C / VEC;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    const int rlanes = kBnThreads / cvt;
    const int cv = blockIdx.y * cvt + threadIdx.x % cvt;
    const int rl = threadIdx.x / cvt;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    __shared__ float red[kBnThreads * (VEC == 8 ? 8 : 1)];
    float s[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) s[i] = 0.f;
    const bool active = (cv < CV) && (rl < rlanes);
    if (active) {
        for (long long r = r0 + rl; r < r1; r += rlanes) {
            Pack<T, VEC> p;
            p.load(x + r * ld + (long long)cv * VEC);
#pragma unroll
            for (int i = 0; i < VEC; ++i) s[i] += p.v[i];
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) red[threadIdx.x * VEC + i] = s[i];
    __syncthreads();
    if (active && rl == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float a = 0.f;
            for (int l = 0; l < rlanes; ++l) a += red[(l * cvt + threadIdx.x) * VEC + i];
            partial[(long long)blockIdx.x * C + cv * VEC + i] = a;
        }
    }
}

__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nslabs, int C, float* __restrict__ out,
                                       int accumulate) {
    int c;
    double tot[1];
    if (!reduce_slabs<1>(partial, nslabs, C, &c, tot)) return;
    const double a = tot[0];
    out[c] = accumulate ? out[c] + (float)a : (float)a;
}

static int g_bn_fused_bwd = 1;     // batch-norm backward as one launch with grid barriers (A/B: denet_bn_set_mode)
static int g_bn_debug = 0;         // profiling knobs of the fused backward (results are garbage): 1 no grid barriers, 2 no phase 2
static int g_bn_wave = 1;          // apply passes as ONE wave of blocks (3 per SM), wide slab totals in the fused backward

// slabs for a grid of at most `max_blocks` co-resident blocks (rows split as evenly as the row-lane quantum allows)
static int bn_slabs_for(long long M, int C, int vec, int max_blocks, int* rows_per_block, int* ychunks) {
    const int CV = C / vec;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    *ychunks = ceil_div(CV, cvt);
    const int rlanes = kBnThreads / cvt;
    long long target = max_blocks / *ychunks;
    if (target < 1) target = 1;
    long long rpb = ceil_div_ll(M, target);
    if (rpb < rlanes * 4) rpb = rlanes * 4;
    *rows_per_block = (int)std::min<long long>(rpb, 1 << 30);
    return (int)ceil_div_ll(M, *rows_per_block);
}

static int bn_slabs(long long M, int C, int vec, int* rows_per_block, int* ychunks) {
    const int CV = C / vec;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    *ychunks = ceil_div(CV, cvt);
    long long target = (long long)num_sms() * 4;
    long long rpb = ceil_div_ll(M, target);
    const int rlanes = kBnThreads / cvt;
    if (rpb < rlanes * 4) rpb = rlanes * 4;
    *rows_per_block = (int)std::min<long long>(rpb, 1 << 30);
    return (int)ceil_div_ll(M, *rows_per_block);
}

// row slabs for the non-reducing passes (apply kernels): ~8 resident blocks per SM, at least kBnUnroll rows per lane
static int ew_slabs(long long M, int C, int vec, int* rows_per_block, int* ychunks) {
    const int CV = C / vec;
    const int cvt = CV < kBnThreads ? CV : kBnThreads;
    *ychunks = ceil_div(CV, cvt);
    const int rlanes = kBnThreads / cvt;
    // one wave (3 resident blocks of 256 threads x 80 registers per SM): every block pays its prologue - per-channel
    // constants, statistics - once and in parallel; with 8 blocks per SM the small tensors (<= 34 MB: one 32-row
    // iteration per block) ran 2.3 waves of prologue + one load latency each, ~17 us for 17 MB
    const long long blocks = g_bn_wave ? std::max<long long>(1, (long long)num_sms() * 3 / *ychunks)
                                       : (long long)num_sms() * 8;
    long long rpb = ceil_div_ll(M, blocks);
    const long long quantum = (long long)rlanes * kBnUnroll;
    rpb = ceil_div_ll(rpb, quantum) * quantum;
    *rows_per_block = (int)std::min<long long>(rpb, 1 << 30);
    return (int)ceil_div_ll(M, *rows_per_block);
}

}  // namespace dn

using namespace dn;

extern "C" size_t denet_bn_workspace_bytes(long long M, int C) {
    int rpb, yc;
    // worst case is the scalar variant (more slabs never happen: slabs depend on M only through rows_per_block)
    const int n1 = bn_slabs(M, C, 1, &rpb, &yc);
    const int n8 = (C % 8 == 0) ? bn_slabs(M, C, 8, &rpb, &yc) : 0;
    const int n = n1 > n8 ? n1 : n8;
    return ((size_t)n * 2 * C + 2 * (size_t)C) * sizeof(float);
}

extern "C" int denet_bn_stats(const void* x, int dtype, long long M, int C, long long ld, float eps, float* mean,
                              float* invstd, float* run_mean, float* run_stdinv, float momentum, float* workspace,
                              size_t workspace_bytes, cudaStream_t stream) {
    DN_REQUIRE(x && mean && invstd && workspace, "bn_stats: null pointer");
    DN_REQUIRE(M > 0 && C > 0, "bn_stats: empty tensor");
    DN_REQUIRE(workspace_bytes >= denet_bn_workspace_bytes(M, C), "bn_stats: workspace too small");
    const bool v = vec8_ok(C, ld, x);
    int rpb, yc;
    const int nslabs = bn_slabs(M, C, v ? 8 : 1, &rpb, &yc);
    DN_DISPATCH(dtype, v, {
        bn_stats_partial_kernel<T, VEC><<<DN_G(dim3(nslabs, yc)), kBnThreads, 0, stream>>>((const T*)x, M, C, ld, rpb, workspace);
        bn_stats_finalize_kernel<T><<<DN_G(ceil_div(C, 32)), kFinThreads, 0, stream>>>((const T*)x, workspace, nslabs, M, C, eps, mean,
                                                                            invstd, run_mean, run_stdinv, momentum);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_bn_apply(const void* x, int dtype, long long M, int C, long long ld, const float* mean,
                              const float* invstd, const float* gamma, const float* beta, const void* residual,
                              int relu, void* y, cudaStream_t stream) {
    DN_REQUIRE(x && y && mean && invstd && gamma && beta, "bn_apply: null pointer");
    const bool v = vec8_ok(C, ld, x, y, residual);
    int rpb, yc;
    const int nslabs = ew_slabs(M, C, v ? 8 : 1, &rpb, &yc);
    BnSums none;
    memset(&none, 0, sizeof(none));
    DN_DISPATCH(dtype, v, {
        launch_pdl(bn_apply_kernel<T, VEC>, DN_G(dim3(nslabs, yc)), dim3(kBnThreads), 0, stream,
                   (const T*)x, M, C, ld, rpb, mean, invstd, gamma, beta, (const T*)residual, relu, (T*)y, none);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_bn_apply_sums(const void* x, int dtype, long long M, int C, long long ld, const float* sum,
                                   const float* sqsum, float eps, const float* gamma, const float* beta,
                                   const void* residual, int relu, void* y, float* mean, float* invstd, float* run_mean,
                                   float* run_stdinv, float momentum, cudaStream_t stream) {
    DN_REQUIRE(x && y && sum && sqsum && mean && invstd && gamma && beta, "bn_apply_sums: null pointer");
    DN_REQUIRE(M > 0, "bn_apply_sums: empty tensor");
    const bool v = vec8_ok(C, ld, x, y, residual);
    int rpb, yc;
    const int nslabs = ew_slabs(M, C, v ? 8 : 1, &rpb, &yc);
    BnSums sm;
    sm.sum = sum; sm.sqsum = sqsum; sm.M = M; sm.eps = eps; sm.mean_out = mean; sm.invstd_out = invstd;
    sm.run_mean = run_mean; sm.run_stdinv = run_stdinv; sm.momentum = momentum;
    DN_DISPATCH(dtype, v, {
        launch_pdl(bn_apply_kernel<T, VEC>, DN_G(dim3(nslabs, yc)), dim3(kBnThreads), 0, stream,
                   (const T*)x, M, C, ld, rpb, nullptr, nullptr, gamma, beta, (const T*)residual, relu, (T*)y, sm);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_bn_set_mode(int mode) {
    g_bn_fused_bwd = mode & 1;       // bit0: one-launch backward (grid barriers); 0 = partial / finalize / apply kernels
    g_bn_wave = (mode >> 1) & 1;     // bit1: one-wave apply grids + wide slab totals (default on)
    g_bn_debug = (mode >> 2) & 3;    // undocumented profiling knobs (scripts/bench_bn.py)
    return 0;
}

extern "C" int denet_bn_inference_invstd(const float* run_stdinv, float eps, float* out, int C, cudaStream_t stream) {
    DN_REQUIRE(run_stdinv && out, "bn_inference_invstd: null pointer");
    bn_inference_invstd_kernel<<<DN_G(ceil_div(C, 128)), 128, 0, stream>>>(run_stdinv, eps, out, C);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_bn_backward(const void* dy, const void* yout, const void* x, int dtype, long long M, int C,
                                 long long ld, const float* mean, const float* invstd, const float* gamma,
                                 const float* beta, int relu, void* dx, void* dres, float* dgamma, float* dbeta,
                                 int accumulate, float* workspace, size_t workspace_bytes, cudaStream_t stream) {
    DN_REQUIRE(dy && x && dx && mean && invstd && gamma && workspace, "bn_backward: null pointer");
    DN_REQUIRE(!relu || yout || beta, "bn_backward: the relu mask needs the forward output, or beta to recompute it");
    DN_REQUIRE(workspace_bytes >= denet_bn_workspace_bytes(M, C), "bn_backward: workspace too small");
    const bool v = vec8_ok(C, ld, dy, x, dx, yout) && vec8_ok(C, ld, dres);
    int rpb, yc;
    const int nslabs = bn_slabs(M, C, v ? 8 : 1, &rpb, &yc);
    float* sums = workspace + (size_t)nslabs * 2 * C;
    if (g_bn_fused_bwd) {
        // one launch: co-resident grid (2 blocks per SM at <= 128 registers), slabs sized for it
        int rpbf, ycf;
        const int nsf = bn_slabs_for(M, C, v ? 8 : 1, 2 * num_sms(), &rpbf, &ycf);
        if (nsf * ycf <= 2 * num_sms()) {
            static int slot_counter = 0;
            const int slot = (slot_counter++) % kBnSyncSlots;
            DN_DISPATCH(dtype, v, {
                constexpr int U = sizeof(T) == 4 ? 2 : 4;
                launch_pdl(bn_bwd_fused_kernel<T, VEC, U>, DN_G(dim3(nsf, ycf)), dim3(kBnThreads), 0, stream,
                    (const T*)dy, (const T*)yout, (const T*)x, M, C, ld, rpbf, mean, invstd, gamma, beta, relu, (T*)dx,
                    (T*)dres, workspace, sums, dgamma, dbeta, accumulate, slot, g_bn_wave, g_bn_debug);
            });
            DN_CHECK_LAUNCH();
            return 0;
        }
    }
    DN_DISPATCH(dtype, v, {
        bn_bwd_partial_kernel<T, VEC><<<DN_G(dim3(nslabs, yc)), kBnThreads, 0, stream>>>(
            (const T*)dy, (const T*)yout, (const T*)x, M, C, ld, rpb, mean, invstd, relu, workspace, gamma, beta);
        bn_bwd_finalize_kernel<<<DN_G(ceil_div(C, 32)), kFinThreads, 0, stream>>>(workspace, nslabs, C, sums, sums + C, dgamma, dbeta,
                                                                     accumulate);
        int rpb2, yc2;
        const int nslabs2 = ew_slabs(M, C, VEC, &rpb2, &yc2);
        bn_bwd_apply_kernel<T, VEC><<<DN_G(dim3(nslabs2, yc2)), kBnThreads, 0, stream>>>(
            (const T*)dy, (const T*)yout, (const T*)x, M, C, ld, rpb2, mean, invstd, gamma, sums, sums + C, relu,
            (T*)dx, (T*)dres, beta, nullptr, nullptr, 0);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_bn_backward_sums(const void* dy, const void* x, int dtype, long long M, int C, long long ld,
                                      const float* mean, const float* invstd, const float* gamma, const float* sum_dy,
                                      const float* sum_dy_xhat, void* dx, float* dgamma, float* dbeta, int accumulate,
                                      cudaStream_t stream) {
    DN_REQUIRE(dy && x && dx && mean && invstd && gamma && sum_dy && sum_dy_xhat, "bn_backward_sums: null pointer");
    DN_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "bn_backward_sums: dgamma and dbeta come together");
    const bool v = vec8_ok(C, ld, dy, x, dx);
    DN_DISPATCH(dtype, v, {
        int rpb2, yc2;
        const int nslabs2 = ew_slabs(M, C, VEC, &rpb2, &yc2);
        bn_bwd_apply_kernel<T, VEC><<<DN_G(dim3(nslabs2, yc2)), kBnThreads, 0, stream>>>(
            (const T*)dy, nullptr, (const T*)x, M, C, ld, rpb2, mean, invstd, gamma, sum_dy, sum_dy_xhat, 0, (T*)dx,
            nullptr, nullptr, dgamma, dbeta, accumulate);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_relu_fwd(const void* x, int dtype, long long M, int C, long long ld, void* y, cudaStream_t stream) {
    DN_REQUIRE(x && y, "relu_fwd: null pointer");
    const bool v = vec8_ok(C, ld, x, y);
    DN_DISPATCH(dtype, v, {
        relu_fwd_kernel<T, VEC><<<DN_G(ew_grid(M * (C / VEC), 256)), 256, 0, stream>>>((const T*)x, M, C, ld, (T*)y);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_relu_bwd(const void* dy, const void* y, int dtype, long long M, int C, long long ld, void* dx,
                              cudaStream_t stream) {
    DN_REQUIRE(dy && y && dx, "relu_bwd: null pointer");
    const bool v = vec8_ok(C, ld, dy, y, dx);
    DN_DISPATCH(dtype, v, {
        relu_bwd_kernel<T, VEC><<<DN_G(ew_grid(M * (C / VEC), 256)), 256, 0, stream>>>((const T*)dy, (const T*)y, M, C, ld,
                                                                                   (T*)dx);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_add(const void* a, const void* b, int dtype, long long M, int C, long long ld, int relu, void* out,
                         cudaStream_t stream) {
    DN_REQUIRE(a && b && out, "add: null pointer");
    const bool v = vec8_ok(C, ld, a, b, out);
    DN_DISPATCH(dtype, v, {
        add_kernel<T, VEC><<<DN_G(ew_grid(M * (C / VEC), 256)), 256, 0, stream>>>((const T*)a, (const T*)b, M, C, ld, relu,
                                                                              (T*)out);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_nchw_to_nhwc(const float* x, int N, int C, int H, int W, void* y, int dtype, long long ld,
                                  cudaStream_t stream) {
    DN_REQUIRE(x && y, "nchw_to_nhwc: null pointer");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)ceil_div_ll(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)N), block(32, 8);
    if (dtype == DENET_F32)
        nchw_to_nhwc_kernel<float><<<DN_G(grid), block, 0, stream>>>(x, C, HW, ld, (float*)y);
    else
        nchw_to_nhwc_kernel<__nv_bfloat16><<<DN_G(grid), block, 0, stream>>>(x, C, HW, ld, (__nv_bfloat16*)y);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_nhwc_to_nchw(const void* x, int dtype, long long ld, int N, int C, int H, int W, float* y,
                                  cudaStream_t stream) {
    DN_REQUIRE(x && y, "nhwc_to_nchw: null pointer");
    const long long HW = (long long)H * W;
    dim3 grid((unsigned)ceil_div_ll(HW, 32), (unsigned)ceil_div(C, 32), (unsigned)N), block(32, 8);
    if (dtype == DENET_F32)
        nhwc_to_nchw_kernel<float><<<DN_G(grid), block, 0, stream>>>((const float*)x, C, HW, ld, y);
    else
        nhwc_to_nchw_kernel<__nv_bfloat16><<<DN_G(grid), block, 0, stream>>>((const __nv_bfloat16*)x, C, HW, ld, y);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_pool_fwd(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int mode, int kh,
                              int kw, int sh, int sw, int ph, int pw, void* y, int Ho, int Wo, long long ldy,
                              uint8_t* argmax, cudaStream_t stream) {
    DN_REQUIRE(x && y, "pool_fwd: null pointer");
    DN_REQUIRE(mode == 0 || mode == 1, "pool_fwd: mode must be 0 (max) or 1 (average_inc_pad)");
    DN_REQUIRE(mode == 1 || argmax, "pool_fwd: max pooling needs an argmax buffer (N*Ho*Wo*C bytes)");
    DN_REQUIRE(kh * kw <= 255, "pool_fwd: window too large");
    const bool v = vec8_ok(C, ldx, x) && vec8_ok(C, ldy, y);
    DN_DISPATCH(dtype, v, {
        const int grid = ew_grid((long long)N * Ho * Wo * (C / VEC), 256);
        if (mode == 0)
            maxpool_fwd_kernel<T, VEC><<<DN_G(grid), 256, 0, stream>>>((const T*)x, N, H, W, C, ldx, kh, kw, sh, sw, ph, pw, Ho,
                                                                  Wo, ldy, (T*)y, argmax);
        else
            avgpool_fwd_kernel<T, VEC><<<DN_G(grid), 256, 0, stream>>>((const T*)x, N, H, W, C, ldx, kh, kw, sh, sw, ph, pw, Ho,
                                                                  Wo, ldy, (T*)y);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_pool_bwd(const void* dy, int dtype, int N, int H, int W, int C, long long ldx, int mode, int kh,
                              int kw, int sh, int sw, int ph, int pw, int Ho, int Wo, long long ldy,
                              const uint8_t* argmax, void* dx, cudaStream_t stream) {
    DN_REQUIRE(dy && dx, "pool_bwd: null pointer");
    DN_REQUIRE(mode == 1 || argmax, "pool_bwd: max pooling needs the argmax buffer of the forward pass");
    const bool v = vec8_ok(C, ldx, dx) && vec8_ok(C, ldy, dy);
    if (mode == 0 && v && dtype == DENET_BF16 && kh == 3 && kw == 3 && sh == 2 && sw == 2 && ph == 1 && pw == 1 &&
        (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8) < 0x7fffffffLL) {
        const int grid = ew_grid((long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8), 256);
        maxpool_bwd_3x3s2_kernel<<<DN_G(grid), 256, 0, stream>>>((const __nv_bfloat16*)dy, argmax, N, H, W, C, ldx, Ho, Wo,
                                                                  ldy, (__nv_bfloat16*)dx);
        DN_CHECK_LAUNCH();
        return 0;
    }
    DN_DISPATCH(dtype, v, {
        const int grid = ew_grid((long long)N * H * W * (C / VEC), 256);
        if (mode == 0)
            maxpool_bwd_kernel<T, VEC><<<DN_G(grid), 256, 0, stream>>>((const T*)dy, argmax, N, H, W, C, ldx, kh, kw, sh, sw, ph,
                                                                  pw, Ho, Wo, ldy, (T*)dx);
        else
            avgpool_bwd_kernel<T, VEC><<<DN_G(grid), 256, 0, stream>>>((const T*)dy, N, H, W, C, ldx, kh, kw, sh, sw, ph, pw, Ho,
                                                                  Wo, ldy, (T*)dx);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_pool_inv_fwd(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int sw, int sh,
                                  void* y, long long ldy, cudaStream_t stream) {
    DN_REQUIRE(x && y, "pool_inv_fwd: null pointer");
    DN_REQUIRE(sw > 0 && sh > 0, "pool_inv_fwd: bad size");
    const bool v = vec8_ok(C, ldx, x) && vec8_ok(C, ldy, y);
    DN_DISPATCH(dtype, v, {
        pool_inv_fwd_kernel<T, VEC><<<DN_G(ew_grid((long long)N * H * sh * W * sw * (C / VEC), 256)), 256, 0, stream>>>(
            (const T*)x, N, H, W, C, ldx, sw, sh, ldy, (T*)y);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_pool_inv_bwd(const void* dy, int dtype, int N, int H, int W, int C, long long ldx, int sw, int sh,
                                  void* dx, long long ldy, cudaStream_t stream) {
    DN_REQUIRE(dy && dx, "pool_inv_bwd: null pointer");
    const bool v = vec8_ok(C, ldx, dx) && vec8_ok(C, ldy, dy);
    DN_DISPATCH(dtype, v, {
        pool_inv_bwd_kernel<T, VEC><<<DN_G(ew_grid((long long)N * H * W * (C / VEC), 256)), 256, 0, stream>>>(
            (const T*)dy, N, H, W, C, ldx, sw, sh, ldy, (T*)dx);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_dilate_add(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int sh, int sw,
                                const void* add, void* y, int Hd, int Wd, long long ldy, cudaStream_t stream) {
    DN_REQUIRE(x && y, "dilate: null pointer");
    DN_REQUIRE(sh >= 1 && sw >= 1, "dilate: bad stride");
    const bool v = vec8_ok(C, ldx, x) && vec8_ok(C, ldy, y, add);
    DN_DISPATCH(dtype, v, {
        dilate_kernel<T, VEC><<<DN_G(ew_grid((long long)N * Hd * Wd * (C / VEC), 256)), 256, 0, stream>>>(
            (const T*)x, N, H, W, C, ldx, sh, sw, Hd, Wd, ldy, (T*)y, (const T*)add);
    });
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_dilate(const void* x, int dtype, int N, int H, int W, int C, long long ldx, int sh, int sw, void* y,
                            int Hd, int Wd, long long ldy, cudaStream_t stream) {
    return denet_dilate_add(x, dtype, N, H, W, C, ldx, sh, sw, nullptr, y, Hd, Wd, ldy, stream);
}

extern "C" int denet_bn_finalize_sums(const float* sum, const float* sqsum, long long M, int C, float eps, float* mean,
                                      float* invstd, float* run_mean, float* run_stdinv, float momentum,
                                      cudaStream_t stream) {
    DN_REQUIRE(sum && sqsum && mean && invstd, "bn_finalize_sums: null pointer");
    bn_finalize_sums_kernel<<<DN_G(ceil_div(C, 128)), 128, 0, stream>>>(sum, sqsum, M, C, eps, mean, invstd, run_mean,
                                                                  run_stdinv, momentum);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_convert(const void* x, int src_dtype, long long M, int C, long long ldx, void* y, int dst_dtype,
                             long long ldy, cudaStream_t stream) {
    DN_REQUIRE(x && y, "convert: null pointer");
    if (M == 0) return 0;
    const bool v = vec8_ok(C, ldx, x) && vec8_ok(C, ldy, y);
    const int grid = ew_grid(M * (C / (v ? 8 : 1)), 256);
#define DN_CONVERT(TI, TO)                                                                                     \
    do {                                                                                                       \
        if (v) convert_kernel<TI, TO, 8><<<DN_G(grid), 256, 0, stream>>>((const TI*)x, M, C, ldx, ldy, (TO*)y);      \
        else convert_kernel<TI, TO, 1><<<DN_G(grid), 256, 0, stream>>>((const TI*)x, M, C, ldx, ldy, (TO*)y);        \
    } while (0)
    if (src_dtype == DENET_F32 && dst_dtype == DENET_BF16) DN_CONVERT(float, __nv_bfloat16);
    else if (src_dtype == DENET_BF16 && dst_dtype == DENET_F32) DN_CONVERT(__nv_bfloat16, float);
    else if (src_dtype == DENET_F32 && dst_dtype == DENET_F32) DN_CONVERT(float, float);
    else DN_CONVERT(__nv_bfloat16, __nv_bfloat16);
#undef DN_CONVERT
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_colsum(const void* x, int dtype, long long M, int C, long long ld, float* out, int accumulate,
                            float* workspace, size_t workspace_bytes, cudaStream_t stream) {
    DN_REQUIRE(x && out && workspace, "colsum: null pointer");
    DN_REQUIRE(M > 0 && C > 0, "colsum: empty tensor");
    DN_REQUIRE(workspace_bytes >= denet_bn_workspace_bytes(M, C), "colsum: workspace too small");
    const bool v = vec8_ok(C, ld, x);
    int rpb, yc;
    const int nslabs = bn_slabs(M, C, v ? 8 : 1, &rpb, &yc);
    DN_DISPATCH(dtype, v, {
        colsum_partial_kernel<T, VEC><<<DN_G(dim3(nslabs, yc)), kBnThreads, 0, stream>>>((const T*)x, M, C, ld, rpb, workspace);
    });
    colsum_finalize_kernel<<<DN_G(ceil_div(C, 32)), kFinThreads, 0, stream>>>(workspace, nslabs, C, out, accumulate);
    DN_CHECK_LAUNCH();
    return 0;
}
