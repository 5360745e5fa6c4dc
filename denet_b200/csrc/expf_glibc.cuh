// Single-precision exp with the exact result bits of glibc's expf (>= 2.27, the ARM optimized-routines algorithm).
//
// Why: the reference scores a box with  float pr = 1.0 / (1.0 + std::exp(fabs(pr_f - pr_t)))
// (denet/layer/denet_sparse.cc:306) where std::exp(float) is libm's expf.  expf is NOT correctly rounded (0.502 ULP),
// so a correctly rounded device exp differs from it in the last float bit for ~0.3 percent of the arguments - enough
// to break bit-exact scores.  The algorithm (third-party: glibc sysdeps/ieee754/flt-32/e_expf.c + e_exp2f_data.c,
// N = 32 table, degree-3 polynomial in double) is restated here: exp(x) = 2^(k/N) * 2^(r/N),
// k = round(x*N/ln2), r = x*N/ln2 - k.  The fused multiply-adds are where gcc contracts them in glibc's FMA build
// (the ifunc variant every x86-64 CPU with FMA selects); verified bit-identical to expf for ALL 1.119e9 floats in
// [0, 90) (tests/test_expf.py re-checks a sample).  Only x >= 0 is needed (x = |pr_f - pr_t|).
#ifndef DENET_EXPF_GLIBC_CUH
#define DENET_EXPF_GLIBC_CUH
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DN_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define DN_HD static inline
#endif

namespace dn {

// T[i] = bits(2^(i/32)) - (i << 47)
#define DN_EXP2F_TABLE                                                                       \
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL

#if defined(__CUDACC__)
static __device__ const uint64_t kExp2fTableDev[32] = {DN_EXP2F_TABLE};
#endif
static const uint64_t kExp2fTableHost[32] = {DN_EXP2F_TABLE};

DN_HD float expf_glibc(float x) {
    const double kN = 32.0;
    const double kInvLn2N = 0x1.71547652b82fep+0 * kN;
    const double kShift = 0x1.8p+52;
    const double kC0 = 0x1.c6af84b912394p-5 / kN / kN / kN;
    const double kC1 = 0x1.ebfce50fac4f3p-3 / kN / kN;
    const double kC2 = 0x1.62e42ff0c52d6p-1 / kN;
    if (x != x) return x;
#if defined(__CUDA_ARCH__)
    if (x > 0x1.62e42ep6f) return __int_as_float(0x7f800000);  // overflow -> +inf
    const double xd = (double)x;
    const double z = __dmul_rn(kInvLn2N, xd);
    double kd = __dadd_rn(z, kShift);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, kShift);
    const double r = __fma_rn(kInvLn2N, xd, -kd);
    const uint64_t t = kExp2fTableDev[ki & 31] + (ki << 47);
    const double s = __longlong_as_double((long long)t);
    const double p = __fma_rn(kC0, r, kC1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(kC2, r, 1.0);
    y = __fma_rn(p, r2, y);
    y = __dmul_rn(y, s);
    return (float)y;
#else
    if (x > 0x1.62e42ep6f) return INFINITY;
    const double xd = (double)x;
    const double z = kInvLn2N * xd;
    double kd = z + kShift;
    uint64_t ki;
    memcpy(&ki, &kd, 8);
    kd -= kShift;
    const double r = fma(kInvLn2N, xd, -kd);
    const uint64_t t = kExp2fTableHost[ki & 31] + (ki << 47);
    double s;
    memcpy(&s, &t, 8);
    const double p = fma(kC0, r, kC1);
    const double r2 = r * r;
    double y = fma(kC2, r, 1.0);
    y = fma(p, r2, y);
    y = y * s;
    return (float)y;
#endif
}

}  // namespace dn
#endif
