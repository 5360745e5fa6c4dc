// Vector pack helpers: kernels are written once over Pack<T, VEC> (VEC = 8 -> 16 B bf16 / 32 B fp32 accesses,
// VEC = 1 -> scalar fallback for channel counts that are not a multiple of 8).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dn {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int VEC>
struct Pack {
    float v[VEC];
    __device__ __forceinline__ void load(const T* p) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = to_f<T>(p[i]);
    }
    __device__ __forceinline__ void store(T* p) const {
#pragma unroll
        for (int i = 0; i < VEC; ++i) p[i] = from_f<T>(v[i]);
    }
};

template <>
struct Pack<float, 8> {
    float v[8];
    __device__ __forceinline__ void load(const float* p) {
        const float4 a = reinterpret_cast<const float4*>(p)[0];
        const float4 b = reinterpret_cast<const float4*>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <>
struct Pack<__nv_bfloat16, 8> {
    float v[8];
    __device__ __forceinline__ void load(const __nv_bfloat16* p) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <>
struct Pack<float, 4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

template <>
struct Pack<__nv_bfloat16, 4> {
    float v[4];
    __device__ __forceinline__ void load(const __nv_bfloat16* p) {
        const uint2 u = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
        v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
};

// Raw<T, VEC>: a pack kept in its memory representation until it is used (8 bf16 = 4 registers instead of the 8 of
// Pack<>): the streaming kernels hold several rows of several tensors in flight per thread, and their occupancy - and
// with it the bytes in flight per SM - is bounded by registers.
template <typename T, int VEC>
struct Raw {
    Pack<T, VEC> p;
    __device__ __forceinline__ void load(const T* ptr) { p.load(ptr); }
    __device__ __forceinline__ float get(int i) const { return p.v[i]; }
};

template <>
struct Raw<__nv_bfloat16, 8> {
    uint4 u;
    __device__ __forceinline__ void load(const __nv_bfloat16* ptr) { u = *reinterpret_cast<const uint4*>(ptr); }
    __device__ __forceinline__ float get(int i) const {
        const uint32_t w = (i >> 1) == 0 ? u.x : ((i >> 1) == 1 ? u.y : ((i >> 1) == 2 ? u.z : u.w));
        return (i & 1) ? __uint_as_float(w & 0xffff0000u) : __uint_as_float(w << 16);
    }
};

// dispatch on dtype code and on whether 8-wide vector access is legal (C % 8 == 0, pitch % 8 == 0, 16B base)
#define DN_DISPATCH(dtype, vec_ok, ...)                                    \
    do {                                                                   \
        if ((dtype) == DENET_F32) {                                        \
            using T = float;                                               \
            if (vec_ok) { constexpr int VEC = 8; __VA_ARGS__; }            \
            else { constexpr int VEC = 1; __VA_ARGS__; }                   \
        } else {                                                           \
            using T = __nv_bfloat16;                                       \
            if (vec_ok) { constexpr int VEC = 8; __VA_ARGS__; }            \
            else { constexpr int VEC = 1; __VA_ARGS__; }                   \
        }                                                                  \
    } while (0)

inline bool vec8_ok(int C, long long ld, const void* p0, const void* p1 = nullptr, const void* p2 = nullptr,
                    const void* p3 = nullptr) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
    return (C % 8 == 0) && (ld % 8 == 0) && al(p0) && al(p1) && al(p2) && al(p3);
}

}  // namespace dn
