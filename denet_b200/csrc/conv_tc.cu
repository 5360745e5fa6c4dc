// tcgen05 implicit-GEMM convolution for sm_100a (NHWC bf16 operands, fp32 accumulation in TMEM).
//
// Replaces the cuDNN calls behind the reference's ConvLayer (reference denet/layer/convolution.py:76-92,
// tensor.nnet.conv2d + its autodiff dgrad/wgrad).
//
//  * conv_fprop_kernel : Y[pixel, co] = sum_{tap, ci} X[pixel + tap - pad, ci] * B[co, tap, ci]
//      - stride-1 R x S correlation. The reference's *true* convolution (filter flip) and the dgrad
//        (swap Cin/Cout, pad' = R-1-pad) are obtained purely by how denet_conv_weight_prep lays out B.
//      - M tile = a TW x TH x TN patch of 128 output pixels, fetched per filter tap as ONE 4-D TMA box
//        shifted by the tap offset; out-of-image rows/cols are zero-filled by TMA, which implements the
//        'half'/'valid'/'full' borders without any im2col buffer.  With R=S=1 this is a plain GEMM.
//      - K loop = terms x taps x 64-channel chunks.  terms=3 is the error-compensated bf16x3 split
//        (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) used by the fp32-parity mode; terms=1 is throughput mode.
//  * conv_wgrad_kernel : dW[co, tap, ci] = sum_{pixel} dY[pixel, co] * X[pixel + tap - pad, ci]
//      - both operands are MN-major (the contraction runs over pixels, channels are contiguous), split-K over
//        pixel blocks into an fp32 workspace, reduced deterministically by wgrad_reduce_kernel.
//
// Structure (both kernels): persistent CTAs, 6 warps: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc),
// warps2-5 = epilogue (TMEM -> registers -> global), mbarrier ring between producer and MMA, double-buffered
// TMEM accumulator between MMA and epilogue.
#include <string.h>
#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace dn {

constexpr int kBM = 128;      // UMMA M (TMEM lanes)
constexpr int kBK = 64;       // K elements per pipeline stage (= 128 B of bf16 = one swizzle row)
constexpr int kThreads = 192; // 6 warps (wgrad kernels): producer, MMA issuer, 4 epilogue warps
constexpr int kThreadsF = 320; // 10 warps (fprop / dgrad kernels): producer, MMA issuer, 8 epilogue warps
constexpr int kEpiWarpsF = 8;

// Division by a run-time constant in two instructions (the producer / MMA roles are single threads: every dependent
// scalar instruction costs ~5 cycles, an integer division by a kernel parameter ~40 of them).
struct FastDiv {
    uint32_t d, mul, shr;
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, mul) >> shr); }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};
static FastDiv make_fastdiv(int d) {   // valid for dividends < 2^31
    FastDiv f;
    f.d = (uint32_t)d;
    f.mul = 0;
    f.shr = 0;
    if (d > 1) {
        int lg = 0;
        while ((1u << lg) < (uint32_t)d) ++lg;
        const int pw = 31 + lg;
        f.mul = (uint32_t)(((1ull << pw) + (uint64_t)d - 1) / (uint64_t)d);
        f.shr = (uint32_t)(pw - 32);
    }
    return f;
}

struct ConvFpropParams {
    CUtensorMap tmA[2];  // input  (C, W, H, N) bf16, [0] = hi, [1] = lo
    CUtensorMap tmB[2];  // weight (KtotPad, Cout) bf16
    int nterms;          // 1 or 3
    int R, S, pad_h, pad_w;
    int stride_h, stride_w;
    int kchunks;         // ceil(Cin / 64)
    int Wo, Ho, No;      // output extent
    int TW, TH, TN;      // patch (TW*TH*TN == 128)
    int tiles_w, tiles_h, tiles_n, tiles_co;
    FastDiv fd_co, fd_w, fd_h, fd_tw, fd_th;   // tile -> (ct, tw, th, tn), row -> (w, h, n) without divisions
    int num_tiles;
    int Cout;
    long long ldy;       // output pixel pitch (elements)
    int y_fp32;
    int relu;
    void* y;
    const float* bias;       // [Cout] or null
    const void* residual;    // same layout/dtype as y, or null
    float* stat_sum;         // optional per-channel sum / sum of squares accumulators (fp32 atomics), or null
    float* stat_sqsum;
    // ---- fused batch-norm backward statistics: this launch is the dgrad whose output dz feeds the backward pass of a
    // batch-norm(+ReLU) layer.  The epilogue masks dz with the layer's ReLU mask, stores the masked gradient and
    // accumulates sum(dz') and sum(dz' * xhat) per channel into stat_sum / stat_sqsum (see fprop_epilogue).
    const void* bnb_x;       // the batch-norm layer's input (layout / dtype of y); null = off
    const void* bnb_yout;    // its forward output (mask = yout > 0) when the forward added a residual; else the mask is
                             // recomputed as (x - mean) * (gamma * invstd) + beta > 0
    const float* bnb_mean;
    const float* bnb_invstd;
    const float* bnb_gamma;
    const float* bnb_beta;
    int bnb_relu;
    int bnb_cpad;            // channel pitch of the shared-memory constant table (Cout rounded up to 32)
    // ---- scattered output: pixel (n, h, w) of this launch is pixel (n, h*osh + ooh, w*osw + oow) of a tensor with
    // Hf x Wf pixels per image (one parity class of a strided convolution's data gradient); default 1, 1, 0, 0, Ho, Wo
    int osh, osw, ooh, oow, Hf, Wf;
    // ---- tap-group variant (conv_fprop_halo_kernel): ONE halo'd A box per 64-channel chunk serves all R*S taps
    int a_loads;             // TMA loads per A stage: 1, or 2 = even / odd input rows of a stride-2 row-folded stem
    int a_dw, a_dh;          // start of load 0 relative to the patch origin (w0*stride_w, h0*stride_h), in map coords
    int a_dh_step;           // start row of load i = a_dh + i * a_dh_step
    uint32_t a_load_bytes;   // bytes one load delivers (the whole box, out-of-image rows are zero filled)
    uint32_t a_load_stride;  // shared-memory distance between the loads of a stage (1024-byte multiple)
    uint32_t a_stage_bytes;  // a_loads * a_load_stride
    uint32_t a_sbo;          // descriptor stride between 8-pixel groups of the M tile (bytes)
    uint32_t tap_off16[64];  // A descriptor start of tap t relative to the stage start, in 16-byte units
    int kmmas;               // K=16 MMAs per 64-wide K block (4, fewer when the tail of the block is structurally zero)
    int a_stages, b_stages;  // ring depths (run-time: the A stage size depends on the patch)
    int b_resident;          // all B tiles stay in shared memory for the life of the CTA (small filters, Cout <= BN)
    uint32_t a_region_bytes; // a_stages * a_stage_bytes
    uint32_t b_region_bytes; // B ring, or the resident filter bank
    int debug;               // profiling knobs (results are garbage): bit0 epilogue = TMEM read only, bit1 no statistics,
                             // bit2 no MMAs, bit3 no A / B loads
    long long* timeline;     // profiling: CTA 0 records clock64() stamps [role][tile][4] (roles: producer, MMA, epilogue)
    // ---- staged epilogue (fprop_epilogue_tma): the bf16 output tile goes through shared memory and leaves with TMA
    // stores; batch-norm statistics are summed from the staged tile
    CUtensorMap tmY;         // output (Cout, Wo, Ho, N) bf16, box 64 x TW x TH x TN, 128B swizzle
    int epi_tma;             // 1 = staged epilogue
    uint32_t stage_off;      // staging tile [BN/64][128 rows][128 B] relative to the 1024-aligned shared-memory base
    uint32_t smem_total;     // dynamic shared memory of the launch (host side only)
};

struct ConvWgradParams {
    CUtensorMap tmDY[2];  // dY (Cout, Wo, Ho, N) bf16
    CUtensorMap tmX[2];   // X  (Cin,  Wi, Hi, N) bf16
    int nterms;
    int R, S, pad_h, pad_w;
    int stride_h, stride_w;
    int TW, TH, TN;       // k-block patch (TW*TH*TN == 64)
    int tiles_w, tiles_h, tiles_n;
    int total_kblocks;
    int splits;
    int co_tiles, ci_tiles;
    FastDiv fd_cot, fd_cit, fd_taps, fd_w, fd_h;   // tile / k-block decoding without divisions
    int log2_tw;          // TW is a power of two
    int num_tiles;
    int Cout, Cin;
    int ldws;             // workspace cin pitch (elements)
    float* ws;            // [splits][Cout][taps][ldws]
    // row-shared variant (conv_wgrad_rows_kernel): one X box with a (S-1)-pixel halo serves the S taps of a filter row
    int xrows;            // rows of one 64-channel X sub-box = (TW + S - 1) * TH * TN
    int xbox_bytes;       // its smem footprint (xrows * 128 rounded up to 1024)
    int a_boxes;          // 64-channel dY sub-boxes actually loaded (1 when Cout <= 64)
    int debug;            // profiling knobs: bit0 skip the MMAs, bit1 skip the TMA loads (results are garbage)
    // row-folded stem variant (conv_wgrad_stem_kernel): all R filter rows per tile
    int x_loads;          // X boxes per stage: 1 (stride_h 1) or 2 (even / odd input rows of a stride-2 stem)
    int xbox_stride;      // shared-memory distance between them (1024-byte multiple)
    int stem_stages;      // ring depth
    int stem_stage_bytes; // a_boxes * 8192 + x_loads * xbox_stride
    int ntap[2];          // filter rows served by X box 0 / 1 (N of its MMA = ntap * 64)
};

template <int BN, int STAGES>
struct SmemLayout {
    static constexpr int kABytes = kBM * kBK * 2;
    static constexpr int kBBytes = BN * kBK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kStatOffset = kBarOffset + 256;     // fprop: per-warp running batch-norm sums [8][BN][2] fp32
    static constexpr int kTotal = kStatOffset + 8 * BN * 8 + 1024;  // + alignment slack (wgrad kernels leave the stat region unused)
};

__device__ __forceinline__ void store_row_chunk(void* y, int y_fp32, long long elem_off, const float (&v)[32],
                                                int ncols_valid) {
    if (y_fp32) {
        float* dst = reinterpret_cast<float*>(y) + elem_off;
        if (ncols_valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < ncols_valid) dst[i] = v[i];
        }
    } else {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(y) + elem_off;
        if (ncols_valid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 u;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i + 0], v[8 * i + 1]);
                __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]);
                __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
                u.x = *reinterpret_cast<uint32_t*>(&p0);
                u.y = *reinterpret_cast<uint32_t*>(&p1);
                u.z = *reinterpret_cast<uint32_t*>(&p2);
                u.w = *reinterpret_cast<uint32_t*>(&p3);
                reinterpret_cast<uint4*>(dst)[i] = u;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i < ncols_valid) dst[i] = __float2bfloat16_rn(v[i]);
        }
    }
}

__device__ __forceinline__ void load_row_chunk(const void* y, int y_fp32, long long elem_off, float (&v)[32],
                                               int ncols_valid) {
    if (y_fp32) {
        const float* src = reinterpret_cast<const float*>(y) + elem_off;
        if (ncols_valid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 f = reinterpret_cast<const float4*>(src)[i];
                v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (i < ncols_valid) ? src[i] : 0.f;
        }
    } else {
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(y) + elem_off;
        if (ncols_valid == 32 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 u = reinterpret_cast<const uint4*>(src)[i];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    v[8 * i + 2 * k] = __uint_as_float(w[k] << 16);
                    v[8 * i + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (i < ncols_valid) ? __bfloat162float(src[i]) : 0.f;
        }
    }
}

// Column sums across the 32 lanes of a warp: lane r holds v[0..31] (row r, 32 columns); on return v[0] of lane l is
// sum over rows of column l.  Each round halves the live values: a lane keeps the half of the columns selected by
// one bit of its lane index and trades the other half with its partner.
__device__ __forceinline__ void butterfly_colsum(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float keep = upper ? v[i + o] : v[i];
            const float send = upper ? v[i] : v[i + o];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
}

// ------------------------------------------------------------------------------------------------ fprop epilogue
// warps 2..5 -> TMEM lane quarter warp%4: TMEM -> registers -> bias / batch-norm statistics / residual / ReLU -> global.
// Shared by conv_fprop_kernel and conv_fprop_halo_kernel (same tile decoding, same accumulator double buffering).
// PAIR: the CTA is one half of a cta_group::2 pair (conv_fprop_halo2_kernel): pair tile `tile` covers the M tiles
// 2*mt and 2*mt+1, this CTA owns 2*mt + rank; the accumulator-empty barrier lives in the leader CTA.
template <int BN, bool PAIR = false>
__device__ __forceinline__ void fprop_epilogue(const ConvFpropParams& p, float2* stat_smem, uint32_t tmem_base,
                                               uint64_t* tfull_bar, uint64_t* tempty_bar, int warp, int lane) {
    // ------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quarter warp%4)
    // Eight epilogue warps: warps w and w+4 read the same TMEM lane quarter (w % 4, the hardware's rule) and split the
    // 32-column chunks of the tile between them (even / odd), which doubles the epilogue throughput of the layers whose
    // tiles are short (64 / 128 output channels, fused statistics).
    const int ew = warp - 2;               // 0..7
    const int q = warp & 3;
    const int half = ew >> 2;              // 0: chunks 0, 2, ..   1: chunks 1, 3, ..
    const int row = q * 32 + lane;
    int it = 0;
    // running per-channel sums of this warp (warp-private shared memory: entry [c][0] = sum, [c][1] = sum of squares)
    float2* sacc = stat_smem + ew * BN;
    if (p.stat_sum)
        for (int c = lane; c < BN; c += 32) sacc[c] = make_float2(0.f, 0.f);
    // fused batch-norm backward: per-channel constants [mean | invstd | gamma*invstd | beta] of ALL output channels
    float* bnb_tab = reinterpret_cast<float*>(stat_smem + kEpiWarpsF * BN);
    if (p.bnb_x) {
        const int cp = p.bnb_cpad;
        for (int c = ew * 32 + lane; c < cp; c += 32 * kEpiWarpsF) {
            const bool ok = c < p.Cout;
            const float mu = ok ? p.bnb_mean[c] : 0.f, is = ok ? p.bnb_invstd[c] : 0.f;
            bnb_tab[c] = mu;
            bnb_tab[cp + c] = is;
            bnb_tab[2 * cp + c] = ok ? p.bnb_gamma[c] * is : 0.f;
            bnb_tab[3 * cp + c] = (ok && p.bnb_beta) ? p.bnb_beta[c] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");       // the eight epilogue warps only
    }
    int acc_ct = -1;
    auto flush_stats = [&]() {
        if (p.stat_sum && acc_ct >= 0) {
            for (int i = half * 32 + lane; i < BN; i += 64) {      // the chunks this warp owns
                const int c = acc_ct * BN + i;
                const float2 a = sacc[i];
                if (c < p.Cout) {
                    atomicAdd(p.stat_sum + c, a.x);
                    atomicAdd(p.stat_sqsum + c, a.y);
                }
                sacc[i] = make_float2(0.f, 0.f);
            }
        }
    };
    const uint32_t pair_rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const int tile0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x, tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        uint32_t uct, mt, tw, th, tn;
        p.fd_co.divmod(tile, mt, uct);
        if (PAIR) mt = 2 * mt + pair_rank;
        const int ct = static_cast<int>(uct);
        if (ct != acc_ct) {
            flush_stats();
            acc_ct = ct;
        }
        p.fd_w.divmod(mt, mt, tw);
        p.fd_h.divmod(mt, tn, th);
        uint32_t rq, rw, rn, rh;
        p.fd_tw.divmod(row, rq, rw);
        p.fd_th.divmod(rq, rn, rh);
        const int w = tw * p.TW + rw;
        const int h = th * p.TH + rh;
        const int n = tn * p.TN + rn;
        const bool row_ok = (w < p.Wo) && (h < p.Ho) && (n < p.No);
        const long long pix = (static_cast<long long>(n) * p.Hf + (h * p.osh + p.ooh)) * p.Wf + (w * p.osw + p.oow);
        const int co0 = ct * BN;

        const bool stamp = p.timeline && blockIdx.x == 0 && warp == 2 && lane == 0 && it < 64;
        if (stamp) p.timeline[(2 * 64 + it) * 4 + 0] = clock64();
        ptx::mbar_wait(&tfull_bar[as], aphase);
        ptx::tc_fence_after();
        if (stamp) p.timeline[(2 * 64 + it) * 4 + 1] = clock64();
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c0;
            ptx::tmem_ld_32x32b_x32(taddr, r);
            ptx::tmem_ld_wait();
            const int co = co0 + c0;
            int nvalid = p.Cout - co;
            nvalid = nvalid > 32 ? 32 : nvalid;
            if (nvalid > 0 && !(p.debug & 1)) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                if (p.bias) {
                    // 8 independent 16-byte loads (32 dependent scalar loads cost ~1600 cycles per chunk)
                    const float* bp = p.bias + co;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
                        float4 b4[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bp) + i);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
                        }
                    } else {
                        float bb[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) bb[i] = (i < nvalid) ? __ldg(bp + i) : 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += bb[i];
                    }
                }
                if (p.stat_sum && !p.bnb_x && !(p.debug & 2)) {
                    // per-channel batch statistics of the (pre-activation) conv output.  Lane = pixel row, v[] = 32
                    // channels: a halving butterfly (16+8+4+2+1 exchanges per quantity instead of 32 x 5) leaves
                    // lane l with the 32-row sum of channel co + l; it is accumulated in registers across the
                    // tiles of this CTA and flushed with one atomic per lane when the channel tile changes.
                    float s[32], sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        s[i] = row_ok ? v[i] : 0.f;
                        sq[i] = s[i] * s[i];
                    }
                    butterfly_colsum(s, lane);
                    butterfly_colsum(sq, lane);
                    float2 a = sacc[c0 + lane];
                    a.x += s[0];
                    a.y += sq[0];
                    sacc[c0 + lane] = a;
                }
                if (p.bnb_x) {
                    // dz = v (+ residual); dz' = dz * [relu mask of the batch-norm layer]; sums of dz' and dz' * xhat.
                    // Lanes of rows outside the tensor contribute zeros (and store nothing).
                    const long long off = pix * p.ldy + co;
                    float xs[32];
                    if (row_ok) {
                        if (p.residual) {
                            float rres[32];
                            load_row_chunk(p.residual, p.y_fp32, off, rres, nvalid);
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] += rres[i];
                        }
                        load_row_chunk(p.bnb_x, p.y_fp32, off, xs, nvalid);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) xs[i] = 0.f;
                    }
                    const int cp = p.bnb_cpad;
                    const float4* t_mu = reinterpret_cast<const float4*>(bnb_tab + co);
                    const float4* t_is = reinterpret_cast<const float4*>(bnb_tab + cp + co);
                    float s[32], sq[32];
                    if (p.bnb_relu && p.bnb_yout) {
                        float yo[32];
                        if (row_ok) {
                            load_row_chunk(p.bnb_yout, p.y_fp32, off, yo, nvalid);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) yo[i] = 0.f;
                        }
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (!(yo[i] > 0.f)) v[i] = 0.f;
                    } else if (p.bnb_relu) {
                        const float4* t_a = reinterpret_cast<const float4*>(bnb_tab + 2 * cp + co);
                        const float4* t_b = reinterpret_cast<const float4*>(bnb_tab + 3 * cp + co);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            const float4 mu = t_mu[g4], a = t_a[g4], b = t_b[g4];
                            // the forward pass's exact expression (bn_apply_kernel): (x - mean) * (gamma*invstd) + beta
                            if (!((xs[4 * g4 + 0] - mu.x) * a.x + b.x > 0.f)) v[4 * g4 + 0] = 0.f;
                            if (!((xs[4 * g4 + 1] - mu.y) * a.y + b.y > 0.f)) v[4 * g4 + 1] = 0.f;
                            if (!((xs[4 * g4 + 2] - mu.z) * a.z + b.z > 0.f)) v[4 * g4 + 2] = 0.f;
                            if (!((xs[4 * g4 + 3] - mu.w) * a.w + b.w > 0.f)) v[4 * g4 + 3] = 0.f;
                        }
                    }
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        const float4 mu = t_mu[g4], is = t_is[g4];
                        const float m4[4] = {mu.x, mu.y, mu.z, mu.w}, i4[4] = {is.x, is.y, is.z, is.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int i = 4 * g4 + k;
                            const float d = row_ok ? v[i] : 0.f;
                            s[i] = d;
                            sq[i] = d * ((xs[i] - m4[k]) * i4[k]);
                        }
                    }
                    butterfly_colsum(s, lane);
                    butterfly_colsum(sq, lane);
                    float2 a2 = sacc[c0 + lane];
                    a2.x += s[0];
                    a2.y += sq[0];
                    sacc[c0 + lane] = a2;
                    if (row_ok) store_row_chunk(p.y, p.y_fp32, off, v, nvalid);
                } else if (row_ok) {
                    const long long off = pix * p.ldy + co;
                    if (p.residual) {
                        float rres[32];
                        load_row_chunk(p.residual, p.y_fp32, off, rres, nvalid);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += rres[i];
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    store_row_chunk(p.y, p.y_fp32, off, v, nvalid);
                }
            }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR)
                ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty_bar[as]), 0));
            else
                ptx::mbar_arrive(&tempty_bar[as]);
        }
        if (stamp) p.timeline[(2 * 64 + it) * 4 + 2] = clock64();
    }
    flush_stats();
}

// ------------------------------------------------------------------------------------------------ staged epilogue
// bf16 outputs without scatter / fused batch-norm backward.  fprop_epilogue above costs ~2500 cycles per 128 x 64 tile
// (two 31-shuffle butterflies per 32-column chunk for the statistics, 16-byte global stores scattered over 128 pixel
// rows) - MORE than the 36 MMAs of a 64-channel 3x3 tile, so the short-tile layers were epilogue-bound.  Here:
//   TMEM -> registers -> (bias / residual / ReLU) -> bf16 -> st.shared.v4 into a staging tile laid out exactly as a
//   TMA box with 128B swizzle (row = pixel of the patch, 64-channel sub-tiles 16 KB apart; 4 lanes per bank group is the
//   512-byte-per-instruction floor), the accumulator is released right after the TMEM reads, ONE thread issues the
//   TMA stores (UTMASTG, clipped at the tensor edge) and meanwhile all 256 epilogue threads sum the statistics from the
//   staged tile: thread = (column pair, row group), 4-byte conflict-free reads, running sums in 4 registers across the
//   tiles of the CTA - no shuffles, no per-tile atomics.  The statistics are those of the bf16-ROUNDED outputs, i.e. of
//   exactly the values the batch-norm apply pass normalises.
template <int BN, bool PAIR>
__device__ __forceinline__ void fprop_epilogue_tma(const ConvFpropParams& p, uint8_t* stage, float* stat_red,
                                                   uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                   int warp, int lane) {
    // Fused batch-norm backward (p.bnb_x, denet_conv2d_dgrad_bnbwd): this launch is the dgrad whose output dz is the
    // gradient wrt the OUTPUT of a batch-norm(+ReLU) layer with input x.  The tile is masked with that layer's ReLU
    // mask before it is staged (the mask inputs - x, or the layer's forward output when a residual was added - are
    // prefetched per lane like the residual), and the two sums of the batch-norm backward's reduction pass are taken
    // from the staged tile: sum(dz') and sum(dz' * xhat), x read column-wise (coalesced 128-byte rows, L2 hits: the
    // lanes have just loaded the same lines).  The separate reduction pass over dz and x (2 x tensor bytes) is gone.
    const bool bnb = p.bnb_x != nullptr;
    float* bnb_tab = stat_red + 2 * 512;              // [mean | invstd | gamma*invstd | beta] x bnb_cpad, after stat_red (4 KB)
    const int ew = warp - 2;               // 0..7
    const int q = warp & 3;
    const int half = ew >> 2;
    const int row = q * 32 + lane;
    const int et = static_cast<int>(threadIdx.x) - 64;     // 0..255
    constexpr int kPairs = BN / 2;               // column pairs of the tile
    constexpr int kGroups = 256 / kPairs;        // row groups summed by different threads
    constexpr int kRowsPerGroup = 128 / kGroups;
    const int cpair = et % kPairs, grp = et / kPairs;
    const uint32_t stage_u32 = ptx::smem_u32(stage);
    // staged word of (row r, column pair cpair): sub-tile cpair / 32, 16-byte chunk (cpair % 32) / 4 swizzled by r & 7
    const uint32_t stat_base = stage_u32 + static_cast<uint32_t>(cpair >> 5) * 16384u + static_cast<uint32_t>(cpair & 3) * 4u;
    const uint32_t stat_chunk = static_cast<uint32_t>((cpair & 31) >> 2);
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    int acc_ct = -1;
    const bool want_stats = p.stat_sum != nullptr;
    auto flush_stats = [&]() {
        if (!want_stats || acc_ct < 0) return;
        // combine the row groups in a fixed order, then one atomic per channel per CTA
        float* red = stat_red + (grp * BN + 2 * cpair) * 2;
        red[0] = s0; red[1] = q0; red[2] = s1; red[3] = q1;
        ptx::named_bar_sync(3, 256);
        if (et < BN) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int g2 = 0; g2 < kGroups; ++g2) {
                a += stat_red[(g2 * BN + et) * 2 + 0];
                b += stat_red[(g2 * BN + et) * 2 + 1];
            }
            const int c = acc_ct * BN + et;
            if (c < p.Cout) {
                atomicAdd(p.stat_sum + c, a);
                atomicAdd(p.stat_sqsum + c, b);
            }
        }
        ptx::named_bar_sync(3, 256);
        s0 = s1 = q0 = q1 = 0.f;
    };
    if (bnb) {
        const int cp = p.bnb_cpad;
        for (int c = et; c < cp; c += 256) {
            const bool ok = c < p.Cout;
            const float mu = ok ? p.bnb_mean[c] : 0.f, is = ok ? p.bnb_invstd[c] : 0.f;
            bnb_tab[c] = mu;
            bnb_tab[cp + c] = is;
            bnb_tab[2 * cp + c] = ok ? p.bnb_gamma[c] * is : 0.f;
            bnb_tab[3 * cp + c] = (ok && p.bnb_beta) ? p.bnb_beta[c] : 0.f;
        }
        ptx::named_bar_sync(3, 256);
    }
    const uint32_t pair_rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const int tile0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x, tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
    int it = 0;
    float mu0 = 0.f, mu1 = 0.f, is0 = 0.f, is1 = 0.f;       // statistics threads, fused batch-norm backward: constants
    for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        uint32_t uct, mt, tw, th, tn;
        p.fd_co.divmod(tile, mt, uct);
        if (PAIR) mt = 2 * mt + pair_rank;
        const int ct = static_cast<int>(uct);
        if (ct != acc_ct) {
            flush_stats();
            acc_ct = ct;
            if (bnb) {
                const int c = ct * BN + 2 * cpair, cp = p.bnb_cpad;
                mu0 = c < cp ? bnb_tab[c] : 0.f;          is0 = c < cp ? bnb_tab[cp + c] : 0.f;
                mu1 = c + 1 < cp ? bnb_tab[c + 1] : 0.f;  is1 = c + 1 < cp ? bnb_tab[cp + c + 1] : 0.f;
            }
        }
        p.fd_w.divmod(mt, mt, tw);
        p.fd_h.divmod(mt, tn, th);
        uint32_t rq, rw, rn, rh;
        p.fd_tw.divmod(row, rq, rw);
        p.fd_th.divmod(rq, rn, rh);
        const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
        const int w = w0 + rw, h = h0 + rh, n = n0 + rn;
        const bool row_ok = (w < p.Wo) && (h < p.Ho) && (n < p.No);
        // (scatter-aware: pixel (n, h, w) of a parity-class launch lives at (n, h*osh + ooh, w*osw + oow) of the full tensor)
        const long long pix = (static_cast<long long>(n) * p.Hf + (h * p.osh + p.ooh)) * p.Wf + (w * p.osw + p.oow);
        const int co0 = ct * BN;

        // residual of the first chunk: issued BEFORE the wait for the accumulator, so that the global-load latency
        // (a full ~1 us per tile otherwise: more than the MMAs of a 64-channel tile) hides behind the MMAs
        const bool res_fast = p.residual && row_ok && (p.Cout % 8 == 0);
        const __nv_bfloat16* res_row = reinterpret_cast<const __nv_bfloat16*>(p.residual) + pix * p.ldy + co0;
        uint4 rraw[4];
        if (res_fast && co0 + half * 32 + 32 <= p.Cout) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rraw[j] = __ldg(reinterpret_cast<const uint4*>(res_row + half * 32) + j);
        }
        // fused batch-norm backward: mask input of the first chunk (x, or the layer's forward output)
        const bool mask_on = bnb && p.bnb_relu;
        const __nv_bfloat16* msk_row = reinterpret_cast<const __nv_bfloat16*>(p.bnb_yout ? p.bnb_yout : p.bnb_x) +
                                       pix * p.ldy + co0;
        const bool msk_fast = mask_on && row_ok && (p.Cout % 8 == 0);
        uint4 mraw[4];
        if (msk_fast && co0 + half * 32 + 32 <= p.Cout) {
#pragma unroll
            for (int j = 0; j < 4; ++j) mraw[j] = __ldg(reinterpret_cast<const uint4*>(msk_row + half * 32) + j);
        }
        ptx::mbar_wait(&tfull_bar[as], aphase);
        ptx::tc_fence_after();
        // the staging tile is free once the previous tile's TMA stores have read it and every thread has finished
        // summing its statistics from it
        if (et == 0 && it > 0) ptx::bulk_wait_read0();
        ptx::named_bar_sync(1, 256);
        const uint32_t srow = stage_u32 + static_cast<uint32_t>(row) * 128u;
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c0;
            ptx::tmem_ld_32x32b_x32(taddr, r);
            ptx::tmem_ld_wait();
            const int co = co0 + c0;
            int nvalid = p.Cout - co;
            nvalid = nvalid > 32 ? 32 : nvalid;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if (nvalid > 0 && row_ok) {
                if (p.bias) {
                    const float* bp = p.bias + co;
                    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0)) {
                        float4 b4[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bp) + i);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            v[4 * i] += b4[i].x; v[4 * i + 1] += b4[i].y; v[4 * i + 2] += b4[i].z; v[4 * i + 3] += b4[i].w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += (i < nvalid) ? __ldg(bp + i) : 0.f;
                    }
                }
                if (p.residual) {
                    if (res_fast && nvalid == 32) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t wv[4] = {rraw[j].x, rraw[j].y, rraw[j].z, rraw[j].w};
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2) {
                                v[8 * j + 2 * k2] += __uint_as_float(wv[k2] << 16);
                                v[8 * j + 2 * k2 + 1] += __uint_as_float(wv[k2] & 0xffff0000u);
                            }
                        }
                        if (c0 + 64 < BN && co + 64 + 32 <= p.Cout) {      // next chunk of this warp: in flight during the stores
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                rraw[j] = __ldg(reinterpret_cast<const uint4*>(res_row + c0 + 64) + j);
                        }
                    } else {
                        float rres[32];
                        load_row_chunk(p.residual, 0, pix * p.ldy + co, rres, nvalid);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] += rres[i];
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (mask_on) {
                    float mv[32];
                    if (msk_fast && nvalid == 32) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t wv[4] = {mraw[j].x, mraw[j].y, mraw[j].z, mraw[j].w};
#pragma unroll
                            for (int k2 = 0; k2 < 4; ++k2) {
                                mv[8 * j + 2 * k2] = __uint_as_float(wv[k2] << 16);
                                mv[8 * j + 2 * k2 + 1] = __uint_as_float(wv[k2] & 0xffff0000u);
                            }
                        }
                        if (c0 + 64 < BN && co + 64 + 32 <= p.Cout) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                mraw[j] = __ldg(reinterpret_cast<const uint4*>(msk_row + c0 + 64) + j);
                        }
                    } else {
                        load_row_chunk(p.bnb_yout ? p.bnb_yout : p.bnb_x, 0, pix * p.ldy + co, mv, nvalid);
                    }
                    if (p.bnb_yout) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (!(mv[i] > 0.f)) v[i] = 0.f;
                    } else {
                        // the forward pass's exact expression (bn_apply_kernel): (x - mean) * (gamma*invstd) + beta > 0
                        const int cp = p.bnb_cpad;
                        const float4* t_mu = reinterpret_cast<const float4*>(bnb_tab + co);
                        const float4* t_a = reinterpret_cast<const float4*>(bnb_tab + 2 * cp + co);
                        const float4* t_b = reinterpret_cast<const float4*>(bnb_tab + 3 * cp + co);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
                            const float4 mu = t_mu[g4], a = t_a[g4], b = t_b[g4];
                            if (!((mv[4 * g4 + 0] - mu.x) * a.x + b.x > 0.f)) v[4 * g4 + 0] = 0.f;
                            if (!((mv[4 * g4 + 1] - mu.y) * a.y + b.y > 0.f)) v[4 * g4 + 1] = 0.f;
                            if (!((mv[4 * g4 + 2] - mu.z) * a.z + b.z > 0.f)) v[4 * g4 + 2] = 0.f;
                            if (!((mv[4 * g4 + 3] - mu.w) * a.w + b.w > 0.f)) v[4 * g4 + 3] = 0.f;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i >= nvalid) v[i] = 0.f;
            } else {
                // rows outside the tensor / channels past Cout: zeros (clipped by the store, neutral for the sums)
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
            }
            uint32_t w16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w16[i] = *reinterpret_cast<uint32_t*>(&t2);
            }
            const uint32_t sub = static_cast<uint32_t>(c0 >> 6) * 16384u;
            const uint32_t cbase = static_cast<uint32_t>((c0 & 63) >> 3);       // first 16-byte chunk of this 32-column run
#pragma unroll
            for (int j = 0; j < 4; ++j)
                ptx::st_shared_v4(srow + sub + (((cbase + j) ^ static_cast<uint32_t>(row & 7)) << 4), w16[4 * j],
                                  w16[4 * j + 1], w16[4 * j + 2], w16[4 * j + 3]);
        }
        // accumulator buffer read completely: hand it back to the MMA issuer before the stores / statistics
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR)
                ptx::mbar_arrive_cluster(ptx::mapa_u32(ptx::smem_u32(&tempty_bar[as]), 0));
            else
                ptx::mbar_arrive(&tempty_bar[as]);
        }
        ptx::fence_proxy_async();              // generic-proxy writes of the staging tile -> visible to the TMA engine
        ptx::named_bar_sync(2, 256);
        if (et == 0) {
#pragma unroll
            for (int sb = 0; sb < BN / 64; ++sb)
                if (co0 + sb * 64 < p.Cout) ptx::tma_store_4d(&p.tmY, stage + sb * 16384, co0 + sb * 64, w0, h0, n0);
            ptx::bulk_commit_group();
        }
        if (want_stats && !bnb) {
#pragma unroll 4
            for (int rr = 0; rr < kRowsPerGroup; ++rr) {
                const uint32_t r2 = static_cast<uint32_t>(grp * kRowsPerGroup + rr);
                const uint32_t wv = ptx::ld_shared_u32(stat_base + r2 * 128u + ((stat_chunk ^ (r2 & 7u)) << 4));
                const float f0 = __uint_as_float(wv << 16), f1 = __uint_as_float(wv & 0xffff0000u);
                s0 += f0; q0 = fmaf(f0, f0, q0);
                s1 += f1; q1 = fmaf(f1, f1, q1);
            }
        } else if (want_stats) {
            // sum(dz') and sum(dz' * xhat): dz' from the staged tile, x column-wise from global memory (8 rows in flight)
            const int cx = co0 + 2 * cpair;
            const bool col_ok = cx + 1 < p.Cout + 1 && cx < p.Cout;
            const __nv_bfloat16* xcol = reinterpret_cast<const __nv_bfloat16*>(p.bnb_x) + cx;
#pragma unroll 1
            for (int r0 = 0; r0 < kRowsPerGroup; r0 += 8) {
                uint32_t xv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t r2 = static_cast<uint32_t>(grp * kRowsPerGroup + r0 + u);
                    uint32_t q2, w2, n2, h2;
                    p.fd_tw.divmod(r2, q2, w2);
                    p.fd_th.divmod(q2, n2, h2);
                    const int ww = w0 + static_cast<int>(w2), hh = h0 + static_cast<int>(h2), nn = n0 + static_cast<int>(n2);
                    const bool ok = col_ok && ww < p.Wo && hh < p.Ho && nn < p.No;
                    const long long px = (static_cast<long long>(nn) * p.Hf + (hh * p.osh + p.ooh)) * p.Wf + (ww * p.osw + p.oow);
                    xv[u] = ok ? __ldg(reinterpret_cast<const uint32_t*>(xcol + px * p.ldy)) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t r2 = static_cast<uint32_t>(grp * kRowsPerGroup + r0 + u);
                    const uint32_t wv = ptx::ld_shared_u32(stat_base + r2 * 128u + ((stat_chunk ^ (r2 & 7u)) << 4));
                    const float d0 = __uint_as_float(wv << 16), d1 = __uint_as_float(wv & 0xffff0000u);
                    const float x0 = __uint_as_float(xv[u] << 16), x1 = __uint_as_float(xv[u] & 0xffff0000u);
                    s0 += d0; q0 = fmaf(d0, (x0 - mu0) * is0, q0);
                    s1 += d1; q1 = fmaf(d1, (x1 - mu1) * is1, q1);
                }
            }
        }
    }
    flush_stats();
    if (et == 0) ptx::bulk_wait0();            // the last stores have completed before the CTA (and its smem) goes away
}

// ------------------------------------------------------------------------------------------------ fprop / dgrad
template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreadsF, 1) conv_fprop_kernel(const __grid_constant__ ConvFpropParams p) {
    using L = SmemLayout<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], kEpiWarpsF);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntaps = p.R * p.S;
    const int num_kb = p.nterms * ntaps * p.kchunks;
    // loop-invariant parameters in registers: the single-instruction-stream roles must not re-read them per k-block
    const int R = p.R, S = p.S, kchunks = p.kchunks, nterms = p.nterms, num_tiles = p.num_tiles, debug = p.debug;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        // The whole warp runs the loops (so that every address / coordinate is warp-uniform and lives in uniform
        // registers); one elected lane issues the TMA instructions.  Running the loops under `if (lane == 0)` makes
        // the compiler wrap every TMA / MMA issue in a register -> uniform-register "waterfall" loop (~200 cycles).
        {
            if (ptx::elect_one()) {
                ptx::tma_prefetch_desc(&p.tmA[0]);
                ptx::tma_prefetch_desc(&p.tmB[0]);
            }
            int stage = 0;
            uint32_t phase = 0;
            const int tile_w = p.TW * p.stride_w, tile_h = p.TH * p.stride_h, pad_w = p.pad_w, pad_h = p.pad_h, TN = p.TN;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                uint32_t ct, mt, tw, th, tn;
                p.fd_co.divmod(tile, mt, ct);
                p.fd_w.divmod(mt, mt, tw);
                p.fd_h.divmod(mt, tn, th);
                // input-space origin of the patch: the tensor map walks the image with element strides (stride_w,
                // stride_h), so a strided convolution is the same box fetch started at (w0*stride + tap offset)
                const int w0 = tw * tile_w - pad_w, h0 = th * tile_h - pad_h;
                const int n0 = tn * TN, co0 = ct * BN;
                for (int term = 0; term < nterms; ++term) {
                    // terms of the fp32-parity split, corrections FIRST: (lo,hi) (hi,lo) then (hi,hi).  The TMEM
                    // accumulator truncates on every accumulation step (measured: relative bias ~ steps x 6e-8), so
                    // the steps taken while the accumulator already holds the full-magnitude sum must be as few as
                    // possible - with the main term last only its own K/16 steps count (3x fewer than main-first)
                    const CUtensorMap* mapA = &p.tmA[(nterms == 3 && term == 0) ? 1 : 0];
                    const CUtensorMap* mapB = &p.tmB[(nterms == 3 && term == 1) ? 1 : 0];
                    int kcol = 0;                                            // K column of the weight operand
                    for (int r = 0; r < R; ++r) {
                        for (int sx = 0; sx < S; ++sx) {
                            for (int kc = 0; kc < kchunks; ++kc, kcol += kBK) {
                                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                                uint8_t* sA = smem + stage * L::kStageBytes;
                                if (ptx::elect_one()) {
                                    if (debug & 8) {
                                        ptx::mbar_arrive(&full_bar[stage]);
                                    } else {
                                        ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
                                        ptx::tma_load_4d(sA, mapA, &full_bar[stage], kc * kBK, w0 + sx, h0 + r, n0);
                                        ptx::tma_load_2d(sA + L::kABytes, mapB, &full_bar[stage], kcol, co0);
                                    }
                                }
                                __syncwarp();
                                if (++stage == STAGES) {
                                    stage = 0;
                                    phase ^= 1;
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BN, 0, 0);
            const uint64_t adesc0 = ptx::make_smem_desc(ptx::smem_u32(smem), 16, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const bool do_mma = !(debug & 4);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
#pragma unroll 1
                for (int kb = 0; kb < num_kb; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    // descriptors of stage 0 + the stage offset in 16-byte units (the address field is addr >> 4)
                    const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * (L::kStageBytes >> 4));
                    const uint64_t bdesc = adesc + (L::kABytes >> 4);
                    if (ptx::elect_one()) {
                        if (do_mma) {
#pragma unroll
                            for (int j = 0; j < kBK / 16; ++j) {
                                // advance 16 bf16 (32 B) along K inside the 128B-swizzled row
                                ptx::umma_f16(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (kb | j) != 0);
                            }
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[as]);
                __syncwarp();
            }
        }
    } else if (p.epi_tma) {
        fprop_epilogue_tma<BN, false>(p, smem + p.stage_off, reinterpret_cast<float*>(smem + L::kStatOffset), tmem_base,
                                      tfull_bar, tempty_bar, warp, lane);
    } else {
        fprop_epilogue<BN>(p, reinterpret_cast<float2*>(smem + L::kStatOffset), tmem_base, tfull_bar, tempty_bar, warp,
                           lane);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ fprop, tap groups
// Stride-1 R x S convolutions (and the row-folded stem) with the A operand fetched ONCE per 64-channel chunk for all
// taps: the patch of 128 output pixels is loaded with its (R-1, S-1) halo as one TMA box, and tap (r, s) reads the box
// through a shared-memory descriptor that starts r*(TW+S-1)+s pixel rows (128 B each) into it, with the stride between
// 8-pixel groups set to one box row (TW = 8, so a group is one image row of the patch).  The 128B-swizzle phase follows
// the absolute shared-memory address, which is how TMA wrote the box, so unaligned starts read consistent data.
// conv_fprop_kernel fetches 9 boxes of 16 KB per chunk for a 3x3 filter; this kernel fetches one of 22.5 KB - the
// L2 -> shared-memory traffic (the bound of the <= 256-channel layers, ~12 TB/s chip-wide) drops by the same factor for
// A.  Filters that fit (64-channel 3x3 layers: 72 KB, the stem: 56 KB) stay resident in shared memory instead of being
// streamed per tile.  Row-folded stem with stride 2: two boxes (even / odd input rows), tap r reads box r&1 at row r>>1.
// NT = number of taps when known at compile time (the tap loop is then fully unrolled with the descriptor offsets in
// registers), 0 = run-time count.  RESIDENT = the filter bank lives in shared memory for the whole kernel.  KM = K=16
// MMAs per 64-wide K block when known at compile time (0 = run-time p.kmmas).
// The producer and MMA roles are single instruction streams whose per-iteration latency bounds the small-N layers (a
// 128 x 64 x 16 MMA executes in 32 cycles): every loop-invariant parameter is hoisted into registers, and the only
// work left per tap is the barrier handshake (streaming filters) and the address adds of the K=16 MMAs.
template <int BN, int NT, bool RESIDENT, int KM>
__global__ void __launch_bounds__(kThreadsF, 1) conv_fprop_halo_kernel(const __grid_constant__ ConvFpropParams p) {
    constexpr int kBBytes = BN * kBK * 2;
    constexpr uint32_t kB16 = kBBytes >> 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_b = smem + p.a_region_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + p.b_region_bytes);
    uint64_t* a_empty = a_full + 8;
    uint64_t* b_full = a_empty + 8;
    uint64_t* b_empty = b_full + 16;
    uint64_t* tfull_bar = b_empty + 16;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* bres_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 2);    // its own 16-byte granule (racecheck pairs
                                                                            // tcgen05.alloc's result slot with neighbouring barrier writes)
    float2* stat_smem = reinterpret_cast<float2*>(smem_b + p.b_region_bytes + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            ptx::mbar_init(&a_full[s], 1);
            ptx::mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 16; ++s) {
            ptx::mbar_init(&b_full[s], 1);
            ptx::mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], kEpiWarpsF);
        }
        ptx::mbar_init(bres_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // loop-invariant parameters (kept in registers: the role loops below must not re-read them)
    const int ntaps = NT > 0 ? NT : p.R * p.S;
    const int kchunks = p.kchunks, nterms = p.nterms, num_tiles = p.num_tiles;
    const int a_stages = p.a_stages, b_stages = p.b_stages;
    const uint32_t a_stage_bytes = p.a_stage_bytes;
    const int debug = p.debug;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (warp-uniform loops, elected issue)
        if (ptx::elect_one()) {
            ptx::tma_prefetch_desc(&p.tmA[0]);
            ptx::tma_prefetch_desc(&p.tmB[0]);
        }
        if (RESIDENT) {
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(bres_bar, p.b_region_bytes);
                const int nb = ntaps * kchunks;
                for (int i = 0; i < nb; ++i)
                    ptx::tma_load_2d(smem_b + i * kBBytes, &p.tmB[0], bres_bar, i * kBK, 0);
            }
            __syncwarp();
        }
        const int a_loads = p.a_loads, a_dh_step = p.a_dh_step;
        const uint32_t a_load_stride = p.a_load_stride, a_tx = p.a_loads * p.a_load_bytes;
        const int tile_w = p.TW * p.stride_w, tile_h = p.TH * p.stride_h, a_dw = p.a_dw, a_dh = p.a_dh, TN = p.TN;
        const int tap_kstep = kchunks * kBK;
        int as = 0, bs = 0;
        uint32_t aphase = 0, bphase = 0;
        int pit = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++pit) {
            if (p.timeline && blockIdx.x == 0 && lane == 0 && pit < 64) p.timeline[(0 * 64 + pit) * 4 + 0] = clock64();
            uint32_t ct, mt, tw, th, tn;
            p.fd_co.divmod(tile, mt, ct);
            p.fd_w.divmod(mt, mt, tw);
            p.fd_h.divmod(mt, tn, th);
            const int w0 = tw * tile_w + a_dw, h0 = th * tile_h + a_dh;
            const int n0 = tn * TN, co0 = ct * BN;
            for (int term = 0; term < nterms; ++term) {
                // terms: (lo,hi) (hi,lo) then (hi,hi) - corrections first, see conv_fprop_kernel
                const CUtensorMap* mapA = &p.tmA[(nterms == 3 && term == 0) ? 1 : 0];
                const CUtensorMap* mapB = &p.tmB[(nterms == 3 && term == 1) ? 1 : 0];
                for (int kc = 0; kc < kchunks; ++kc) {
                    ptx::mbar_wait(&a_empty[as], aphase ^ 1);
                    uint8_t* sA = smem + as * a_stage_bytes;
                    if (ptx::elect_one()) {
                        if (debug & 8) {
                            ptx::mbar_arrive(&a_full[as]);
                        } else {
                            ptx::mbar_expect_tx(&a_full[as], a_tx);
                            for (int i = 0; i < a_loads; ++i)
                                ptx::tma_load_4d(sA + i * a_load_stride, mapA, &a_full[as], kc * kBK, w0,
                                                 h0 + i * a_dh_step, n0);
                        }
                    }
                    __syncwarp();
                    if (p.timeline && blockIdx.x == 0 && lane == 0 && pit < 64 && kc == 0)
                        p.timeline[(0 * 64 + pit) * 4 + 1] = clock64();
                    if (++as == a_stages) {
                        as = 0;
                        aphase ^= 1;
                    }
                    if (!RESIDENT) {
                        int kcol = kc * kBK;                       // B column of (tap 0, chunk kc)
#pragma unroll 1
                        for (int t = 0; t < ntaps; ++t, kcol += tap_kstep) {
                            ptx::mbar_wait(&b_empty[bs], bphase ^ 1);
                            if (ptx::elect_one()) {
                                if (debug & 8) {
                                    ptx::mbar_arrive(&b_full[bs]);
                                } else {
                                    ptx::mbar_expect_tx(&b_full[bs], kBBytes);
                                    ptx::tma_load_2d(smem_b + bs * kBBytes, mapB, &b_full[bs], kcol, co0);
                                }
                            }
                            __syncwarp();
                            if (++bs == b_stages) {
                                bs = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BN, 0, 0);
        const uint64_t adesc0 = ptx::make_smem_desc(ptx::smem_u32(smem), 16, p.a_sbo);
        const uint64_t bdesc0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, 1024);
        const uint32_t a_stage16 = a_stage_bytes >> 4;
        const int kmmas = p.kmmas;
        const bool do_mma = !(debug & 4);
        constexpr int NTR = NT > 0 ? NT : 1;
        uint32_t toff[NTR];
#pragma unroll
        for (int t = 0; t < NTR; ++t) toff[t] = p.tap_off16[t];
        if (RESIDENT) {
            ptx::mbar_wait(bres_bar, 0);
            ptx::tc_fence_after();
        }
        int as = 0, bs = 0;
        uint32_t aphase = 0, bphase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t accphase = (it >> 1) & 1;
            const bool stamp = p.timeline && blockIdx.x == 0 && lane == 0 && it < 64;
            if (stamp) p.timeline[(1 * 64 + it) * 4 + 0] = clock64();
            ptx::mbar_wait(&tempty_bar[acc], accphase ^ 1);
            ptx::tc_fence_after();
            if (stamp) p.timeline[(1 * 64 + it) * 4 + 1] = clock64();
            const uint32_t d_tmem = tmem_base + acc * BN;
            uint32_t accflag = 0;                                   // 0 only for the first MMA of the tile
            for (int term = 0; term < nterms; ++term) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    ptx::mbar_wait(&a_full[as], aphase);
                    ptx::tc_fence_after();
                    if (stamp && kc == 0 && term == 0) p.timeline[(1 * 64 + it) * 4 + 2] = clock64();
                    const uint64_t ad = adesc0 + static_cast<uint64_t>(as * a_stage16);
                    if (NT > 0 && RESIDENT) {
                        // the whole chunk (NT taps x 4 MMAs) is one straight-line issue sequence
                        const uint64_t bd = bdesc0 + static_cast<uint64_t>(kc * kB16);
                        const uint32_t bstep = kchunks * kB16;
                        if (ptx::elect_one()) {
                            if (do_mma) {
#pragma unroll
                                for (int t = 0; t < NTR; ++t) {
#pragma unroll
                                    for (int j = 0; j < (KM > 0 ? KM : 4); ++j)
                                        if (KM > 0 || j < kmmas)
                                            ptx::umma_f16(d_tmem, ad + toff[t] + 2 * j, bd + t * bstep + 2 * j, idesc,
                                                          (t | j) != 0 ? 1u : accflag);
                                }
                            }
                            ptx::umma_commit(&a_empty[as]);
                        }
                        __syncwarp();
                    } else if (NT > 0) {
#pragma unroll
                        for (int t = 0; t < NTR; ++t) {
                            ptx::mbar_wait(&b_full[bs], bphase);
                            ptx::tc_fence_after();
                            const uint64_t bd = bdesc0 + static_cast<uint64_t>(bs * kB16);
                            if (ptx::elect_one()) {
                                if (do_mma) {
#pragma unroll
                                    for (int j = 0; j < (KM > 0 ? KM : 4); ++j)
                                        if (KM > 0 || j < kmmas)
                                            ptx::umma_f16(d_tmem, ad + toff[t] + 2 * j, bd + 2 * j, idesc,
                                                          (t | j) != 0 ? 1u : accflag);
                                }
                                ptx::umma_commit(&b_empty[bs]);
                                if (t == NTR - 1) ptx::umma_commit(&a_empty[as]);
                            }
                            __syncwarp();
                            if (++bs == b_stages) {
                                bs = 0;
                                bphase ^= 1;
                            }
                        }
                    } else {
                        // run-time tap count (5x5, 7x7 ... filters): same protocol, offsets read from the parameters
#pragma unroll 1
                        for (int t = 0; t < ntaps; ++t) {
                            uint64_t bd;
                            if (RESIDENT) {
                                bd = bdesc0 + static_cast<uint64_t>((t * kchunks + kc) * kB16);
                            } else {
                                ptx::mbar_wait(&b_full[bs], bphase);
                                ptx::tc_fence_after();
                                bd = bdesc0 + static_cast<uint64_t>(bs * kB16);
                            }
                            const uint64_t at = ad + p.tap_off16[t];
                            if (ptx::elect_one()) {
                                if (do_mma) {
#pragma unroll
                                    for (int j = 0; j < (KM > 0 ? KM : 4); ++j)
                                        if (KM > 0 || j < kmmas)
                                            ptx::umma_f16(d_tmem, at + 2 * j, bd + 2 * j, idesc,
                                                          (t | j) != 0 ? 1u : accflag);
                                }
                                if (!RESIDENT) ptx::umma_commit(&b_empty[bs]);
                                if (t == ntaps - 1) ptx::umma_commit(&a_empty[as]);
                            }
                            __syncwarp();
                            if (!RESIDENT && ++bs == b_stages) {
                                bs = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                    accflag = 1;
                    if (++as == a_stages) {
                        as = 0;
                        aphase ^= 1;
                    }
                }
            }
            if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[acc]);
            __syncwarp();
            if (stamp) p.timeline[(1 * 64 + it) * 4 + 3] = clock64();
        }
    } else if (p.epi_tma) {
        fprop_epilogue_tma<BN, false>(p, smem + p.stage_off, reinterpret_cast<float*>(stat_smem), tmem_base, tfull_bar,
                                      tempty_bar, warp, lane);
    } else {
        fprop_epilogue<BN>(p, stat_smem, tmem_base, tfull_bar, tempty_bar, warp, lane);
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ fprop, CTA pairs
// conv_fprop_halo_kernel on CTA PAIRS (tcgen05 cta_group::2, cluster of 2 on one TPC): one MMA instruction of M = 256
// covers two adjacent 128-pixel patches (CTA rank r owns patch 2*mt + r: its own halo'd A box, its own 128 accumulator
// lanes, its own epilogue) and reads the filter tile split in halves - rank r keeps rows [r*BN/2, (r+1)*BN/2) of every
// (tap, chunk) B tile in ITS shared memory.  What that buys:
//   * the L2 -> shared-memory fill of the filters per SM halves.  The 256-channel layers stream 9 x 32 KB of B per
//     36 MMAs (4608 tensor cycles) per CTA = ~67 B/clk/SM, 9.9 KB/clk chip-wide against the ~6.3 KB/clk the L2 delivers
//     (B300_MICROARCH.md "LTS throughput cap"); with halves it is 36 B/clk/SM;
//   * one issuing thread feeds TWO tensor cores: the N = 64 / 128 layers were bound by the ~50 cycles the single thread
//     needs per tcgen05.mma against the 32 / 64 cycles the instruction occupies the pipe.
// Protocol: "full" barriers (A stage, B stage) live in the leader (rank 0) and count one arrive.expect_tx per CTA plus
// the bytes of both CTAs' TMA loads (cp.async.bulk.tensor ... .cta_group::2 signals the leader's barrier from either
// CTA); the leader's MMA thread multicasts its commits to the "empty" / "accumulator full" barriers of both CTAs; the
// epilogue warps of both CTAs release the accumulator on the leader's "accumulator empty" barrier (count 16).
template <int BN, int NT, bool RESIDENT, int KM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsF, 1)
    conv_fprop_halo2_kernel(const __grid_constant__ ConvFpropParams p) {
    constexpr int kBHalfBytes = (BN / 2) * kBK * 2;          // this CTA's half of a B tile
    constexpr uint32_t kB16 = kBHalfBytes >> 4;
    // taps per filter stage (halo2_taps_per_stage on the host): with N <= 128 the per-stage handshake of the issuing
    // thread (wait, fence, elect, commit, ~150 cycles) plus four ~50-cycle MMA issues exceeds the 4 x 64 cycles the MMAs
    // occupy the tensor pipe - one stage per filter ROW (12 MMAs per handshake) makes those layers MMA-bound again
    constexpr int TG = (!RESIDENT && NT == 9 && BN <= 128) ? 3 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_b = smem + p.a_region_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + p.b_region_bytes);
    uint64_t* a_empty = a_full + 8;
    uint64_t* b_full = a_empty + 8;
    uint64_t* b_empty = b_full + 16;
    uint64_t* tfull_bar = b_empty + 16;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* bres_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 2);    // its own 16-byte granule (racecheck pairs
                                                                            // tcgen05.alloc's result slot with neighbouring barrier writes)
    float2* stat_smem = reinterpret_cast<float2*>(smem_b + p.b_region_bytes + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            ptx::mbar_init(&a_full[s], 2);        // one arrive.expect_tx per CTA of the pair
            ptx::mbar_init(&a_empty[s], 1);       // multicast commit of the leader's MMA thread
        }
        for (int s = 0; s < 16; ++s) {
            ptx::mbar_init(&b_full[s], 2);
            ptx::mbar_init(&b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 2 * kEpiWarpsF);     // the epilogue warps of both CTAs
        }
        ptx::mbar_init(bres_bar, 2);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair(tmem_slot, kTmemCols);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    __syncthreads();                              // (CTA-level ordering of the TMEM address slot; the cluster barrier below
                                                  // implies it, but compute-sanitizer's racecheck only models bar.sync)
    ptx::cluster_sync_all();                      // barriers of BOTH CTAs initialised before any remote arrive
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kchunks = p.kchunks, nterms = p.nterms, num_tiles = p.num_tiles;
    const int a_stages = p.a_stages, b_stages = p.b_stages;
    const uint32_t a_stage_bytes = p.a_stage_bytes;
    const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;
    const int debug = p.debug;        // profiling knobs (results are garbage): bit2 no MMAs, bit3 no A / B loads

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (both CTAs: own A box, own half of B)
        if (ptx::elect_one()) {
            ptx::tma_prefetch_desc(&p.tmA[0]);
            ptx::tma_prefetch_desc(&p.tmB[0]);
        }
        const uint32_t a_full0 = ptx::mapa_u32(ptx::smem_u32(a_full), 0);     // the leader's barriers
        const uint32_t b_full0 = ptx::mapa_u32(ptx::smem_u32(b_full), 0);
        const int brow = static_cast<int>(rank) * (BN / 2);
        if (RESIDENT) {
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx_cluster(ptx::mapa_u32(ptx::smem_u32(bres_bar), 0), p.b_region_bytes);
                const int nb = NT * kchunks;
                for (int i = 0; i < nb; ++i)
                    ptx::tma_load_2d_pair(smem_b + i * kBHalfBytes, &p.tmB[0], ptx::mapa_u32(ptx::smem_u32(bres_bar), 0),
                                          i * kBK, brow);
            }
            __syncwarp();
        }
        const int a_loads = p.a_loads, a_dh_step = p.a_dh_step;
        const uint32_t a_load_stride = p.a_load_stride, a_tx = p.a_loads * p.a_load_bytes;
        const int tile_w = p.TW * p.stride_w, tile_h = p.TH * p.stride_h, a_dw = p.a_dw, a_dh = p.a_dh, TN = p.TN;
        const int tap_kstep = kchunks * kBK;
        int as = 0, bs = 0;
        uint32_t aphase = 0, bphase = 0;
        griddep_wait();                      // the activations come from the previous kernel (the filters do not)
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            uint32_t ct, mt, tw, th, tn;
            p.fd_co.divmod(tile, mt, ct);
            mt = 2 * mt + rank;                   // an M tile past the end is all TMA zero fill (n0 >= N) and stores nothing
            p.fd_w.divmod(mt, mt, tw);
            p.fd_h.divmod(mt, tn, th);
            const int w0 = tw * tile_w + a_dw, h0 = th * tile_h + a_dh;
            const int n0 = tn * TN, co0 = ct * BN + brow;
            for (int term = 0; term < nterms; ++term) {
                const CUtensorMap* mapA = &p.tmA[(nterms == 3 && term == 0) ? 1 : 0];
                const CUtensorMap* mapB = &p.tmB[(nterms == 3 && term == 1) ? 1 : 0];
                for (int kc = 0; kc < kchunks; ++kc) {
                    ptx::mbar_wait(&a_empty[as], aphase ^ 1);
                    uint8_t* sA = smem + as * a_stage_bytes;
                    if (ptx::elect_one()) {
                        const uint32_t bar = a_full0 + as * 8;
                        if (debug & 8) {
                            ptx::mbar_arrive_cluster(bar);
                        } else {
                            ptx::mbar_expect_tx_cluster(bar, a_tx);
                            for (int i = 0; i < a_loads; ++i)
                                ptx::tma_load_4d_pair(sA + i * a_load_stride, mapA, bar, kc * kBK, w0, h0 + i * a_dh_step,
                                                      n0);
                        }
                    }
                    __syncwarp();
                    if (++as == a_stages) {
                        as = 0;
                        aphase ^= 1;
                    }
                    if (!RESIDENT) {
                        int kcol = kc * kBK;
#pragma unroll 1
                        for (int t = 0; t < NT; t += TG, kcol += TG * tap_kstep) {
                            ptx::mbar_wait(&b_empty[bs], bphase ^ 1);
                            if (ptx::elect_one()) {
                                const uint32_t bar = b_full0 + bs * 8;
                                if (debug & 8) {
                                    ptx::mbar_arrive_cluster(bar);
                                } else {
                                    ptx::mbar_expect_tx_cluster(bar, TG * kBHalfBytes);
#pragma unroll
                                    for (int u = 0; u < TG; ++u)
                                        ptx::tma_load_2d_pair(smem_b + (bs * TG + u) * kBHalfBytes, mapB, bar,
                                                              kcol + u * tap_kstep, co0);
                                }
                            }
                            __syncwarp();
                            if (++bs == b_stages) {
                                bs = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                }
            }
        }
        griddep_launch();                         // all loads issued: the next kernel may take the SMs this grid leaves
    } else if (warp == 1 && rank == 0) {
        // ------------------------------------------------ MMA issuer (leader CTA only)
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BN, 0, 0, 256);
        const uint64_t adesc0 = ptx::make_smem_desc(ptx::smem_u32(smem), 16, p.a_sbo);
        const uint64_t bdesc0 = ptx::make_smem_desc(ptx::smem_u32(smem_b), 16, 1024);
        const uint32_t a_stage16 = a_stage_bytes >> 4;
        uint32_t toff[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) toff[t] = p.tap_off16[t];
        if (RESIDENT) {
            ptx::mbar_wait(bres_bar, 0);
            ptx::tc_fence_after();
        }
        int as = 0, bs = 0;
        uint32_t aphase = 0, bphase = 0;
        int it = 0;
        const bool do_mma = !(debug & 4);
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            const int acc = it & 1;
            const uint32_t accphase = (it >> 1) & 1;
            ptx::mbar_wait(&tempty_bar[acc], accphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            uint32_t accflag = 0;
            for (int term = 0; term < nterms; ++term) {
                for (int kc = 0; kc < kchunks; ++kc) {
                    ptx::mbar_wait(&a_full[as], aphase);
                    ptx::tc_fence_after();
                    const uint64_t ad = adesc0 + static_cast<uint64_t>(as * a_stage16);
                    if (RESIDENT) {
                        const uint64_t bd = bdesc0 + static_cast<uint64_t>(kc * kB16);
                        const uint32_t bstep = kchunks * kB16;
                        if (ptx::elect_one()) {
                            if (do_mma) {
#pragma unroll
                                for (int t = 0; t < NT; ++t) {
#pragma unroll
                                    for (int j = 0; j < KM; ++j)
                                        ptx::umma_f16_pair(d_tmem, ad + toff[t] + 2 * j, bd + t * bstep + 2 * j, idesc,
                                                           (t | j) != 0 ? 1u : accflag);
                                }
                            }
                            ptx::umma_commit_pair(&a_empty[as]);
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int t = 0; t < NT; t += TG) {
                            ptx::mbar_wait(&b_full[bs], bphase);
                            ptx::tc_fence_after();
                            const uint64_t bd = bdesc0 + static_cast<uint64_t>(bs * (TG * kB16));
                            if (ptx::elect_one()) {
                                if (do_mma) {
#pragma unroll
                                    for (int u = 0; u < TG; ++u)
#pragma unroll
                                        for (int j = 0; j < KM; ++j)
                                            ptx::umma_f16_pair(d_tmem, ad + toff[t + u] + 2 * j, bd + u * kB16 + 2 * j, idesc,
                                                               (t | u | j) != 0 ? 1u : accflag);
                                }
                                ptx::umma_commit_pair(&b_empty[bs]);
                                if (t + TG >= NT) ptx::umma_commit_pair(&a_empty[as]);
                            }
                            __syncwarp();
                            if (++bs == b_stages) {
                                bs = 0;
                                bphase ^= 1;
                            }
                        }
                    }
                    accflag = 1;
                    if (++as == a_stages) {
                        as = 0;
                        aphase ^= 1;
                    }
                }
            }
            if (ptx::elect_one()) ptx::umma_commit_pair(&tfull_bar[acc]);
            __syncwarp();
        }
    } else if (warp >= 2) {
        griddep_wait();                      // bias / residual reads, output and statistics writes
        if (p.epi_tma)
            fprop_epilogue_tma<BN, true>(p, smem + p.stage_off, reinterpret_cast<float*>(stat_smem), tmem_base, tfull_bar,
                                         tempty_bar, warp, lane);
        else
            fprop_epilogue<BN, true>(p, stat_smem, tmem_base, tfull_bar, tempty_bar, warp, lane);
    }

    // the peer's shared memory and barriers must outlive every MMA read / remote arrive of the pair
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad
// KS = k-blocks (64 pixels each) per pipeline stage: one full / empty handshake of the issuing thread (~150 cycles) per
// 4 * KS MMAs.  With BN <= 128 four MMAs occupy the tensor pipe for <= 256 cycles, less than the handshake plus four
// ~50-cycle issues: two k-blocks per stage make those instances MMA-bound (STAGES counts stages of KS k-blocks).
template <int BN, int STAGES, int KS>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const __grid_constant__ ConvWgradParams p) {
    using L = SmemLayout<BN, STAGES * KS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
    constexpr int kBoxBytes = kBK * 128;  // one [64 pixels][64 channels] bf16 box

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntaps = p.R * p.S;
    // loop-invariant parameters in registers (the producer / MMA roles are single instruction streams)
    const int nterms = p.nterms, debug = p.debug, num_tiles = p.num_tiles;

    // tile -> (co tile, ci tile, tap, split); split is the slowest index so that concurrently running CTAs
    // share the same pixel range (L2 reuse of dY / X boxes).
    auto decode = [&](int tile, int& cot, int& cit, int& tap, int& kb0, int& kb1) {
        uint32_t t, a, b, c;
        p.fd_cot.divmod(tile, t, a);
        p.fd_cit.divmod(t, t, b);
        p.fd_taps.divmod(t, t, c);
        cot = a; cit = b; tap = c;
        const int split = t;
        kb0 = static_cast<int>(static_cast<long long>(p.total_kblocks) * split / p.splits);
        kb1 = static_cast<int>(static_cast<long long>(p.total_kblocks) * (split + 1) / p.splits);
    };

    if (warp == 0) {
        {
            if (ptx::elect_one()) {
                ptx::tma_prefetch_desc(&p.tmDY[0]);
                ptx::tma_prefetch_desc(&p.tmX[0]);
            }
            int stage = 0;
            uint32_t phase = 0;
            const int TW = p.TW, TH = p.TH, TN = p.TN, a_boxes = p.a_boxes, stride_w = p.stride_w, stride_h = p.stride_h;
            const uint32_t tiles_w = p.tiles_w, tiles_h = p.tiles_h;
            const uint32_t stage_tx = L::kStageBytes - (2 - a_boxes) * kBoxBytes;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int cot, cit, tap, kb0, kb1;
                decode(tile, cot, cit, tap, kb0, kb1);
                const int dh = tap / p.S - p.pad_h;
                const int dw = tap % p.S - p.pad_w;
                const int cA = cot * kBM, cB = cit * BN;
                // terms of the fp32-parity split are the OUTER loop, corrections first: (lo,hi) (hi,lo) then (hi,hi)
                // (the accumulator truncates on every step; see conv_fprop_kernel)
                for (int term = 0; term < nterms; ++term) {
                const int ai = (nterms == 3 && term == 0) ? 1 : 0;
                const int bi = (nterms == 3 && term == 1) ? 1 : 0;
                // k-block kb0 -> patch position, then advanced incrementally (no division in the loop)
                uint32_t tw, th, tn, t2;
                p.fd_w.divmod(kb0, t2, tw);
                p.fd_h.divmod(t2, tn, th);
#pragma unroll 1
                for (int kb = kb0; kb < kb1; kb += KS) {
                    const int nsub = kb1 - kb < KS ? kb1 - kb : KS;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (ptx::elect_one()) {
                        if (debug & 2)
                            ptx::mbar_arrive(&full_bar[stage]);
                        else
                            ptx::mbar_expect_tx(&full_bar[stage], nsub * stage_tx);
                    }
#pragma unroll
                    for (int u = 0; u < KS; ++u) {
                        if (u < nsub) {
                            const int w0 = tw * TW, h0 = th * TH, n0 = tn * TN;
                            uint8_t* sA = smem + (stage * KS + u) * L::kStageBytes;
                            uint8_t* sB = sA + L::kABytes;
                            if (!(debug & 2) && ptx::elect_one()) {
                                // Cout <= 64: the second 64-channel dY box would be all TMA zero fill; skip it (its
                                // accumulator rows are never read)
                                ptx::tma_load_4d(sA, &p.tmDY[ai], &full_bar[stage], cA, w0, h0, n0);
                                if (a_boxes > 1)
                                    ptx::tma_load_4d(sA + kBoxBytes, &p.tmDY[ai], &full_bar[stage], cA + 64, w0, h0, n0);
#pragma unroll
                                for (int i = 0; i < BN / 64; ++i)
                                    ptx::tma_load_4d(sB + i * kBoxBytes, &p.tmX[bi], &full_bar[stage],
                                                     cB + i * 64, w0 * stride_w + dw, h0 * stride_h + dh, n0);
                            }
                            if (++tw == tiles_w) {
                                tw = 0;
                                if (++th == tiles_h) {
                                    th = 0;
                                    ++tn;
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                }
            }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BN, 1, 1);
            const uint64_t adesc0 = ptx::make_smem_desc(ptx::smem_u32(smem), kBoxBytes, 1024);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const bool do_mma = !(debug & 1);
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                int cot, cit, tap, kb0, kb1;
                decode(tile, cot, cit, tap, kb0, kb1);
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                uint32_t accflag = 0;                       // 0 only for the first MMA of the tile
                for (int term = 0; term < nterms; ++term) {
#pragma unroll 1
                    for (int kb = kb0; kb < kb1; kb += KS) {
                        const int nsub = kb1 - kb < KS ? kb1 - kb : KS;
                        ptx::mbar_wait(&full_bar[stage], phase);
                        ptx::tc_fence_after();
                        // MN-major: LBO = next 64-channel box, SBO = next group of 8 pixel rows
                        const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * KS * (L::kStageBytes >> 4));
                        const uint64_t bdesc = adesc + (L::kABytes >> 4);
                        if (ptx::elect_one()) {
                            if (do_mma) {
#pragma unroll
                                for (int u = 0; u < KS; ++u) {
                                    if (u < nsub) {
#pragma unroll
                                        for (int j = 0; j < kBK / 16; ++j) {
                                            // advance 16 pixel rows = 2048 B; next k-block of the stage = kStageBytes on
                                            ptx::umma_f16(d_tmem, adesc + u * (L::kStageBytes >> 4) + 128 * j,
                                                          bdesc + u * (L::kStageBytes >> 4) + 128 * j, idesc,
                                                          (u | j) != 0 ? 1u : accflag);
                                        }
                                    }
                                }
                            }
                            ptx::umma_commit(&empty_bar[stage]);
                        }
                        __syncwarp();
                        accflag = 1;
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[as]);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            int cot, cit, tap, kb0, kb1;
            decode(tile, cot, cit, tap, kb0, kb1);
            const int split = static_cast<int>(p.fd_taps.div(p.fd_cit.div(p.fd_cot.div(tile))));
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int co = cot * kBM + row;
            const bool row_ok = co < p.Cout;
            float* dst_row =
                p.ws + ((static_cast<long long>(split) * p.Cout + co) * ntaps + tap) * static_cast<long long>(p.ldws);
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c0;
                ptx::tmem_ld_32x32b_x32(taddr, r);
                ptx::tmem_ld_wait();
                const int ci = cit * BN + c0;
                int nvalid = p.Cin - ci;
                nvalid = nvalid > 32 ? 32 : nvalid;
                if (row_ok && nvalid > 0 && !(p.debug & 4)) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    store_row_chunk(dst_row, 1, ci, v, nvalid);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad, CTA pairs
// conv_wgrad_kernel<256> on CTA PAIRS (tcgen05 cta_group::2): the two CTAs of a cluster take two adjacent Cout tiles
// (rank r: rows [(2 cp + r) * 128, +128) of the filter gradient - its own dY boxes, its own accumulator lanes, its own
// epilogue) of the SAME (Cin tile, tap, split), so they share the X operand: one MMA of M = 256 reads N / 2 = 128 input
// channels of the X box from each CTA's shared memory.  Per k-block a CTA fetches 16 KB of dY + 16 KB of X instead of
// 16 + 32 KB: the one-CTA kernel needs 148 x 48 KB per 512 tensor cycles = ~21 TB/s of L2 -> shared-memory fill to keep
// the tensor pipe busy, the L2 delivers ~14 (measured: 958 TFLOP/s = 0.68 of the MMA rate on the 256-channel layers).
// Barrier protocol as in conv_fprop_halo2_kernel: "full" barriers in the leader (2 arrive.expect_tx + the bytes of both
// CTAs' loads), multicast commits to the "empty" / "accumulator full" barriers of both CTAs, both epilogues release
// the accumulator on the leader's barrier.
template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
    conv_wgrad2_kernel(const __grid_constant__ ConvWgradParams p) {
    constexpr int BN = 256;
    constexpr int kBoxBytes = kBK * 128;                  // one [64 pixels][64 channels] bf16 box
    constexpr int kABytes = 2 * kBoxBytes;                // dY: this CTA's 128 output channels
    constexpr int kBBytes = (BN / 128) * kBoxBytes;       // X: this CTA's half (128 input channels)
    constexpr int kStageBytes = kABytes + kBBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    constexpr uint32_t kTmemCols = 2 * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 2);              // one arrive.expect_tx per CTA of the pair
            ptx::mbar_init(&empty_bar[s], 1);             // multicast commit of the leader's MMA thread
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 2 * 4);        // the epilogue warps of both CTAs
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_pair(tmem_slot, kTmemCols);
        ptx::tmem_relinquish_pair();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();                              // barriers of BOTH CTAs initialised before any remote arrive
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ntaps = p.R * p.S;
    const int nterms = p.nterms, debug = p.debug, num_tiles = p.num_tiles;
    const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

    // pair tile -> (co tile PAIR, ci tile, tap, split): p.fd_cot divides by the number of pairs
    auto decode = [&](int tile, int& cop, int& cit, int& tap, int& split, int& kb0, int& kb1) {
        uint32_t t, a, b, c;
        p.fd_cot.divmod(tile, t, a);
        p.fd_cit.divmod(t, t, b);
        p.fd_taps.divmod(t, t, c);
        cop = a; cit = b; tap = c;
        split = t;
        kb0 = static_cast<int>(static_cast<long long>(p.total_kblocks) * split / p.splits);
        kb1 = static_cast<int>(static_cast<long long>(p.total_kblocks) * (split + 1) / p.splits);
    };

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (both CTAs: own dY boxes, own half of X)
        if (ptx::elect_one()) {
            ptx::tma_prefetch_desc(&p.tmDY[0]);
            ptx::tma_prefetch_desc(&p.tmX[0]);
        }
        const uint32_t full0 = ptx::mapa_u32(ptx::smem_u32(full_bar), 0);      // the leader's barriers
        int stage = 0;
        uint32_t phase = 0;
        const int TW = p.TW, TH = p.TH, TN = p.TN, stride_w = p.stride_w, stride_h = p.stride_h;
        const uint32_t tiles_w = p.tiles_w, tiles_h = p.tiles_h;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int cop, cit, tap, split, kb0, kb1;
            decode(tile, cop, cit, tap, split, kb0, kb1);
            const int dh = tap / p.S - p.pad_h;
            const int dw = tap % p.S - p.pad_w;
            const int cA = (2 * cop + static_cast<int>(rank)) * kBM;
            const int cB = cit * BN + static_cast<int>(rank) * (BN / 2);
            for (int term = 0; term < nterms; ++term) {
                // terms of the fp32-parity split, corrections first (see conv_wgrad_kernel)
                const int ai = (nterms == 3 && term == 0) ? 1 : 0;
                const int bi = (nterms == 3 && term == 1) ? 1 : 0;
                uint32_t tw, th, tn, t2;
                p.fd_w.divmod(kb0, t2, tw);
                p.fd_h.divmod(t2, tn, th);
#pragma unroll 1
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int w0 = tw * TW, h0 = th * TH, n0 = tn * TN;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * kStageBytes;
                    uint8_t* sB = sA + kABytes;
                    if (ptx::elect_one()) {
                        const uint32_t bar = full0 + stage * 8;
                        if (debug & 2) {
                            ptx::mbar_arrive_cluster(bar);
                        } else {
                            ptx::mbar_expect_tx_cluster(bar, kStageBytes);
                            ptx::tma_load_4d_pair(sA, &p.tmDY[ai], bar, cA, w0, h0, n0);
                            ptx::tma_load_4d_pair(sA + kBoxBytes, &p.tmDY[ai], bar, cA + 64, w0, h0, n0);
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i)
                                ptx::tma_load_4d_pair(sB + i * kBoxBytes, &p.tmX[bi], bar, cB + i * 64,
                                                      w0 * stride_w + dw, h0 * stride_h + dh, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (++tw == tiles_w) {
                        tw = 0;
                        if (++th == tiles_h) {
                            th = 0;
                            ++tn;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ------------------------------------------------ MMA issuer (leader CTA only)
        constexpr uint32_t idesc = ptx::make_idesc_bf16(BN, 1, 1, 256);
        const uint64_t adesc0 = ptx::make_smem_desc(ptx::smem_u32(smem), kBoxBytes, 1024);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        const bool do_mma = !(debug & 1);
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            int cop, cit, tap, split, kb0, kb1;
            decode(tile, cop, cit, tap, split, kb0, kb1);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * BN;
            const int nk = (kb1 - kb0) * nterms;
#pragma unroll 1
            for (int kb = 0; kb < nk; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                // MN-major: LBO = next 64-channel box, SBO = next group of 8 pixel rows
                const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * (kStageBytes >> 4));
                const uint64_t bdesc = adesc + (kABytes >> 4);
                if (ptx::elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int j = 0; j < kBK / 16; ++j)      // advance 16 pixel rows = 2048 B
                            ptx::umma_f16_pair(d_tmem, adesc + 128 * j, bdesc + 128 * j, idesc, (kb | j) != 0);
                    }
                    ptx::umma_commit_pair(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (ptx::elect_one()) ptx::umma_commit_pair(&tfull_bar[as]);
            __syncwarp();
        }
    } else if (warp >= 2) {
        // ------------------------------------------------ epilogue (both CTAs: own 128 rows x 256 columns)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t tempty0 = ptx::mapa_u32(ptx::smem_u32(tempty_bar), 0);
        int it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
            int cop, cit, tap, split, kb0, kb1;
            decode(tile, cop, cit, tap, split, kb0, kb1);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int co = (2 * cop + static_cast<int>(rank)) * kBM + row;
            const bool row_ok = co < p.Cout;
            float* dst_row =
                p.ws + ((static_cast<long long>(split) * p.Cout + co) * ntaps + tap) * static_cast<long long>(p.ldws);
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + c0;
                ptx::tmem_ld_32x32b_x32(taddr, r);
                ptx::tmem_ld_wait();
                const int ci = cit * BN + c0;
                int nvalid = p.Cin - ci;
                nvalid = nvalid > 32 ? 32 : nvalid;
                if (row_ok && nvalid > 0 && !(p.debug & 4)) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                    store_row_chunk(dst_row, 1, ci, v, nvalid);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(tempty0 + as * 8);
        }
    }

    // the peer's shared memory and barriers must outlive every MMA read / remote arrive of the pair
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad, row-shared
// Stride-1 filter gradient with the S taps of one filter row computed from ONE pair of TMA loads per k-block: the X
// box is fetched with an (S-1)-pixel halo along W and tap s reads it through a descriptor that starts s rows (s*128 B)
// into the box.  Tile = (Cout tile, Cin tile, filter row r, split); accumulators = S x BN TMEM columns.  Compared
// with conv_wgrad_kernel (one tap per tile) the L2 -> shared-memory traffic of a 3x3 layer drops 3x, which is what
// bounds the small-channel layers.
template <int BN>
struct RowsLayout {
    static constexpr int kABytes = kBM * kBK * 2;                  // dY: 2 sub-boxes [64 px][64 co]
    static constexpr int kBMaxBytes = (BN / 64) * 10240;           // X: BN/64 sub-boxes of <= 80 rows
    static constexpr int kStageBytes = kABytes + kBMaxBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kBarOffset = kStages * kStageBytes;
    static constexpr int kTotal = kBarOffset + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_rows_kernel(const __grid_constant__ ConvWgradParams p) {
    using L = RowsLayout<BN>;
    constexpr int STAGES = L::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 512;
    constexpr int kBoxBytes = kBK * 128;
    const int nbuf = (2 * p.S * BN <= 512) ? 2 : 1;    // accumulator sets (S * BN columns each)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tfull_bar[s], 1);
            ptx::mbar_init(&tempty_bar[s], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int tile, int& cot, int& cit, int& r, int& kb0, int& kb1, int& split) {
        uint32_t t, a, b, c;
        p.fd_cot.divmod(tile, t, a);
        p.fd_cit.divmod(t, t, b);
        p.fd_taps.divmod(t, t, c);      // here: filter rows (R)
        cot = a; cit = b; r = c;
        split = t;
        kb0 = static_cast<int>(static_cast<long long>(p.total_kblocks) * split / p.splits);
        kb1 = static_cast<int>(static_cast<long long>(p.total_kblocks) * (split + 1) / p.splits);
    };
    const uint32_t stage_tx = p.a_boxes * kBoxBytes + (BN / 64) * p.xrows * 128;

    if (warp == 0) {
        {
            if (ptx::elect_one()) {
                ptx::tma_prefetch_desc(&p.tmDY[0]);
                ptx::tma_prefetch_desc(&p.tmX[0]);
            }
            int stage = 0;
            uint32_t phase = 0;
            const bool two_a = p.a_boxes > 1;
            const int TW = p.TW, TH = p.TH, TN = p.TN, nterms = p.nterms, debug = p.debug, pad_w = p.pad_w;
            const int num_tiles = p.num_tiles, xbox_bytes = p.xbox_bytes;
            const uint32_t tiles_w = p.tiles_w, tiles_h = p.tiles_h;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int cot, cit, r, kb0, kb1, split;
                decode(tile, cot, cit, r, kb0, kb1, split);
                const int dh = r - p.pad_h;
                const int cA = cot * kBM, cB = cit * BN;
                for (int term = 0; term < nterms; ++term) {      // corrections first, (hi,hi) last (see conv_fprop_kernel)
                const CUtensorMap* mapA = &p.tmDY[(nterms == 3 && term == 0) ? 1 : 0];
                const CUtensorMap* mapB = &p.tmX[(nterms == 3 && term == 1) ? 1 : 0];
                uint32_t tw, th, tn, t2;
                p.fd_w.divmod(kb0, t2, tw);
                p.fd_h.divmod(t2, tn, th);
#pragma unroll 1
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int w0 = tw * TW, h0 = th * TH, n0 = tn * TN;
                    {
                        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sA = smem + stage * L::kStageBytes;
                        uint8_t* sB = sA + L::kABytes;
                        if (ptx::elect_one()) {
                            if (debug & 2) {
                                ptx::mbar_arrive(&full_bar[stage]);
                            } else {
                                ptx::mbar_expect_tx(&full_bar[stage], stage_tx);
                                ptx::tma_load_4d(sA, mapA, &full_bar[stage], cA, w0, h0, n0);
                                if (two_a)
                                    ptx::tma_load_4d(sA + kBoxBytes, mapA, &full_bar[stage], cA + 64, w0, h0, n0);
#pragma unroll
                                for (int i = 0; i < BN / 64; ++i)
                                    ptx::tma_load_4d(sB + i * xbox_bytes, mapB, &full_bar[stage], cB + i * 64,
                                                     w0 - pad_w, h0 + dh, n0);
                            }
                        }
                        __syncwarp();
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (++tw == tiles_w) {
                        tw = 0;
                        if (++th == tiles_h) {
                            th = 0;
                            ++tn;
                        }
                    }
                }
                }
            }
        }
    } else if (warp == 1) {
        {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(BN, 1, 1);
            const int xw = p.TW + p.S - 1;                 // X rows per image row of the patch
            // one K=16 MMA = two 8-row groups of dY; with TW == 8 they sit in consecutive image rows of the X box
            const uint32_t sbo_x = (p.TW >= 16) ? 1024u : static_cast<uint32_t>(xw) * 128u;
            // descriptors of stage 0; everything else is an offset in 16-byte units added to the address field:
            //   stage        -> stage * kStageBytes / 16
            //   MMA j (K=16) -> dY: 16 rows = 2048 B;  X: rows of pixel (16j) of the patch inside the halo'd box
            //   tap s        -> X: s rows = s * 128 B.   The base-offset field stays 0: the hardware derives the
            //   128B-swizzle phase from the absolute shared-memory address, which is also how TMA wrote the box
            //   (measured: with the field set the result is wrong).
            const uint32_t smem0 = ptx::smem_u32(smem);
            const uint64_t adesc0 = ptx::make_smem_desc(smem0, kBoxBytes, 1024);
            const uint64_t bdesc0 = ptx::make_smem_desc(smem0 + L::kABytes, p.xbox_bytes, sbo_x);
            uint32_t xoff[kBK / 16];
#pragma unroll
            for (int j = 0; j < kBK / 16; ++j)
                xoff[j] = static_cast<uint32_t>(((16 * j) >> p.log2_tw) * xw + ((16 * j) & (p.TW - 1))) * 8u;
            const int S = p.S;
            const bool do_mma = !(p.debug & 1);
            // Cin tile of 64 channels and S * 64 <= 256: ONE MMA per K step covers all S taps.  The B operand is
            // described as S column groups of 64 channels whose group stride (LBO) is one pixel row of the box
            // (128 B), i.e. group g is the same X box shifted by g pixels = tap g; its accumulator columns are
            // [g*64, g*64+64), exactly where the per-tap MMAs put them.  3x fewer MMA instructions for a 3x3 filter
            // (the single issuing thread is what bounds N = 64 MMAs).
            const bool stacked = (BN == 64) && (S * 64 <= 256) && !(p.debug & 8);
            const uint32_t idesc_stacked = ptx::make_idesc_bf16(S * 64, 1, 1);
            const uint64_t bdesc0_stacked = ptx::make_smem_desc(smem0 + L::kABytes, 128, sbo_x);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const int nterms = p.nterms, num_tiles = p.num_tiles;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                int cot, cit, r, kb0, kb1, split;
                decode(tile, cot, cit, r, kb0, kb1, split);
                const int as = (nbuf == 2) ? (it & 1) : 0;
                const uint32_t aphase = (nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * (S * BN);
                const int nk = (kb1 - kb0) * nterms;
#pragma unroll 1
                for (int kb = 0; kb < nk; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint64_t soff = static_cast<uint64_t>(stage * (L::kStageBytes >> 4));
                    const uint64_t adesc = adesc0 + soff;
                    const uint64_t bdesc = bdesc0 + soff;
                    if (ptx::elect_one()) {
                        if (do_mma && stacked) {
                            const uint64_t bst = bdesc0_stacked + soff;
#pragma unroll
                            for (int j = 0; j < kBK / 16; ++j)
                                ptx::umma_f16(d_tmem, adesc + 128 * j, bst + xoff[j], idesc_stacked, (kb | j) != 0);
                        } else if (do_mma) {
                            if (S == 3) {          // the 3x3 layers: 12 MMAs as one straight-line sequence
#pragma unroll
                                for (int j = 0; j < kBK / 16; ++j) {
                                    const uint64_t bj = bdesc + xoff[j];
#pragma unroll
                                    for (int sx = 0; sx < 3; ++sx)
                                        ptx::umma_f16(d_tmem + sx * BN, adesc + 128 * j, bj + 8 * sx, idesc,
                                                      (kb | j) != 0);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < kBK / 16; ++j) {
                                    const uint64_t bj = bdesc + xoff[j];
                                    for (int sx = 0; sx < S; ++sx)
                                        ptx::umma_f16(d_tmem + sx * BN, adesc + 128 * j, bj + 8 * sx, idesc,
                                                      (kb | j) != 0);
                                }
                            }
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[as]);
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            int cot, cit, r, kb0, kb1, split;
            decode(tile, cot, cit, r, kb0, kb1, split);
            const int as = (nbuf == 2) ? (it & 1) : 0;
            const uint32_t aphase = (nbuf == 2) ? ((it >> 1) & 1) : (it & 1);
            const int co = cot * kBM + row;
            const bool row_ok = co < p.Cout;
            ptx::mbar_wait(&tfull_bar[as], aphase);
            ptx::tc_fence_after();
            for (int s = 0; s < p.S; ++s) {
                float* dst_row = p.ws + ((static_cast<long long>(split) * p.Cout + co) * (p.R * p.S) + r * p.S + s) *
                                            static_cast<long long>(p.ldws);
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t rr[32];
                    const uint32_t taddr =
                        tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * (p.S * BN) + s * BN + c0;
                    ptx::tmem_ld_32x32b_x32(taddr, rr);
                    ptx::tmem_ld_wait();
                    const int ci = cit * BN + c0;
                    int nvalid = p.Cin - ci;
                    nvalid = nvalid > 32 ? 32 : nvalid;
                    if (row_ok && nvalid > 0 && !(p.debug & 4)) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
                        store_row_chunk(dst_row, 1, ci, v, nvalid);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ wgrad, stem
// Filter gradient of the row-folded stem (the S taps of a filter row are folded into K = 64 columns): ALL R filter rows
// in one tile.  Per k-block of 8 x 8 output pixels the dY box is fetched once (conv_wgrad_kernel fetches it once per
// filter row) together with the input rows the block touches - one box for stride 1, the even and the odd input rows as
// two boxes for stride 2 - and filter row r = p + stride * j reads box p starting j image rows (TW * 128 B) in.  The j's
// of one box are stacked along N: the B operand is described as ntap column groups of 64 whose group stride (LBO) is
// one image row of the box, so ONE MMA per K step computes every filter row of that parity (N = 256 + N = 192 for the
// 7-row stem instead of 7 MMAs of N = 64, and 2 instead of 7 barrier round trips per 64 pixels).
// Accumulators: R * 64 TMEM columns, column block of row r = (r % stride) * ntap[0] * 64 + (r / stride) * 64.
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_stem_kernel(const __grid_constant__ ConvWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int STAGES = p.stem_stages;
    const int stage_bytes = p.stem_stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tfull_bar = empty_bar + 8;
    uint64_t* tempty_bar = tfull_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 512;
    constexpr int kBoxBytes = kBK * 128;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 8; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(tfull_bar, 1);
        ptx::mbar_init(tempty_bar, 4);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(tmem_slot, kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.num_tiles, nterms = p.nterms, a_boxes = p.a_boxes, x_loads = p.x_loads;
    const int a_bytes = a_boxes * kBoxBytes, xbox_stride = p.xbox_stride;
    // tile -> (co tile, split)
    auto decode = [&](int tile, int& cot, int& split, int& kb0, int& kb1) {
        uint32_t t, a;
        p.fd_cot.divmod(tile, t, a);
        cot = a;
        split = t;
        kb0 = static_cast<int>(static_cast<long long>(p.total_kblocks) * split / p.splits);
        kb1 = static_cast<int>(static_cast<long long>(p.total_kblocks) * (split + 1) / p.splits);
    };

    if (warp == 0) {
        if (ptx::elect_one()) {
            ptx::tma_prefetch_desc(&p.tmDY[0]);
            ptx::tma_prefetch_desc(&p.tmX[0]);
        }
        int stage = 0;
        uint32_t phase = 0;
        const int TW = p.TW, TH = p.TH, TN = p.TN, stride_h = p.stride_h, debug = p.debug;
        const uint32_t tiles_w = p.tiles_w, tiles_h = p.tiles_h;
        const uint32_t stage_tx = a_bytes + x_loads * (p.xrows * 128);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int cot, split, kb0, kb1;
            decode(tile, cot, split, kb0, kb1);
            const int cA = cot * kBM;
            for (int term = 0; term < nterms; ++term) {          // corrections first, (hi,hi) last (see conv_fprop_kernel)
            const CUtensorMap* mapA = &p.tmDY[(nterms == 3 && term == 0) ? 1 : 0];
            const CUtensorMap* mapB = &p.tmX[(nterms == 3 && term == 1) ? 1 : 0];
            uint32_t tw, th, tn, t2;
            p.fd_w.divmod(kb0, t2, tw);
            p.fd_h.divmod(t2, tn, th);
#pragma unroll 1
            for (int kb = kb0; kb < kb1; ++kb) {
                const int w0 = tw * TW, h0 = th * TH, n0 = tn * TN;
                {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * stage_bytes;
                    uint8_t* sB = sA + a_bytes;
                    if (ptx::elect_one()) {
                        if (debug & 2) {
                            ptx::mbar_arrive(&full_bar[stage]);
                        } else {
                            ptx::mbar_expect_tx(&full_bar[stage], stage_tx);
                            ptx::tma_load_4d(sA, mapA, &full_bar[stage], cA, w0, h0, n0);
                            if (a_boxes > 1) ptx::tma_load_4d(sA + kBoxBytes, mapA, &full_bar[stage], cA + 64, w0, h0, n0);
                            for (int i = 0; i < x_loads; ++i)
                                ptx::tma_load_4d(sB + i * xbox_stride, mapB, &full_bar[stage], 0, w0, h0 * stride_h + i, n0);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (++tw == tiles_w) {
                    tw = 0;
                    if (++th == tiles_h) {
                        th = 0;
                        ++tn;
                    }
                }
            }
            }
        }
    } else if (warp == 1) {
        const uint32_t smem0 = ptx::smem_u32(smem);
        // dY: MN-major, 64-channel sub-boxes kBoxBytes apart, 8-pixel groups 1024 B apart.  X: MN-major, column group g
        // = filter row j = g of this box, one image row (TW * 128 B) further in; 8-pixel groups 1024 B apart (the box
        // has no halo along W, so the 64 pixels of the k-block are contiguous rows of it).
        const uint64_t adesc0 = ptx::make_smem_desc(smem0, kBoxBytes, 1024);
        const uint64_t bdesc0 = ptx::make_smem_desc(smem0 + a_bytes, p.TW * 128, 1024);
        const uint32_t idesc0 = ptx::make_idesc_bf16(p.ntap[0] * 64, 1, 1);
        const uint32_t idesc1 = ptx::make_idesc_bf16(p.ntap[1] > 0 ? p.ntap[1] * 64 : 64, 1, 1);
        const uint32_t col1 = p.ntap[0] * 64;
        const uint32_t xb16 = static_cast<uint32_t>(xbox_stride) >> 4;
        const bool two = x_loads > 1;
        const bool do_mma = !(p.debug & 1);
        const uint32_t stage16 = static_cast<uint32_t>(stage_bytes) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int cot, split, kb0, kb1;
            decode(tile, cot, split, kb0, kb1);
            ptx::mbar_wait(tempty_bar, (it & 1) ^ 1);
            ptx::tc_fence_after();
            const int nk = (kb1 - kb0) * nterms;
#pragma unroll 1
            for (int kb = 0; kb < nk; ++kb) {
                ptx::mbar_wait(&full_bar[stage], phase);
                ptx::tc_fence_after();
                const uint64_t soff = static_cast<uint64_t>(stage * stage16);
                const uint64_t adesc = adesc0 + soff;
                const uint64_t bdesc = bdesc0 + soff;
                if (ptx::elect_one()) {
                    if (do_mma) {
#pragma unroll
                        for (int j = 0; j < kBK / 16; ++j) {        // 16 pixels = 2048 B in both operands
                            ptx::umma_f16(tmem_base, adesc + 128 * j, bdesc + 128 * j, idesc0, (kb | j) != 0);
                            if (two)
                                ptx::umma_f16(tmem_base + col1, adesc + 128 * j, bdesc + xb16 + 128 * j, idesc1,
                                              (kb | j) != 0);
                        }
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (ptx::elect_one()) ptx::umma_commit(tfull_bar);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int R = p.R, stride_h = p.stride_h;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            int cot, split, kb0, kb1;
            decode(tile, cot, split, kb0, kb1);
            const int co = cot * kBM + row;
            const bool row_ok = co < p.Cout;
            ptx::mbar_wait(tfull_bar, it & 1);
            ptx::tc_fence_after();
            for (int r = 0; r < R; ++r) {
                const uint32_t col = static_cast<uint32_t>((r % stride_h) * p.ntap[0] * 64 + (r / stride_h) * 64);
                float* dst_row = p.ws + ((static_cast<long long>(split) * p.Cout + co) * R + r) * static_cast<long long>(p.ldws);
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t rr[32];
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col + c0;
                    ptx::tmem_ld_32x32b_x32(taddr, rr);
                    ptx::tmem_ld_wait();
                    if (row_ok && !(p.debug & 4)) {
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rr[i]);
                        store_row_chunk(dst_row, 1, c0, v, 32);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// Sum the split-K partials and scatter into the reference filter layout (Cout, Cin, R, S) of the *true*
// convolution (tap (r,s) of the correlation is element (R-1-r, S-1-s) of the reference filter).
// accumulate != 0 adds into dw (used when a layer's weight receives gradient from two paths).
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int Cout, int Cin,
                                    int R, int S, int ldws, int accumulate) {
    const long long total = static_cast<long long>(Cout) * Cin * R * S;
    const int ntaps = R * S;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        // idx enumerates (co, tap, ci) so that workspace reads are coalesced
        const int ci = static_cast<int>(idx % Cin);
        long long t = idx / Cin;
        const int tap = static_cast<int>(t % ntaps);
        const int co = static_cast<int>(t / ntaps);
        float acc = 0.f;
        for (int s = 0; s < splits; ++s)
            acc += ws[((static_cast<long long>(s) * Cout + co) * ntaps + tap) * ldws + ci];
        const int r = tap / S, sx = tap % S;
        const long long o = ((static_cast<long long>(co) * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - sx);
        dw[o] = accumulate ? dw[o] + acc : acc;
    }
}

// Weight preparation: reference fp32 filters (Cout, Cin, R, S) [true convolution] -> GEMM B operand, bf16 hi (+lo).
//   mode 0 (fprop): B[co][tap=(r,s)][ci] = W[co][ci][R-1-r][S-1-s]     rows = Cout, K = taps * CinP
//   mode 1 (dgrad): B[ci][tap=(r,s)][co] = W[co][ci][r][s]             rows = Cin,  K = taps * CoutP
__global__ void weight_prep_kernel(const float* __restrict__ w, int Cout, int Cin, int R, int S, int mode,
                                   __nv_bfloat16* __restrict__ b_hi, __nv_bfloat16* __restrict__ b_lo) {
    const int rows = mode == 0 ? Cout : Cin;
    const int kin = mode == 0 ? Cin : Cout;
    const int kp = (kin + 63) / 64 * 64;
    const int ntaps = R * S;
    const long long total = static_cast<long long>(rows) * ntaps * kp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(idx % kp);
        long long t = idx / kp;
        const int tap = static_cast<int>(t % ntaps);
        const int row = static_cast<int>(t / ntaps);
        float v = 0.f;
        if (k < kin) {
            const int r = tap / S, s = tap % S;
            if (mode == 0)
                v = w[((static_cast<long long>(row) * Cin + k) * R + (R - 1 - r)) * S + (S - 1 - s)];
            else
                v = w[((static_cast<long long>(k) * Cin + row) * R + r) * S + s];
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        b_hi[idx] = hi;
        if (b_lo) b_lo[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// All conv layers' operands in ONE launch (86 small launches per step otherwise): block i prepares elements
// [block_offset[i], +kPrepChunk) of operand block_entry[i] of a device-resident table.
// mode 0 / 1 as weight_prep_kernel, mode 2 = row-folded stem operand [Cout][R][64] (weight_prep_rowfold_kernel).
struct PrepEntry {
    const float* w;
    __nv_bfloat16* hi;
    __nv_bfloat16* lo;
    long long total;      // work items = operand rows x padded K columns (rows x 64 for the row-folded operand)
    int Cout, Cin, R, S, mode, Cp;
};
constexpr int kPrepChunk = 4096;

__global__ void __launch_bounds__(256) weight_prep_multi_kernel(const PrepEntry* __restrict__ entries,
                                                                 const int* __restrict__ block_entry,
                                                                 const long long* __restrict__ block_offset) {
    // One work item = one (operand row, K column) pair; the item writes that column for every filter tap.  Reading the
    // R*S taps of a (co, ci) filter as one contiguous run keeps the fp32 reads coalesced (a per-element mapping reads
    // 4 bytes out of every 36), consecutive items write consecutive bf16 of each tap row; all index math is 32-bit.
    const PrepEntry e = entries[block_entry[blockIdx.x]];
    const unsigned items = (unsigned)e.total;
    const unsigned start = (unsigned)block_offset[blockIdx.x];
    const unsigned end = start + kPrepChunk < items ? start + kPrepChunk : items;
    const unsigned R = e.R, S = e.S, Cin = e.Cin, Cout = e.Cout;
    if (e.mode == 2) {
        const unsigned Cp = e.Cp;
        for (unsigned it = start + threadIdx.x; it < end; it += blockDim.x) {
            const unsigned k = it & 63u, co = it >> 6;
            const unsigned sx = k / Cp, c = k - sx * Cp;
            const bool live = sx < S && c < Cin;
            const float* src = e.w + ((size_t)co * Cin + c) * R * S + (S - 1 - sx);
            for (unsigned r = 0; r < R; ++r) {
                const float v = live ? src[(R - 1 - r) * S] : 0.f;
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                const size_t o = ((size_t)co * R + r) * 64 + k;
                e.hi[o] = hi;
                if (e.lo) e.lo[o] = __float2bfloat16_rn(v - __bfloat162float(hi));
            }
        }
        return;
    }
    if (e.mode == 3) {
        // parity class of a strided convolution's data gradient: B[ci][(tr, ts)][co] = W[co][ci][r0 + sh*tr][s0 + sw*ts]
        // Cp packs r0 | s0 << 4 | sh << 8 | sw << 12 | Rc << 16 | Sc << 20 (Rc x Sc taps of this class)
        const unsigned r0 = e.Cp & 15u, s0 = (e.Cp >> 4) & 15u, sh = (e.Cp >> 8) & 15u, sw = (e.Cp >> 12) & 15u;
        const unsigned Rc = (e.Cp >> 16) & 15u, Sc = (e.Cp >> 20) & 15u;
        const unsigned kp = (Cout + 63u) / 64u * 64u;
        for (unsigned it = start + threadIdx.x; it < end; it += blockDim.x) {
            const unsigned row = it / kp, k = it - row * kp;        // row = ci, k = co
            const float* run = e.w + ((size_t)k * Cin + row) * R * S;
            size_t o = (size_t)row * Rc * Sc * kp + k;
            for (unsigned tr = 0; tr < Rc; ++tr)
                for (unsigned ts = 0; ts < Sc; ++ts, o += kp) {
                    const float v = k < Cout ? run[(r0 + sh * tr) * S + (s0 + sw * ts)] : 0.f;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                    e.hi[o] = hi;
                    if (e.lo) e.lo[o] = __float2bfloat16_rn(v - __bfloat162float(hi));
                }
        }
        return;
    }
    const unsigned ntaps = R * S;
    const unsigned kin = e.mode == 0 ? Cin : Cout;
    const unsigned kp = (kin + 63u) / 64u * 64u;
    for (unsigned it = start + threadIdx.x; it < end; it += blockDim.x) {
        const unsigned row = it / kp, k = it - row * kp;
        const bool live = k < kin;
        // mode 0: B[co=row][tap][ci=k] = W[row][k][R-1-r][S-1-s] = run[ntaps-1-tap];  mode 1: B[ci=row][tap][co=k] = W[k][row][tap]
        const float* run = e.w + (e.mode == 0 ? ((size_t)row * Cin + k) : ((size_t)k * Cin + row)) * ntaps;
        size_t o = (size_t)row * ntaps * kp + k;
        for (unsigned t = 0; t < ntaps; ++t, o += kp) {
            const float v = live ? run[e.mode == 0 ? ntaps - 1 - t : t] : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            e.hi[o] = hi;
            if (e.lo) e.lo[o] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

// The split-K reductions of ALL filter gradients in ONE launch (one small launch per conv layer otherwise): block i
// reduces elements [block_offset[i], +kReduceChunk) of entry block_entry[i].  Same fixed summation order and the same
// output layouts as wgrad_reduce_kernel (mode 0) / wgrad_reduce_rowfold_kernel (mode 1).
struct ReduceEntry {
    const float* ws;
    float* dw;
    long long total;    // Cout * Cin * R * S
    int splits, Cout, Cin, R, S, ldws, mode, Cp, accumulate, pad_;
};
constexpr int kReduceChunk = 1024;      // elementwise path: elements per block
constexpr int kReduceItems = 1;         // tiled path: (co, channel group) items per block
constexpr int kReduceGroup = 128;       // tiled path: input channels per item

// 1x1 filters, vector form: workspace [split][co][ci] and gradient [co][ci] enumerate alike; a thread owns V consecutive
// ci (start, Cin and ldws are multiples of V) with 8 splits in flight, summed in split order.
template <typename VT, int V>
__device__ __forceinline__ void reduce_1x1_vec(const ReduceEntry& e, long long start, long long end) {
    const long long step = static_cast<long long>(e.Cout) * e.ldws;
    for (long long idx = start + V * threadIdx.x; idx < end; idx += V * blockDim.x) {
        const int ci = static_cast<int>(idx % e.Cin);
        const int co = static_cast<int>(idx / e.Cin);
        const float* src = e.ws + static_cast<long long>(co) * e.ldws + ci;
        VT accv;
        float* acc = reinterpret_cast<float*>(&accv);
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = 0.f;
        int sp = 0;
        for (; sp + 8 <= e.splits; sp += 8) {
            VT a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ldcs(reinterpret_cast<const VT*>(src + j * step));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* f = reinterpret_cast<const float*>(&a[j]);
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] += f[v];
            }
            src += 8 * step;
        }
        for (; sp < e.splits; ++sp, src += step) {
            const VT a = __ldcs(reinterpret_cast<const VT*>(src));
            const float* f = reinterpret_cast<const float*>(&a);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] += f[v];
        }
        VT* dst = reinterpret_cast<VT*>(e.dw + idx);
        if (e.accumulate) {
            const VT o = *dst;
            const float* f = reinterpret_cast<const float*>(&o);
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] += f[v];
        }
        *dst = accv;
    }
}

constexpr int kReduceThreads = 288;     // 9 warps: one per tap of a 3x3 filter (a second, one-warp round cost a full latency)
__global__ void __launch_bounds__(kReduceThreads) wgrad_reduce_multi_kernel(const ReduceEntry* __restrict__ entries,
                                                                  const int* __restrict__ block_entry,
                                                                  const long long* __restrict__ block_offset) {
    const ReduceEntry e = entries[block_entry[blockIdx.x]];
    const int R = e.R, S = e.S, Cin = e.Cin, Cout = e.Cout;
    const int ntaps = R * S;
    if (e.mode == 0 && ntaps > 1) {
        // Tiled path (multi-tap filters): the workspace is [split][co][tap][ci], the reference gradient [co][ci][tap'].
        // A block owns (co, kReduceGroup = 128 consecutive ci, all taps): warp = tap, lane = 4 consecutive ci read as
        // ONE 16-byte load per split (512 contiguous bytes per warp instruction, 8 splits in flight: the pass is pure
        // streaming, the bytes in flight per SM and the contiguous run per request set its speed - 128-byte rows with
        // 4 loads in flight ran at 1.9 TB/s), summed in split order, transposed through shared memory, and the
        // 128 * ntaps contiguous output floats are written coalesced.
        __shared__ float tile[kReduceGroup][65];
        const int groups = (Cin + kReduceGroup - 1) / kReduceGroup;
        const long long item = block_offset[blockIdx.x];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const long long step = static_cast<long long>(Cout) * ntaps * e.ldws;
        const int co = static_cast<int>(item / groups);
        const int ci0 = static_cast<int>(item % groups) * kReduceGroup;
        const int ci = ci0 + lane * 4;
        for (int t = warp; t < ntaps; t += kReduceThreads / 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ci < Cin) {         // ldws is a multiple of 4 and the pad columns of the workspace rows are never read back
                const float* src = e.ws + (static_cast<long long>(co) * ntaps + t) * e.ldws + ci;
                int sp = 0;
                for (; sp + 8 <= e.splits; sp += 8) {
                    float4 a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = __ldcs(reinterpret_cast<const float4*>(src + j * step));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                                      // fixed (split) order
                        acc.x += a[j].x; acc.y += a[j].y; acc.z += a[j].z; acc.w += a[j].w;
                    }
                    src += 8 * step;
                }
                for (; sp < e.splits; ++sp, src += step) {
                    const float4 a = __ldcs(reinterpret_cast<const float4*>(src));
                    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
                }
            }
            const int tt = ntaps - 1 - t;           // tap (r, s) of the correlation = element (R-1-r, S-1-s)
            tile[lane * 4 + 0][tt] = acc.x;
            tile[lane * 4 + 1][tt] = acc.y;
            tile[lane * 4 + 2][tt] = acc.z;
            tile[lane * 4 + 3][tt] = acc.w;
        }
        __syncthreads();
        const int nci = Cin - ci0 < kReduceGroup ? Cin - ci0 : kReduceGroup;
        float* dst = e.dw + (static_cast<long long>(co) * Cin + ci0) * ntaps;
        for (int i = threadIdx.x; i < nci * ntaps; i += blockDim.x) {
            const float v = tile[i / ntaps][i % ntaps];
            dst[i] = e.accumulate ? dst[i] + v : v;
        }
        return;
    }
    const long long start = block_offset[blockIdx.x];
    const long long end = start + kReduceChunk < e.total ? start + kReduceChunk : e.total;
    if (e.mode == 0 && (Cin & 1) == 0) {
        if ((Cin & 3) == 0)
            reduce_1x1_vec<float4, 4>(e, start, end);
        else
            reduce_1x1_vec<float2, 2>(e, start, end);
        return;
    }
    for (long long idx = start + threadIdx.x; idx < end; idx += blockDim.x) {
        const float* src;
        long long step, o;
        if (e.mode == 0) {
            // 1x1 filters: workspace [split][co][ci] and gradient [co][ci] enumerate alike
            const int ci = static_cast<int>(idx % Cin);
            const int co = static_cast<int>(idx / Cin);
            src = e.ws + static_cast<long long>(co) * e.ldws + ci;
            step = static_cast<long long>(Cout) * e.ldws;
            o = idx;
        } else {
            const int sf = static_cast<int>(idx % S);
            const int rf = static_cast<int>((idx / S) % R);
            const int c = static_cast<int>((idx / (static_cast<long long>(S) * R)) % Cin);
            const int co = static_cast<int>(idx / (static_cast<long long>(S) * R * Cin));
            src = e.ws + (static_cast<long long>(co) * R + (R - 1 - rf)) * e.ldws + (S - 1 - sf) * e.Cp + c;
            step = static_cast<long long>(Cout) * R * e.ldws;
            o = idx;
        }
        float acc = 0.f;
        int sp = 0;
        for (; sp + 8 <= e.splits; sp += 8) {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __ldcs(src + j * step);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += a[j];
            src += 8 * step;
        }
        for (; sp < e.splits; ++sp, src += step) acc += __ldcs(src);
        e.dw[o] = e.accumulate ? e.dw[o] + acc : acc;
    }
}

// fp32 -> bf16 hi/lo split of an activation tensor (parity mode operand preparation).
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float v = x[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ------------------------------------------------------------------------------------------------ host side
static void pick_patch(int W, int H, int N, int target, int& TW, int& TH, int& TN) {
    // powers of two with TW*TH*TN == target, covering W first, then H, then batch
    TW = 1;
    while (TW < W && TW < target) TW <<= 1;
    TH = 1;
    while (TH < H && TW * TH < target) TH <<= 1;
    TN = target / (TW * TH);
    (void)N;
}

template <int BN, int STAGES>
static int launch_fprop(const ConvFpropParams& p, cudaStream_t stream) {
    using L = SmemLayout<BN, STAGES>;
    auto kern = conv_fprop_kernel<BN, STAGES>;
    // + batch-norm constant table [4][cpad] fp32, or (staged epilogue) the 1024-aligned staging tile after the ring
    const int smem = p.epi_tma ? (int)p.smem_total : L::kTotal + (p.bnb_x ? 16 * p.bnb_cpad : 0);
    DN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    kern<<<DN_G(grid), kThreadsF, smem, stream>>>(p);
    DN_CHECK_LAUNCH();
    return 0;
}

template <int BN, int STAGES, int KS>
static int launch_wgrad(const ConvWgradParams& p, cudaStream_t stream) {
    using L = SmemLayout<BN, STAGES * KS>;
    auto kern = conv_wgrad_kernel<BN, STAGES, KS>;
    DN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    kern<<<DN_G(grid), kThreads, L::kTotal, stream>>>(p);
    DN_CHECK_LAUNCH();
    return 0;
}

template <int BN>
static int launch_wgrad_rows(const ConvWgradParams& p, cudaStream_t stream) {
    using L = RowsLayout<BN>;
    auto kern = conv_wgrad_rows_kernel<BN>;
    DN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    kern<<<DN_G(grid), kThreads, L::kTotal, stream>>>(p);
    DN_CHECK_LAUNCH();
    return 0;
}

// sw/sh: traversal (element) strides along W/H.  With a stride s the box must span TW*s input columns to deliver
// TW elements (cuTensorMapEncodeTiled loads ceil(box/stride) elements per dimension).
static int make_act_map(CUtensorMap* tm, const void* base, int C, int W, int H, int N, long long ld, int TW, int TH,
                        int TN, int sw = 1, int sh = 1) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * W, (uint64_t)ld * 2 * W * H};
    uint32_t box[4] = {64, (uint32_t)(TW * sw), (uint32_t)(TH * sh), (uint32_t)TN};
    uint32_t estr[4] = {1, (uint32_t)sw, (uint32_t)sh, 1};
    if (box[1] > 256 || box[2] > 256) return set_error(DENET_ERR_ARG, "conv: strided patch exceeds the TMA box limit");
    return encode_tmap_bf16(tm, base, 4, dims, strides, box, estr);
}


static uint32_t staged_epilogue_bytes(ConvFpropParams& p, int BN, int* rc);

// weight tensor map(s) + kernel dispatch, shared by the generic and the row-folded entry points.  p must hold the
// A map(s), the tiling and the epilogue fields; p.tiles_co / p.num_tiles are filled here.
int fprop_finish(ConvFpropParams& p, const void* b_hi, const void* b_lo, cudaStream_t stream) {
    const int Cout = p.Cout;
    const int BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    p.tiles_co = ceil_div(Cout, BN);
    p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_co;
    p.fd_co = make_fastdiv(p.tiles_co);
    p.fd_w = make_fastdiv(p.tiles_w);
    p.fd_h = make_fastdiv(p.tiles_h);
    p.fd_tw = make_fastdiv(p.TW);
    p.fd_th = make_fastdiv(p.TH);
    int rc;
    const uint64_t ktot = (uint64_t)p.R * p.S * p.kchunks * 64;
    uint64_t dims[2] = {ktot, (uint64_t)Cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    if ((rc = encode_tmap_bf16(&p.tmB[0], b_hi, 2, dims, strides, box, nullptr))) return rc;
    if (b_lo && (rc = encode_tmap_bf16(&p.tmB[1], b_lo, 2, dims, strides, box, nullptr))) return rc;
    // staged epilogue for the narrow tiles (the parity-class launches of the strided data gradients, the 1x1 / strided
    // layers with <= 128 output channels: K loops of 1-4 taps, where the register epilogue's ~2500 cycles per 128 x 64
    // tile are exposed); the 256-wide tiles keep four stages and the register epilogue (long K loops hide it)
    // 256-wide tiles with K <= 1536 (the head's 1x1 layers after the first): the register epilogue with statistics takes
    // ~10k cycles per 128 x 256 tile, more than the tile's <= 96 MMAs - staged epilogue on a three-stage ring
    p.epi_tma = 0;
    const int num_kb = p.nterms * p.R * p.S * p.kchunks;
    if (!p.bnb_x && (BN <= 128 || num_kb <= 24)) {
        const uint32_t stage_bytes = staged_epilogue_bytes(p, BN, &rc);
        if (rc) return rc;
        if (stage_bytes) {
            const uint32_t front = BN == 64    ? SmemLayout<64, 8>::kStatOffset + 8 * 64 * 8
                                   : BN == 128 ? SmemLayout<128, 5>::kStatOffset + 8 * 128 * 8
                                               : SmemLayout<256, 3>::kStatOffset + 8 * 256 * 8;
            p.stage_off = (front + 1023) / 1024 * 1024;
            p.smem_total = 1024 + p.stage_off + stage_bytes;
            return BN == 64 ? launch_fprop<64, 8>(p, stream)
                            : (BN == 128 ? launch_fprop<128, 5>(p, stream) : launch_fprop<256, 3>(p, stream));
        }
    }
    switch (BN) {
        case 64: return launch_fprop<64, 8>(p, stream);
        case 128: return launch_fprop<128, 6>(p, stream);
        default: return launch_fprop<256, 4>(p, stream);
    }
}

// ---- tap-group (halo) fprop: ring sizing, weight map, dispatch.  The caller has filled the A map(s), the patch
// geometry (TW/TH/TN, tiles_w/h/n) and the a_* / tap-offset fields.  Returns 1 when the layer does not fit (caller falls
// back to conv_fprop_kernel), 0 on success, < 0 on error.
static int g_fprop_mode = 15;     // bit0: tap-group kernel where eligible, bit1: resident filters, bit2: CTA pairs,
                                  // bit3: staged epilogue (TMA stores, statistics from the staged tile)
static int g_fprop_debug = 0;     // ConvFpropParams::debug
static long long* g_fprop_timeline = nullptr;

template <int BN, int NT, bool RESIDENT, int KM>
static int launch_fprop_halo(const ConvFpropParams& p, cudaStream_t stream) {
    const size_t smem = p.smem_total;
    auto kern = conv_fprop_halo_kernel<BN, NT, RESIDENT, KM>;
    DN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    kern<<<DN_G(grid), kThreadsF, smem, stream>>>(p);
    DN_CHECK_LAUNCH();
    return 0;
}

// specialised instances: 3x3 filters (9 taps, 4 MMAs per K block) and the 7-row stride-2 stem (K = 28 -> 2 MMAs);
// everything else runs the run-time-count instance
template <int BN>
static int dispatch_fprop_halo(const ConvFpropParams& p, cudaStream_t stream) {
    const int ntaps = p.R * p.S;
    constexpr int BR = BN <= 128 ? BN : 128;      // resident filters exist for BN <= 128 only
    // (2 and 4 taps: the parity classes of a 3x3 / stride-2 data gradient - 8-32 MMAs per tile, where the run-time-count
    // instance's ~100 cycles per issued MMA, 220 per (chunk, tap) iteration even without MMAs, are the whole tile time)
    if (p.b_resident) {
        if (BN == 64 && ntaps == 9 && p.kmmas == 4) return launch_fprop_halo<64, 9, true, 4>(p, stream);
        if (BN == 64 && ntaps == 7 && p.kmmas == 2) return launch_fprop_halo<64, 7, true, 2>(p, stream);
        if (BN == 64 && ntaps == 4 && p.kmmas == 4) return launch_fprop_halo<64, 4, true, 4>(p, stream);
        if (BN == 64 && ntaps == 2 && p.kmmas == 4) return launch_fprop_halo<64, 2, true, 4>(p, stream);
        return launch_fprop_halo<BR, 0, true, 0>(p, stream);
    }
    if (ntaps == 9 && p.kmmas == 4) return launch_fprop_halo<BN, 9, false, 4>(p, stream);
    if (BN <= 128 && ntaps == 4 && p.kmmas == 4) return launch_fprop_halo<BR, 4, false, 4>(p, stream);
    if (BN <= 128 && ntaps == 2 && p.kmmas == 4) return launch_fprop_halo<BR, 2, false, 4>(p, stream);
    if (BN == 64 && ntaps == 7 && p.kmmas == 2) return launch_fprop_halo<64, 7, false, 2>(p, stream);
    return launch_fprop_halo<BN, 0, false, 0>(p, stream);
}

// staged (TMA store) epilogue: applicable to bf16 outputs written in place (no scatter, no fused batch-norm backward).
// Fills tmY / epi_tma; returns the bytes of the staging tile (0 = not applicable).  The caller places it with
// place_stage() once the A / B regions are sized.
static uint32_t staged_epilogue_bytes(ConvFpropParams& p, int BN, int* rc) {
    *rc = 0;
    p.epi_tma = 0;
    // (the statistics are those of the conv output BEFORE a fused residual / ReLU: the staged tile holds the final
    // values, so that combination - which no layer of the model uses - keeps the register epilogue)
    const bool ok = (g_fprop_mode & 8) && !p.y_fp32 && (!p.bnb_x || p.Cout % 2 == 0) && p.TW <= 256 && p.TH <= 256 &&
                    p.TN <= 256 && !(p.stat_sum && (p.residual || p.relu));
    if (!ok) return 0;
    // The store map describes the pixels of THIS launch: all of them, or - for a parity class of a strided data gradient
    // - every osw-th / osh-th pixel from (ooh, oow) of the Hf x Wf tensor, as a strided view (plain byte strides, no
    // element strides), so that the scatter costs nothing.
    uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.No};
    uint64_t strides[3] = {(uint64_t)p.ldy * 2 * p.osw, (uint64_t)p.ldy * 2 * p.Wf * p.osh, (uint64_t)p.ldy * 2 * p.Wf * p.Hf};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, (uint32_t)p.TN};
    const uint8_t* ybase = reinterpret_cast<const uint8_t*>(p.y) + ((size_t)p.ooh * p.Wf + p.oow) * p.ldy * 2;
    if ((*rc = encode_tmap_bf16(&p.tmY, ybase, 4, dims, strides, box, nullptr))) return 0;
    p.epi_tma = 1;
    return (uint32_t)(128 * BN * 2);
}

static size_t place_stage(ConvFpropParams& p, int BN, uint32_t stage_bytes) {
    const size_t front = (size_t)p.a_region_bytes + p.b_region_bytes + 512 + 8 * (size_t)BN * 8;
    if (!p.epi_tma) return 1024 + front + (p.bnb_x ? 16 * p.bnb_cpad : 0);
    // staged epilogue: [.. | 512 barriers | stat_red 4 KB + batch-norm constant table 16 * cpad | -> 1024 | staging tile]
    const size_t front2 = (size_t)p.a_region_bytes + p.b_region_bytes + 512 + 4096 + (p.bnb_x ? 16 * p.bnb_cpad : 0);
    p.stage_off = (uint32_t)(((front2 > front ? front2 : front) + 1023) / 1024 * 1024);
    return 1024 + (size_t)p.stage_off + stage_bytes;
}

template <int BN, int NT, bool RESIDENT, int KM>
static int launch_fprop_halo2(const ConvFpropParams& p, cudaStream_t stream) {
    const size_t smem = p.smem_total;
    auto kern = conv_fprop_halo2_kernel<BN, NT, RESIDENT, KM>;
    DN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int pairs = p.num_tiles < num_sms() / 2 ? p.num_tiles : num_sms() / 2;
    DN_CHECK_CUDA(launch_pdl(kern, dim3(DN_G(2 * pairs)), dim3(kThreadsF), smem, stream, p));
    DN_CHECK_LAUNCH();
    return 0;
}

// CTA-pair variant of halo_finish (3x3 filters only).  Returns 1 when not applicable.
static int halo2_finish(ConvFpropParams& p, const void* b_hi, const void* b_lo, cudaStream_t stream) {
    const int Cout = p.Cout;
    const int BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    const int ntaps = p.R * p.S;
    if ((ntaps != 9 && ntaps != 4 && ntaps != 2) || p.kmmas != 4 || p.a_loads != 1) return 1;
    if (ntaps != 9 && BN == 256) return 1;        // (2 / 4 taps = parity classes of a strided data gradient: <= 128 wide)
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    if (m_tiles < 2) return 1;
    const uint32_t bh_bytes = (BN / 2) * 128;
    int rc;
    const uint32_t stage_bytes = staged_epilogue_bytes(p, BN, &rc);
    if (rc) return rc;
    const uint32_t budget = 200 * 1024 - stage_bytes;
    p.a_stage_bytes = p.a_loads * p.a_load_stride;
    p.b_resident = (g_fprop_mode & 2) && p.nterms == 1 && BN == 64 && Cout <= BN &&
                   (uint32_t)(ntaps * p.kchunks) * bh_bytes <= 80 * 1024;
    if (p.b_resident) {
        p.b_region_bytes = ntaps * p.kchunks * bh_bytes;
        p.b_stages = 1;
        p.a_stages = (int)((budget - p.b_region_bytes) / p.a_stage_bytes);
        if (p.a_stages > 6) p.a_stages = 6;
    } else {
        // (stages of TG taps each: one filter row per stage for the 9-tap, <= 128-wide instances - see the kernel)
        const int TG = (ntaps == 9 && BN <= 128) ? 3 : 1;
        p.b_stages = BN == 256 ? (stage_bytes ? 4 : 8) : 12 / TG;
        p.b_region_bytes = p.b_stages * TG * bh_bytes;
        p.a_stages = (int)((budget - p.b_region_bytes) / p.a_stage_bytes);
        if (p.a_stages > 4) p.a_stages = 4;
    }
    if (p.a_stages < 2) return 1;
    p.a_region_bytes = p.a_stages * p.a_stage_bytes;
    p.smem_total = (uint32_t)place_stage(p, BN, stage_bytes);
    p.tiles_co = ceil_div(Cout, BN);
    p.num_tiles = ceil_div(m_tiles, 2) * p.tiles_co;          // PAIR tiles
    p.fd_co = make_fastdiv(p.tiles_co);
    p.fd_w = make_fastdiv(p.tiles_w);
    p.fd_h = make_fastdiv(p.tiles_h);
    p.fd_tw = make_fastdiv(p.TW);
    p.fd_th = make_fastdiv(p.TH);
    const uint64_t ktot = (uint64_t)ntaps * p.kchunks * 64;
    uint64_t dims[2] = {ktot, (uint64_t)Cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64, (uint32_t)(BN / 2)};
    if ((rc = encode_tmap_bf16(&p.tmB[0], b_hi, 2, dims, strides, box, nullptr))) return rc;
    if (b_lo && (rc = encode_tmap_bf16(&p.tmB[1], b_lo, 2, dims, strides, box, nullptr))) return rc;
    if (ntaps == 4) {
        if (BN == 64) return p.b_resident ? launch_fprop_halo2<64, 4, true, 4>(p, stream)
                                          : launch_fprop_halo2<64, 4, false, 4>(p, stream);
        return launch_fprop_halo2<128, 4, false, 4>(p, stream);
    }
    if (ntaps == 2) {
        if (BN == 64) return p.b_resident ? launch_fprop_halo2<64, 2, true, 4>(p, stream)
                                          : launch_fprop_halo2<64, 2, false, 4>(p, stream);
        return launch_fprop_halo2<128, 2, false, 4>(p, stream);
    }
    switch (BN) {
        case 64:
            return p.b_resident ? launch_fprop_halo2<64, 9, true, 4>(p, stream)
                                : launch_fprop_halo2<64, 9, false, 4>(p, stream);
        case 128: return launch_fprop_halo2<128, 9, false, 4>(p, stream);
        default: return launch_fprop_halo2<256, 9, false, 4>(p, stream);
    }
}

int halo_finish(ConvFpropParams& p, const void* b_hi, const void* b_lo, cudaStream_t stream) {
    if (g_fprop_mode & 4) {
        ConvFpropParams q = p;
        const int rc2 = halo2_finish(q, b_hi, b_lo, stream);
        if (rc2 <= 0) return rc2;
    }
    const int Cout = p.Cout;
    const int BN = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    const uint32_t b_bytes = BN * 128;
    int rc;
    uint32_t stage_bytes = 0;
    p.epi_tma = 0;
    if (BN <= 128) {                    // (a 256-wide staging tile does not fit next to four 32 KB filter stages)
        stage_bytes = staged_epilogue_bytes(p, BN, &rc);
        if (rc) return rc;
    }
    const uint32_t budget = 200 * 1024 - stage_bytes;
    const int ntaps = p.R * p.S;
    p.a_stage_bytes = p.a_loads * p.a_load_stride;
    p.b_resident = (g_fprop_mode & 2) && p.nterms == 1 && BN <= 128 && Cout <= BN &&
                   (uint32_t)(ntaps * p.kchunks) * b_bytes <= 80 * 1024;
    if (p.b_resident) {
        p.b_region_bytes = ntaps * p.kchunks * b_bytes;
        p.b_stages = 1;
        p.a_stages = (int)((budget - p.b_region_bytes) / p.a_stage_bytes);
        if (p.a_stages > 6) p.a_stages = 6;
    } else {
        p.b_stages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
        p.b_region_bytes = p.b_stages * b_bytes;
        p.a_stages = (int)((budget - p.b_region_bytes) / p.a_stage_bytes);
        if (p.a_stages > 4) p.a_stages = 4;
    }
    if (p.a_stages < 2) return 1;
    p.a_region_bytes = p.a_stages * p.a_stage_bytes;
    p.smem_total = (uint32_t)place_stage(p, BN, stage_bytes);
    p.tiles_co = ceil_div(Cout, BN);
    p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_co;
    p.fd_co = make_fastdiv(p.tiles_co);
    p.fd_w = make_fastdiv(p.tiles_w);
    p.fd_h = make_fastdiv(p.tiles_h);
    p.fd_tw = make_fastdiv(p.TW);
    p.fd_th = make_fastdiv(p.TH);
    const uint64_t ktot = (uint64_t)ntaps * p.kchunks * 64;
    uint64_t dims[2] = {ktot, (uint64_t)Cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64, (uint32_t)BN};
    if ((rc = encode_tmap_bf16(&p.tmB[0], b_hi, 2, dims, strides, box, nullptr))) return rc;
    if (b_lo && (rc = encode_tmap_bf16(&p.tmB[1], b_lo, 2, dims, strides, box, nullptr))) return rc;
    switch (BN) {
        case 64: return dispatch_fprop_halo<64>(p, stream);
        case 128: return dispatch_fprop_halo<128>(p, stream);
        default: return dispatch_fprop_halo<256>(p, stream);
    }
}

// Split-K factor of the filter gradient.  A persistent grid of num_sms() CTAs runs the base_tiles * splits tiles in
// ceil(tiles / SMs) rounds of ceil(total_kblocks / splits) k-blocks each, so a tile count just above a multiple of the
// SM count wastes most of a round (the former rule ceil(2*SMs / base_tiles) produced 297..306 tiles on 148 SMs for the
// ResNet layers: 3 rounds for 2.0x waves).  Pick the factor that minimises
//     rounds * kblocks_per_split * t_kblock  +  splits * t_reduce
// where t_kblock is the larger of the MMA time and the L2 -> shared-memory time of one k-block on one SM and t_reduce
// the workspace write + read of one split (the fixed-order reduction pass).  Ties go to fewer splits.
static int g_wgrad_legacy_splits = 0;
static int g_wgrad_rows_max_cin = 64;     // widest layer (input channels) that takes the row-shared wgrad kernel
int wgrad_splits(int total_kblocks, int base_tiles, int bn_cols, int stage_bytes, long long out_elems) {
    const int sms = num_sms();
    if (g_wgrad_legacy_splits) {
        int splits = ceil_div(2 * sms, base_tiles);
        if (splits > total_kblocks) splits = total_kblocks;
        return splits < 1 ? 1 : splits;
    }
    const double t_mma = 2.0 * 128 * bn_cols * 64 / 10.0e6;          // us: ~10 TFLOP/s of bf16 MMA per SM
    const double t_l2 = stage_bytes / 84.0e3;                         // us: ~12.4 TB/s chip-wide over 148 SMs
    const double t_kb = t_mma > t_l2 ? t_mma : t_l2;
    const double t_red = 8.0 * out_elems / 4.0e6;                     // us per split: fp32 write + read at ~4 TB/s
    int max_splits = ceil_div(8 * sms, base_tiles);
    if (max_splits > total_kblocks) max_splits = total_kblocks;
    if (max_splits < 1) max_splits = 1;
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= max_splits; ++s) {
        const int rounds = ceil_div(base_tiles * s, sms);
        const double cost = rounds * (ceil_div(total_kblocks, s) * t_kb + 0.3) + s * t_red;
        if (cost < best_cost * 0.999) {
            best_cost = cost;
            best = s;
        }
    }
    return best;
}

static int g_wgrad_pairs = 1;        // CTA-pair kernel for the 256-wide tiles (A/B: denet_conv2d_wgrad_set_mode bit 7 = off)

// dispatch of the wgrad GEMM; p holds the maps, the k-block tiling, Cout/Cin (GEMM N extent), R, S and the workspace
int wgrad_launch(ConvWgradParams& p, cudaStream_t stream) {
    const int BN = p.Cin <= 64 ? 64 : (p.Cin <= 128 ? 128 : 256);
    if (BN == 256 && g_wgrad_pairs && p.co_tiles % 2 == 0 && p.a_boxes == 2) {
        constexpr int kStages = 6;                              // 6 x (16 KB dY + 16 KB X half)
        ConvWgradParams q = p;
        q.fd_cot = make_fastdiv(p.co_tiles / 2);
        q.num_tiles = p.num_tiles / 2;                          // PAIR tiles
        const int smem = kStages * 32768 + 256 + 1024;
        auto kern = conv_wgrad2_kernel<kStages>;
        DN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int pairs = q.num_tiles < num_sms() / 2 ? q.num_tiles : num_sms() / 2;
        kern<<<DN_G(2 * pairs), kThreads, smem, stream>>>(q);
        DN_CHECK_LAUNCH();
        return 0;
    }
    switch (BN) {
        // (KS = 2 measured: the 128-wide instance issues faster - 43 -> 37 us with the loads switched off - but three
        // 64 KB stages expose the L2 -> shared-memory fill, 590 MB per launch, which is the floor: 43 -> 45 us)
        case 64: return launch_wgrad<64, 8, 1>(p, stream);
        case 128: return launch_wgrad<128, 6, 1>(p, stream);
        default: return launch_wgrad<256, 4, 1>(p, stream);
    }
}

}  // namespace dn

using namespace dn;

extern "C" int denet_conv_weight_prep(const float* w, int Cout, int Cin, int R, int S, int mode, void* b_hi, void* b_lo,
                                      cudaStream_t stream) {
    DN_REQUIRE(w && b_hi, "weight_prep: null pointer");
    DN_REQUIRE(mode == 0 || mode == 1, "weight_prep: mode must be 0 (fprop) or 1 (dgrad)");
    const int rows = mode == 0 ? Cout : Cin;
    const int kin = mode == 0 ? Cin : Cout;
    const long long total = static_cast<long long>(rows) * R * S * ((kin + 63) / 64 * 64);
    const int block = 256;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, block), 148LL * 16);
    weight_prep_kernel<<<DN_G(grid), block, 0, stream>>>(w, Cout, Cin, R, S, mode, (__nv_bfloat16*)b_hi, (__nv_bfloat16*)b_lo);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_weight_prep_entry_bytes(void) { return (int)sizeof(PrepEntry); }
extern "C" int denet_weight_prep_chunk(void) { return kPrepChunk; }

extern "C" int denet_conv_weight_prep_multi(const void* entries, const int* block_entry, const long long* block_offset,
                                            int nblocks, cudaStream_t stream) {
    DN_REQUIRE(entries && block_entry && block_offset, "weight_prep_multi: null pointer");
    if (nblocks == 0) return 0;
    weight_prep_multi_kernel<<<DN_G(nblocks), 256, 0, stream>>>((const PrepEntry*)entries, block_entry, block_offset);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_wgrad_reduce_entry_bytes(void) { return (int)sizeof(ReduceEntry); }
extern "C" int denet_wgrad_reduce_chunk(void) { return kReduceChunk; }
extern "C" int denet_wgrad_reduce_items(void) { return kReduceItems; }
extern "C" int denet_wgrad_reduce_group(void) { return kReduceGroup; }

extern "C" int denet_wgrad_reduce_multi(const void* entries, const int* block_entry, const long long* block_offset,
                                        int nblocks, cudaStream_t stream) {
    DN_REQUIRE(entries && block_entry && block_offset, "wgrad_reduce_multi: null pointer");
    if (nblocks == 0) return 0;
    wgrad_reduce_multi_kernel<<<DN_G(nblocks), kReduceThreads, 0, stream>>>((const ReduceEntry*)entries, block_entry, block_offset);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_split_bf16(const float* x, void* hi, void* lo, long long n, cudaStream_t stream) {
    DN_REQUIRE(x && hi, "split_bf16: null pointer");
    if (n == 0) return 0;
    const int block = 256;
    const int grid = (int)std::min<long long>(ceil_div_ll(n, block), 148LL * 32);
    split_bf16_kernel<<<DN_G(grid), block, 0, stream>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
    DN_CHECK_LAUNCH();
    return 0;
}

namespace dn {
struct OutScatter {
    int Hf, Wf, osh, osw, ooh, oow;
};
struct BnBwdArgs {
    const void* x;
    const void* yout;
    const float* mean;
    const float* invstd;
    const float* gamma;
    const float* beta;
    int relu;
};
}  // namespace dn

static int conv2d_fprop_impl(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin, long long ldx,
                             const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w,
                             int stride_h, int stride_w, void* y, int y_dtype, long long ldy, int Ho, int Wo,
                             const float* bias, const void* residual, int relu, float* stat_sum, float* stat_sqsum,
                             const BnBwdArgs* bnb, const OutScatter* sc, cudaStream_t stream);

extern "C" int denet_conv2d_fprop(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin, long long ldx,
                                  const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w,
                                  int stride_h, int stride_w, void* y, int y_dtype, long long ldy, int Ho, int Wo,
                                  const float* bias,
                                  const void* residual, int relu, float* stat_sum, float* stat_sqsum,
                                  cudaStream_t stream) {
    return conv2d_fprop_impl(x_hi, x_lo, N, Hi, Wi, Cin, ldx, b_hi, b_lo, Cout, R, S, pad_h, pad_w, stride_h, stride_w, y,
                             y_dtype, ldy, Ho, Wo, bias, residual, relu, stat_sum, stat_sqsum, nullptr, nullptr, stream);
}

extern "C" int denet_conv2d_fprop_scatter(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin,
                                          long long ldx, const void* b_hi, const void* b_lo, int Cout, int R, int S,
                                          int pad_h, int pad_w, void* y, int y_dtype, long long ldy, int Ho, int Wo,
                                          int Hf, int Wf, int osh, int osw, int ooh, int oow, const void* residual,
                                          cudaStream_t stream) {
    DN_REQUIRE(osh >= 1 && osw >= 1 && ooh >= 0 && oow >= 0 && (Ho - 1) * osh + ooh < Hf && (Wo - 1) * osw + oow < Wf,
               "conv2d_fprop_scatter: the scattered output does not fit a %d x %d image", Hf, Wf);
    OutScatter sc;
    sc.Hf = Hf; sc.Wf = Wf; sc.osh = osh; sc.osw = osw; sc.ooh = ooh; sc.oow = oow;
    return conv2d_fprop_impl(x_hi, x_lo, N, Hi, Wi, Cin, ldx, b_hi, b_lo, Cout, R, S, pad_h, pad_w, 1, 1, y, y_dtype, ldy,
                             Ho, Wo, nullptr, residual, 0, nullptr, nullptr, nullptr, &sc, stream);
}

extern "C" int denet_conv2d_dgrad_bnbwd(const void* dy_hi, const void* dy_lo, int N, int Hi, int Wi, int Cin,
                                        long long lddy, const void* b_hi, const void* b_lo, int Cout, int R, int S,
                                        int pad_h, int pad_w, void* dz, int dz_dtype, long long lddz, int Ho, int Wo,
                                        const void* residual, const void* bn_x, const void* bn_yout,
                                        const float* bn_mean, const float* bn_invstd, const float* bn_gamma,
                                        const float* bn_beta, int bn_relu, float* sum_dz, float* sum_dz_xhat,
                                        cudaStream_t stream) {
    DN_REQUIRE(bn_x && bn_mean && bn_invstd && bn_gamma && sum_dz && sum_dz_xhat, "conv2d_dgrad_bnbwd: null pointer");
    DN_REQUIRE(!bn_relu || bn_yout || bn_beta, "conv2d_dgrad_bnbwd: the relu mask needs the forward output or beta");
    DN_REQUIRE(Cout <= 512, "conv2d_dgrad_bnbwd: at most 512 channels (shared-memory constant table), got %d", Cout);
    BnBwdArgs a;
    a.x = bn_x; a.yout = bn_yout; a.mean = bn_mean; a.invstd = bn_invstd; a.gamma = bn_gamma; a.beta = bn_beta;
    a.relu = bn_relu;
    return conv2d_fprop_impl(dy_hi, dy_lo, N, Hi, Wi, Cin, lddy, b_hi, b_lo, Cout, R, S, pad_h, pad_w, 1, 1, dz, dz_dtype,
                             lddz, Ho, Wo, nullptr, residual, 0, sum_dz, sum_dz_xhat, &a, nullptr, stream);
}

static int conv2d_fprop_impl(const void* x_hi, const void* x_lo, int N, int Hi, int Wi, int Cin, long long ldx,
                             const void* b_hi, const void* b_lo, int Cout, int R, int S, int pad_h, int pad_w,
                             int stride_h, int stride_w, void* y, int y_dtype, long long ldy, int Ho, int Wo,
                             const float* bias, const void* residual, int relu, float* stat_sum, float* stat_sqsum,
                             const BnBwdArgs* bnb, const OutScatter* sc, cudaStream_t stream) {
    DN_REQUIRE(x_hi && b_hi && y, "conv2d_fprop: null pointer");
    DN_REQUIRE((x_lo == nullptr) == (b_lo == nullptr), "conv2d_fprop: x_lo and b_lo must both be given or both null");
    DN_REQUIRE(ldx % 8 == 0, "conv2d_fprop: input pixel pitch must be a multiple of 8 elements (16 B), got %lld", ldx);
    DN_REQUIRE(y_dtype == DENET_F32 || y_dtype == DENET_BF16, "conv2d_fprop: bad y_dtype %d", y_dtype);
    DN_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && Cin > 0 && Cout > 0 && Ho > 0 && Wo > 0, "conv2d_fprop: empty tensor");
    DN_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), "conv2d_fprop: stat pointers must come in pairs");
    DN_REQUIRE(stride_h >= 1 && stride_h <= 8 && stride_w >= 1 && stride_w <= 8, "conv2d_fprop: stride must be in [1,8]");

    ConvFpropParams p;
    memset(&p, 0, sizeof(p));
    p.nterms = x_lo ? 3 : 1;
    p.debug = g_fprop_debug;
    p.timeline = g_fprop_timeline;
    p.R = R; p.S = S; p.pad_h = pad_h; p.pad_w = pad_w;
    p.stride_h = stride_h; p.stride_w = stride_w;
    p.kchunks = ceil_div(Cin, 64);
    p.Wo = Wo; p.Ho = Ho; p.No = N;
    pick_patch(Wo, Ho, N, 128, p.TW, p.TH, p.TN);
    p.tiles_w = ceil_div(Wo, p.TW);
    p.tiles_h = ceil_div(Ho, p.TH);
    p.tiles_n = ceil_div(N, p.TN);
    p.Cout = Cout;
    p.ldy = ldy;
    p.y_fp32 = (y_dtype == DENET_F32);
    p.relu = relu;
    p.y = y;
    p.bias = bias;
    p.residual = residual;
    p.stat_sum = stat_sum;
    p.stat_sqsum = stat_sqsum;
    p.osh = p.osw = 1; p.ooh = p.oow = 0; p.Hf = Ho; p.Wf = Wo;
    if (sc) {
        p.osh = sc->osh; p.osw = sc->osw; p.ooh = sc->ooh; p.oow = sc->oow; p.Hf = sc->Hf; p.Wf = sc->Wf;
    }
    if (bnb) {
        p.bnb_x = bnb->x; p.bnb_yout = bnb->yout; p.bnb_mean = bnb->mean; p.bnb_invstd = bnb->invstd;
        p.bnb_gamma = bnb->gamma; p.bnb_beta = bnb->beta; p.bnb_relu = bnb->relu;
        p.bnb_cpad = (Cout + 31) / 32 * 32;
    }

    int rc;
    // stride-1 multi-tap filters: one halo'd A box per channel chunk serves all taps (conv_fprop_halo_kernel).  The
    // patch is 8 x 16 pixels of ONE image so that an 8-pixel group of the M tile is one row of the box.
    if ((g_fprop_mode & 1) && stride_h == 1 && stride_w == 1 && R * S > 1 && R <= 8 && S <= 8 && Ho >= 12 && Wo >= 8) {
        ConvFpropParams h = p;
        h.TW = 8; h.TH = 16; h.TN = 1;
        h.tiles_w = ceil_div(Wo, 8);
        h.tiles_h = ceil_div(Ho, 16);
        h.tiles_n = N;
        const int bw = 8 + S - 1, bh = 16 + R - 1;
        h.a_loads = 1;
        h.a_dw = -pad_w; h.a_dh = -pad_h; h.a_dh_step = 0;
        h.a_load_bytes = (uint32_t)bw * bh * 128;
        h.a_load_stride = (h.a_load_bytes + 1023) / 1024 * 1024;
        h.a_sbo = (uint32_t)bw * 128;
        for (int r = 0; r < R; ++r)
            for (int sx = 0; sx < S; ++sx) h.tap_off16[r * S + sx] = (uint32_t)(r * bw + sx) * 8;   // 128 B per box row
        h.kmmas = 4;
        if ((rc = make_act_map(&h.tmA[0], x_hi, Cin, Wi, Hi, N, ldx, bw, bh, 1))) return rc;
        if (x_lo && (rc = make_act_map(&h.tmA[1], x_lo, Cin, Wi, Hi, N, ldx, bw, bh, 1))) return rc;
        rc = halo_finish(h, b_hi, b_lo, stream);
        if (rc <= 0) return rc;
    }
    if ((rc = make_act_map(&p.tmA[0], x_hi, Cin, Wi, Hi, N, ldx, p.TW, p.TH, p.TN, stride_w, stride_h))) return rc;
    if (x_lo && (rc = make_act_map(&p.tmA[1], x_lo, Cin, Wi, Hi, N, ldx, p.TW, p.TH, p.TN, stride_w, stride_h)))
        return rc;
    return fprop_finish(p, b_hi, b_lo, stream);
}

// profiling: device buffer of 3 * 64 * 4 int64 receiving CTA 0's per-tile clock stamps (NULL = off)
extern "C" int denet_conv2d_fprop_set_timeline(void* buf) {
    g_fprop_timeline = (long long*)buf;
    return 0;
}

extern "C" int denet_conv2d_fprop_set_mode(int mode) {
    g_fprop_mode = mode & 15;
    g_fprop_debug = mode >> 4;        // undocumented profiling knobs, see ConvFpropParams::debug
    return 0;
}

namespace dn {

// Tiling plan of the filter gradient.  rows_path: the row-shared kernel (stride 1, S > 1, patch >= 8 pixels wide,
// halo box within the shared-memory budget); otherwise one tap per tile.
struct WgradPlan {
    int TW, TH, TN, total_kb, BN, co_tiles, ci_tiles, base_tiles, splits, rows_path, xrows;
};

static WgradPlan plan_wgrad(int N, int Ho, int Wo, int Cout, int Cin, int R, int S, int stride_h, int stride_w,
                            int want_rows) {
    WgradPlan pl;
    pick_patch(Wo, Ho, N, 64, pl.TW, pl.TH, pl.TN);
    pl.total_kb = ceil_div(Wo, pl.TW) * ceil_div(Ho, pl.TH) * ceil_div(N, pl.TN);
    pl.BN = Cin <= 64 ? 64 : (Cin <= 128 ? 128 : 256);
    pl.xrows = (pl.TW + S - 1) * pl.TH * pl.TN;
    // measured on B200 (scripts/bench_wgrad.py): the row-shared kernel wins for Cin <= 64 (266 vs 181 TFLOP/s at
    // 64->64 3x3 128^2); from 128 channels on, its smaller Cin tile and single accumulator set lose to one tap per tile
    pl.rows_path = want_rows && stride_h == 1 && stride_w == 1 && S > 1 && S <= 8 && pl.TW >= 8 && pl.xrows <= 80 &&
                   Cin <= g_wgrad_rows_max_cin;
    if (pl.rows_path)
        while (S * pl.BN > 512) pl.BN /= 2;
    pl.co_tiles = ceil_div(Cout, 128);
    pl.ci_tiles = ceil_div(Cin, pl.BN);
    pl.base_tiles = pl.co_tiles * pl.ci_tiles * R * (pl.rows_path ? 1 : S);
    {
        const int a_boxes = Cout <= 64 ? 1 : 2;
        const int stage_bytes = pl.rows_path ? a_boxes * 8192 + (pl.BN / 64) * pl.xrows * 128
                                             : a_boxes * 8192 + pl.BN * 128;
        pl.splits = wgrad_splits(pl.total_kb, pl.base_tiles, pl.rows_path ? S * pl.BN : pl.BN, stage_bytes,
                                 (long long)Cout * Cin * R * S);
    }
    return pl;
}

static int g_wgrad_rows = 1;         // 0: always one tap per tile (A/B switch for tests / profiling)

}  // namespace dn

static int g_wgrad_debug = 0;
extern "C" int denet_conv2d_wgrad_set_mode(int row_shared) {
    g_wgrad_rows = row_shared & 1;
    g_wgrad_debug = ((row_shared >> 1) & 7) | (((row_shared >> 5) & 1) << 3);   // undocumented profiling knobs (bit1: no
                                                  // MMA, bit2: no TMA, bit3: no store; bit5: per-tap MMAs in the rows kernel)
    g_wgrad_legacy_splits = (row_shared >> 4) & 1;   // bit4: the former split-K rule (A/B measurements)
    g_wgrad_rows_max_cin = ((row_shared >> 6) & 1) ? 4096 : 64;   // bit6: row-shared kernel for every channel count
    g_wgrad_pairs = ((row_shared >> 7) & 1) ? 0 : 1;              // bit7: no CTA pairs for the 256-wide tiles
    return 0;
}

extern "C" size_t denet_conv2d_wgrad_workspace(int N, int Ho, int Wo, int Cout, int Cin, int R, int S) {
    // the stride is not known here: size for whichever plan needs more
    const WgradPlan a = plan_wgrad(N, Ho, Wo, Cout, Cin, R, S, 1, 1, 1);
    const WgradPlan b = plan_wgrad(N, Ho, Wo, Cout, Cin, R, S, 2, 2, 0);
    const int splits = a.splits > b.splits ? a.splits : b.splits;
    const int ldws = (Cin + 3) / 4 * 4;
    return (size_t)splits * Cout * R * S * ldws * sizeof(float);
}

extern "C" int denet_conv2d_wgrad_splits(int N, int Ho, int Wo, int Cout, int Cin, int R, int S, int stride_h,
                                         int stride_w) {
    return plan_wgrad(N, Ho, Wo, Cout, Cin, R, S, stride_h, stride_w, g_wgrad_rows).splits;
}

extern "C" int denet_conv2d_wgrad(const void* dy_hi, const void* dy_lo, int N, int Ho, int Wo, int Cout, long long lddy,
                                  const void* x_hi, const void* x_lo, int Hi, int Wi, int Cin, long long ldx, int R,
                                  int S, int pad_h, int pad_w, int stride_h, int stride_w, float* dw, int accumulate,
                                  float* workspace, size_t workspace_bytes, cudaStream_t stream) {
    DN_REQUIRE(dy_hi && x_hi && workspace, "conv2d_wgrad: null pointer");
    DN_REQUIRE((dy_lo == nullptr) == (x_lo == nullptr), "conv2d_wgrad: dy_lo and x_lo must both be given or both null");
    DN_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0, "conv2d_wgrad: pixel pitches must be multiples of 8 elements");
    DN_REQUIRE(stride_h >= 1 && stride_h <= 8 && stride_w >= 1 && stride_w <= 8, "conv2d_wgrad: stride must be in [1,8]");
    const WgradPlan pl = plan_wgrad(N, Ho, Wo, Cout, Cin, R, S, stride_h, stride_w, g_wgrad_rows);
    ConvWgradParams p;
    memset(&p, 0, sizeof(p));
    p.nterms = dy_lo ? 3 : 1;
    p.R = R; p.S = S; p.pad_h = pad_h; p.pad_w = pad_w;
    p.stride_h = stride_h; p.stride_w = stride_w;
    p.TW = pl.TW; p.TH = pl.TH; p.TN = pl.TN;
    p.tiles_w = ceil_div(Wo, p.TW);
    p.tiles_h = ceil_div(Ho, p.TH);
    p.tiles_n = ceil_div(N, p.TN);
    p.total_kblocks = pl.total_kb;
    p.co_tiles = pl.co_tiles;
    p.ci_tiles = pl.ci_tiles;
    p.splits = pl.splits;
    p.num_tiles = pl.base_tiles * pl.splits;
    p.Cout = Cout; p.Cin = Cin;
    p.ldws = (Cin + 3) / 4 * 4;
    p.ws = workspace;
    p.fd_cot = make_fastdiv(p.co_tiles);
    p.fd_cit = make_fastdiv(p.ci_tiles);
    p.fd_taps = make_fastdiv(pl.rows_path ? R : R * S);
    p.fd_w = make_fastdiv(p.tiles_w);
    p.fd_h = make_fastdiv(p.tiles_h);
    for (p.log2_tw = 0; (1 << p.log2_tw) < p.TW; ++p.log2_tw) {
    }
    p.xrows = pl.xrows;
    p.xbox_bytes = (pl.xrows * 128 + 1023) / 1024 * 1024;
    p.a_boxes = Cout <= 64 ? 1 : 2;
    p.debug = g_wgrad_debug;
    const size_t need = (size_t)pl.splits * Cout * R * S * p.ldws * sizeof(float);
    DN_REQUIRE(workspace_bytes >= need, "conv2d_wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);

    int rc;
    if ((rc = make_act_map(&p.tmDY[0], dy_hi, Cout, Wo, Ho, N, lddy, p.TW, p.TH, p.TN))) return rc;
    if (dy_lo && (rc = make_act_map(&p.tmDY[1], dy_lo, Cout, Wo, Ho, N, lddy, p.TW, p.TH, p.TN))) return rc;
    if (pl.rows_path) {
        const int xw = p.TW + S - 1;
        if ((rc = make_act_map(&p.tmX[0], x_hi, Cin, Wi, Hi, N, ldx, xw, p.TH, p.TN))) return rc;
        if (x_lo && (rc = make_act_map(&p.tmX[1], x_lo, Cin, Wi, Hi, N, ldx, xw, p.TH, p.TN))) return rc;
        switch (pl.BN) {
            case 64: rc = launch_wgrad_rows<64>(p, stream); break;
            case 128: rc = launch_wgrad_rows<128>(p, stream); break;
            default: rc = launch_wgrad_rows<256>(p, stream); break;
        }
    } else {
        if ((rc = make_act_map(&p.tmX[0], x_hi, Cin, Wi, Hi, N, ldx, p.TW, p.TH, p.TN, stride_w, stride_h))) return rc;
        if (x_lo && (rc = make_act_map(&p.tmX[1], x_lo, Cin, Wi, Hi, N, ldx, p.TW, p.TH, p.TN, stride_w, stride_h)))
            return rc;
        rc = wgrad_launch(p, stream);
    }
    if (rc) return rc;
    if (!dw) return 0;      // partial sums only: the caller reduces them later (denet_wgrad_reduce_multi)
    const long long total = (long long)Cout * Cin * R * S;
    const int block = 256;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, block), 148LL * 16);
    wgrad_reduce_kernel<<<DN_G(grid), block, 0, stream>>>(workspace, dw, pl.splits, Cout, Cin, R, S, p.ldws, accumulate);
    DN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------ row-folded conv
// Convolutions with very few input channels (the 3-channel image stem, reference ConvLayer C.B[64,7,2] of
// examples/resnet34-imagenet.sh:7): a 64-channel K chunk per filter tap would be >90 % zero padding and an explicit
// im2col matrix costs ~0.6 GB of HBM traffic per pass.  Instead the image is kept zero-PADDED in NHWC with Cp (4 or 8)
// channels, and the A operand is described to TMA as an OVERLAPPING-window tensor:
//     dim0 = the S*Cp contiguous elements of one filter row under an output pixel   (K of one "tap" = filter row r)
//     dim1 = output column wo, stride = stride_w * Cp elements (16 B)               (windows overlap)
//     dim2 = padded input row, walked with element stride stride_h;  dim3 = image
// so the same fprop / wgrad kernels run an R-tap "1-D" convolution with K = 64 per filter row (columns >= S*Cp are
// TMA zero fill or multiply zero weights).  No im2col buffer exists in either direction.
namespace dn {

static int rowfold_k(int S, int Cp) { return (S * Cp + 7) / 8 * 8; }   // dim0 extent (elements), <= 64

// TH = rows the box delivers (the box spans TH * stride_h input rows walked with element stride stride_h)
static int make_rowfold_map(CUtensorMap* tm, const void* base, int Cp, int S, int Wo, int Hp, int Wp, int N,
                            int stride_w, int stride_h, int TW, int TH, int TN) {
    const int kf = rowfold_k(S, Cp);
    uint64_t dims[4] = {(uint64_t)kf, (uint64_t)Wo, (uint64_t)Hp, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)stride_w * Cp * 2, (uint64_t)Wp * Cp * 2, (uint64_t)Hp * Wp * Cp * 2};
    uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)(TH * stride_h), (uint32_t)TN};
    uint32_t estr[4] = {1, 1, (uint32_t)stride_h, 1};
    if (box[2] > 256) return set_error(DENET_ERR_ARG, "rowfold conv: strided patch exceeds the TMA box limit");
    return encode_tmap_bf16(tm, base, 4, dims, strides, box, estr);
}

static int rowfold_check(const char* what, int Cin, int Cp, int R, int S, int stride_h, int stride_w, int Ho, int Wo,
                         int Hp, int Wp) {
    if (!(Cp == 4 || Cp == 8) || Cin > Cp) return set_error(DENET_ERR_ARG, "%s: Cp must be 4 or 8 and >= Cin", what);
    if ((stride_w * Cp) % 8 != 0) return set_error(DENET_ERR_ARG, "%s: stride_w*Cp must be a multiple of 8", what);
    if ((Wp * Cp) % 8 != 0) return set_error(DENET_ERR_ARG, "%s: Wp*Cp must be a multiple of 8", what);
    if (rowfold_k(S, Cp) > 64) return set_error(DENET_ERR_ARG, "%s: S*Cp must be <= 64", what);
    if (stride_h < 1 || stride_h > 8 || stride_w < 1) return set_error(DENET_ERR_ARG, "%s: bad stride", what);
    if ((Ho - 1) * stride_h + R > Hp) return set_error(DENET_ERR_ARG, "%s: padded height %d too small", what, Hp);
    if ((long long)(Wo - 1) * stride_w * Cp + rowfold_k(S, Cp) > (long long)Wp * Cp)
        return set_error(DENET_ERR_ARG, "%s: padded width %d too small", what, Wp);
    return 0;
}

// B[co][r][k], k = s*Cp + c  <-  W[co][c][R-1-r][S-1-s]  (true convolution -> correlation taps), 64 columns per row r
__global__ void weight_prep_rowfold_kernel(const float* __restrict__ w, int Cout, int Cin, int R, int S, int Cp,
                                           __nv_bfloat16* __restrict__ b_hi, __nv_bfloat16* __restrict__ b_lo) {
    const long long total = (long long)Cout * R * 64;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(idx % 64);
        const int r = (int)((idx / 64) % R);
        const int co = (int)(idx / (64LL * R));
        const int s = k / Cp, c = k % Cp;
        float v = 0.f;
        if (s < S && c < Cin) v = w[(((long long)co * Cin + c) * R + (R - 1 - r)) * S + (S - 1 - s)];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        b_hi[idx] = hi;
        if (b_lo) b_lo[idx] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

// split-K partials [split][Cout][R][ldws] (column k = s*Cp + c) -> reference filter gradient (Cout, Cin, R, S)
__global__ void wgrad_reduce_rowfold_kernel(const float* __restrict__ ws, float* __restrict__ dw, int splits, int Cout,
                                            int Cin, int R, int S, int Cp, int ldws, int accumulate) {
    const long long total = (long long)Cout * Cin * R * S;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int sf = (int)(idx % S);
        const int rf = (int)((idx / S) % R);
        const int c = (int)((idx / ((long long)S * R)) % Cin);
        const int co = (int)(idx / ((long long)S * R * Cin));
        const int r = R - 1 - rf, s = S - 1 - sf;
        float acc = 0.f;
        for (int sp = 0; sp < splits; ++sp)
            acc += ws[(((long long)sp * Cout + co) * R + r) * ldws + s * Cp + c];
        dw[idx] = accumulate ? dw[idx] + acc : acc;
    }
}

// NCHW fp32 image -> interior of a zero-padded NHWC-Cp bf16 buffer (hi [+ lo]); borders and pad channels untouched
__global__ void nchw_to_padded_kernel(const float* __restrict__ x, int N, int C, int H, int W, int Cp, int ph, int pw,
                                      int Hp, int Wp, __nv_bfloat16* __restrict__ y_hi,
                                      __nv_bfloat16* __restrict__ y_lo) {
    const long long total = (long long)N * H * W;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(idx % W);
        const int h = (int)((idx / W) % H);
        const int n = (int)(idx / ((long long)W * H));
        const long long o = (((long long)n * Hp + h + ph) * Wp + w + pw) * Cp;
        for (int c = 0; c < Cp; ++c) {
            const float v = c < C ? x[(((long long)n * C + c) * H + h) * W + w] : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            y_hi[o + c] = hi;
            if (y_lo) y_lo[o + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
    }
}

}  // namespace dn

extern "C" int denet_nchw_to_padded_nhwc(const float* x, int N, int C, int H, int W, int Cp, int pad_h, int pad_w,
                                         int Hp, int Wp, void* y_hi, void* y_lo, cudaStream_t stream) {
    DN_REQUIRE(x && y_hi, "nchw_to_padded_nhwc: null pointer");
    DN_REQUIRE(C <= Cp && H + pad_h <= Hp && W + pad_w <= Wp, "nchw_to_padded_nhwc: image does not fit the buffer");
    const long long total = (long long)N * H * W;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 32);
    nchw_to_padded_kernel<<<DN_G(grid), 256, 0, stream>>>(x, N, C, H, W, Cp, pad_h, pad_w, Hp, Wp, (__nv_bfloat16*)y_hi,
                                                          (__nv_bfloat16*)y_lo);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_conv_weight_prep_rowfold(const float* w, int Cout, int Cin, int R, int S, int Cp, void* b_hi,
                                              void* b_lo, cudaStream_t stream) {
    DN_REQUIRE(w && b_hi, "weight_prep_rowfold: null pointer");
    DN_REQUIRE(Cin <= Cp && rowfold_k(S, Cp) <= 64, "weight_prep_rowfold: S*Cp must be <= 64 and Cp >= Cin");
    const long long total = (long long)Cout * R * 64;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 16);
    weight_prep_rowfold_kernel<<<DN_G(grid), 256, 0, stream>>>(w, Cout, Cin, R, S, Cp, (__nv_bfloat16*)b_hi,
                                                               (__nv_bfloat16*)b_lo);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_conv2d_rowfold_fprop(const void* x_hi, const void* x_lo, int N, int Hp, int Wp, int Cp, int Cin,
                                          const void* b_hi, const void* b_lo, int Cout, int R, int S, int stride_h,
                                          int stride_w, void* y, int y_dtype, long long ldy, int Ho, int Wo,
                                          const float* bias, int relu, float* stat_sum, float* stat_sqsum,
                                          cudaStream_t stream) {
    DN_REQUIRE(x_hi && b_hi && y, "conv2d_rowfold_fprop: null pointer");
    DN_REQUIRE((x_lo == nullptr) == (b_lo == nullptr), "conv2d_rowfold_fprop: x_lo and b_lo must come together");
    DN_REQUIRE(y_dtype == DENET_F32 || y_dtype == DENET_BF16, "conv2d_rowfold_fprop: bad y_dtype %d", y_dtype);
    DN_REQUIRE((stat_sum == nullptr) == (stat_sqsum == nullptr), "conv2d_rowfold_fprop: stat pointers come in pairs");
    int rc;
    if ((rc = rowfold_check("conv2d_rowfold_fprop", Cin, Cp, R, S, stride_h, stride_w, Ho, Wo, Hp, Wp))) return rc;
    ConvFpropParams p;
    memset(&p, 0, sizeof(p));
    p.nterms = x_lo ? 3 : 1;
    p.debug = g_fprop_debug;
    p.timeline = g_fprop_timeline;
    p.R = R; p.S = 1; p.pad_h = 0; p.pad_w = 0;       // taps = filter rows; the padding lives in the buffer
    p.stride_h = stride_h; p.stride_w = 1;             // the column stride lives in the tensor map
    p.kchunks = 1;
    p.Wo = Wo; p.Ho = Ho; p.No = N;
    p.osh = p.osw = 1; p.ooh = p.oow = 0; p.Hf = Ho; p.Wf = Wo;
    pick_patch(Wo, Ho, N, 128, p.TW, p.TH, p.TN);
    p.tiles_w = ceil_div(Wo, p.TW);
    p.tiles_h = ceil_div(Ho, p.TH);
    p.tiles_n = ceil_div(N, p.TN);
    p.Cout = Cout;
    p.ldy = ldy;
    p.y_fp32 = (y_dtype == DENET_F32);
    p.relu = relu;
    p.y = y;
    p.bias = bias;
    p.stat_sum = stat_sum;
    p.stat_sqsum = stat_sqsum;
    // filter rows sharing one fetch of the input rows (conv_fprop_halo_kernel): the rows 2*ho + r of a stride-2 stem
    // are loaded as two boxes, even and odd input rows; filter row r reads box r & 1 starting r >> 1 rows in
    if ((g_fprop_mode & 1) && stride_h <= 2 && R > 1 && R <= 16 && Wo >= 8) {
        ConvFpropParams h = p;
        h.TW = Wo >= 16 ? 16 : 8; h.TH = 128 / h.TW; h.TN = 1;
        h.tiles_w = ceil_div(Wo, h.TW);
        h.tiles_h = ceil_div(Ho, h.TH);
        h.tiles_n = N;
        const int rows = h.TH + (R - 1) / stride_h;
        h.a_loads = stride_h < R ? stride_h : R;
        h.a_dw = 0; h.a_dh = 0; h.a_dh_step = 1;
        h.a_load_bytes = (uint32_t)rows * h.TW * 128;
        h.a_load_stride = (h.a_load_bytes + 1023) / 1024 * 1024;
        h.a_sbo = 1024;
        for (int r = 0; r < R; ++r)      // filter row r: box r % stride_h, r / stride_h rows of TW pixels in
            h.tap_off16[r] = ((uint32_t)(r % stride_h) * h.a_load_stride + (uint32_t)(r / stride_h) * h.TW * 128) >> 4;
        h.kmmas = (rowfold_k(S, Cp) + 15) / 16;
        if (rows * stride_h <= 256) {
            if ((rc = make_rowfold_map(&h.tmA[0], x_hi, Cp, S, Wo, Hp, Wp, N, stride_w, stride_h, h.TW, rows, 1))) return rc;
            if (x_lo && (rc = make_rowfold_map(&h.tmA[1], x_lo, Cp, S, Wo, Hp, Wp, N, stride_w, stride_h, h.TW, rows, 1)))
                return rc;
            rc = halo_finish(h, b_hi, b_lo, stream);
            if (rc <= 0) return rc;
        }
    }
    if ((rc = make_rowfold_map(&p.tmA[0], x_hi, Cp, S, Wo, Hp, Wp, N, stride_w, stride_h, p.TW, p.TH, p.TN))) return rc;
    if (x_lo && (rc = make_rowfold_map(&p.tmA[1], x_lo, Cp, S, Wo, Hp, Wp, N, stride_w, stride_h, p.TW, p.TH, p.TN)))
        return rc;
    return fprop_finish(p, b_hi, b_lo, stream);
}

namespace dn {

// Tiling plan of the row-folded filter gradient.  stem: conv_wgrad_stem_kernel (all R filter rows per tile, 8 x 8 pixel
// k-blocks of one image); otherwise conv_wgrad_kernel with one filter row per tile.
struct RowfoldWgradPlan {
    int stem, TW, TH, TN, total_kb, base_tiles, splits, rows, x_loads, ntap0, ntap1, xbox_stride, stage_bytes, stages;
};

static RowfoldWgradPlan plan_rowfold_wgrad(int N, int Ho, int Wo, int Cout, int R, int stride_h) {
    RowfoldWgradPlan pl;
    memset(&pl, 0, sizeof(pl));
    const int a_boxes = Cout <= 64 ? 1 : 2;
    const int co_tiles = ceil_div(Cout, 128);
    pl.ntap0 = ceil_div(R, stride_h);
    pl.ntap1 = stride_h == 2 ? R / 2 : 0;
    pl.stem = g_wgrad_rows && R > 1 && (stride_h == 1 || stride_h == 2) && pl.ntap0 * 64 <= 256 && R * 64 <= 512 &&
              Wo >= 8 && Ho >= 4;
    if (pl.stem) {
        pl.TW = 8; pl.TH = 8; pl.TN = 1;
        pl.total_kb = ceil_div(Wo, 8) * ceil_div(Ho, 8) * N;
        pl.base_tiles = co_tiles;
        pl.rows = pl.TH + (R - 1) / stride_h;
        pl.x_loads = stride_h < R ? stride_h : R;
        pl.xbox_stride = (pl.rows * pl.TW * 128 + 1023) / 1024 * 1024;
        pl.stage_bytes = a_boxes * 8192 + pl.x_loads * pl.xbox_stride;
        pl.stages = (200 * 1024) / pl.stage_bytes;
        if (pl.stages > 8) pl.stages = 8;
        if (pl.stages < 2) pl.stem = 0;
    }
    if (pl.stem) {
        pl.splits = wgrad_splits(pl.total_kb, pl.base_tiles, R * 64, pl.stage_bytes, (long long)Cout * R * 64);
    } else {
        pick_patch(Wo, Ho, N, 64, pl.TW, pl.TH, pl.TN);
        pl.total_kb = ceil_div(Wo, pl.TW) * ceil_div(Ho, pl.TH) * ceil_div(N, pl.TN);
        pl.base_tiles = co_tiles * R;
        pl.splits = wgrad_splits(pl.total_kb, pl.base_tiles, 64, a_boxes * 8192 + 8192, (long long)Cout * R * 64);
    }
    return pl;
}

}  // namespace dn

extern "C" size_t denet_conv2d_rowfold_wgrad_workspace(int N, int Ho, int Wo, int Cout, int R, int stride_h) {
    const RowfoldWgradPlan pl = plan_rowfold_wgrad(N, Ho, Wo, Cout, R, stride_h);
    return (size_t)pl.splits * Cout * R * 64 * sizeof(float);
}

extern "C" int denet_conv2d_rowfold_wgrad_splits(int N, int Ho, int Wo, int Cout, int R, int stride_h) {
    return plan_rowfold_wgrad(N, Ho, Wo, Cout, R, stride_h).splits;
}

extern "C" int denet_conv2d_rowfold_wgrad(const void* dy_hi, const void* dy_lo, int N, int Ho, int Wo, int Cout,
                                          long long lddy, const void* x_hi, const void* x_lo, int Hp, int Wp, int Cp,
                                          int Cin, int R, int S, int stride_h, int stride_w, float* dw, int accumulate,
                                          float* workspace, size_t workspace_bytes, cudaStream_t stream) {
    DN_REQUIRE(dy_hi && x_hi && workspace, "conv2d_rowfold_wgrad: null pointer");
    DN_REQUIRE((dy_lo == nullptr) == (x_lo == nullptr), "conv2d_rowfold_wgrad: dy_lo and x_lo must come together");
    DN_REQUIRE(lddy % 8 == 0, "conv2d_rowfold_wgrad: dy pitch must be a multiple of 8 elements");
    int rc;
    if ((rc = rowfold_check("conv2d_rowfold_wgrad", Cin, Cp, R, S, stride_h, stride_w, Ho, Wo, Hp, Wp))) return rc;
    ConvWgradParams p;
    memset(&p, 0, sizeof(p));
    p.nterms = dy_lo ? 3 : 1;
    p.R = R; p.S = 1; p.pad_h = 0; p.pad_w = 0;
    p.stride_h = stride_h; p.stride_w = 1;
    const RowfoldWgradPlan pl = plan_rowfold_wgrad(N, Ho, Wo, Cout, R, stride_h);
    p.TW = pl.TW; p.TH = pl.TH; p.TN = pl.TN;
    p.tiles_w = ceil_div(Wo, p.TW);
    p.tiles_h = ceil_div(Ho, p.TH);
    p.tiles_n = ceil_div(N, p.TN);
    p.total_kblocks = pl.total_kb;
    p.co_tiles = ceil_div(Cout, 128);
    p.ci_tiles = 1;
    p.splits = pl.splits;
    p.num_tiles = pl.base_tiles * pl.splits;
    p.debug = g_wgrad_debug;
    p.fd_cot = make_fastdiv(p.co_tiles);
    p.fd_cit = make_fastdiv(1);
    p.fd_taps = make_fastdiv(R);
    p.fd_w = make_fastdiv(p.tiles_w);
    p.fd_h = make_fastdiv(p.tiles_h);
    p.a_boxes = Cout <= 64 ? 1 : 2;
    p.Cout = Cout; p.Cin = 64;             // GEMM N extent: the 64 folded columns of a filter row
    p.ldws = 64;
    p.ws = workspace;
    const size_t need = (size_t)p.splits * Cout * R * 64 * sizeof(float);
    DN_REQUIRE(workspace_bytes >= need, "conv2d_rowfold_wgrad: workspace too small (%zu < %zu)", workspace_bytes, need);
    if ((rc = make_act_map(&p.tmDY[0], dy_hi, Cout, Wo, Ho, N, lddy, p.TW, p.TH, p.TN))) return rc;
    if (dy_lo && (rc = make_act_map(&p.tmDY[1], dy_lo, Cout, Wo, Ho, N, lddy, p.TW, p.TH, p.TN))) return rc;
    const int xbox_rows = pl.stem ? pl.rows : p.TH;
    if ((rc = make_rowfold_map(&p.tmX[0], x_hi, Cp, S, Wo, Hp, Wp, N, stride_w, stride_h, p.TW, xbox_rows, p.TN))) return rc;
    if (x_lo && (rc = make_rowfold_map(&p.tmX[1], x_lo, Cp, S, Wo, Hp, Wp, N, stride_w, stride_h, p.TW, xbox_rows, p.TN)))
        return rc;
    if (pl.stem) {
        p.x_loads = pl.x_loads;
        p.xbox_stride = pl.xbox_stride;
        p.xrows = pl.rows * pl.TW;
        p.stem_stages = pl.stages;
        p.stem_stage_bytes = pl.stage_bytes;
        p.ntap[0] = pl.ntap0;
        p.ntap[1] = pl.ntap1;
        const size_t smem = 1024 + (size_t)pl.stages * pl.stage_bytes + 256;
        DN_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
        conv_wgrad_stem_kernel<<<DN_G(grid), kThreads, smem, stream>>>(p);
        DN_CHECK_LAUNCH();
    } else if ((rc = wgrad_launch(p, stream))) {
        return rc;
    }
    if (!dw) return 0;      // partial sums only (see denet_wgrad_reduce_multi)
    const long long total = (long long)Cout * Cin * R * S;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 16);
    wgrad_reduce_rowfold_kernel<<<DN_G(grid), 256, 0, stream>>>(workspace, dw, p.splits, Cout, Cin, R, S, Cp, p.ldws,
                                                                accumulate);
    DN_CHECK_LAUNCH();
    return 0;
}
