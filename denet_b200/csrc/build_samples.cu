// Directed sparse sampling: corner log-probability maps -> ranked RoI set, on the device.
//
// Replaces the reference's C++ CPython extension entry point build_samples (denet/layer/denet_sparse.cc:559-668):
// run_build_samples :489-557 (threshold scan, optional local-max test, top-max_corners per corner type),
// search_corners :321-374 (TL x BR and TR x BL pairing with bbox-hash dedupe), get_sample :271-308 (score) and the
// final top sample_num^2 by score (:547-549).  SURVEY.md §8 row a10.  The reference runs one std::thread per image
// on the host after a device->host copy of corner_pr; here the maps never leave HBM.
//
// Kernel 1 (corner_select_kernel, one CTA per image x corner type): ordered (row-major) stream compaction of the
//   positions with logp > ln(threshold); if more than max_corners survive, a 4-pass radix select on the
//   order-preserving bit pattern keeps the max_corners most probable (ties at the cut: lowest position first).
// Kernel 2 (pair_select_kernel, one CTA per image): enumerates corner pairs warp-strided, scores them with the
//   reference's exact fp32 sequence, and selects the sample_num^2 best with an MSB-first radix select over a unique
//   64-bit key (score distance bits : box), then sorts the survivors in shared memory (bitonic) - fully
//   deterministic, no hash table: a TR x BL box duplicates a TL x BR box iff its (x0,y0) is a TL corner and its
//   (x1,y1) a BR corner, which two position bitmaps in shared memory answer.
//
// Ranking: the reference sorts by pr = 1/(1+exp(d)), d = |pr_f - pr_t|, descending, with an unstable sort; ranking
// by d ascending is the same order wherever the reference's order is defined, and breaks its ties by box.
#include "common.cuh"
#include "expf_glibc.cuh"

namespace dn {

constexpr int kBsThreads = 1024;
constexpr int kRadixBits = 11;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kSortCap = 4096;  // survivors sorted in shared memory (>= sample_num^2)

// block-wide exclusive scan of one int per thread (kBsThreads threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc - w;  // exclusive
        if (lane == 31) warp_sums[32] = winc;
    }
    __syncthreads();
    const int r = warp_sums[warp] + inc - v;
    *total = warp_sums[32];
    __syncthreads();
    return r;
}

__device__ __forceinline__ uint32_t float_order_key(float f) {
    // ascending unsigned order == ascending float order
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// reference get_local_max (:474-487), including its exclusive upper loop bounds
__device__ __forceinline__ float local_max_at(const float* __restrict__ map, int H, int W, int y, int x, int lm) {
    const int x0 = max(0, x - lm), y0 = max(0, y - lm);
    const int x1 = min(W - 1, x + lm), y1 = min(H - 1, y + lm);
    float m = -100000.f;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) m = fmaxf(m, map[yy * W + xx]);
    return m;
}

__device__ __forceinline__ bool is_corner(const float* __restrict__ map, int H, int W, int pos, float thr, int lm,
                                          float* lp_out) {
    const float lp = map[pos];
    *lp_out = lp;
    if (!(lp > thr)) return false;
    if (lm > 0 && lp < local_max_at(map, H, W, pos / W, pos % W, lm)) return false;
    return true;
}

// corners: [B][CN][max_corners] packed (y << 16 | x); counts: [B][CN]
__global__ void __launch_bounds__(kBsThreads) corner_select_kernel(const float* __restrict__ corner_pr, int CN, int H,
                                                                    int W, float thr, int max_corners, int local_max,
                                                                    uint32_t* __restrict__ corners,
                                                                    int* __restrict__ counts) {
    __shared__ int warp_sums[33];
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int s_need;
    const int b = blockIdx.x / CN, ci = blockIdx.x % CN;      // CN = 4 corner types, 5 with the centre map (DNC.C)
    const int HW = H * W;
    const float* map = corner_pr + (((long long)b * 2 + 1) * CN + ci) * HW;
    uint32_t* out = corners + ((long long)b * CN + ci) * max_corners;

    // pass 0: count
    int mine = 0;
    for (int pos = threadIdx.x; pos < HW; pos += kBsThreads) {
        float lp;
        mine += is_corner(map, H, W, pos, thr, local_max, &lp) ? 1 : 0;
    }
    int total;
    block_excl_scan(mine, warp_sums, &total);

    uint32_t cut_key = 0;  // keep keys > cut_key, plus the first `need_ties` positions with key == cut_key
    int need_ties = 0;
    const bool select = total > max_corners;
    if (select) {
        // radix select (MSB first, 8 bits per pass) of the max_corners-th LARGEST key
        if (threadIdx.x == 0) {
            s_prefix = 0;
            s_need = max_corners;
        }
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            const uint32_t mask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
            for (int pos = threadIdx.x; pos < HW; pos += kBsThreads) {
                float lp;
                if (is_corner(map, H, W, pos, thr, local_max, &lp)) {
                    const uint32_t k = float_order_key(lp);
                    if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int need = s_need, bin = 255;
                for (; bin > 0; --bin) {
                    if (hist[bin] >= need) break;
                    need -= hist[bin];
                }
                s_prefix = prefix | ((uint32_t)bin << shift);
                s_need = need;
            }
            __syncthreads();
        }
        cut_key = s_prefix;
        need_ties = s_need;
    }

    // ordered compaction, 1024 positions per round (row-major order == the reference's scan order)
    int base = 0, tie_base = 0;
    for (int p0 = 0; p0 < HW; p0 += kBsThreads) {
        const int pos = p0 + threadIdx.x;
        float lp = 0.f;
        bool ok = pos < HW && is_corner(map, H, W, pos, thr, local_max, &lp);
        bool tie = false;
        if (ok && select) {
            const uint32_t k = float_order_key(lp);
            tie = (k == cut_key);
            ok = k > cut_key;
        }
        int ntie;
        const int tie_rank = block_excl_scan(tie ? 1 : 0, warp_sums, &ntie);
        if (tie && tie_base + tie_rank < need_ties) ok = true;
        int n;
        const int rank = block_excl_scan(ok ? 1 : 0, warp_sums, &n);
        if (ok) out[base + rank] = ((uint32_t)(pos / W) << 16) | (uint32_t)(pos % W);
        base += n;
        tie_base += ntie;
    }
    if (threadIdx.x == 0) counts[b * CN + ci] = base;
}

struct PairCtx {
    const float* cp;  // corner_pr[b] = (2,CN,H,W)
    int CN, H, W, HW;
    const uint32_t* c0;  // TL
    const uint32_t* c1;  // TR
    const uint32_t* c2;  // BL
    const uint32_t* c3;  // BR
    const uint32_t* c4;  // centres (CN == 5)
    int n0, n1, n2, n3, n4;
    const uint32_t* bm_tl;
    const uint32_t* bm_br;
    const uint32_t* bm_tr;   // CN == 5 only
    const uint32_t* bm_bl;
};

__device__ __forceinline__ bool bm_test(const uint32_t* bm, int p) { return (bm[p >> 5] >> (p & 31)) & 1u; }

// candidate idx -> unique 64-bit rank key (d bits << 32 | box); false if the pair is not a sample
// Centre phase (:377-468, CN == 5): candidate = (centre m, corner type, corner a), centre-major like the reference's
// loops.  The box is the corner reflected through the centre, so a box has exactly ONE possible centre
// ((x0+x1)/2, (y0+y1)/2 with even extents): the reference's hash set can only have seen it before through the TL x BR
// search, the TR x BL search, or an EARLIER corner type of the same centre - three questions the four position bitmaps
// answer without a hash table.
__device__ __forceinline__ bool centre_box(const PairCtx& c, long long j, int* bx0, int* by0, int* bx1, int* by1) {
    const int nsum = c.n0 + c.n1 + c.n2 + c.n3;
    const int m = (int)(j / nsum);
    int a = (int)(j % nsum);
    const uint32_t ctr = c.c4[m];
    const int cx = ctr & 0xffff, cy = ctr >> 16;
    int type = 0;
    if (a >= c.n0) { a -= c.n0; type = 1; if (a >= c.n1) { a -= c.n1; type = 2; if (a >= c.n2) { a -= c.n2; type = 3; } } }
    const uint32_t v = type == 0 ? c.c0[a] : (type == 1 ? c.c1[a] : (type == 2 ? c.c2[a] : c.c3[a]));
    const int px = v & 0xffff, py = v >> 16;
    int x0, y0, x1, y1;
    if (type == 0) { x0 = px; y0 = py; x1 = x0 + 2 * (cx - x0); y1 = y0 + 2 * (cy - y0); }
    else if (type == 1) { x1 = px; y0 = py; x0 = x1 - 2 * (x1 - cx); y1 = y0 + 2 * (cy - y0); }
    else if (type == 2) { x0 = px; y1 = py; x1 = x0 + 2 * (cx - x0); y0 = y1 - 2 * (y1 - cy); }
    else { x1 = px; y1 = py; x0 = x1 - 2 * (x1 - cx); y0 = y1 - 2 * (y1 - cy); }
    if (x0 < 0 || y0 < 0 || x1 >= c.W || y1 >= c.H || x1 <= x0 || y1 <= y0) return false;
    const int p00 = y0 * c.W + x0, p01 = y0 * c.W + x1, p10 = y1 * c.W + x0, p11 = y1 * c.W + x1;
    const bool tl = bm_test(c.bm_tl, p00), tr = bm_test(c.bm_tr, p01), bl = bm_test(c.bm_bl, p10),
               br = bm_test(c.bm_br, p11);
    if ((tl && br) || (tr && bl)) return false;                         // found by the corner-pair searches
    if ((type >= 1 && tl) || (type >= 2 && tr) || (type >= 3 && bl)) return false;   // earlier type, same centre
    *bx0 = x0; *by0 = y0; *bx1 = x1; *by1 = y1;
    return true;
}

__device__ __forceinline__ bool pair_key(const PairCtx& c, long long idx, long long nA, long long nB, uint64_t* key) {
    int x0, y0, x1, y1;
    if (idx >= nB) {
        if (!centre_box(c, idx - nB, &x0, &y0, &x1, &y1)) return false;
    } else if (idx < nA) {  // top-left x bottom-right (:337-353)
        const uint32_t tl = c.c0[idx / c.n3], br = c.c3[idx % c.n3];
        x0 = tl & 0xffff; y0 = tl >> 16; x1 = br & 0xffff; y1 = br >> 16;
        if (x1 <= x0 || y1 <= y0) return false;
    } else {  // top-right x bottom-left (:357-373); skipped when the TL x BR search already produced the box
        const long long j = idx - nA;
        const uint32_t tr = c.c1[j / c.n2], bl = c.c2[j % c.n2];
        x1 = tr & 0xffff; y0 = tr >> 16; x0 = bl & 0xffff; y1 = bl >> 16;
        if (x1 <= x0 || y1 <= y0) return false;
        const int p00 = y0 * c.W + x0, p11 = y1 * c.W + x1;
        if (bm_test(c.bm_tl, p00) && bm_test(c.bm_br, p11)) return false;
    }
    // get_sample (:280-294): fp32 sums in the reference's order, starting from 0
    const float* f = c.cp;
    const float* t = c.cp + c.CN * c.HW;
    const int p00 = y0 * c.W + x0, p01 = y0 * c.W + x1, p10 = y1 * c.W + x0, p11 = y1 * c.W + x1;
    float pr_f = __fadd_rn(0.f, f[p00]);
    pr_f = __fadd_rn(pr_f, f[c.HW + p01]);
    pr_f = __fadd_rn(pr_f, f[2 * c.HW + p10]);
    pr_f = __fadd_rn(pr_f, f[3 * c.HW + p11]);
    float pr_t = __fadd_rn(0.f, t[p00]);
    pr_t = __fadd_rn(pr_t, t[c.HW + p01]);
    pr_t = __fadd_rn(pr_t, t[2 * c.HW + p10]);
    pr_t = __fadd_rn(pr_t, t[3 * c.HW + p11]);
    if (c.CN == 5) {     // :296-303 centre map at the box centre (integer division)
        const int pc = ((y0 + y1) / 2) * c.W + (x0 + x1) / 2;
        pr_f = __fadd_rn(pr_f, f[4 * c.HW + pc]);
        pr_t = __fadd_rn(pr_t, t[4 * c.HW + pc]);
    }
    const float d = fabsf(__fsub_rn(pr_f, pr_t));
    const uint32_t box = ((uint32_t)x0 << 24) | ((uint32_t)y0 << 16) | ((uint32_t)x1 << 8) | (uint32_t)y1;
    *key = ((uint64_t)__float_as_uint(d) << 32) | box;  // d >= 0 (or NaN, which sorts last like the reference's pr=NaN)
    return true;
}

// corner lists of image b -> shared memory, position bitmaps (TL, BR; TR, BL for the centre phase); all threads
__device__ __forceinline__ void setup_pair_ctx(PairCtx& c, const float* corner_pr, int CN, int H, int W, int max_corners,
                                               const uint32_t* corners, const int* counts, int b, uint32_t* s_c,
                                               uint32_t* s_bm, int bm_words) {
    const int HW = H * W;
    uint32_t* s_bm_tl = s_bm;
    uint32_t* s_bm_br = s_bm + bm_words;
    uint32_t* s_bm_tr = s_bm + 2 * bm_words;
    uint32_t* s_bm_bl = s_bm + 3 * bm_words;
    c.cp = corner_pr + (long long)b * 2 * CN * HW;
    c.CN = CN; c.H = H; c.W = W; c.HW = HW;
    c.n0 = counts[b * CN + 0]; c.n1 = counts[b * CN + 1]; c.n2 = counts[b * CN + 2]; c.n3 = counts[b * CN + 3];
    c.n4 = CN == 5 ? counts[b * CN + 4] : 0;
    c.c0 = s_c; c.c1 = s_c + max_corners; c.c2 = s_c + 2 * max_corners; c.c3 = s_c + 3 * max_corners;
    c.c4 = s_c + 4 * max_corners;
    c.bm_tl = s_bm_tl; c.bm_br = s_bm_br; c.bm_tr = s_bm_tr; c.bm_bl = s_bm_bl;
    const uint32_t* gc = corners + (long long)b * CN * max_corners;
    for (int i = threadIdx.x; i < CN * max_corners; i += kBsThreads) {
        const int ci = i / max_corners, j = i % max_corners;
        const int n = counts[b * CN + ci];
        s_c[i] = j < n ? gc[i] : 0u;
    }
    for (int i = threadIdx.x; i < 4 * bm_words; i += kBsThreads) s_bm[i] = 0u;
    __syncthreads();
    for (int ci = 0; ci < 4; ++ci) {
        if (CN != 5 && (ci == 1 || ci == 2)) continue;          // TR / BL bitmaps serve the centre phase only
        const int n = ci == 0 ? c.n0 : (ci == 1 ? c.n1 : (ci == 2 ? c.n2 : c.n3));
        uint32_t* bm = ci == 0 ? s_bm_tl : (ci == 1 ? s_bm_tr : (ci == 2 ? s_bm_bl : s_bm_br));
        for (int i = threadIdx.x; i < n; i += kBsThreads) {
            const uint32_t v = s_c[ci * max_corners + i];
            const int p = (int)(v >> 16) * W + (int)(v & 0xffff);
            atomicOr(&bm[p >> 5], 1u << (p & 31));
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kBsThreads) pair_select_kernel(const float* __restrict__ corner_pr, int CN, int H,
                                                                  int W, int max_corners, int K,
                                                                  const uint32_t* __restrict__ corners,
                                                                  const int* __restrict__ counts,
                                                                  float* __restrict__ out_pr,
                                                                  float* __restrict__ out_bbox,
                                                                  int* __restrict__ out_ibox,
                                                                  int* __restrict__ out_count,
                                                                  int* __restrict__ out_ncand) {
    extern __shared__ __align__(16) uint8_t bs_smem[];
    const int b = blockIdx.x;
    const int HW = H * W;
    const int bm_words = (HW + 31) / 32;
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(bs_smem);                   // kSortCap
    uint32_t* s_c = reinterpret_cast<uint32_t*>(s_keys + kSortCap);            // CN * max_corners
    uint32_t* s_bm_tl = s_c + CN * max_corners;                                // bm_words each
    uint32_t* s_bm_br = s_bm_tl + bm_words;
    uint32_t* s_bm_tr = s_bm_br + bm_words;                                    // (used when CN == 5)
    uint32_t* s_bm_bl = s_bm_tr + bm_words;
    int* s_hist = reinterpret_cast<int*>(s_bm_bl + bm_words);                  // kRadixBins
    __shared__ int s_cnt;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_below, s_done;

    PairCtx c;
    setup_pair_ctx(c, corner_pr, CN, H, W, max_corners, corners, counts, b, s_c, s_bm_tl, bm_words);
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();

    const long long nA = (long long)c.n0 * c.n3;
    const long long nB = nA + (long long)c.n1 * c.n2;
    const long long nP = nB + (long long)c.n4 * ((long long)c.n0 + c.n1 + c.n2 + c.n3);

    // total number of samples (unique valid boxes)
    {
        int mine = 0;
        for (long long idx = threadIdx.x; idx < nP; idx += kBsThreads) {
            uint64_t key;
            mine += pair_key(c, idx, nA, nB, &key) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
        __syncthreads();
    }
    const int total = s_cnt;
    __syncthreads();

    // MSB-first radix select of the K-th smallest key; stop as soon as everything up to the selected bin fits in
    // the shared-memory sort buffer.  prefix/shift describe the accepted key range: (key >> shift) <= prefix.
    int shift = 64;
    unsigned long long prefix = 0;
    if (total > kSortCap) {
        if (threadIdx.x == 0) {
            s_prefix = 0;
            s_need = K;
            s_below = 0;
            s_done = 0;
        }
        __syncthreads();
        while (true) {
            const int bits = shift > 32 ? ((shift - 32) >= kRadixBits ? kRadixBits : shift - 32)
                                        : (shift >= kRadixBits ? kRadixBits : shift);
            const int nshift = shift - bits;
            for (int i = threadIdx.x; i < kRadixBins; i += kBsThreads) s_hist[i] = 0;
            __syncthreads();
            const unsigned long long pfx = s_prefix;
            for (long long idx = threadIdx.x; idx < nP; idx += kBsThreads) {
                uint64_t key;
                if (pair_key(c, idx, nA, nB, &key)) {
                    if (shift == 64 || (key >> shift) == pfx) atomicAdd(&s_hist[(key >> nshift) & ((1u << bits) - 1)], 1);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int need = s_need, bin = 0;
                const int nb = 1 << bits;
                for (; bin < nb - 1; ++bin) {
                    if (s_hist[bin] >= need) break;
                    need -= s_hist[bin];
                    s_below += s_hist[bin];
                }
                s_prefix = (pfx << bits) | (unsigned long long)bin;
                s_need = need;
                s_done = (s_below + s_hist[bin] <= kSortCap) || nshift == 0;
            }
            __syncthreads();
            shift = nshift;
            if (s_done) break;
        }
        prefix = s_prefix;
    }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();

    // collect survivors
    for (long long idx = threadIdx.x; idx < nP; idx += kBsThreads) {
        uint64_t key;
        if (pair_key(c, idx, nA, nB, &key)) {
            if (shift == 64 || (key >> shift) <= prefix) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < kSortCap) s_keys[slot] = key;
            }
        }
    }
    __syncthreads();
    const int nsurv = s_cnt < kSortCap ? s_cnt : kSortCap;
    int npow = 1;
    while (npow < nsurv) npow <<= 1;
    for (int i = nsurv + threadIdx.x; i < npow; i += kBsThreads) s_keys[i] = ~0ull;
    __syncthreads();
    // bitonic sort ascending
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += kBsThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = s_keys[i], bb = s_keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > bb) == up) {
                        s_keys[i] = bb;
                        s_keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
    const int nout = nsurv < K ? nsurv : K;
    for (int i = threadIdx.x; i < K; i += kBsThreads) {
        const long long o = (long long)b * K + i;
        if (i < nout) {
            const uint64_t key = s_keys[i];
            const float d = __uint_as_float((uint32_t)(key >> 32));
            const uint32_t box = (uint32_t)key;
            const int x0 = box >> 24, y0 = (box >> 16) & 255, x1 = (box >> 8) & 255, y1 = box & 255;
            // :306  float pr = 1.0 / (1.0 + std::exp(fabs(pr_f - pr_t)))  - float exp, double division, rounded to float
            const float e = expf_glibc(d);  // libm expf bit for bit (expf_glibc.cuh)
            out_pr[o] = (float)(1.0 / (1.0 + (double)e));
            // :307  (double)x0 / width ... rounded to float by the SampleType constructor
            out_bbox[o * 4 + 0] = (float)((double)x0 / (double)W);
            out_bbox[o * 4 + 1] = (float)((double)y0 / (double)H);
            out_bbox[o * 4 + 2] = (float)((double)(x1 + 1) / (double)W);
            out_bbox[o * 4 + 3] = (float)((double)(y1 + 1) / (double)H);
            out_ibox[o * 4 + 0] = x0; out_ibox[o * 4 + 1] = y0; out_ibox[o * 4 + 2] = x1; out_ibox[o * 4 + 3] = y1;
        } else {
            out_pr[o] = 0.f;
            out_bbox[o * 4 + 0] = out_bbox[o * 4 + 1] = out_bbox[o * 4 + 2] = out_bbox[o * 4 + 3] = 0.f;
            out_ibox[o * 4 + 0] = out_ibox[o * 4 + 1] = out_ibox[o * 4 + 2] = out_ibox[o * 4 + 3] = 0;
        }
    }
    if (threadIdx.x == 0) {
        out_count[b] = nout;
        if (out_ncand) out_ncand[b] = total;
    }
}

// ------------------------------------------------------------------------------------------------ corner clustering
// apply_cluster (denet_sparse.cc:165-242), run by the reference on images whose corner search found more than
// sample_num^2 boxes when the sparse layer's nmsThreshold is < 1.  The reference processes the samples one by one on
// the CPU: a sample joins the last cluster of a list that it overlaps (IoU > threshold with any member), the other
// overlapping clusters are merged into that one.  What that sequential procedure computes is order-free except for the
// list order, and both have closed forms that parallelise:
//   * the clusters are the CONNECTED COMPONENTS of the graph "IoU(i, j) > threshold" over the input samples;
//   * a merged cluster keeps the list position of the youngest cluster it absorbed, so the final list is ordered by
//     the largest "creator" of each component - a creator being a sample with no edge to an earlier sample.
// Three kernels: collect (CTA per image: the input set - all candidates in enumeration order, or the 10 K best by score
// when there are more - with their scores / float boxes, in global memory), edges (grid over sample rows: IoU tests,
// lock-free union-find with atomicCAS hooking, creator flags), finish (CTA per image: component sizes and list
// positions, the clusters that survive the cap, 1 + floor(size * ratio) best samples of each, final ranking).
// Sorts are bitonic over global memory (n <= 32768 per image).  Equal scores: see pair_select_kernel.
struct ClusterWs {                   // per-image scratch, all arrays of capacity `cap` (power of two >= 10 * K)
    unsigned long long* key;         // (d bits << 32 | box): ascending = descending score
    unsigned long long* aux;         // sort scratch
    float4* box;                     // normalised float box of the sample at processing position p
    float* pr;                       // score
    int* parent;                     // union-find
    int* prank;                      // rank of position p in score order
    int* flags;                      // bit0 creator
    int* tmp;                        // sizes / list keys / selections
    int* n;                          // [1] number of input samples (0: image not clustered)
};

__device__ __forceinline__ ClusterWs cluster_ws(void* base, int b, int cap) {
    // layout per image: key, aux (u64) | box (float4) | pr, parent, prank, flags (4 B) | tmp (4 x cap ints) | n
    const size_t per = (size_t)cap * (8 + 8 + 16 + 4 * 4 + 16) + 16;
    uint8_t* p = reinterpret_cast<uint8_t*>(base) + per * (size_t)b;
    ClusterWs w;
    w.key = reinterpret_cast<unsigned long long*>(p); p += (size_t)cap * 8;
    w.aux = reinterpret_cast<unsigned long long*>(p); p += (size_t)cap * 8;
    w.box = reinterpret_cast<float4*>(p); p += (size_t)cap * 16;
    w.pr = reinterpret_cast<float*>(p); p += (size_t)cap * 4;
    w.parent = reinterpret_cast<int*>(p); p += (size_t)cap * 4;
    w.prank = reinterpret_cast<int*>(p); p += (size_t)cap * 4;
    w.flags = reinterpret_cast<int*>(p); p += (size_t)cap * 4;
    w.tmp = reinterpret_cast<int*>(p); p += (size_t)cap * 16;
    w.n = reinterpret_cast<int*>(p);
    return w;
}
static size_t cluster_ws_bytes(int B, int cap) { return ((size_t)cap * (8 + 8 + 16 + 4 * 4 + 16) + 16) * (size_t)B; }

// ascending bitonic sort of a[0..npow) (npow a power of two, padded by the caller), whole CTA
__device__ __forceinline__ void bitonic_sort_global(unsigned long long* a, int npow) {
    for (int k = 2; k <= npow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += kBsThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) {
                        a[i] = y;
                        a[ixj] = x;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void key_to_sample(unsigned long long key, int H, int W, float* pr, float4* box) {
    const float d = __uint_as_float((uint32_t)(key >> 32));
    const uint32_t bx = (uint32_t)key;
    const int x0 = bx >> 24, y0 = (bx >> 16) & 255, x1 = (bx >> 8) & 255, y1 = bx & 255;
    *pr = (float)(1.0 / (1.0 + (double)expf_glibc(d)));
    *box = make_float4((float)((double)x0 / (double)W), (float)((double)y0 / (double)H),
                       (float)((double)(x1 + 1) / (double)W), (float)((double)(y1 + 1) / (double)H));
}

__global__ void __launch_bounds__(kBsThreads) cluster_collect_kernel(const float* __restrict__ corner_pr, int CN, int H,
                                                                      int W, int max_corners, int K, int cap,
                                                                      const uint32_t* __restrict__ corners,
                                                                      const int* __restrict__ counts, void* ws_base) {
    extern __shared__ __align__(16) uint8_t bs_smem[];
    const int b = blockIdx.x;
    const int bm_words = (H * W + 31) / 32;
    uint32_t* s_c = reinterpret_cast<uint32_t*>(bs_smem);
    uint32_t* s_bm = s_c + CN * max_corners;
    int* s_hist = reinterpret_cast<int*>(s_bm + 4 * bm_words);
    __shared__ int warp_sums[33];
    __shared__ int s_cnt;
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need;
    ClusterWs w = cluster_ws(ws_base, b, cap);
    PairCtx c;
    setup_pair_ctx(c, corner_pr, CN, H, W, max_corners, corners, counts, b, s_c, s_bm, bm_words);
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const long long nA = (long long)c.n0 * c.n3;
    const long long nB = nA + (long long)c.n1 * c.n2;
    const long long nP = nB + (long long)c.n4 * ((long long)c.n0 + c.n1 + c.n2 + c.n3);
    {
        int mine = 0;
        for (long long idx = threadIdx.x; idx < nP; idx += kBsThreads) {
            uint64_t key;
            mine += pair_key(c, idx, nA, nB, &key) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
        __syncthreads();
    }
    const int total = s_cnt;
    __syncthreads();
    const int input_num = 10 * K;                       // cluster_snum (:497)
    if (total <= K) {                                   // (:540) not over-full: pair_select_kernel's result stands
        if (threadIdx.x == 0) *w.n = 0;
        return;
    }
    const int n = total < input_num ? total : input_num;
    if (total <= input_num) {
        // all candidates, in the reference's enumeration order (= idx order): ordered compaction
        int base = 0;
        for (long long i0 = 0; i0 < nP; i0 += kBsThreads) {
            const long long idx = i0 + threadIdx.x;
            uint64_t key = 0;
            const bool ok = idx < nP && pair_key(c, idx, nA, nB, &key);
            int cnt;
            const int rank = block_excl_scan(ok ? 1 : 0, warp_sums, &cnt);
            if (ok) w.key[base + rank] = key;
            base += cnt;
        }
    } else {
        // the input_num best by score: MSB radix select of the input_num-th smallest (unique) key, then collect + sort
        if (threadIdx.x == 0) {
            s_prefix = 0;
            s_need = input_num;
        }
        __syncthreads();
        int shift = 64;
        while (shift > 0) {
            const int bits = shift >= kRadixBits ? kRadixBits : shift;
            const int nshift = shift - bits;
            for (int i = threadIdx.x; i < kRadixBins; i += kBsThreads) s_hist[i] = 0;
            __syncthreads();
            const unsigned long long pfx = s_prefix;
            for (long long idx = threadIdx.x; idx < nP; idx += kBsThreads) {
                uint64_t key;
                if (pair_key(c, idx, nA, nB, &key)) {
                    if (shift == 64 || (key >> shift) == pfx) atomicAdd(&s_hist[(key >> nshift) & ((1u << bits) - 1)], 1);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int need = s_need, bin = 0;
                const int nb = 1 << bits;
                for (; bin < nb - 1; ++bin) {
                    if (s_hist[bin] >= need) break;
                    need -= s_hist[bin];
                }
                s_prefix = (pfx << bits) | (unsigned long long)bin;
                s_need = need;
            }
            __syncthreads();
            shift = nshift;
        }
        const unsigned long long kth = s_prefix;        // keys are unique: exactly input_num keys are <= kth
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (long long idx = threadIdx.x; idx < nP; idx += kBsThreads) {
            uint64_t key;
            if (pair_key(c, idx, nA, nB, &key) && key <= kth) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot < cap) w.key[slot] = key;
            }
        }
        __syncthreads();
        int npow = 1;
        while (npow < n) npow <<= 1;
        for (int i = n + threadIdx.x; i < npow; i += kBsThreads) w.key[i] = ~0ull;
        __syncthreads();
        bitonic_sort_global(w.key, npow);               // ascending key = descending score (:170-173)
    }
    __syncthreads();
    // scores, float boxes, union-find initial state; score rank of every processing position
    for (int i = threadIdx.x; i < n; i += kBsThreads) {
        key_to_sample(w.key[i], H, W, &w.pr[i], &w.box[i]);
        w.parent[i] = i;
        w.flags[i] = 1;                                 // creator until an edge to an earlier sample is found
    }
    if (total <= input_num) {
        int npow = 1;
        while (npow < n) npow <<= 1;
        // (key with its low 32 box bits replaced by the position would lose uniqueness of equal d: sort (d, position))
        for (int i = threadIdx.x; i < npow; i += kBsThreads)
            w.aux[i] = i < n ? ((w.key[i] & 0xffffffff00000000ull) | (unsigned)i) : ~0ull;
        __syncthreads();
        bitonic_sort_global(w.aux, npow);
        for (int r = threadIdx.x; r < n; r += kBsThreads) w.prank[(int)(w.aux[r] & 0xffffffffu)] = r;
    } else {
        for (int i = threadIdx.x; i < n; i += kBsThreads) w.prank[i] = i;
    }
    if (threadIdx.x == 0) *w.n = n;
}

__device__ __forceinline__ float sample_iou(const float4 a, const float4 b) {
    // SampleType::overlap / overlap_iou (:90-101), fp32 operation for operation
    const float dx = fmaxf(0.0f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
    const float dy = fmaxf(0.0f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
    const float ai = __fmul_rn(dx, dy);
    const float aa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float ab = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(ai, __fsub_rn(__fadd_rn(aa, ab), ai));
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
    // loads bypass L1 (__ldcg): other SMs hook roots concurrently and L1 is not coherent - a stale root would make
    // the hooking CAS below fail forever
    while (true) {
        const int p = __ldcg(parent + x);
        if (p == x) return x;
        const int g = __ldcg(parent + p);
        if (g != p) atomicCAS(&parent[x], p, g);        // path halving (benign race: g is always an ancestor of x)
        x = p;
    }
}

// grid (row blocks, B): thread = sample i of the block's row range, tests every earlier sample j
constexpr int kEdgeThreads = 256;
__global__ void __launch_bounds__(kEdgeThreads) cluster_edges_kernel(void* ws_base, int cap, float threshold) {
    const int b = blockIdx.y;
    ClusterWs w = cluster_ws(ws_base, b, cap);
    const int n = *w.n;
    __shared__ float4 s_box[kEdgeThreads];
    const int i = blockIdx.x * kEdgeThreads + threadIdx.x;
    if (blockIdx.x * kEdgeThreads >= n) return;
    const float4 bi = i < n ? w.box[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    bool creator = true;
    const int jmax = min(n, (blockIdx.x + 1) * kEdgeThreads);
    for (int j0 = 0; j0 < jmax; j0 += kEdgeThreads) {
        __syncthreads();
        if (j0 + threadIdx.x < n) s_box[threadIdx.x] = w.box[j0 + threadIdx.x];
        __syncthreads();
        const int jend = min(kEdgeThreads, min(i, n) - j0);           // j < i only
        for (int jj = 0; jj < jend; ++jj) {
            if (sample_iou(bi, s_box[jj]) > threshold) {
                creator = false;
                // union(i, j0 + jj): hook the larger root under the smaller
                int a = i, c2 = j0 + jj;
                while (true) {
                    a = uf_find(w.parent, a);
                    c2 = uf_find(w.parent, c2);
                    if (a == c2) break;
                    const int hi = a > c2 ? a : c2, lo = a > c2 ? c2 : a;
                    if (atomicCAS(&w.parent[hi], hi, lo) == hi) break;
                }
            }
        }
    }
    if (i < n && !creator) w.flags[i] = 0;
}

__global__ void __launch_bounds__(kBsThreads) cluster_finish_kernel(void* ws_base, int cap, int K, int H, int W,
                                                                     float* __restrict__ out_pr,
                                                                     float* __restrict__ out_bbox,
                                                                     int* __restrict__ out_ibox,
                                                                     int* __restrict__ out_count) {
    const int b = blockIdx.x;
    ClusterWs w = cluster_ws(ws_base, b, cap);
    const int n = *w.n;
    if (n == 0) return;                                  // image not clustered: pair_select_kernel's output stands
    __shared__ int warp_sums[33];
    __shared__ int s_nsel;
    int* size = w.tmp;                                   // [cap] members of the component rooted at r
    int* lkey = w.tmp + cap;                             // [cap] list position key of the component rooted at r
    int* keep = w.tmp + 2 * cap;                         // [cap] 1 if the component rooted at r survives the cap
    int* take = w.tmp + 3 * cap;                         // [cap] samples the component rooted at r contributes
    for (int i = threadIdx.x; i < n; i += kBsThreads) {
        size[i] = 0;
        lkey[i] = -1;
        keep[i] = 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kBsThreads) {  // flatten (the edges kernel has completed: no more hooks)
        int r = i;
        while (w.parent[r] != r) r = w.parent[r];
        w.parent[i] = r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kBsThreads) {
        const int r = w.parent[i];
        atomicAdd(&size[r], 1);
        if (w.flags[i] & 1) atomicMax(&lkey[r], i);      // youngest creator = list position of the merged cluster
    }
    __syncthreads();
    // clusters (roots) in list order; ncl of them
    int ncl;
    {
        int base = 0;
        for (int i0 = 0; i0 < n; i0 += kBsThreads) {
            const int i = i0 + threadIdx.x;
            int cnt;
            block_excl_scan((i < n && w.parent[i] == i) ? 1 : 0, warp_sums, &cnt);
            base += cnt;
        }
        ncl = base;
    }
    if (ncl > K) {
        // :211-221 keep the K clusters with the most samples; std::list::sort is stable: ties keep list order.
        // sort key: (n - size) << 32 | list key  (ascending)
        int npow = 1;
        while (npow < n) npow <<= 1;
        for (int i = threadIdx.x; i < npow; i += kBsThreads)
            w.aux[i] = (i < n && w.parent[i] == i)
                           ? (((unsigned long long)(unsigned)(n - size[i]) << 32) | (unsigned)lkey[i]) : ~0ull;
        __syncthreads();
        bitonic_sort_global(w.aux, npow);
        // the cluster with list key v is the one whose youngest creator is sample v: its root is parent[v]
        for (int r = threadIdx.x; r < K; r += kBsThreads) keep[w.parent[(int)(w.aux[r] & 0xffffffffu)]] = 1;
        ncl = K;
    } else {
        for (int i = threadIdx.x; i < n; i += kBsThreads)
            if (w.parent[i] == i) keep[i] = 1;
    }
    __syncthreads();
    // :225 cluster_ratio; every kept cluster contributes its 1 + floor(size * ratio) best samples (:230)
    const double ratio = (double)(K - ncl) / (double)(n - ncl);
    for (int i = threadIdx.x; i < n; i += kBsThreads) {
        if (w.parent[i] == i && keep[i]) {
            long long t = 1 + (long long)floor((double)size[i] * ratio);
            take[i] = (int)(t < size[i] ? t : size[i]);
        } else if (w.parent[i] == i) {
            take[i] = 0;
        }
    }
    __syncthreads();
    // rank of every sample inside its cluster by score: sort (root, score rank), position inside the root's segment
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int i = threadIdx.x; i < npow; i += kBsThreads)
        w.aux[i] = i < n ? (((unsigned long long)(unsigned)w.parent[i] << 32) | (unsigned)w.prank[i]) : ~0ull;
    __syncthreads();
    bitonic_sort_global(w.aux, npow);
    // segment starts: first sorted position of each root -> lkey (reused)
    for (int r = threadIdx.x; r < n; r += kBsThreads) {
        const int root = (int)(w.aux[r] >> 32);
        if (r == 0 || (int)(w.aux[r - 1] >> 32) != root) lkey[root] = r;
    }
    __syncthreads();
    // selected samples, keyed for the final ranking by their score key
    if (threadIdx.x == 0) s_nsel = 0;
    __syncthreads();
    // prank -> processing position map in flags (score rank r belongs to position flags[r] >> 1)
    for (int i = threadIdx.x; i < n; i += kBsThreads) atomicOr(&w.flags[w.prank[i]], i << 1);
    __syncthreads();
    // (size[] of a root is read before any slot of it is overwritten?  no: slots and roots share the array, so the
    // selected positions go to the sort scratch of the SECOND half of aux, which the n <= cap/2 ... cap layout leaves
    // free only when npow < cap; use prank[] instead - it is not needed any more once flags[] holds the inverse map)
    __syncthreads();
    int* selpos = w.prank;
    for (int r = threadIdx.x; r < n; r += kBsThreads) {
        const int root = (int)(w.aux[r] >> 32);
        const int within = r - lkey[root];
        if (keep[root] && within < take[root]) {
            const int pos = w.flags[(int)(w.aux[r] & 0xffffffffu)] >> 1;     // processing position of that score rank
            selpos[atomicAdd(&s_nsel, 1)] = pos;
        }
    }
    __syncthreads();
    const int nsel = s_nsel;                              // <= K by construction
    int spow = 1;
    while (spow < nsel) spow <<= 1;
    for (int i = threadIdx.x; i < spow; i += kBsThreads) w.aux[i] = i < nsel ? w.key[selpos[i]] : ~0ull;
    __syncthreads();
    bitonic_sort_global(w.aux, spow);                    // :547 final ranking, descending score
    const int nout = nsel < K ? nsel : K;
    for (int i = threadIdx.x; i < K; i += kBsThreads) {
        const long long o = (long long)b * K + i;
        if (i < nout) {
            const unsigned long long key = w.aux[i];
            float pr;
            float4 bx;
            key_to_sample(key, H, W, &pr, &bx);
            const uint32_t q = (uint32_t)key;
            out_pr[o] = pr;
            out_bbox[o * 4 + 0] = bx.x; out_bbox[o * 4 + 1] = bx.y; out_bbox[o * 4 + 2] = bx.z; out_bbox[o * 4 + 3] = bx.w;
            out_ibox[o * 4 + 0] = q >> 24; out_ibox[o * 4 + 1] = (q >> 16) & 255;
            out_ibox[o * 4 + 2] = (q >> 8) & 255; out_ibox[o * 4 + 3] = q & 255;
        } else {
            out_pr[o] = 0.f;
            out_bbox[o * 4 + 0] = out_bbox[o * 4 + 1] = out_bbox[o * 4 + 2] = out_bbox[o * 4 + 3] = 0.f;
            out_ibox[o * 4 + 0] = out_ibox[o * 4 + 1] = out_ibox[o * 4 + 2] = out_ibox[o * 4 + 3] = 0;
        }
    }
    if (threadIdx.x == 0) out_count[b] = nout;
}

static size_t pair_smem_bytes(int CN, int H, int W, int max_corners) {
    const size_t bm_words = ((size_t)H * W + 31) / 32;
    return sizeof(uint64_t) * kSortCap + sizeof(uint32_t) * CN * max_corners + sizeof(uint32_t) * 4 * bm_words +
           sizeof(int) * kRadixBins;
}

}  // namespace dn

using namespace dn;

extern "C" size_t denet_build_samples_workspace(int B, int H, int W, int max_corners) {
    (void)H; (void)W;          // sized for 5 corner types (DNC.C); 4 need less
    return (size_t)B * 5 * max_corners * sizeof(uint32_t) + (size_t)B * 5 * sizeof(int);
}

extern "C" int denet_build_samples(const float* corner_pr, int B, int H, int W, float corner_threshold, int sample_num,
                                   int max_corners, int local_max, float* out_pr, float* out_bbox, int* out_ibox,
                                   int* out_count, int* out_ncand, void* workspace, size_t workspace_bytes,
                                   cudaStream_t stream) {
    return denet_build_samples_cn(corner_pr, B, 4, H, W, corner_threshold, sample_num, max_corners, local_max, out_pr,
                                  out_bbox, out_ibox, out_count, out_ncand, workspace, workspace_bytes, stream);
}

static int cluster_cap(int sample_num) {
    int cap = 1;
    while (cap < 10 * sample_num * sample_num) cap <<= 1;
    return cap;
}

extern "C" size_t denet_build_samples_cluster_workspace(int B, int H, int W, int max_corners, int sample_num) {
    const size_t base = (denet_build_samples_workspace(B, H, W, max_corners) + 255) / 256 * 256;
    return base + cluster_ws_bytes(B, cluster_cap(sample_num));
}

extern "C" int denet_build_samples_cluster(const float* corner_pr, int B, int corner_num, int H, int W,
                                           float corner_threshold, int sample_num, int max_corners, int local_max,
                                           float cluster_threshold, float* out_pr, float* out_bbox, int* out_ibox,
                                           int* out_count, int* out_ncand, void* workspace, size_t workspace_bytes,
                                           cudaStream_t stream) {
    const bool cluster = cluster_threshold < 1.0f;
    DN_REQUIRE(!cluster || workspace_bytes >= denet_build_samples_cluster_workspace(B, H, W, max_corners, sample_num),
               "build_samples: workspace too small for clustering");
    int rc = denet_build_samples_cn(corner_pr, B, corner_num, H, W, corner_threshold, sample_num, max_corners, local_max,
                                    out_pr, out_bbox, out_ibox, out_count, out_ncand, workspace, workspace_bytes, stream);
    if (rc || !cluster) return rc;
    // images with more than sample_num^2 boxes are re-done through apply_cluster (denet_sparse.cc:540-541)
    const int CN = corner_num, K = sample_num * sample_num, cap = cluster_cap(sample_num);
    DN_REQUIRE(cap <= 65536, "build_samples: clustering supports sample_num <= 80");
    uint32_t* corners = reinterpret_cast<uint32_t*>(workspace);
    int* counts = reinterpret_cast<int*>(corners + (size_t)B * CN * max_corners);
    void* cws = reinterpret_cast<uint8_t*>(workspace) + (denet_build_samples_workspace(B, H, W, max_corners) + 255) / 256 * 256;
    const size_t bm_words = ((size_t)H * W + 31) / 32;
    const size_t smem = sizeof(uint32_t) * CN * max_corners + sizeof(uint32_t) * 4 * bm_words + sizeof(int) * kRadixBins;
    DN_CHECK_CUDA(cudaFuncSetAttribute(cluster_collect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cluster_collect_kernel<<<DN_G(B), kBsThreads, smem, stream>>>(corner_pr, CN, H, W, max_corners, K, cap, corners, counts,
                                                                  cws);
    DN_CHECK_LAUNCH();
    cluster_edges_kernel<<<DN_G(dim3(ceil_div(10 * K, kEdgeThreads), B)), kEdgeThreads, 0, stream>>>(cws, cap,
                                                                                                   cluster_threshold);
    DN_CHECK_LAUNCH();
    cluster_finish_kernel<<<DN_G(B), kBsThreads, 0, stream>>>(cws, cap, K, H, W, out_pr, out_bbox, out_ibox, out_count);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_build_samples_cn(const float* corner_pr, int B, int corner_num, int H, int W,
                                      float corner_threshold, int sample_num, int max_corners, int local_max,
                                      float* out_pr, float* out_bbox, int* out_ibox, int* out_count, int* out_ncand,
                                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int CN = corner_num;
    DN_REQUIRE(CN == 4 || CN == 5, "build_samples: corner_num must be 4 or 5 (centre map), got %d", CN);
    DN_REQUIRE(corner_pr && out_pr && out_bbox && out_ibox && out_count && workspace, "build_samples: null pointer");
    DN_REQUIRE(B > 0 && H > 0 && W > 0 && H <= 256 && W <= 256, "build_samples: map size must be in [1,256]");
    DN_REQUIRE(sample_num > 0 && sample_num * sample_num <= kSortCap, "build_samples: sample_num^2 must be <= %d",
               kSortCap);
    DN_REQUIRE(max_corners > 0 && max_corners <= 4096, "build_samples: max_corners must be in [1,4096]");
    DN_REQUIRE(workspace_bytes >= denet_build_samples_workspace(B, H, W, max_corners),
               "build_samples: workspace too small");
    uint32_t* corners = reinterpret_cast<uint32_t*>(workspace);
    int* counts = reinterpret_cast<int*>(corners + (size_t)B * CN * max_corners);
    const float thr = logf(corner_threshold);  // std::log(float), denet_sparse.cc:504
    corner_select_kernel<<<DN_G(B * CN), kBsThreads, 0, stream>>>(corner_pr, CN, H, W, thr, max_corners, local_max, corners,
                                                                  counts);
    DN_CHECK_LAUNCH();
    const size_t smem = pair_smem_bytes(CN, H, W, max_corners);
    DN_REQUIRE(smem <= 200 * 1024, "build_samples: shared memory budget exceeded");
    DN_CHECK_CUDA(cudaFuncSetAttribute(pair_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pair_select_kernel<<<DN_G(B), kBsThreads, smem, stream>>>(corner_pr, CN, H, W, max_corners, sample_num * sample_num,
                                                              corners, counts, out_pr, out_bbox, out_ibox, out_count,
                                                              out_ncand);
    DN_CHECK_LAUNCH();
    return 0;
}
