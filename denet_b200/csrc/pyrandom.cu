// Host-side helper (no device code): the python-`random` post-processing of the ranked RoIs that the reference does in
// DeNetSparseLayer.get_target (denet/layer/denet_sparse.py:184-201), for the whole batch in one native call and with
// EXACTLY the reference's consumption of CPython's Mersenne-Twister stream, so that a run seeded like the reference
// draws the same RoIs:
//
//     if len(samples) > n_keep: samples = random.sample(samples, n_keep)        # CPython Lib/random.py sample()
//     while len(samples) < K:   x0, y0 = random.uniform(0,1) x2; x1 = random.uniform(x0,1); y1 = random.uniform(y0,1)
//
// Third-party algorithms restated here (public, pinned by the interpreter the reference runs on):
//   MT19937 genrand_uint32 / genrand_res53     CPython Modules/_randommodule.c
//   Random.sample (pool / set variants), Random._randbelow_with_getrandbits, Random.uniform   CPython Lib/random.py (3.12)
// tests/test_host.py checks every function against the interpreter's own `random` module.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace dn {

struct PyMT {
    uint32_t* mt;   // 624 words
    int* pos;       // index of the next word (624 = regenerate)
    uint32_t next() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t MATRIX_A = 0x9908b0dfU, UPPER = 0x80000000U, LOWER = 0x7fffffffU;
        if (*pos >= N) {
            int kk;
            uint32_t y;
            for (kk = 0; kk < N - M; kk++) {
                y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1U) ? MATRIX_A : 0U);
            }
            for (; kk < N - 1; kk++) {
                y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1U) ? MATRIX_A : 0U);
            }
            y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1U) ? MATRIX_A : 0U);
            *pos = 0;
        }
        uint32_t y = mt[(*pos)++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680U;
        y ^= (y << 15) & 0xefc60000U;
        y ^= (y >> 18);
        return y;
    }
};

// The same stream read from a buffer of tempered words generated AHEAD of time (denet_pyrandom_ahead): the host steps the
// generator while the GPU is still running the forward trunk, so that the post-processing between the two graphs only
// converts words.  `over` is set when the buffer runs out (the caller then redoes the step on the live generator).
struct PyWords {
    const uint32_t* w;
    long long n, i;
    bool over;
    uint32_t next() {
        if (i >= n) {
            over = true;
            return 0;
        }
        return w[i++];
    }
};

template <class G>
static double py_random(G& g) {   // genrand_res53
    const uint32_t a = g.next() >> 5, b = g.next() >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}
template <class G>
static uint32_t py_randbelow(G& g, uint32_t n) {   // _randbelow_with_getrandbits, n >= 1 (n < 2^31)
    int k = 0;
    for (uint32_t v = n; v; v >>= 1) ++k;          // n.bit_length()
    uint32_t r = g.next() >> (32 - k);             // getrandbits(k), k <= 32
    int guard = 0;
    while (r >= n && ++guard < (1 << 20)) r = g.next() >> (32 - k);
    return r < n ? r : 0;
}
// random.sample(range(n), k) -> indices
template <class G>
static void py_sample(G& g, int n, int k, std::vector<int>& out) {
    out.resize(k);
    long long setsize = 21;
    if (k > 5) setsize += (long long)pow(4.0, ceil(log((double)k * 3.0) / log(4.0)));
    if (n <= setsize) {
        std::vector<int> pool(n);
        for (int i = 0; i < n; ++i) pool[i] = i;
        for (int i = 0; i < k; ++i) {
            const uint32_t j = py_randbelow(g, (uint32_t)(n - i));
            out[i] = pool[j];
            pool[j] = pool[n - i - 1];
        }
    } else {
        std::unordered_set<uint32_t> selected;
        for (int i = 0; i < k; ++i) {
            uint32_t j = py_randbelow(g, (uint32_t)n);
            int guard = 0;
            while (selected.count(j) && ++guard < (1 << 20)) j = py_randbelow(g, (uint32_t)n);
            selected.insert(j);
            out[i] = (int)j;
        }
    }
}

template <class G>
static void sparse_postprocess(G& g, const float* pr32, const float* bbox32, const long long* count, int B, int K,
                               int n_keep, double* pr, double* bbox) {
    std::vector<int> keep;
    for (int b = 0; b < B; ++b) {
        const float* p32 = pr32 + (size_t)b * K;
        const float* b32 = bbox32 + (size_t)b * K * 4;
        double* p = pr + (size_t)b * K;
        double* bb = bbox + (size_t)b * K * 4;
        int cnt = (int)std::min<long long>(std::max<long long>(count[b], 0), K);
        if (cnt > n_keep) {                                   // denet_sparse.py:184-187
            py_sample(g, cnt, n_keep, keep);
            for (int i = 0; i < n_keep; ++i) {
                p[i] = p32[keep[i]];
                for (int c = 0; c < 4; ++c) bb[i * 4 + c] = b32[keep[i] * 4 + c];
            }
            cnt = n_keep;
        } else {
            for (int i = 0; i < cnt; ++i) {
                p[i] = p32[i];
                for (int c = 0; c < 4; ++c) bb[i * 4 + c] = b32[i * 4 + c];
            }
        }
        for (int i = cnt; i < K; ++i) {                       // :190-196, random.uniform(a, b) = a + (b - a) * random()
            const double x0 = 0.0 + (1.0 - 0.0) * py_random(g);
            const double y0 = 0.0 + (1.0 - 0.0) * py_random(g);
            const double x1 = x0 + (1.0 - x0) * py_random(g);
            const double y1 = y0 + (1.0 - y0) * py_random(g);
            p[i] = 0.0;
            bb[i * 4 + 0] = x0; bb[i * 4 + 1] = y0; bb[i * 4 + 2] = x1; bb[i * 4 + 3] = y1;
        }
    }
}

}  // namespace dn

extern "C" int denet_pyrandom_sample(uint32_t* mt_state, int* mt_pos, int n, int k, int* out_index) {
    DN_REQUIRE(mt_state && mt_pos && out_index, "pyrandom_sample: null pointer");
    DN_REQUIRE(0 <= k && k <= n, "pyrandom_sample: sample larger than population or is negative");
    dn::PyMT g{mt_state, mt_pos};
    std::vector<int> idx;
    dn::py_sample(g, n, k, idx);
    memcpy(out_index, idx.data(), sizeof(int) * (size_t)k);
    return 0;
}

extern "C" int denet_pyrandom_random(uint32_t* mt_state, int* mt_pos, long long n, double* out) {
    DN_REQUIRE(mt_state && mt_pos && (out || n == 0), "pyrandom_random: null pointer");
    dn::PyMT g{mt_state, mt_pos};
    for (long long i = 0; i < n; ++i) out[i] = dn::py_random(g);
    return 0;
}

extern "C" int denet_sparse_postprocess(uint32_t* mt_state, int* mt_pos, const float* pr32, const float* bbox32,
                                        const long long* count, int B, int K, int n_keep, double* pr, double* bbox) {
    DN_REQUIRE(mt_state && mt_pos && pr32 && bbox32 && count && pr && bbox, "sparse_postprocess: null pointer");
    DN_REQUIRE(B >= 0 && K > 0 && n_keep >= 0 && n_keep <= K, "sparse_postprocess: bad sizes");
    dn::PyMT g{mt_state, mt_pos};
    dn::sparse_postprocess(g, pr32, bbox32, count, B, K, n_keep, pr, bbox);
    return 0;
}

// Steps a COPY of the interpreter's generator `nblocks` regenerations ahead: out_words receives the tempered words in
// stream order - the (624 - mt_pos) words left in the current block, then 624 per block - and out_states the raw
// 624-word state after each regeneration, so that the interpreter can be set to "u words consumed" afterwards
// (state of the block word u falls in, position inside it).  The interpreter's own state is not touched.
extern "C" int denet_pyrandom_ahead(const uint32_t* mt_state, int mt_pos, int nblocks, uint32_t* out_states,
                                    uint32_t* out_words, long long* out_nwords) {
    DN_REQUIRE(mt_state && out_states && out_words && out_nwords && nblocks >= 0 && mt_pos >= 0 && mt_pos <= 624,
               "pyrandom_ahead: bad arguments");
    uint32_t st[624];
    memcpy(st, mt_state, sizeof(st));
    int pos = mt_pos;
    dn::PyMT g{st, &pos};
    long long n = 0;
    for (int i = mt_pos; i < 624; ++i) out_words[n++] = g.next();
    for (int b = 0; b < nblocks; ++b) {
        out_words[n++] = g.next();                      // regenerates the block
        memcpy(out_states + (size_t)b * 624, st, sizeof(st));
        for (int i = 1; i < 624; ++i) out_words[n++] = g.next();
    }
    *out_nwords = n;
    return 0;
}

// denet_sparse_postprocess on words generated by denet_pyrandom_ahead; *used = words consumed.  Returns 1 (results
// invalid) when the buffer ran out: the caller repeats the call on the live generator.
extern "C" int denet_sparse_postprocess_ahead(const uint32_t* words, long long nwords, long long* used, const float* pr32,
                                              const float* bbox32, const long long* count, int B, int K, int n_keep,
                                              double* pr, double* bbox) {
    DN_REQUIRE(words && used && pr32 && bbox32 && count && pr && bbox, "sparse_postprocess_ahead: null pointer");
    DN_REQUIRE(B >= 0 && K > 0 && n_keep >= 0 && n_keep <= K, "sparse_postprocess_ahead: bad sizes");
    dn::PyWords g{words, nwords, 0, false};
    dn::sparse_postprocess(g, pr32, bbox32, count, B, K, n_keep, pr, bbox);
    *used = g.i;
    return g.over ? 1 : 0;
}
