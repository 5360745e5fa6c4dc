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
    double random() {   // genrand_res53
        const uint32_t a = next() >> 5, b = next() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    uint32_t randbelow(uint32_t n) {   // _randbelow_with_getrandbits, n >= 1 (n < 2^31)
        int k = 0;
        for (uint32_t v = n; v; v >>= 1) ++k;          // n.bit_length()
        uint32_t r = next() >> (32 - k);               // getrandbits(k), k <= 32
        while (r >= n) r = next() >> (32 - k);
        return r;
    }
    // random.sample(range(n), k) -> indices
    void sample(int n, int k, std::vector<int>& out) {
        out.resize(k);
        long long setsize = 21;
        if (k > 5) setsize += (long long)pow(4.0, ceil(log((double)k * 3.0) / log(4.0)));
        if (n <= setsize) {
            std::vector<int> pool(n);
            for (int i = 0; i < n; ++i) pool[i] = i;
            for (int i = 0; i < k; ++i) {
                const uint32_t j = randbelow((uint32_t)(n - i));
                out[i] = pool[j];
                pool[j] = pool[n - i - 1];
            }
        } else {
            std::unordered_set<uint32_t> selected;
            for (int i = 0; i < k; ++i) {
                uint32_t j = randbelow((uint32_t)n);
                while (selected.count(j)) j = randbelow((uint32_t)n);
                selected.insert(j);
                out[i] = (int)j;
            }
        }
    }
};

}  // namespace dn

extern "C" int denet_pyrandom_sample(uint32_t* mt_state, int* mt_pos, int n, int k, int* out_index) {
    DN_REQUIRE(mt_state && mt_pos && out_index, "pyrandom_sample: null pointer");
    DN_REQUIRE(0 <= k && k <= n, "pyrandom_sample: sample larger than population or is negative");
    dn::PyMT g{mt_state, mt_pos};
    std::vector<int> idx;
    g.sample(n, k, idx);
    memcpy(out_index, idx.data(), sizeof(int) * (size_t)k);
    return 0;
}

extern "C" int denet_pyrandom_random(uint32_t* mt_state, int* mt_pos, long long n, double* out) {
    DN_REQUIRE(mt_state && mt_pos && (out || n == 0), "pyrandom_random: null pointer");
    dn::PyMT g{mt_state, mt_pos};
    for (long long i = 0; i < n; ++i) out[i] = g.random();
    return 0;
}

extern "C" int denet_sparse_postprocess(uint32_t* mt_state, int* mt_pos, const float* pr32, const float* bbox32,
                                        const long long* count, int B, int K, int n_keep, double* pr, double* bbox) {
    DN_REQUIRE(mt_state && mt_pos && pr32 && bbox32 && count && pr && bbox, "sparse_postprocess: null pointer");
    DN_REQUIRE(B >= 0 && K > 0 && n_keep >= 0 && n_keep <= K, "sparse_postprocess: bad sizes");
    dn::PyMT g{mt_state, mt_pos};
    std::vector<int> keep;
    for (int b = 0; b < B; ++b) {
        const float* p32 = pr32 + (size_t)b * K;
        const float* b32 = bbox32 + (size_t)b * K * 4;
        double* p = pr + (size_t)b * K;
        double* bb = bbox + (size_t)b * K * 4;
        int cnt = (int)std::min<long long>(std::max<long long>(count[b], 0), K);
        if (cnt > n_keep) {                                   // denet_sparse.py:184-187
            g.sample(cnt, n_keep, keep);
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
            const double x0 = 0.0 + (1.0 - 0.0) * g.random();
            const double y0 = 0.0 + (1.0 - 0.0) * g.random();
            const double x1 = x0 + (1.0 - x0) * g.random();
            const double y1 = y0 + (1.0 - y0) * g.random();
            p[i] = 0.0;
            bb[i * 4 + 0] = x0; bb[i * 4 + 1] = y0; bb[i * 4 + 2] = x1; bb[i * 4 + 3] = y1;
        }
    }
    return 0;
}
