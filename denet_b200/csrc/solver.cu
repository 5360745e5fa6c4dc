// Parameter update rules of ModelCNN.build_train_func (reference denet/model/model_cnn.py:282-305, 320-324) as ONE
// multi-tensor kernel launch per step: every CTA owns a fixed 4096-element chunk of one parameter tensor, looked up
// from a device-resident table, so ~150 per-tensor elementwise launches collapse into one HBM-bound pass.
//   sgd              m <- rho*m + (1-rho)*g ; p <- p - lr*m
//   torch / nesterov m <- rho*m + g         ; p <- p - lr*(g + mu*m)
//   adam             m,v moments with bias correction
// rho = mu for iteration > 0 and 0 at iteration 0 (:284,291);  g <- g + decay*p for weights only (:322-324).
#include "common.cuh"

namespace dn {

struct SolverEntry {
    float* p;
    const float* g;
    float* m;
    float* v;
    long long n;
    int is_weight;
    int pad;
};

constexpr int kChunk = 4096;

__global__ void __launch_bounds__(256) solver_update_kernel(const SolverEntry* __restrict__ entries,
                                                             const int* __restrict__ block_tensor,
                                                             const long long* __restrict__ block_offset, int solver,
                                                             float lr, float mu, float mu2, float rho, float decay,
                                                             int iteration, int bias_decay, float grad_scale,
                                                             const float* __restrict__ hp) {
    if (hp) {   // hyper-parameters of THIS step from device memory: a captured CUDA graph replays with new values
        lr = hp[0];
        mu = hp[1];
        mu2 = hp[2];
        decay = hp[3];
        iteration = (int)hp[4];
        grad_scale = hp[5];
        rho = iteration > 0 ? mu : 0.f;
    }
    const SolverEntry e = entries[block_tensor[blockIdx.x]];
    const long long start = block_offset[blockIdx.x];
    const long long end = start + kChunk < e.n ? start + kChunk : e.n;
    const float wd = (e.is_weight || bias_decay) ? decay : 0.f;
    float c1 = 1.f, c2 = 1.f;
    if (solver == 2) {
        c1 = 1.0f / (1.0f - powf(mu, (float)(iteration + 1)));
        c2 = 1.0f / (1.0f - powf(mu2, (float)(iteration + 1)));
    }
    for (long long i = start + threadIdx.x; i < end; i += blockDim.x) {
        float p = e.p[i];
        const float g = e.g[i] * grad_scale + wd * p;
        float m = e.m[i];
        if (solver == 1) {
            m = rho * m + g;
            p = p - lr * (g + mu * m);
        } else if (solver == 2) {
            m = mu * m + (1.0f - mu) * g;
            const float v = mu2 * e.v[i] + (1.0f - mu2) * (g * g);
            e.v[i] = v;
            p = p - lr * (m * c1) / (sqrtf(v * c2) + 1e-8f);
        } else {
            m = rho * m + (1.0f - rho) * g;
            p = p - lr * m;
        }
        e.m[i] = m;
        e.p[i] = p;
    }
}

}  // namespace dn

using namespace dn;

extern "C" int denet_solver_entry_bytes(void) { return (int)sizeof(SolverEntry); }
extern "C" int denet_solver_chunk(void) { return kChunk; }

extern "C" int denet_solver_update(const void* entries, const int* block_tensor, const long long* block_offset,
                                   int nblocks, int solver, float lr, float momentum0, float momentum1, float decay,
                                   int iteration, int bias_decay, float grad_scale, cudaStream_t stream) {
    DN_REQUIRE(entries && block_tensor && block_offset, "solver_update: null pointer");
    DN_REQUIRE(solver >= 0 && solver <= 2, "solver_update: solver must be 0 (sgd), 1 (nesterov/torch) or 2 (adam)");
    if (nblocks == 0) return 0;
    const float rho = iteration > 0 ? momentum0 : 0.0f;
    solver_update_kernel<<<DN_G(nblocks), 256, 0, stream>>>((const SolverEntry*)entries, block_tensor, block_offset, solver, lr,
                                                      momentum0, momentum1, rho, decay, iteration, bias_decay,
                                                      grad_scale, nullptr);
    DN_CHECK_LAUNCH();
    return 0;
}

extern "C" int denet_solver_update_dev(const void* entries, const int* block_tensor, const long long* block_offset,
                                       int nblocks, int solver, const float* hp, int bias_decay, cudaStream_t stream) {
    DN_REQUIRE(entries && block_tensor && block_offset && hp, "solver_update_dev: null pointer");
    DN_REQUIRE(solver >= 0 && solver <= 2, "solver_update_dev: solver must be 0 (sgd), 1 (nesterov/torch) or 2 (adam)");
    if (nblocks == 0) return 0;
    solver_update_kernel<<<DN_G(nblocks), 256, 0, stream>>>((const SolverEntry*)entries, block_tensor, block_offset, solver,
                                                            0.f, 0.f, 0.f, 0.f, 0.f, 0, bias_decay, 1.f, hp);
    DN_CHECK_LAUNCH();
    return 0;
}
