// Multi-tensor Adam with coupled L2 weight decay in ONE launch (torch.optim.Adam(params, lr, weight_decay=wd), the
// optimiser net_run/get_optimizer.py:16-17 builds for every FPL+ .cfg):
//     g  = grad + wd * p
//     m  = m + (1 - b1) * (g - m)                    (torch: exp_avg.lerp_(grad, 1 - beta1))
//     v  = b2 * v + (1 - b2) * g * g
//     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// over a DEVICE table of tensors (parameter, gradient, exp_avg, exp_avg_sq, element count, step counter).  The learning
// rate is read from device memory when lr_dev != NULL, so the same launch can be replayed from a CUDA graph while
// MultiStepLR (get_optimizer.py:50-54) changes it; the per-tensor step counters (torch keeps one per parameter: a BN of
// the domain that was absent from a step is skipped and keeps its own bias correction) are bumped by the last block to
// finish.  HBM streaming: 16 B/element read + 12 B/element written.
#include <math.h>

#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

struct __align__(16) AdamSeg {      // one row of the device table (48 bytes); built by the host from tensor pointers
    float* p;
    const float* g;
    float* m;
    float* v;
    float* step;
    int numel;
    int pad;
};
static_assert(sizeof(AdamSeg) == 48, "AdamSeg layout is part of the C ABI (fpl_adam_multi_tensor)");

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 4096;    // elements per work item: 256 threads x 4 x float4

__global__ void __launch_bounds__(kAdamThreads) adam_multi_tensor_kernel(const AdamSeg* __restrict__ segs, int nseg,
                                                                         const int2* __restrict__ chunks, int nchunks,
                                                                         const float* __restrict__ lr_dev, double lr_host,
                                                                         double beta1, double beta2, float eps, float wd,
                                                                         unsigned int* done_counter) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    // torch evaluates 1 - beta, beta ** step and lr / bias_correction in Python doubles and hands the kernels the fp32
    // roundings of those: 1.0f - 0.999f differs from float(1 - 0.999) by 1.3e-5 relative
    const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2);
    __shared__ float s_step_size, s_bc2_sqrt;
    __shared__ int s_last;
    for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const int2 ch = __ldg(chunks + c);
        const AdamSeg sg = segs[ch.x];
        __syncthreads();
        if (threadIdx.x == 0) {
            const double t = (double)(*sg.step) + 1.0;
            const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
            const double lr = lr_dev != nullptr ? (double)__ldg(lr_dev) : lr_host;
            s_step_size = (float)(lr / bc1);
            s_bc2_sqrt = (float)sqrt(bc2);
        }
        __syncthreads();
        const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
        const int end = min(sg.numel, ch.y + kAdamChunk);
        const int n4_end = ch.y + ((end - ch.y) & ~3);
        for (int i = ch.y + 4 * threadIdx.x; i < n4_end; i += 4 * kAdamThreads) {
            float4 p = *reinterpret_cast<const float4*>(sg.p + i);
            const float4 g4 = ld_stream_f4(sg.g + i);
            float4 m = *reinterpret_cast<const float4*>(sg.m + i);
            float4 v = *reinterpret_cast<const float4*>(sg.v + i);
            float* pp = reinterpret_cast<float*>(&p);
            const float* gp = reinterpret_cast<const float*>(&g4);
            float* mp = reinterpret_cast<float*>(&m);
            float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float g = gp[k] + wd * pp[k];
                mp[k] = mp[k] + omb1 * (g - mp[k]);
                vp[k] = b2 * vp[k] + (omb2 * g) * g;
                const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
                pp[k] = pp[k] - step_size * (mp[k] / denom);
            }
            *reinterpret_cast<float4*>(sg.p + i) = p;
            *reinterpret_cast<float4*>(sg.m + i) = m;
            *reinterpret_cast<float4*>(sg.v + i) = v;
        }
        for (int i = n4_end + threadIdx.x; i < end; i += kAdamThreads) {      // tail of a tensor whose size is not 4k
            const float g = sg.g[i] + wd * sg.p[i];
            const float m = sg.m[i] + omb1 * (g - sg.m[i]);
            const float v = b2 * sg.v[i] + (omb2 * g) * g;
            sg.m[i] = m;
            sg.v[i] = v;
            sg.p[i] = sg.p[i] - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
        }
    }
    // the last block to finish bumps every tensor's step counter (all blocks have read theirs by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        for (int s = threadIdx.x; s < nseg; s += kAdamThreads) *segs[s].step += 1.0f;
        if (threadIdx.x == 0) *done_counter = 0u;
    }
}

}  // namespace

extern "C" int fpl_adam_chunk_elems(void) { return kAdamChunk; }

extern "C" int fpl_adam_multi_tensor(const void* d_segs, int nseg, const int* d_chunks, int nchunks, const float* lr_dev,
                                     double lr_host, double beta1, double beta2, double eps, double weight_decay,
                                     unsigned int* d_done_counter, void* stream) {
    FPL_REQUIRE(nseg >= 0 && nchunks >= 0, "fpl_adam_multi_tensor: negative counts");
    if (nseg == 0 || nchunks == 0) return 0;
    FPL_REQUIRE(d_segs != nullptr && d_chunks != nullptr && d_done_counter != nullptr, "fpl_adam_multi_tensor: NULL argument");
    FPL_REQUIRE((reinterpret_cast<uintptr_t>(d_segs) & 15) == 0, "fpl_adam_multi_tensor: table must be 16-byte aligned");
    FPL_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0,
                "fpl_adam_multi_tensor: betas (%g, %g) / eps %g out of range", beta1, beta2, eps);
    int grid = nchunks;
    if (grid > FPL_NUM_SMS * 8) grid = FPL_NUM_SMS * 8;
    fpl_launch(adam_multi_tensor_kernel, grid, kAdamThreads, 0, (cudaStream_t)stream, 
        reinterpret_cast<const AdamSeg*>(d_segs), nseg, reinterpret_cast<const int2*>(d_chunks), nchunks, lr_dev, lr_host,
        beta1, beta2, (float)eps, (float)weight_decay, d_done_counter);
    FPL_LAUNCH_CHECK();
    return 0;
}
