// Gradient accumulation across the domain passes of one optimiser step (agent_seg.py:459-495: the source and the target
// batch are two forward/backward passes through the SAME weights, summed by autograd's AccumulateGrad -- ~65 tiny add
// kernels per step).  Every backward pass of fplplus_b200 writes its gradients into one flat fp32 buffer; this kernel
// adds the segments of that buffer into the network's persistent master gradient buffer (the storage behind p.grad)
// in ONE launch.  The adds are atomics: the two passes run on two streams.
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

// table[3*s + {0,1,2}] = {source offset, destination offset, element count} (all multiples of 4 floats except the count)
__global__ void __launch_bounds__(256) grad_scatter_add_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                               const int* __restrict__ table) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int s = blockIdx.y;
    const int so = __ldg(table + 3 * s), dof = __ldg(table + 3 * s + 1), n = __ldg(table + 3 * s + 2);
    const float* sp = src + so;
    float* dp = dst + dof;
    const int n4 = n >> 2;
    // batches of 4 loads per thread before the first reduction is issued (one load -> red per iteration made the largest
    // segments, 1.8 M floats on 64 blocks, 27 memory round trips long)
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k * stride < n4) v[k] = ld_stream_f4(sp + 4 * (int64_t)(i0 + k * stride));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k * stride < n4) {
                float* d = dp + 4 * (int64_t)(i0 + k * stride);
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v[k].x), "f"(v[k].y), "f"(v[k].z), "f"(v[k].w) : "memory");
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) atomicAdd(dp + 4 * n4 + threadIdx.x, __ldg(sp + 4 * n4 + threadIdx.x));
}

}  // namespace

extern "C" int fpl_grad_scatter_add(float* dst, const float* src, const int* d_table, int segments, int max_numel, void* stream) {
    FPL_REQUIRE(segments >= 0 && segments <= 65535, "fpl_grad_scatter_add: %d segments out of range", segments);
    if (segments == 0) return 0;
    FPL_REQUIRE(dst != nullptr && src != nullptr && d_table != nullptr, "fpl_grad_scatter_add: NULL argument");
    FPL_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0,
                "fpl_grad_scatter_add: buffers must be 16-byte aligned");
    int bx = (max_numel / 4 + 1023) / 1024;
    if (bx < 1) bx = 1;
    if (bx > 128) bx = 128;
    fpl_launch(grad_scatter_add_kernel, dim3(bx, segments), 256, 0, (cudaStream_t)stream, dst, src, d_table);
    FPL_LAUNCH_CHECK();
    return 0;
}
