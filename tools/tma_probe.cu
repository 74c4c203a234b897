// Microbenchmark: steady-state cadence of TMA tile loads (cp.async.bulk.tensor, SWIZZLE_NONE, the C8-planar activation
// layout [N][D][C/8][H][W][8] bf16) issued by ONE producer thread per CTA into a ring of stages, as a function of the box
// shape (row bytes x rows x channel groups x planes), the ring depth, CTAs per SM and the working-set size (L2 resident or
// streamed from HBM).  Development tool behind the conv kernels' tile-shape choices; timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/tma_probe tools/tma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../fpl-plus_b200/csrc/tc_ptx.cuh"

struct TCfg {
    int box_w, box_h, box_c8, box_d;     // voxels per row, rows, channel groups, planes per box
    int stages, boxes;                   // ring depth, boxes per CTA
    int W, H, C8, D, N;                  // tensor extents
    int split;                           // 1: one box per stage; 2: the box of a stage is fetched as two half-height boxes
    int spin_warps;                      // extra warps polling an mbarrier that completes at the end (the epilogue warps of the conv kernels)
    int commit;                          // 1: the consumer frees slots with tcgen05.commit instead of mbarrier.arrive
    int warp_consumer;                   // 1: all 32 lanes of the consumer warp poll the full barrier
};

__global__ void __launch_bounds__(256) tma_probe_kernel(const __grid_constant__ CUtensorMap map, TCfg c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full_bar[16], empty_bar[16], end_bar;
    const uint32_t box_bytes = (uint32_t)(c.box_w * c.box_h * c.box_c8 * c.box_d * 16);
    if (threadIdx.x == 0) {
        for (int s = 0; s < c.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&end_bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int tiles_w = c.W / c.box_w, tiles_h = c.H / c.box_h, dsteps = c.D / c.box_d;
    if (threadIdx.x == 0) {                 // producer
        int stage = 0; uint32_t phase = 0;
        // consecutive boxes of a CTA walk the depth first (like the conv kernels); counters instead of divisions so that the
        // loop measures the TMA path and not the producer thread's index arithmetic
        int t = blockIdx.x * c.boxes;
        int dz = t % dsteps; t /= dsteps;
        int tw = t % tiles_w; t /= tiles_w;
        int th = t % tiles_h; t /= tiles_h;
        int n = t % c.N;
        for (int i = 0; i < c.boxes; ++i) {
            if (i > 0 && ++dz == dsteps) {
                dz = 0;
                if (++tw == tiles_w) { tw = 0; if (++th == tiles_h) { th = 0; if (++n == c.N) n = 0; } }
            }
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], box_bytes);
            uint8_t* dst = smem + (size_t)stage * box_bytes;
            if (c.split == 1) {
                tma_load_5d(dst, &map, &full_bar[stage], tw * c.box_w * 2, th * c.box_h, 0, dz * c.box_d, n);
            } else {     // two boxes of box_c8/2 channel groups each (the tensor map was encoded with the half box)
                tma_load_5d(dst, &map, &full_bar[stage], tw * c.box_w * 2, th * c.box_h, 0, dz * c.box_d, n);
                tma_load_5d(dst + box_bytes / 2, &map, &full_bar[stage], tw * c.box_w * 2, th * c.box_h, c.box_c8 / 2, dz * c.box_d, n);
            }
            if (++stage == c.stages) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x >= 32 && threadIdx.x < 64 && (c.warp_consumer || threadIdx.x == 32)) {
        // consumer: frees the slot as soon as the box has landed
        int stage = 0; uint32_t phase = 0;
        long long t0 = 0;
        const bool leader = threadIdx.x == 32;
        for (int i = 0; i < c.boxes; ++i) {
            mbar_wait(&full_bar[stage], phase);
            if (i == c.stages + 1) t0 = clock64();     // steady state: after the initial burst
            if (leader) {
                if (c.commit) umma_commit(&empty_bar[stage]);
                else mbar_arrive(&empty_bar[stage]);
            }
            if (c.warp_consumer) __syncwarp();
            if (++stage == c.stages) { stage = 0; phase ^= 1; }
        }
        if (leader) {
            out[blockIdx.x] = (clock64() - t0) / (c.boxes - c.stages - 2);
            mbar_arrive(&end_bar);
        }
    } else if (threadIdx.x >= 64 && threadIdx.x < 64 + 32 * c.spin_warps) {
        mbar_wait(&end_bar, 0);
    }
}

static double run(const TCfg& cin, int ctas_per_sm, void* d_x, long long* d_out, EncodeTiledFn encode) {
    TCfg c = cin;
    CUtensorMap map;
    cuuint64_t gdim[5] = {(cuuint64_t)c.W * 2, (cuuint64_t)c.H, (cuuint64_t)c.C8, (cuuint64_t)c.D, (cuuint64_t)c.N};
    cuuint64_t gstr[4] = {(cuuint64_t)c.W * 16, (cuuint64_t)c.H * c.W * 16, (cuuint64_t)c.C8 * c.H * c.W * 16,
                          (cuuint64_t)c.D * c.C8 * c.H * c.W * 16};
    cuuint32_t box[5] = {(cuuint32_t)c.box_w * 2, (cuuint32_t)c.box_h, (cuuint32_t)(c.box_c8 / c.split), (cuuint32_t)c.box_d, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, d_x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return -1; }
    const int grid = 148 * ctas_per_sm;
    const size_t smem = (size_t)c.stages * c.box_w * c.box_h * c.box_c8 * c.box_d * 16 + 1024;
    // pad the allocation so that exactly ctas_per_sm CTAs fit on an SM
    size_t want = ctas_per_sm == 1 ? 120 * 1024 : 100 * 1024;
    if (want < smem) want = smem;
    cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    double best = 1e30;
    for (int trial = 0; trial < 3; ++trial) {
        tma_probe_kernel<<<grid, 256, want>>>(map, c, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (long long v : h) mean += (double)v;
        mean /= grid;
        if (mean < best) best = mean;
    }
    return best;
}

int main() {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    // 16 channels x 4 x 32 x 128 x 128 voxels = 67 MB (the full-resolution 16-channel activation); small = 1 sample x 4 planes
    void* d_x;
    cudaMalloc(&d_x, (size_t)4 * 32 * 4 * 128 * 128 * 16);
    cudaMemset(d_x, 0, (size_t)4 * 32 * 4 * 128 * 128 * 16);
    long long* d_out;
    cudaMalloc(&d_out, 512 * sizeof(long long));
    printf("%-58s %10s %10s %10s\n", "box (voxels/row x rows x c8 x planes) | tensor | stages", "cyc/box", "B/clk/CTA", "clk/row");
    struct Row { const char* name; TCfg c; int ctas; };
    auto mk = [](int bw, int bh, int bc, int bd, int stages, int c8, int d, int n, int split) {
        TCfg c; c.box_w = bw; c.box_h = bh; c.box_c8 = bc; c.box_d = bd; c.stages = stages; c.boxes = 64;
        c.W = 128; c.H = 128; c.C8 = c8; c.D = d; c.N = n; c.split = split; c.spin_warps = 0; c.commit = 0; c.warp_consumer = 0; return c;
    };
    std::vector<Row> rows = {
        {"10x18x2x1 67MB s6 (dfold 16ch today)        2 CTA/SM", mk(10, 18, 2, 1, 6, 2, 32, 4, 1), 2},
        {"10x18x2x1 67MB s6                           1 CTA/SM", mk(10, 18, 2, 1, 6, 2, 32, 4, 1), 1},
        {"10x18x2x1 67MB s2                           2 CTA/SM", mk(10, 18, 2, 1, 2, 2, 32, 4, 1), 2},
        {"10x18x2x1 4MB (L2 resident) s6              2 CTA/SM", mk(10, 18, 2, 1, 6, 2, 2, 4, 1), 2},
        {"8x16x2x1  67MB s6 (aligned 128 B rows)      2 CTA/SM", mk(8, 16, 2, 1, 6, 2, 32, 4, 1), 2},
        {"8x16x2x1  4MB s6                            2 CTA/SM", mk(8, 16, 2, 1, 6, 2, 2, 4, 1), 2},
        {"16x16x2x1 67MB s6 (256 B rows)              2 CTA/SM", mk(16, 16, 2, 1, 6, 2, 32, 4, 1), 2},
        {"32x8x2x1  67MB s6 (512 B rows)              2 CTA/SM", mk(32, 8, 2, 1, 6, 2, 32, 4, 1), 2},
        {"64x4x2x1  67MB s6 (1 KB rows)               2 CTA/SM", mk(64, 4, 2, 1, 6, 2, 32, 4, 1), 2},
        {"128x2x2x1 67MB s6 (2 KB rows)               2 CTA/SM", mk(128, 2, 2, 1, 6, 2, 32, 4, 1), 2},
        {"8x16x2x2  67MB s4 (2 planes per box)        2 CTA/SM", mk(8, 16, 2, 2, 4, 2, 32, 4, 1), 2},
        {"8x16x2x4  67MB s3 (4 planes per box)        2 CTA/SM", mk(8, 16, 2, 4, 3, 2, 32, 4, 1), 2},
        {"32x8x2x4  67MB s2 (wgrad-like 32 KB box)    2 CTA/SM", mk(32, 8, 2, 4, 2, 2, 32, 4, 1), 2},
        {"32x8x2x4  67MB s4                           1 CTA/SM", mk(32, 8, 2, 4, 4, 2, 32, 4, 1), 1},
        {"8x16x2x1  67MB s6, two 1-c8 boxes per stage 2 CTA/SM", mk(8, 16, 2, 1, 6, 2, 32, 4, 2), 2},
        {"8x8x2x1   67MB s6 (half height)             2 CTA/SM", mk(8, 8, 2, 1, 6, 2, 32, 4, 1), 2},
        {"8x32x2x1  67MB s6 (double height)           2 CTA/SM", mk(8, 32, 2, 1, 6, 2, 32, 4, 1), 2},
        {"8x16x1x1  67MB s6 (one c8 group)            2 CTA/SM", mk(8, 16, 1, 1, 6, 2, 32, 4, 1), 2},
    };
    {
        TCfg c = mk(10, 18, 2, 1, 6, 2, 32, 4, 1);
        c.spin_warps = 4; rows.push_back({"10x18x2x1 67MB s6 + 4 warps polling an mbarrier      2 CTA/SM", c, 2});
        c.spin_warps = 0; c.commit = 1; rows.push_back({"10x18x2x1 67MB s6, slots freed by tcgen05.commit     2 CTA/SM", c, 2});
        c.commit = 0; c.warp_consumer = 1; rows.push_back({"10x18x2x1 67MB s6, 32-lane consumer polling          2 CTA/SM", c, 2});
        c.commit = 1; c.spin_warps = 4; rows.push_back({"10x18x2x1 67MB s6, all three (the dfold structure)   2 CTA/SM", c, 2});
    }
    for (const Row& r : rows) {
        const double cyc = run(r.c, r.ctas, d_x, d_out, encode);
        const double bytes = (double)r.c.box_w * r.c.box_h * r.c.box_c8 * r.c.box_d * 16;
        printf("%-58s %10.0f %10.1f %10.1f\n", r.name, cyc, bytes / cyc, cyc / (r.c.box_h * r.c.box_c8 * r.c.box_d));
    }
    return 0;
}
