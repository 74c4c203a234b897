// Microbenchmark: SM cycles per tcgen05.mma (kind::f16, bf16 -> fp32, both operands in shared memory) as a function of
// the tile shape (M, N), the operand layout (MN-major no-swizzle = the C8-planar activation tiles of the wgrad kernel,
// K-major no-swizzle = the fwd / dgrad kernels, K-major 128-byte swizzle) and the alignment of the operand start address
// and strides.  Development tool behind DESIGN.md's "operand feed" model; timing only (operands are arbitrary bytes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_build/umma_feed_probe tools/umma_feed_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../fpl-plus_b200/csrc/tc_ptx.cuh"

struct Cfg {
    const char* name;
    int m, n;
    int a_mn, b_mn;            // 1 = MN-major operand, 0 = K-major
    int swz;                   // 0 none, 2 = 128-byte swizzle (descriptor layout type)
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_off[9];         // start-address byte offsets of the nmma A operands of one step (the tap shifts)
    uint32_t b_off;
    int nmma;                  // MMAs per step (<= 9), each into its own accumulator of n columns (wrapping inside 512)
    uint32_t a_step, b_step;   // byte advance of both operands per step (wraps inside the window)
    int steps_wrap;
    int commit_every;          // > 0: a tcgen05.commit (to a barrier nobody waits on) after every commit_every steps
};

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, int swz) {
    uint64_t d = make_desc(saddr, lbo, sbo);
    d |= (uint64_t)(swz & 7) << 61;
    return d;
}

template <int NM>
__global__ void __launch_bounds__(128) probe_kernel(Cfg c, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, side_bar;
    __shared__ uint32_t tmem_slot;
    // small finite bf16 values (0x3c00 + noise)
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + ((uint32_t)i * 2654435761u & 0x007f007fu);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&side_bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32) {      // the whole warp runs the loop (descriptors stay in uniform registers), one lane issues
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)c.a_mn << 15) | ((uint32_t)c.b_mn << 16) |
                               ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(c.m >> 4) << 24);
        const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem) + 128 * 1024;
        const int nacc = 512 / c.n < NM ? 512 / c.n : NM;
        // descriptors of step 0 in registers; a step adds (bytes >> 4) to the 14-bit start-address field
        uint64_t ad[NM];
        uint32_t acc[NM];
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            ad[i] = desc_sw(a_base + c.a_off[i], c.a_lbo, c.a_sbo, c.swz);
            acc[i] = tmem + (uint32_t)((i % nacc) * c.n);
        }
        const uint64_t bd = desc_sw(b_base + c.b_off, c.b_lbo, c.b_sbo, c.swz);
        const uint64_t a_inc = c.a_step >> 4, b_inc = c.b_step >> 4;
        const int wrap = c.steps_wrap;
        uint32_t phase = 0;
        long long best = 1ll << 62;
        for (int trial = 0; trial < 3; ++trial) {
            const long long t0 = clock64();
            int s = 0, since = 0;
            const int every = c.commit_every > 0 ? c.commit_every : 0x7fffffff;
            uint64_t a_add = 0, b_add = 0;
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int i = 0; i < NM; ++i)
                    if (leader) umma_bf16(acc[i], ad[i] + a_add, bd + b_add, idesc, 1);
                a_add += a_inc; b_add += b_inc;
                if (++since == every) {
                    since = 0;
                    if (leader) umma_commit(&side_bar);
                    __syncwarp();
                }
                if (++s == wrap) { s = 0; a_add = 0; b_add = 0; }
            }
            if (leader) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, phase);
            phase ^= 1;
            const long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        if (leader) out[blockIdx.x] = best;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NM>
static void launch(const Cfg& c, int grid, int reps, long long* d_out) {
    cudaFuncSetAttribute(probe_kernel<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
    probe_kernel<NM><<<grid, 128, 201 * 1024 + 1024>>>(c, reps, d_out);
}

static double run(const Cfg& c, int grid, int reps, long long* d_out) {
    switch (c.nmma) {
        case 1: launch<1>(c, grid, reps, d_out); break;
        case 2: launch<2>(c, grid, reps, d_out); break;
        case 3: launch<3>(c, grid, reps, d_out); break;
        case 4: launch<4>(c, grid, reps, d_out); break;
        case 5: launch<5>(c, grid, reps, d_out); break;
        case 6: launch<6>(c, grid, reps, d_out); break;
        case 8: launch<8>(c, grid, reps, d_out); break;
        default: launch<9>(c, grid, reps, d_out); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA error %s\n", c.name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long v : h) mx = v > mx ? v : mx;
    return (double)mx / ((double)reps * c.nmma);
}

static Cfg mn_cfg(const char* name, int m, int n, int tw, int th, int nmma, bool aligned, bool shifted) {
    // the wgrad tile: halo'd x planes of (th+2) x (tw+2) voxels x 16 B, dy planes th x tw x 16 B; K step = two 8-voxel
    // row segments one row apart (LBO = row pitch), M / N groups at the plane stride (SBO)
    Cfg c = {};
    c.name = name; c.m = m; c.n = n; c.a_mn = 1; c.b_mn = 1; c.swz = 0;
    uint32_t pitch_x = (uint32_t)(tw + 2) * 16, plane_x = (uint32_t)(th + 2) * pitch_x;
    if (aligned) { pitch_x = (pitch_x + 127) & ~127u; plane_x = ((uint32_t)(th + 2) * pitch_x + 127) & ~127u; }
    c.a_lbo = pitch_x; c.a_sbo = plane_x;
    c.b_lbo = (uint32_t)tw * 16; c.b_sbo = (uint32_t)th * tw * 16;
    for (int i = 0; i < 9; ++i) c.a_off[i] = shifted ? (uint32_t)(i / 3) * pitch_x + (uint32_t)(i % 3) * 16 : 0;
    c.nmma = nmma;
    c.a_step = 128; c.b_step = 128; c.steps_wrap = tw / 8;
    // keep the M / N windows inside the 128 KB / 72 KB regions
    while ((uint32_t)(m / 8) * c.a_sbo + 4 * pitch_x > 120 * 1024) c.a_sbo -= 128 * 8;
    while ((uint32_t)(n / 8) * c.b_sbo > 64 * 1024) c.b_sbo -= 128 * 8;
    return c;
}

static Cfg k_cfg(const char* name, int m, int n, int swz, int nmma, uint32_t a_misalign) {
    Cfg c = {};
    c.name = name; c.m = m; c.n = n; c.a_mn = 0; c.b_mn = 0; c.swz = swz;
    if (swz == 0) {      // core matrices (8 rows x 16 B) contiguous along M; K groups M*16 bytes apart
        c.a_lbo = (uint32_t)m * 16; c.a_sbo = 128; c.b_lbo = (uint32_t)n * 16; c.b_sbo = 128;
        c.a_step = (uint32_t)m * 32; c.b_step = (uint32_t)n * 32; c.steps_wrap = 4;
    } else {             // 128-byte rows (64 bf16 of K), 8-row atoms 1024 bytes apart; a K step advances 32 bytes
        c.a_lbo = 16; c.a_sbo = 1024; c.b_lbo = 16; c.b_sbo = 1024;
        c.a_step = 32; c.b_step = 32; c.steps_wrap = 4;
    }
    for (int i = 0; i < 9; ++i) c.a_off[i] = a_misalign * (uint32_t)i;
    c.nmma = nmma;
    return c;
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 256 * sizeof(long long));
    std::vector<Cfg> cfgs;
    // ---- MN-major no-swizzle (wgrad): shapes of today's kernel and of the candidates
    cfgs.push_back(mn_cfg("MN nosw M64  N32  (16->16 today) shifted", 64, 32, 32, 8, 9, false, true));
    cfgs.push_back(mn_cfg("MN nosw M64  N32  unshifted", 64, 32, 32, 8, 9, false, false));
    cfgs.push_back(mn_cfg("MN nosw M64  N32  128B-aligned strides, unshifted", 64, 32, 32, 8, 9, true, false));
    cfgs.push_back(mn_cfg("MN nosw M64  N32  128B-aligned strides, shifted", 64, 32, 32, 8, 9, true, true));
    cfgs.push_back(mn_cfg("MN nosw M128 N32  (32->16 today) shifted", 128, 32, 16, 8, 9, false, true));
    cfgs.push_back(mn_cfg("MN nosw M128 N32  unshifted", 128, 32, 16, 8, 9, false, false));
    cfgs.push_back(mn_cfg("MN nosw M128 N16  shifted", 128, 16, 16, 8, 9, false, true));
    cfgs.push_back(mn_cfg("MN nosw M128 N64  shifted", 128, 64, 16, 8, 6, false, true));
    cfgs.push_back(mn_cfg("MN nosw M128 N64  unshifted", 128, 64, 16, 8, 6, false, false));
    cfgs.push_back(mn_cfg("MN nosw M128 N64  aligned, unshifted", 128, 64, 16, 8, 6, true, false));
    cfgs.push_back(mn_cfg("MN nosw M128 N96  shifted", 128, 96, 16, 8, 5, false, true));
    cfgs.push_back(mn_cfg("MN nosw M128 N128 shifted", 128, 128, 16, 8, 3, false, true));
    cfgs.push_back(mn_cfg("MN nosw M128 N128 aligned, unshifted", 128, 128, 16, 8, 3, true, false));
    cfgs.push_back(mn_cfg("MN nosw M128 N256 shifted", 128, 256, 8, 8, 2, false, true));
    cfgs.push_back(mn_cfg("MN nosw M64  N64  shifted", 64, 64, 16, 8, 6, false, true));
    cfgs.push_back(mn_cfg("MN nosw M64  N128 shifted", 64, 128, 16, 8, 3, false, true));
    // ---- K-major (fwd / dgrad)
    for (int n : {16, 32, 48, 64, 96, 128, 256}) {
        static char names[3][8][64];
        static int k = 0;
        snprintf(names[0][k], 64, "K  nosw M128 N%-3d", n);
        cfgs.push_back(k_cfg(names[0][k], 128, n, 0, 512 / n < 9 ? 512 / n : 9, 0));
        snprintf(names[1][k], 64, "K  sw128 M128 N%-3d", n);
        cfgs.push_back(k_cfg(names[1][k], 128, n, 2, 512 / n < 9 ? 512 / n : 9, 0));
        ++k;
    }
    cfgs.push_back(k_cfg("K  nosw M128 N48  A start +16 B per tap", 128, 48, 0, 9, 16));
    cfgs.push_back(k_cfg("K  nosw M64  N64", 64, 64, 0, 8, 0));
    cfgs.push_back(k_cfg("K  sw128 M64  N64", 64, 64, 2, 8, 0));

    {   // the exact operand geometry of conv3d_wgrad_hs_kernel (Cin 16, tw 32): slabs one row pitch apart, K along the row
        auto hs = [](const char* name, uint32_t sx, uint32_t sdy, uint32_t lbo) {
            Cfg c = {};
            c.name = name; c.m = 128; c.n = 64; c.a_mn = 1; c.b_mn = 1; c.swz = 0;
            c.a_lbo = lbo; c.a_sbo = sx; c.b_lbo = lbo; c.b_sbo = sdy;
            for (int i = 0; i < 6; ++i) c.a_off[i] = (uint32_t)(i / 2) * 16 + (uint32_t)(i % 2) * 2 * 8 * sx;
            c.nmma = 6; c.a_step = 256; c.b_step = 256; c.steps_wrap = 2;
            return c;
        };
        cfgs.push_back(hs("hs  M128 N64  sx 544 sdy 512 (kernel)", 544, 512, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 576 sdy 512", 576, 512, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 544 sdy 528", 544, 528, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 576 sdy 576", 576, 576, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 560 sdy 560", 560, 560, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 640 sdy 640", 640, 640, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 288 sdy 256 (tw 16)", 288, 256, 128));
        cfgs.push_back(hs("hs  M128 N64  sx 272 sdy 272", 272, 272, 128));
    }
    for (int nm : {1, 3, 9}) {
        for (int every : {0, 1, 2, 4}) {
            static char nm2[24][64];
            static int k2 = 0;
            snprintf(nm2[k2], 64, "K  nosw M128 N48  %d MMAs/step, commit every %d", nm, every);
            Cfg c = k_cfg(nm2[k2], 128, 48, 0, nm, 0);
            c.commit_every = every;
            cfgs.push_back(c);
            ++k2;
        }
    }
    printf("%-52s %9s %9s %8s %9s %8s\n", "config (148 CTAs | 1 CTA)", "cyc/MMA", "1-CTA", "A+B KB", "B/clk", "math cyc");
    for (const Cfg& c : cfgs) {
        const double cyc = run(c, 148, 4000, d_out);
        const double cyc1 = run(c, 1, 4000, d_out);
        const double kb = (c.m + c.n) * 32.0 / 1024.0;
        printf("%-52s %9.1f %9.1f %8.1f %9.1f %8.1f\n", c.name, cyc, cyc1, kb, kb * 1024.0 / cyc, 128 * c.n / 256.0);
    }
    return 0;
}
