// nn.Upsample(scale_factor=2, mode='trilinear' | 'bilinear', align_corners=True) of UpBlock in `bilinear=True` mode
// (PyMIC/pymic/net/net3d/unet2d5_dsbn.py:149-150, 170-176) on C8-planar bf16, forward and backward.
//   source index = dst * (in - 1) / (out - 1)  (fp32, as ATen's area_pixel_compute_source_index with align_corners),
//   i0 = floor, lambda = src - i0, i1 = i0 + (i0 < in - 1).
// kd2 = 2: depth is interpolated too (trilinear); kd2 = 1: in-plane only (the 2-D blocks of the 2.5-D network).
// Forward writes straight into the second half of the concat buffer (y_c8tot / y_c8off).  Backward is a GATHER over the
// output voxels that read an input voxel (deterministic, no atomics); `bilinear=True` is not on the benchmarked path
// (every shipped .cfg uses ConvTranspose), so these kernels favour simplicity over bandwidth.
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

constexpr int kThreadsU = 256;

struct UpParams {
    const bf16x8* x;         // forward: low-res input; backward: gradient wrt the up-sampled tensor
    int x_c8tot, x_c8off;
    bf16x8* y;               // forward: up-sampled output; backward: gradient wrt the low-res input
    int y_c8tot, y_c8off;
    int N, D, H, W, C8;      // LOW-resolution geometry
    int kd2;
    float sd, sh, sw;        // (in - 1) / (out - 1) per axis (0 when out == 1)
};

__device__ __forceinline__ void src_index(float scale, int dst, int in, int& i0, int& i1, float& l1) {
    const float s = scale * (float)dst;
    i0 = (int)s;
    if (i0 > in - 1) i0 = in - 1;
    l1 = s - (float)i0;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
}

// one thread per OUTPUT vector
__global__ void __launch_bounds__(kThreadsU) upsample2x_fwd_kernel(const UpParams P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int Do = P.D * P.kd2, Ho = 2 * P.H, Wo = 2 * P.W;
    const int64_t total = (int64_t)P.N * Do * P.C8 * Ho * Wo;
    const int64_t HW = (int64_t)P.H * P.W, HWo = (int64_t)Ho * Wo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = i;
        const int wo = (int)(t % Wo); t /= Wo;
        const int ho = (int)(t % Ho); t /= Ho;
        const int c8 = (int)(t % P.C8); t /= P.C8;
        const int dout = (int)(t % Do);
        const int n = (int)(t / Do);
        int d0 = dout, d1 = dout, h0, h1, w0, w1;
        float ld = 0.0f, lh, lw;
        if (P.kd2 == 2) src_index(P.sd, dout, P.D, d0, d1, ld);
        src_index(P.sh, ho, P.H, h0, h1, lh);
        src_index(P.sw, wo, P.W, w0, w1, lw);
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dd = (q & 4) ? d1 : d0, hh = (q & 2) ? h1 : h0, ww = (q & 1) ? w1 : w0;
            const float wt = ((q & 4) ? ld : 1.0f - ld) * ((q & 2) ? lh : 1.0f - lh) * ((q & 1) ? lw : 1.0f - lw);
            if (wt == 0.0f) continue;
            float f[8];
            bf16x8_to_float(ldg_bf16x8(P.x + (((int64_t)n * P.D + dd) * P.x_c8tot + P.x_c8off + c8) * HW + (int64_t)hh * P.W + ww), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(wt, f[k], acc[k]);
        }
        st_bf16x8(P.y + (((int64_t)n * Do + dout) * P.y_c8tot + P.y_c8off + c8) * HWo + (int64_t)ho * Wo + wo, acc);
    }
}

// weights with which output index o reads input index i along one axis (0 when it does not)
__device__ __forceinline__ float axis_weight(float scale, int o, int in, int i) {
    int i0, i1;
    float l1;
    src_index(scale, o, in, i0, i1, l1);
    return (i0 == i ? 1.0f - l1 : 0.0f) + (i1 == i ? l1 : 0.0f);
}

// one thread per INPUT (low-res) vector: gathers from the <= 4 output indices per axis that read it
__global__ void __launch_bounds__(kThreadsU) upsample2x_bwd_kernel(const UpParams P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int Do = P.D * P.kd2, Ho = 2 * P.H, Wo = 2 * P.W;
    const int64_t total = (int64_t)P.N * P.D * P.C8 * P.H * P.W;
    const int64_t HW = (int64_t)P.H * P.W, HWo = (int64_t)Ho * Wo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t t = i;
        const int w = (int)(t % P.W); t /= P.W;
        const int h = (int)(t % P.H); t /= P.H;
        const int c8 = (int)(t % P.C8); t /= P.C8;
        const int d = (int)(t % P.D);
        const int n = (int)(t / P.D);
        // candidate output indices: with scale (in-1)/(2in-1) just under 1/2, the readers of i lie in [2i-2, 2i+2]
        float wd[5], wh[5], ww[5];
        int od[5], oh[5], ow[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            od[k] = P.kd2 == 2 ? 2 * d - 2 + k : d;
            wd[k] = P.kd2 == 2 ? ((od[k] >= 0 && od[k] < Do) ? axis_weight(P.sd, od[k], P.D, d) : 0.0f) : (k == 0 ? 1.0f : 0.0f);
            oh[k] = 2 * h - 2 + k;
            wh[k] = (oh[k] >= 0 && oh[k] < Ho) ? axis_weight(P.sh, oh[k], P.H, h) : 0.0f;
            ow[k] = 2 * w - 2 + k;
            ww[k] = (ow[k] >= 0 && ow[k] < Wo) ? axis_weight(P.sw, ow[k], P.W, w) : 0.0f;
        }
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        for (int a = 0; a < 5; ++a) {
            if (wd[a] == 0.0f) continue;
            for (int b = 0; b < 5; ++b) {
                if (wh[b] == 0.0f) continue;
                const float wab = wd[a] * wh[b];
                const bf16x8* row = P.x + (((int64_t)n * Do + od[a]) * P.x_c8tot + P.x_c8off + c8) * HWo + (int64_t)oh[b] * Wo;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    if (ww[c] == 0.0f) continue;
                    const float wt = wab * ww[c];
                    float f[8];
                    bf16x8_to_float(ldg_bf16x8(row + ow[c]), f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] = fmaf(wt, f[k], acc[k]);
                }
            }
        }
        st_bf16x8(P.y + (((int64_t)n * P.D + d) * P.y_c8tot + P.y_c8off + c8) * HW + (int64_t)h * P.W + w, acc);
    }
}

int fill(UpParams& P, const void* x, int x_c8tot, int x_c8off, void* y, int y_c8tot, int y_c8off, int n, int d, int h, int w,
         int c, int kd2) {
    FPL_REQUIRE(c > 0 && c % 8 == 0, "fpl_upsample2x: channels (%d) must be a multiple of 8", c);
    FPL_REQUIRE(kd2 == 1 || kd2 == 2, "fpl_upsample2x: kd2=%d must be 1 or 2", kd2);
    FPL_REQUIRE(x != nullptr && y != nullptr && n > 0 && d > 0 && h > 0 && w > 0, "fpl_upsample2x: bad arguments");
    P.x = (const bf16x8*)x; P.x_c8tot = x_c8tot; P.x_c8off = x_c8off; P.y = (bf16x8*)y; P.y_c8tot = y_c8tot; P.y_c8off = y_c8off;
    P.N = n; P.D = d; P.H = h; P.W = w; P.C8 = c / 8; P.kd2 = kd2;
    P.sd = (kd2 == 2 && d > 1) ? (float)(d - 1) / (float)(2 * d - 1) : 0.0f;
    P.sh = h > 1 ? (float)(h - 1) / (float)(2 * h - 1) : 0.0f;
    P.sw = w > 1 ? (float)(w - 1) / (float)(2 * w - 1) : 0.0f;
    return 0;
}

int grid_u(int64_t items) {
    int64_t b = (items + kThreadsU - 1) / kThreadsU;
    const int64_t cap = (int64_t)FPL_NUM_SMS * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int fpl_upsample2x_c8(const void* x, int x_c8tot, int x_c8off, void* y, int y_c8tot, int y_c8off, int n, int d,
                                 int h, int w, int c, int kd2, void* stream) {
    UpParams P;
    if (int rc = fill(P, x, x_c8tot, x_c8off, y, y_c8tot, y_c8off, n, d, h, w, c, kd2)) return rc;
    fpl_launch(upsample2x_fwd_kernel, grid_u((int64_t)n * d * kd2 * (c / 8) * 4 * h * w), kThreadsU, 0, (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_upsample2x_c8_bwd(const void* gy, int gy_c8tot, int gy_c8off, void* gx, int gx_c8tot, int gx_c8off, int n,
                                     int d, int h, int w, int c, int kd2, void* stream) {
    UpParams P;
    if (int rc = fill(P, gy, gy_c8tot, gy_c8off, gx, gx_c8tot, gx_c8off, n, d, h, w, c, kd2)) return rc;
    fpl_launch(upsample2x_bwd_kernel, grid_u((int64_t)n * d * (c / 8) * h * w), kThreadsU, 0, (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}
