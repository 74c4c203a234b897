// CUDA-core convolution kernels on C8-planar bf16 activations.
//   * fpl_conv3d_direct / fpl_conv3d_wgrad : general k(kd,3,3) conv, its dgrad (transpose_flip) and wgrad.
//     The tensor-core path (conv_tc.cu) supersedes fwd/dgrad wherever its shape constraints hold;
//     these stay as the cover for the remaining shapes and as the on-device cross-check.
//   * stem (image fp32 NCDHW, in_chns <= 8) and head ((1,3,3) conv to class logits, fp32 NCDHW):
//     PyMIC/pymic/net/net3d/unet2d5_dsbn.py:75 (first conv3d_1) and :293-294.
//   * ConvTranspose3d k2 s2 (unet2d5_dsbn.py:152,181) forward / dgrad / wgrad.
// Accumulation is fp32 everywhere; weights stay fp32 (optionally rounded to bf16 for cross-checks).
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

__device__ __forceinline__ float maybe_round(float v, int round_bf16) {
    return round_bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}

// ------------------------------------------------------------------------------------
// general direct conv: thread = one output voxel x 8 output channels
// ------------------------------------------------------------------------------------
struct ConvP {
    const bf16x8* x; int x_c8tot, x_c8off;
    const float* w; const float* bias;
    bf16x8* y; int y_c8tot, y_c8off;
    double* stats;
    int N, D, H, W, cin, cout, kd, transpose_flip, round_w;
};

__global__ void __launch_bounds__(128) conv3d_direct_kernel(ConvP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int T = P.kd * 9;
    __shared__ __align__(16) float wsm[27][8][8];   // [tap][ci][co]
    __shared__ float red[4][16];
    const int HW = P.H * P.W;
    const int hw = blockIdx.x * 128 + threadIdx.x;
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int co8 = blockIdx.z;
    const bool active = hw < HW;
    const int h = active ? hw / P.W : 0, w = active ? hw % P.W : 0;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    const int pad_d = P.kd / 2;
    for (int ci8 = 0; ci8 < P.cin / 8; ++ci8) {
        __syncthreads();
        for (int e = threadIdx.x; e < T * 64; e += 128) {
            int t = e / 64, ci = (e / 8) % 8, co = e % 8;
            int o = co8 * 8 + co, i = ci8 * 8 + ci;
            float v = P.transpose_flip ? P.w[((int64_t)i * P.cout + o) * T + (T - 1 - t)]
                                       : P.w[((int64_t)o * P.cin + i) * T + t];
            wsm[t][ci][co] = maybe_round(v, P.round_w);
        }
        __syncthreads();
        if (!active) continue;
        for (int kd = 0; kd < P.kd; ++kd) {
            int dz = d + kd - pad_d;
            if (dz < 0 || dz >= P.D) continue;
            const bf16x8* plane = P.x + (((int64_t)n * P.D + dz) * P.x_c8tot + P.x_c8off + ci8) * HW;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                int hy = h + kh - 1;
                if (hy < 0 || hy >= P.H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    int wx = w + kw - 1;
                    if (wx < 0 || wx >= P.W) continue;
                    float xf[8];
                    bf16x8_to_float(ldg_bf16x8(plane + hy * P.W + wx), xf);
                    const int t = (kd * 3 + kh) * 3 + kw;
#pragma unroll
                    for (int ci = 0; ci < 8; ++ci) {
                        float4 wa = *reinterpret_cast<const float4*>(&wsm[t][ci][0]);
                        float4 wb = *reinterpret_cast<const float4*>(&wsm[t][ci][4]);
                        acc[0] = fmaf(xf[ci], wa.x, acc[0]); acc[1] = fmaf(xf[ci], wa.y, acc[1]);
                        acc[2] = fmaf(xf[ci], wa.z, acc[2]); acc[3] = fmaf(xf[ci], wa.w, acc[3]);
                        acc[4] = fmaf(xf[ci], wb.x, acc[4]); acc[5] = fmaf(xf[ci], wb.y, acc[5]);
                        acc[6] = fmaf(xf[ci], wb.z, acc[6]); acc[7] = fmaf(xf[ci], wb.w, acc[7]);
                    }
                }
            }
        }
    }
    if (P.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += P.bias[co8 * 8 + i];
    }
    if (active) st_bf16x8(&P.y[(((int64_t)n * P.D + d) * P.y_c8tot + P.y_c8off + co8) * HW + hw], acc);
    if (P.stats != nullptr) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = active ? acc[i] : 0.0f;
            float s = warp_sum(v), q = warp_sum(v * v);
            if (lane == 0) { red[wid][i] = s; red[wid][8 + i] = q; }
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            float t = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
            int i = threadIdx.x & 7;
            atomicAdd(P.stats + (threadIdx.x < 8 ? 0 : P.cout) + co8 * 8 + i, (double)t);
        }
    }
}

// ------------------------------------------------------------------------------------
// general wgrad / convT wgrad: block = (plane, 32x32 (co,ci) tile, tap), 256 threads, thread = 1 co x 4 ci
//   MODE 0: conv      dW[co][ci][tap]  += sum_v dy[v][co] * x[v + tap - pad][ci]
//   MODE 1: convT k2  dW[ci][co][abc]  += sum_v x[v][ci]  * dy[2v + abc][co]     (v on the low-res grid)
// ------------------------------------------------------------------------------------
struct WgradP {
    const bf16x8* x; int x_c8tot, x_c8off;
    const bf16x8* dy; int dy_c8tot, dy_c8off;
    float* dw;
    int N, D, H, W, cin, cout, kd;   // D,H,W: grid of v (conv: the common grid; convT: low-res grid)
};

template <int MODE>
__global__ void __launch_bounds__(256) wgrad_kernel(WgradP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    __shared__ __align__(16) float dys[64][33];
    __shared__ __align__(16) float xs[64][36];
    const int T = MODE == 0 ? P.kd * 9 : P.kd * 4;
    const int tap = blockIdx.z;
    const int nd = blockIdx.x, n = nd / P.D, d = nd % P.D;
    const int ci_tiles = (P.cin + 31) / 32;
    const int co0 = (blockIdx.y / ci_tiles) * 32, ci0 = (blockIdx.y % ci_tiles) * 32;
    const int HW = P.H * P.W;
    const int tco = threadIdx.x / 8, tci = (threadIdx.x % 8) * 4;
    // tap decomposition
    int td, th, tw;
    if (MODE == 0) { td = tap / 9 - P.kd / 2; th = (tap / 3) % 3 - 1; tw = tap % 3 - 1; }
    else { td = tap / 4; th = (tap / 2) % 2; tw = tap % 2; }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // loader role: thread -> (voxel lv = tid/4, channel group lg = tid%4)
    const int lv = threadIdx.x / 4, lg = threadIdx.x % 4;
    for (int base = 0; base < HW; base += 64) {
        int hw = base + lv;
        float fy[8], fx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { fy[i] = 0.f; fx[i] = 0.f; }
        if (hw < HW) {
            int h = hw / P.W, w = hw % P.W;
            if (MODE == 0) {
                if (co0 + lg * 8 < P.cout)
                    bf16x8_to_float(ldg_bf16x8(P.dy + (((int64_t)n * P.D + d) * P.dy_c8tot + P.dy_c8off + co0 / 8 + lg) * HW + hw), fy);
                int dz = d + td, hy = h + th, wx = w + tw;
                if (ci0 + lg * 8 < P.cin && dz >= 0 && dz < P.D && hy >= 0 && hy < P.H && wx >= 0 && wx < P.W)
                    bf16x8_to_float(ldg_bf16x8(P.x + (((int64_t)n * P.D + dz) * P.x_c8tot + P.x_c8off + ci0 / 8 + lg) * HW + hy * P.W + wx), fx);
            } else {
                const int Do = P.D * P.kd, Ho = P.H * 2, Wo = P.W * 2;
                int dz = d * P.kd + td, hy = h * 2 + th, wx = w * 2 + tw;
                if (co0 + lg * 8 < P.cout)
                    bf16x8_to_float(ldg_bf16x8(P.dy + (((int64_t)n * Do + dz) * P.dy_c8tot + P.dy_c8off + co0 / 8 + lg) * ((int64_t)Ho * Wo) + (int64_t)hy * Wo + wx), fy);
                if (ci0 + lg * 8 < P.cin)
                    bf16x8_to_float(ldg_bf16x8(P.x + (((int64_t)n * P.D + d) * P.x_c8tot + P.x_c8off + ci0 / 8 + lg) * HW + hw), fx);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) { dys[lv][lg * 8 + i] = fy[i]; xs[lv][lg * 8 + i] = fx[i]; }
        __syncthreads();
#pragma unroll 8
        for (int v = 0; v < 64; ++v) {
            float g = dys[v][tco];
            float4 xv = *reinterpret_cast<const float4*>(&xs[v][tci]);
            acc[0] = fmaf(g, xv.x, acc[0]); acc[1] = fmaf(g, xv.y, acc[1]);
            acc[2] = fmaf(g, xv.z, acc[2]); acc[3] = fmaf(g, xv.w, acc[3]);
        }
    }
    int co = co0 + tco;
    if (co < P.cout) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = ci0 + tci + j;
            if (ci < P.cin) {
                int64_t o = MODE == 0 ? ((int64_t)co * P.cin + ci) * T + tap : ((int64_t)ci * P.cout + co) * T + tap;
                atomicAdd(P.dw + o, acc[j]);
            }
        }
    }
}

// per-channel sum of a C8-planar tensor (bias gradients): grid (chunks, planes)
__global__ void __launch_bounds__(256) channel_sum_kernel(const bf16x8* __restrict__ g, int c8tot, int c8off, float* out,
                                                          int ND, int C8, int HW) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int plane = blockIdx.y, c8 = plane % C8, nd = plane / C8;
    float s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
    for (int hw = blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += gridDim.x * blockDim.x) {
        float f[8];
        bf16x8_to_float(ldg_bf16x8(g + ((int64_t)nd * c8tot + c8off + c8) * HW + hw), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
    __shared__ float sm[8][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float t = warp_sum(s[i]);
        if (lane == 0) sm[wid][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
        atomicAdd(out + c8 * 8 + threadIdx.x, t);
    }
}

// ------------------------------------------------------------------------------------
// stem: fp32 NCDHW image -> C8-planar bf16
// ------------------------------------------------------------------------------------
struct StemP {
    const float* x; const float* w; const float* bias;
    bf16x8* y; int y_c8tot, y_c8off; double* stats;
    int N, cin, D, H, W, cout, kd;
};

__global__ void __launch_bounds__(128) stem_conv_fwd_kernel(StemP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float wsm_dyn[];   // [ci][tap][8 co] for this co8
    __shared__ float red[4][16];
    const int T = P.kd * 9;
    const int HW = P.H * P.W;
    const int hw = blockIdx.x * 128 + threadIdx.x;
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int co8 = blockIdx.z;
    for (int e = threadIdx.x; e < P.cin * T * 8; e += 128) {
        int co = e % 8, t = (e / 8) % T, ci = e / (8 * T);
        wsm_dyn[e] = P.w[((int64_t)(co8 * 8 + co) * P.cin + ci) * T + t];
    }
    __syncthreads();
    const bool active = hw < HW;
    const int h = active ? hw / P.W : 0, w = active ? hw % P.W : 0;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = P.bias != nullptr ? P.bias[co8 * 8 + i] : 0.f;
    if (active) {
        for (int ci = 0; ci < P.cin; ++ci) {
            const float* img = P.x + ((int64_t)n * P.cin + ci) * P.D * HW;
            for (int kd = 0; kd < P.kd; ++kd) {
                int dz = d + kd - P.kd / 2;
                if (dz < 0 || dz >= P.D) continue;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {
                    int hy = h + kh - 1;
                    if (hy < 0 || hy >= P.H) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        int wx = w + kw - 1;
                        if (wx < 0 || wx >= P.W) continue;
                        float xv = __ldg(img + (int64_t)dz * HW + hy * P.W + wx);
                        const float* wp = wsm_dyn + (ci * T + (kd * 3 + kh) * 3 + kw) * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv, wp[i], acc[i]);
                    }
                }
            }
        }
        st_bf16x8(&P.y[(((int64_t)n * P.D + d) * P.y_c8tot + P.y_c8off + co8) * HW + hw], acc);
    }
    if (P.stats != nullptr) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = active ? acc[i] : 0.0f;
            float s = warp_sum(v), q = warp_sum(v * v);
            if (lane == 0) { red[wid][i] = s; red[wid][8 + i] = q; }
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            float t = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
            atomicAdd(P.stats + (threadIdx.x < 8 ? 0 : P.cout) + co8 * 8 + (threadIdx.x & 7), (double)t);
        }
    }
}

// stem wgrad: block = one (n,d,h) row segment of 64 voxels; thread = one (co, ci, tap) weight element
struct StemWP {
    const float* x; const bf16x8* dy; int dy_c8tot, dy_c8off; float* dw;
    int N, cin, D, H, W, cout, kd;
};

__global__ void __launch_bounds__(256) stem_conv_wgrad_kernel(StemWP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float sm_dyn[];
    const int T = P.kd * 9;
    const int HW = P.H * P.W;
    const int rows_per_block = 4;                       // 4 rows x W voxels per iteration
    float* dys = sm_dyn;                                // [rows*W][cout]  (+1 pad)
    const int ldy = P.cout + 1;
    float* xs = sm_dyn + rows_per_block * P.W * ldy;    // [cin][kd][rows+2][W+2]
    const int xw = P.W + 2, xrows = rows_per_block + 2;
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int nelem = P.cout * P.cin * T;
    constexpr int kMaxPerThread = 8;
    float acc[kMaxPerThread];
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) acc[i] = 0.f;
    for (int h0 = blockIdx.x * rows_per_block; h0 < P.H; h0 += gridDim.x * rows_per_block) {
        __syncthreads();
        // stage dy rows (zero beyond H)
        for (int e = threadIdx.x; e < rows_per_block * P.W * (P.cout / 8); e += 256) {
            int c8 = e % (P.cout / 8), v = e / (P.cout / 8);
            int r = v / P.W, w = v % P.W, h = h0 + r;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = 0.f;
            if (h < P.H)
                bf16x8_to_float(ldg_bf16x8(P.dy + (((int64_t)n * P.D + d) * P.dy_c8tot + P.dy_c8off + c8) * HW + h * P.W + w), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) dys[v * ldy + c8 * 8 + i] = f[i];
        }
        // stage x halo tile
        for (int e = threadIdx.x; e < P.cin * P.kd * xrows * xw; e += 256) {
            int xx = e % xw, rr = (e / xw) % xrows, kd = (e / (xw * xrows)) % P.kd, ci = e / (xw * xrows * P.kd);
            int dz = d + kd - P.kd / 2, hy = h0 + rr - 1, wx = xx - 1;
            float v = 0.f;
            if (dz >= 0 && dz < P.D && hy >= 0 && hy < P.H && wx >= 0 && wx < P.W)
                v = __ldg(P.x + (((int64_t)n * P.cin + ci) * P.D + dz) * HW + hy * P.W + wx);
            xs[e] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kMaxPerThread; ++k) {
            int e = threadIdx.x + k * 256;
            if (e < nelem) {
                int t = e % T, ci = (e / T) % P.cin, co = e / (T * P.cin);
                int kd = t / 9, kh = (t / 3) % 3, kw = t % 3;
                const float* xb = xs + ((ci * P.kd + kd) * xrows + kh) * xw + kw;
                float a = 0.f;
                for (int r = 0; r < rows_per_block; ++r)
                    for (int w = 0; w < P.W; ++w) a = fmaf(dys[(r * P.W + w) * ldy + co], xb[r * xw + w], a);
                acc[k] += a;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxPerThread; ++k) {
        int e = threadIdx.x + k * 256;
        if (e < nelem) atomicAdd(P.dw + e, acc[k]);
    }
}

// ------------------------------------------------------------------------------------
// head: (1,3,3) conv to fp32 NCDHW logits, and its backward
// ------------------------------------------------------------------------------------
constexpr int kMaxClasses = 8;

struct HeadP {
    const bf16x8* x; int x_c8tot, x_c8off;
    const float* w; const float* bias; float* logits;
    const float* dlogits; bf16x8* dx; int dx_c8tot, dx_c8off; float* dw; float* db;
    int N, D, H, W, cin, classes;
};

__global__ void __launch_bounds__(128) head_conv_fwd_kernel(HeadP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float wsm_dyn[];   // [tap 9][ci][cls]
    const int HW = P.H * P.W;
    for (int e = threadIdx.x; e < 9 * P.cin * P.classes; e += 128) {
        int cls = e % P.classes, ci = (e / P.classes) % P.cin, t = e / (P.classes * P.cin);
        wsm_dyn[e] = P.w[((int64_t)cls * P.cin + ci) * 9 + t];
    }
    __syncthreads();
    const int hw = blockIdx.x * 128 + threadIdx.x;
    if (hw >= HW) return;
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int h = hw / P.W, w = hw % P.W;
    float acc[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) acc[c] = (c < P.classes && P.bias != nullptr) ? P.bias[c] : 0.f;
    for (int ci8 = 0; ci8 < P.cin / 8; ++ci8) {
        const bf16x8* plane = P.x + ((int64_t)nd * P.x_c8tot + P.x_c8off + ci8) * HW;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            int hy = h + kh - 1;
            if (hy < 0 || hy >= P.H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int wx = w + kw - 1;
                if (wx < 0 || wx >= P.W) continue;
                float xf[8];
                bf16x8_to_float(ldg_bf16x8(plane + hy * P.W + wx), xf);
                const float* wp = wsm_dyn + ((kh * 3 + kw) * P.cin + ci8 * 8) * P.classes;
#pragma unroll
                for (int ci = 0; ci < 8; ++ci)
#pragma unroll
                    for (int c = 0; c < kMaxClasses; ++c)
                        if (c < P.classes) acc[c] = fmaf(xf[ci], wp[ci * P.classes + c], acc[c]);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
        if (c < P.classes) P.logits[(((int64_t)n * P.classes + c) * P.D + d) * HW + hw] = acc[c];
}

// dgrad: thread = voxel x 8 input channels
__global__ void __launch_bounds__(128) head_conv_dgrad_kernel(HeadP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float wsm_dyn[];   // [tap 9][cls][8 ci] for this ci8
    const int HW = P.H * P.W;
    const int ci8 = blockIdx.z;
    for (int e = threadIdx.x; e < 9 * P.classes * 8; e += 128) {
        int ci = e % 8, cls = (e / 8) % P.classes, t = e / (8 * P.classes);
        wsm_dyn[e] = P.w[((int64_t)cls * P.cin + ci8 * 8 + ci) * 9 + t];
    }
    __syncthreads();
    const int hw = blockIdx.x * 128 + threadIdx.x;
    if (hw >= HW) return;
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int h = hw / P.W, w = hw % P.W;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // dx[v] = sum_t dlogits[v - (t-1)] * w[t]
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        int hy = h - (kh - 1);
        if (hy < 0 || hy >= P.H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            int wx = w - (kw - 1);
            if (wx < 0 || wx >= P.W) continue;
            for (int c = 0; c < P.classes; ++c) {
                float g = __ldg(P.dlogits + (((int64_t)n * P.classes + c) * P.D + d) * HW + hy * P.W + wx);
                const float* wp = wsm_dyn + ((kh * 3 + kw) * P.classes + c) * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(g, wp[i], acc[i]);
            }
        }
    }
    st_bf16x8(&P.dx[((int64_t)nd * P.dx_c8tot + P.dx_c8off + ci8) * HW + hw], acc);
}

// wgrad + bias grad: block = (row group, plane); thread = (cls, ci, tap) element
__global__ void __launch_bounds__(256) head_conv_wgrad_kernel(HeadP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float sm_dyn[];
    const int HW = P.H * P.W;
    const int rows = 2;
    const int xw = P.W + 2, xrows = rows + 2;
    float* gs = sm_dyn;                                 // [cls][rows*W]
    float* xs = sm_dyn + P.classes * rows * P.W;        // [ci][xrows][xw]
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int nelem = P.classes * P.cin * 9;
    constexpr int kMaxPerThread = 8;
    float acc[kMaxPerThread];
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) acc[i] = 0.f;
    float bacc = 0.f;
    for (int h0 = blockIdx.x * rows; h0 < P.H; h0 += gridDim.x * rows) {
        __syncthreads();
        for (int e = threadIdx.x; e < P.classes * rows * P.W; e += 256) {
            int v = e % (rows * P.W), c = e / (rows * P.W);
            int h = h0 + v / P.W, w = v % P.W;
            gs[e] = h < P.H ? __ldg(P.dlogits + (((int64_t)n * P.classes + c) * P.D + d) * HW + h * P.W + w) : 0.f;
        }
        for (int e = threadIdx.x; e < (P.cin / 8) * xrows * xw; e += 256) {
            int xx = e % xw, rr = (e / xw) % xrows, c8 = e / (xw * xrows);
            int hy = h0 + rr - 1, wx = xx - 1;
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = 0.f;
            if (hy >= 0 && hy < P.H && wx >= 0 && wx < P.W)
                bf16x8_to_float(ldg_bf16x8(P.x + ((int64_t)nd * P.x_c8tot + P.x_c8off + c8) * HW + hy * P.W + wx), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) xs[((c8 * 8 + i) * xrows + rr) * xw + xx] = f[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kMaxPerThread; ++k) {
            int e = threadIdx.x + k * 256;
            if (e < nelem) {
                int t = e % 9, ci = (e / 9) % P.cin, c = e / (9 * P.cin);
                const float* xb = xs + (ci * xrows + t / 3) * xw + t % 3;
                const float* gb = gs + c * rows * P.W;
                float a = 0.f;
                for (int r = 0; r < rows; ++r)
                    for (int w = 0; w < P.W; ++w) a = fmaf(gb[r * P.W + w], xb[r * xw + w], a);
                acc[k] += a;
            }
        }
        if (threadIdx.x < P.classes) {
            const float* gb = gs + threadIdx.x * rows * P.W;
            for (int v = 0; v < rows * P.W; ++v) bacc += gb[v];
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxPerThread; ++k) {
        int e = threadIdx.x + k * 256;
        if (e < nelem) atomicAdd(P.dw + e, acc[k]);
    }
    if (threadIdx.x < P.classes && P.db != nullptr) atomicAdd(P.db + threadIdx.x, bacc);
}

// ------------------------------------------------------------------------------------
// ConvTranspose k2 s2
// ------------------------------------------------------------------------------------
struct ConvTP {
    const bf16x8* x; int x_c8tot, x_c8off;
    const float* w; const float* bias;
    bf16x8* y; int y_c8tot, y_c8off;          // fwd output / bwd: dy
    bf16x8* dx; int dx_c8tot, dx_c8off;
    int N, D, H, W, cin, cout, kd2;           // D,H,W: LOW-res grid
};

// fwd: thread = output voxel x 8 co.  block = 128 consecutive output w of one output row.
__global__ void __launch_bounds__(128) convt_fwd_kernel(ConvTP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float wsm_dyn[];   // [ci][c parity 2][8 co]
    const int T = P.kd2 * 4;
    const int Ho = P.H * 2, Wo = P.W * 2, Do = P.D * P.kd2;
    const int co8 = blockIdx.z;
    const int row = blockIdx.y;                       // (n, do, ho)
    const int ho = row % Ho, dq = (row / Ho) % Do, n = row / (Ho * Do);
    const int a = P.kd2 == 2 ? (dq & 1) : 0, b = ho & 1;
    const int d = P.kd2 == 2 ? dq >> 1 : dq, h = ho >> 1;
    for (int e = threadIdx.x; e < P.cin * 16; e += 128) {
        int co = e % 8, c = (e / 8) % 2, ci = e / 16;
        wsm_dyn[e] = P.w[((int64_t)ci * P.cout + co8 * 8 + co) * T + (a * 2 + b) * 2 + c];
    }
    __syncthreads();
    const int wo = blockIdx.x * 128 + threadIdx.x;
    if (wo >= Wo) return;
    const int w = wo >> 1, c = wo & 1;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = P.bias != nullptr ? P.bias[co8 * 8 + i] : 0.f;
    const int HW = P.H * P.W;
    for (int ci8 = 0; ci8 < P.cin / 8; ++ci8) {
        float xf[8];
        bf16x8_to_float(ldg_bf16x8(P.x + (((int64_t)n * P.D + d) * P.x_c8tot + P.x_c8off + ci8) * HW + h * P.W + w), xf);
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) {
            const float* wp = wsm_dyn + ((ci8 * 8 + ci) * 2 + c) * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(xf[ci], wp[i], acc[i]);
        }
    }
    st_bf16x8(&P.y[(((int64_t)n * Do + dq) * P.y_c8tot + P.y_c8off + co8) * ((int64_t)Ho * Wo) + (int64_t)ho * Wo + wo], acc);
}

// dgrad: thread = low-res voxel x 8 ci; loops the kd2*4 positions and all co
__global__ void __launch_bounds__(128) convt_dgrad_kernel(ConvTP P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    extern __shared__ float wsm_dyn[];   // [pos T][co][8 ci]
    const int T = P.kd2 * 4;
    const int ci8 = blockIdx.z;
    for (int e = threadIdx.x; e < T * P.cout * 8; e += 128) {
        int ci = e % 8, co = (e / 8) % P.cout, t = e / (8 * P.cout);
        wsm_dyn[e] = P.w[((int64_t)(ci8 * 8 + ci) * P.cout + co) * T + t];
    }
    __syncthreads();
    const int HW = P.H * P.W;
    const int hw = blockIdx.x * 128 + threadIdx.x;
    if (hw >= HW) return;
    const int nd = blockIdx.y, n = nd / P.D, d = nd % P.D;
    const int h = hw / P.W, w = hw % P.W;
    const int Ho = P.H * 2, Wo = P.W * 2, Do = P.D * P.kd2;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int t = 0; t < T; ++t) {
        int a = t / 4, b = (t / 2) % 2, c = t % 2;
        int dq = d * P.kd2 + a, ho = h * 2 + b, wo = w * 2 + c;
        for (int co8 = 0; co8 < P.cout / 8; ++co8) {
            float g[8];
            bf16x8_to_float(ldg_bf16x8(P.y + (((int64_t)n * Do + dq) * P.y_c8tot + P.y_c8off + co8) * ((int64_t)Ho * Wo) + (int64_t)ho * Wo + wo), g);
#pragma unroll
            for (int co = 0; co < 8; ++co) {
                const float* wp = wsm_dyn + (t * P.cout + co8 * 8 + co) * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(g[co], wp[i], acc[i]);
            }
        }
    }
    st_bf16x8(&P.dx[((int64_t)nd * P.dx_c8tot + P.dx_c8off + ci8) * HW + hw], acc);
}

// fp32 NCDHW [N][C][D][H][W] -> C8-planar bf16 channel groups (zero padded to a multiple of 8 channels),
// optional per-channel sums (the bias gradient when the tensor is dlogits).  grid: (chunks of H*W, N*D)
__global__ void __launch_bounds__(256) pack_ncdhw_c8_kernel(const float* __restrict__ x, int C, bf16x8* out, int c8tot,
                                                            int c8off, int groups, float* chan_sum, int D, int HW) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int nd = blockIdx.y, n = nd / D, d = nd - n * D;
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0.f;
    for (int hw = blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += gridDim.x * blockDim.x) {
        for (int g = 0; g < groups; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = g * 8 + i;
                f[i] = c < C ? __ldg(x + (((int64_t)n * C + c) * D + d) * HW + hw) : 0.f;
                if (g < 2) s[g * 8 + i] += f[i];
            }
            st_bf16x8(out + ((int64_t)nd * c8tot + c8off + g) * HW + hw, f);
        }
    }
    if (chan_sum != nullptr) {
        __shared__ float sm[8][16];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float t = warp_sum(s[i]);
            if (lane == 0) sm[wid][i] = t;
        }
        __syncthreads();
        if (threadIdx.x < 16 && threadIdx.x < C) {
            float t = 0.f;
            for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
            atomicAdd(chan_sum + threadIdx.x, t);
        }
    }
}

// 1-channel fp32 image [N][1][D][H][W] -> C8-planar bf16 with 16 channels: channel t9 = kh*3+kw holds the in-plane
// neighbour x[d][h+kh-1][w+kw-1] (zero outside the plane), channels 9..15 are zero.  The stem conv k(3,3,3) then is a
// k(3,1,1) conv over these 16 channels, which the tensor-core kernels run with ONE in-plane tap.
__global__ void __launch_bounds__(256) patch9_kernel(const float* __restrict__ x, bf16x8* out, int D, int H, int W, int split) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int nd = blockIdx.y;
    const int HW = H * W;
    const float* plane = x + (int64_t)nd * HW;
    for (int hw = blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += gridDim.x * blockDim.x) {
        const int h = hw / W, w = hw - h * W;
        float f[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) f[t] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hy = h + kh - 1;
            if (hy < 0 || hy >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int wx = w + kw - 1;
                if (wx >= 0 && wx < W) f[kh * 3 + kw] = __ldg(plane + hy * W + wx);
            }
        }
        const int groups = split ? 4 : 2;
        st_bf16x8(out + ((int64_t)nd * groups + 0) * HW + hw, f);
        st_bf16x8(out + ((int64_t)nd * groups + 1) * HW + hw, f + 8);
        if (split) {
            // second half: the bf16 rounding residual, so that hi + lo carries 16 mantissa bits of the image
#pragma unroll
            for (int t = 0; t < 16; ++t) f[t] -= __bfloat162float(__float2bfloat16_rn(f[t]));
            st_bf16x8(out + ((int64_t)nd * 4 + 2) * HW + hw, f);
            st_bf16x8(out + ((int64_t)nd * 4 + 3) * HW + hw, f + 8);
        }
    }
}

}  // namespace

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" int fpl_patch9_c8(const float* x, void* out, int split_hi_lo, int n, int d, int h, int w, void* stream) {
    FPL_REQUIRE((int64_t)n * d <= 65535, "fpl_patch9_c8: too many planes");
    int chunks = (h * w + 1023) / 1024;
    fpl_launch(patch9_kernel, dim3(chunks, n * d), 256, 0, (cudaStream_t)stream, x, (bf16x8*)out, d, h, w, split_hi_lo);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_pack_ncdhw_to_c8(const float* x, int c, void* out, int out_c8tot, int out_c8off, int groups,
                                    float* chan_sum, int n, int d, int h, int w, void* stream) {
    FPL_REQUIRE(c >= 1 && groups >= 1 && c <= groups * 8, "fpl_pack_ncdhw_to_c8: %d channels do not fit %d groups", c, groups);
    FPL_REQUIRE(chan_sum == nullptr || c <= 16, "fpl_pack_ncdhw_to_c8: channel sums need c <= 16");
    FPL_REQUIRE((int64_t)n * d <= 65535, "fpl_pack_ncdhw_to_c8: too many planes");
    int chunks = (h * w + 1023) / 1024;
    fpl_launch(pack_ncdhw_c8_kernel, dim3(chunks, n * d), 256, 0, (cudaStream_t)stream, x, c, (bf16x8*)out, out_c8tot, out_c8off,
                                                                               groups, chan_sum, d, h * w);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_conv3d_direct(const void* x, int x_c8tot, int x_c8off, const float* w, const float* bias, void* y,
                                 int y_c8tot, int y_c8off, double* stats, int n, int d, int h, int w_, int cin,
                                 int cout, int kd, int transpose_flip, int round_w_bf16, void* stream) {
    FPL_REQUIRE(cin % 8 == 0 && cout % 8 == 0, "fpl_conv3d_direct: channels (%d,%d) must be multiples of 8", cin, cout);
    FPL_REQUIRE(kd == 1 || kd == 3, "fpl_conv3d_direct: kd=%d must be 1 or 3", kd);
    FPL_REQUIRE((int64_t)n * d <= 65535 && cout / 8 <= 65535, "fpl_conv3d_direct: grid too large");
    ConvP P{(const bf16x8*)x, x_c8tot, x_c8off, w, bias, (bf16x8*)y, y_c8tot, y_c8off, stats,
            n, d, h, w_, cin, cout, kd, transpose_flip, round_w_bf16};
    dim3 grid((h * w_ + 127) / 128, n * d, cout / 8);
    fpl_launch(conv3d_direct_kernel, grid, 128, 0, (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_conv3d_wgrad(const void* x, int x_c8tot, int x_c8off, const void* dy, int dy_c8tot, int dy_c8off,
                                float* dw, int n, int d, int h, int w, int cin, int cout, int kd, void* stream) {
    FPL_REQUIRE(cin % 8 == 0 && cout % 8 == 0, "fpl_conv3d_wgrad: channels (%d,%d) must be multiples of 8", cin, cout);
    FPL_REQUIRE(kd == 1 || kd == 3, "fpl_conv3d_wgrad: kd=%d must be 1 or 3", kd);
    WgradP P{(const bf16x8*)x, x_c8tot, x_c8off, (const bf16x8*)dy, dy_c8tot, dy_c8off, dw, n, d, h, w, cin, cout, kd};
    dim3 grid(n * d, ((cout + 31) / 32) * ((cin + 31) / 32), kd * 9);
    fpl_launch(wgrad_kernel<0>, grid, 256, 0, (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

bool fpl_stem_tc_eligible(int n, int cin, int d, int h, int w, int cout, int kd);
int fpl_stem_fwd_tc_launch(const float* x, const float* w, const float* bias, void* y, int y_c8tot, int y_c8off, double* stats, int n,
                           int d, int h, int w_, void* stream);
int fpl_stem_wgrad_tc_launch(const float* x, const void* dy, int dy_c8tot, int dy_c8off, float* dw, int n, int d, int h, int w_,
                             void* stream);

extern "C" int fpl_stem_conv_fwd(const float* x, const float* w, const float* bias, void* y, int y_c8tot, int y_c8off,
                                 double* stats, int n, int cin, int d, int h, int w_, int cout, int kd, void* stream) {
    FPL_REQUIRE(cin >= 1 && cin <= 8, "fpl_stem_conv_fwd: in_chns=%d not in [1,8]", cin);
    FPL_REQUIRE(cout % 8 == 0, "fpl_stem_conv_fwd: cout=%d must be a multiple of 8", cout);
    FPL_REQUIRE(kd == 1 || kd == 3, "fpl_stem_conv_fwd: kd=%d must be 1 or 3", kd);
    if (fpl_stem_tc_eligible(n, cin, d, h, w_, cout, kd) && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {   // tensor cores (stem_tc.cu)
        if (int rc = fpl_stem_fwd_tc_launch(x, w, bias, y, y_c8tot, y_c8off, stats, n, d, h, w_, stream)) return rc;
        FPL_LAUNCH_CHECK();
        return 0;
    }
    StemP P{x, w, bias, (bf16x8*)y, y_c8tot, y_c8off, stats, n, cin, d, h, w_, cout, kd};
    dim3 grid((h * w_ + 127) / 128, n * d, cout / 8);
    size_t smem = (size_t)cin * kd * 9 * 8 * sizeof(float);
    fpl_launch(stem_conv_fwd_kernel, grid, 128, smem, (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_stem_conv_wgrad(const float* x, const void* dy, int dy_c8tot, int dy_c8off, float* dw, int n,
                                   int cin, int d, int h, int w_, int cout, int kd, void* stream) {
    FPL_REQUIRE(cin >= 1 && cin <= 8 && cout % 8 == 0, "fpl_stem_conv_wgrad: bad channels (%d,%d)", cin, cout);
    FPL_REQUIRE(cout * cin * kd * 9 <= 256 * 8, "fpl_stem_conv_wgrad: %d weights exceed the per-block budget", cout * cin * kd * 9);
    if (fpl_stem_tc_eligible(n, cin, d, h, w_, cout, kd)) {                                                    // tensor cores (stem_tc.cu)
        if (int rc = fpl_stem_wgrad_tc_launch(x, dy, dy_c8tot, dy_c8off, dw, n, d, h, w_, stream)) return rc;
        FPL_LAUNCH_CHECK();
        return 0;
    }
    StemWP P{x, (const bf16x8*)dy, dy_c8tot, dy_c8off, dw, n, cin, d, h, w_, cout, kd};
    size_t smem = ((size_t)4 * w_ * (cout + 1) + (size_t)cin * kd * 6 * (w_ + 2)) * sizeof(float);
    FPL_REQUIRE(smem <= 200 * 1024, "fpl_stem_conv_wgrad: row tile needs %zu B of shared memory", smem);
    FPL_CHECK_CUDA(cudaFuncSetAttribute(stem_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int rows_blocks = (h + 3) / 4;
    if (rows_blocks > 8) rows_blocks = 8;
    dim3 grid(rows_blocks, n * d);
    fpl_launch(stem_conv_wgrad_kernel, grid, 256, smem, (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_head_conv_fwd(const void* x, int x_c8tot, int x_c8off, const float* w, const float* bias,
                                 float* logits, int n, int d, int h, int w_, int cin, int classes, void* stream) {
    FPL_REQUIRE(cin % 8 == 0, "fpl_head_conv_fwd: cin=%d must be a multiple of 8", cin);
    FPL_REQUIRE(classes >= 1 && classes <= kMaxClasses, "fpl_head_conv_fwd: class_num=%d not in [1,%d]", classes, kMaxClasses);
    HeadP P{};
    P.x = (const bf16x8*)x; P.x_c8tot = x_c8tot; P.x_c8off = x_c8off; P.w = w; P.bias = bias; P.logits = logits;
    P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.classes = classes;
    dim3 grid((h * w_ + 127) / 128, n * d);
    fpl_launch(head_conv_fwd_kernel, grid, 128, (size_t)9 * cin * classes * sizeof(float), (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_head_conv_bwd(const void* x, int x_c8tot, int x_c8off, const float* w, const float* dlogits,
                                 void* dx, int dx_c8tot, int dx_c8off, float* dw, float* db, int n, int d, int h,
                                 int w_, int cin, int classes, void* stream) {
    FPL_REQUIRE(cin % 8 == 0, "fpl_head_conv_bwd: cin=%d must be a multiple of 8", cin);
    FPL_REQUIRE(classes >= 1 && classes <= kMaxClasses, "fpl_head_conv_bwd: class_num=%d not in [1,%d]", classes, kMaxClasses);
    FPL_REQUIRE(classes * cin * 9 <= 256 * 8, "fpl_head_conv_bwd: too many weights");
    HeadP P{};
    P.x = (const bf16x8*)x; P.x_c8tot = x_c8tot; P.x_c8off = x_c8off; P.w = w; P.dlogits = dlogits;
    P.dx = (bf16x8*)dx; P.dx_c8tot = dx_c8tot; P.dx_c8off = dx_c8off; P.dw = dw; P.db = db;
    P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.classes = classes;
    if (dx != nullptr) {
        dim3 grid((h * w_ + 127) / 128, n * d, cin / 8);
        fpl_launch(head_conv_dgrad_kernel, grid, 128, (size_t)9 * classes * 8 * sizeof(float), (cudaStream_t)stream, P);
        FPL_LAUNCH_CHECK();
    }
    if (dw != nullptr) {
        size_t smem = ((size_t)classes * 2 * w_ + (size_t)cin * 4 * (w_ + 2)) * sizeof(float);
        FPL_REQUIRE(smem <= 200 * 1024, "fpl_head_conv_bwd: row tile needs %zu B of shared memory", smem);
        FPL_CHECK_CUDA(cudaFuncSetAttribute(head_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int rb = (h + 1) / 2;
        if (rb > 16) rb = 16;
        dim3 grid(rb, n * d);
        fpl_launch(head_conv_wgrad_kernel, grid, 256, smem, (cudaStream_t)stream, P);
        FPL_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int fpl_convt_k2s2_fwd(const void* x, int x_c8tot, int x_c8off, const float* w, const float* bias, void* y,
                                  int y_c8tot, int y_c8off, int n, int d, int h, int w_, int cin, int cout, int kd2,
                                  void* stream) {
    FPL_REQUIRE(cin % 8 == 0 && cout % 8 == 0, "fpl_convt_k2s2_fwd: channels (%d,%d) must be multiples of 8", cin, cout);
    FPL_REQUIRE(kd2 == 1 || kd2 == 2, "fpl_convt_k2s2_fwd: kd2=%d must be 1 or 2", kd2);
    ConvTP P{};
    P.x = (const bf16x8*)x; P.x_c8tot = x_c8tot; P.x_c8off = x_c8off; P.w = w; P.bias = bias;
    P.y = (bf16x8*)y; P.y_c8tot = y_c8tot; P.y_c8off = y_c8off;
    P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.cout = cout; P.kd2 = kd2;
    int64_t rows = (int64_t)n * d * kd2 * h * 2;
    FPL_REQUIRE(rows <= 65535 * 32, "fpl_convt_k2s2_fwd: too many rows");
    FPL_REQUIRE(rows <= 65535, "fpl_convt_k2s2_fwd: too many output rows (%lld)", (long long)rows);
    dim3 grid((w_ * 2 + 127) / 128, (unsigned)rows, cout / 8);
    fpl_launch(convt_fwd_kernel, grid, 128, (size_t)cin * 16 * sizeof(float), (cudaStream_t)stream, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_convt_k2s2_bwd(const void* x, int x_c8tot, int x_c8off, const float* w, const void* dy, int dy_c8tot,
                                  int dy_c8off, void* dx, int dx_c8tot, int dx_c8off, float* dw, float* db, int n,
                                  int d, int h, int w_, int cin, int cout, int kd2, void* stream) {
    FPL_REQUIRE(cin % 8 == 0 && cout % 8 == 0, "fpl_convt_k2s2_bwd: channels (%d,%d) must be multiples of 8", cin, cout);
    FPL_REQUIRE(kd2 == 1 || kd2 == 2, "fpl_convt_k2s2_bwd: kd2=%d must be 1 or 2", kd2);
    const int T = kd2 * 4;
    if (dx != nullptr) {
        ConvTP P{};
        P.w = w; P.y = (bf16x8*)const_cast<void*>(dy); P.y_c8tot = dy_c8tot; P.y_c8off = dy_c8off;
        P.dx = (bf16x8*)dx; P.dx_c8tot = dx_c8tot; P.dx_c8off = dx_c8off;
        P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.cout = cout; P.kd2 = kd2;
        size_t smem = (size_t)T * cout * 8 * sizeof(float);
        FPL_REQUIRE(smem <= 200 * 1024, "fpl_convt_k2s2_bwd: cout=%d too large", cout);
        FPL_CHECK_CUDA(cudaFuncSetAttribute(convt_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((h * w_ + 127) / 128, n * d, cin / 8);
        fpl_launch(convt_dgrad_kernel, grid, 128, smem, (cudaStream_t)stream, P);
        FPL_LAUNCH_CHECK();
    }
    if (dw != nullptr) {
        WgradP P{(const bf16x8*)x, x_c8tot, x_c8off, (const bf16x8*)dy, dy_c8tot, dy_c8off, dw, n, d, h, w_, cin, cout, kd2};
        dim3 grid(n * d, ((cout + 31) / 32) * ((cin + 31) / 32), T);
        fpl_launch(wgrad_kernel<1>, grid, 256, 0, (cudaStream_t)stream, P);
        FPL_LAUNCH_CHECK();
    }
    if (db != nullptr) {
        int HWo = h * 2 * w_ * 2;
        int chunks = (HWo + 1023) / 1024;
        dim3 grid(chunks, n * d * kd2 * (cout / 8));
        fpl_launch(channel_sum_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16x8*)dy, dy_c8tot, dy_c8off, db, n * d * kd2,
                                                                  cout / 8, HWo);
        FPL_LAUNCH_CHECK();
    }
    return 0;
}

/* per-channel sums of a C8-planar bf16 tensor, ACCUMULATED into out[c] (bias gradient of a conv whose output gradient
 * is `g`): the 1x1 conv of UpBlock's bilinear mode. */
extern "C" int fpl_channel_sum_c8(const void* g, int g_c8tot, int g_c8off, float* out, int n, int d, int h, int w, int c,
                                  void* stream) {
    FPL_REQUIRE(c > 0 && c % 8 == 0 && g != nullptr && out != nullptr, "fpl_channel_sum_c8: bad arguments");
    FPL_REQUIRE((int64_t)n * d * (c / 8) <= 65535, "fpl_channel_sum_c8: too many planes");
    const int HW = h * w;
    int chunks = (HW + 1023) / 1024;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, n * d * (c / 8));
    fpl_launch(channel_sum_kernel, grid, 256, 0, (cudaStream_t)stream, (const bf16x8*)g, g_c8tot, g_c8off, out, n * d, c / 8, HW);
    FPL_LAUNCH_CHECK();
    return 0;
}
