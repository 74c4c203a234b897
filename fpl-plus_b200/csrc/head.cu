// The (1,3,3) output convolution of UNet2D5_dsbn (unet2d5_dsbn.py:301-308: nn.Conv3d(ft0, class_num, (1,3,3), padding (0,1,1)))
// on CUDA cores.  With 2-8 classes against 16-32 feature channels this layer is 0.5 % of the network's FLOPs; padded to a
// 16-wide tensor-core N it costs 9 shared-memory-fed MMAs per 128 voxels and a zero-padded 16-channel bf16 copy of the
// logit gradient (65 + 61 + 55 us per pass at 4x32x128x128).  Here a thread owns VPT consecutive voxels of a row, so every
// weight fetched from shared memory feeds VPT FMAs per class / channel and the kernels are FMA-issue bound, not LDS bound:
//   forward : 9 taps x ft0 channels from a bf16 halo tile in shared memory, class_num accumulators, fp32 NCDHW logits
//   dgrad   : 9 taps x class_num logit gradients from an fp32 halo tile, ft0 accumulators -> C8-planar bf16; the same pass
//             writes the 8-channel bf16 copy of the logit gradient that the tensor-core wgrad reads and the bias gradient
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

constexpr int kTH = 8, kTX = 16;                 // thread grid of a block: 8 rows x 16 threads = 128 threads
constexpr int kThreadsH = kTH * kTX;
constexpr int kHH = kTH + 2;
constexpr int kMaxClasses = 8;

struct HeadParams {
    const bf16x8* x;           // forward input
    int x_c8tot, x_c8off;
    const float* w;            // [classes][cin][1][3][3]
    const float* bias;         // forward
    float* logits;             // forward output [N][classes][D][H][W]
    const float* dlogits;      // dgrad input   [N][classes][D][H][W]
    bf16x8* g;                 // dgrad output, C8-planar
    int g_c8tot, g_c8off;
    bf16x8* dl8;               // bf16 copy of dlogits, classes padded to 8 (one channel group), C8-planar
    int dl_c8tot, dl_c8off;
    float* dbias;              // [classes], ACCUMULATED
    int N, D, H, W, cin, classes;
};

__device__ __forceinline__ float bf16_lane(const int4& v, int i) {        // channel i (0..7) of a packed bf16x8
    const uint32_t word = reinterpret_cast<const uint32_t*>(&v)[i >> 1];
    return __uint_as_float((i & 1) ? (word & 0xffff0000u) : (word << 16));
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad: dX[v][c] = sum_{kh,kw} sum_k dL[v - (kh-1, kw-1)][k] * W[k][c][kh][kw]
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int VPT>
__global__ void __launch_bounds__(kThreadsH) head_dgrad_kernel(const HeadParams P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    constexpr int kTW = kTX * VPT, kHW = kTW + 2;
    __shared__ float sdl[kMaxClasses][kHH][kHW + 1];
    __shared__ __align__(16) float sw[9 * kMaxClasses * CIN];      // [tap][k][c]
    __shared__ float sred[kThreadsH / 32][kMaxClasses];
    const int K = P.classes;
    const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
    const int nd = blockIdx.z;
    const int n = nd / P.D, d = nd - n * P.D;
    const int h0 = blockIdx.y * kTH, w0 = blockIdx.x * kTW;
    const int64_t HW = (int64_t)P.H * P.W;
    for (int i = threadIdx.x; i < 9 * K * CIN; i += kThreadsH) {
        const int c = i % CIN, k = (i / CIN) % K, tap = i / (CIN * K);
        sw[i] = __ldg(P.w + ((int64_t)k * CIN + c) * 9 + tap);
    }
    // halo tile: all loads of a batch of 8 are issued before the first shared-memory store (one DRAM latency per batch)
    for (int base = 0; base < K * kHH * kHW; base += 8 * kThreadsH) {
        float tmp[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * kThreadsH + threadIdx.x;
            const int cc = i % kHW, rr = (i / kHW) % kHH, k = i / (kHW * kHH);
            const int h = h0 + rr - 1, w = w0 + cc - 1;
            tmp[u] = 0.0f;
            if (k < K && h >= 0 && h < P.H && w >= 0 && w < P.W)
                tmp[u] = __ldg(P.dlogits + (((int64_t)n * K + k) * P.D + d) * HW + (int64_t)h * P.W + w);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * kThreadsH + threadIdx.x;
            if (i < K * kHH * kHW) (&sdl[0][0][0])[(i / kHW) * (kHW + 1) + i % kHW] = tmp[u];
        }
    }
    __syncthreads();
    const int h = h0 + ty;                                         // this thread's voxels: (h, w0 + tx + kTX * v), v < VPT:
                                                                   // consecutive threads = consecutive voxels (no bank conflicts,
                                                                   // coalesced stores)
    float acc[VPT][CIN];
#pragma unroll
    for (int v = 0; v < VPT; ++v)
#pragma unroll
        for (int c = 0; c < CIN; ++c) acc[v][c] = 0.0f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        for (int k = 0; k < K; ++k) {
            float vals[VPT][3];                                    // halo columns tx + kTX*v + (2 - kw)
#pragma unroll
            for (int v = 0; v < VPT; ++v)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) vals[v][kw] = sdl[k][ty + 2 - kh][tx + kTX * v + 2 - kw];
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const float4* wt = reinterpret_cast<const float4*>(sw + ((kh * 3 + kw) * K + k) * CIN);
#pragma unroll
                for (int q4 = 0; q4 < CIN / 4; ++q4) {
                    const float4 ww = wt[q4];                      // warp-uniform address: broadcast
#pragma unroll
                    for (int v = 0; v < VPT; ++v) {
                        const float s = vals[v][kw];
                        acc[v][4 * q4 + 0] = fmaf(s, ww.x, acc[v][4 * q4 + 0]);
                        acc[v][4 * q4 + 1] = fmaf(s, ww.y, acc[v][4 * q4 + 1]);
                        acc[v][4 * q4 + 2] = fmaf(s, ww.z, acc[v][4 * q4 + 2]);
                        acc[v][4 * q4 + 3] = fmaf(s, ww.w, acc[v][4 * q4 + 3]);
                    }
                }
            }
        }
    }
    float bsum[kMaxClasses];
#pragma unroll
    for (int k = 0; k < kMaxClasses; ++k) bsum[k] = 0.0f;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int w = w0 + tx + kTX * v;
        const bool valid = h < P.H && w < P.W;
        float centre[kMaxClasses];
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) {
            centre[k] = (k < K && valid) ? sdl[k][ty + 1][tx + kTX * v + 1] : 0.0f;
            bsum[k] += centre[k];
        }
        if (valid) {
            const int64_t hw = (int64_t)h * P.W + w;
#pragma unroll
            for (int j = 0; j < CIN / 8; ++j)
                st_bf16x8(P.g + ((int64_t)nd * P.g_c8tot + P.g_c8off + j) * HW + hw, &acc[v][8 * j]);
            if (P.dl8 != nullptr) st_bf16x8(P.dl8 + ((int64_t)nd * P.dl_c8tot + P.dl_c8off) * HW + hw, centre);
        }
    }
    if (P.dbias != nullptr) {
        // bias gradient: sum of the logit gradient over the tile's voxels
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < kMaxClasses; ++k) {
            if (k < K) {
                const float s = warp_sum(bsum[k]);
                if (lane == 0) sred[wid][k] = s;
            }
        }
        __syncthreads();
        if (threadIdx.x < K) {
            float s = 0.0f;
#pragma unroll
            for (int r = 0; r < kThreadsH / 32; ++r) s += sred[r][threadIdx.x];
            atomicAdd(P.dbias + threadIdx.x, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: logits[v][k] = bias[k] + sum_{kh,kw} sum_c x[v + (kh-1, kw-1)][c] * W[k][c][kh][kw]
// ---------------------------------------------------------------------------------------------------------------
template <int CIN, int KP>                      // KP: classes padded to 2, 4 or 8 accumulators per voxel
__global__ void __launch_bounds__(kThreadsH) head_fwd_kernel(const HeadParams P) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    constexpr int VPT = 4, kTW = kTX * VPT, kHW = kTW + 2;
    extern __shared__ __align__(16) uint8_t head_smem[];
    int4* sx = reinterpret_cast<int4*>(head_smem);                                        // [CIN/8][kHH][kHW]
    float* sw = reinterpret_cast<float*>(head_smem + (size_t)(CIN / 8) * kHH * kHW * 16); // [tap][c][KP]
    const int K = P.classes;
    const int tx = threadIdx.x % kTX, ty = threadIdx.x / kTX;
    const int nd = blockIdx.z;
    const int n = nd / P.D, d = nd - n * P.D;
    const int h0 = blockIdx.y * kTH, w0 = blockIdx.x * kTW;
    const int64_t HW = (int64_t)P.H * P.W;
    for (int i = threadIdx.x; i < 9 * CIN * KP; i += kThreadsH) {
        const int k = i % KP, c = (i / KP) % CIN, tap = i / (KP * CIN);
        sw[i] = k < K ? __ldg(P.w + ((int64_t)k * CIN + c) * 9 + tap) : 0.0f;
    }
    // halo tile: all loads of a batch are issued before the first shared-memory store (one DRAM latency per batch)
    constexpr int kTotalX = (CIN / 8) * kHH * kHW, kBatch = 6;
    for (int base = 0; base < kTotalX; base += kBatch * kThreadsH) {
        int4 tmp[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = base + u * kThreadsH + threadIdx.x;
            const int cc = i % kHW, rr = (i / kHW) % kHH, j = i / (kHW * kHH);
            const int h = h0 + rr - 1, w = w0 + cc - 1;
            tmp[u] = make_int4(0, 0, 0, 0);
            if (i < kTotalX && h >= 0 && h < P.H && w >= 0 && w < P.W)
                tmp[u] = ld_stream16(P.x + ((int64_t)nd * P.x_c8tot + P.x_c8off + j) * HW + (int64_t)h * P.W + w);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = base + u * kThreadsH + threadIdx.x;
            if (i < kTotalX) sx[i] = tmp[u];
        }
    }
    __syncthreads();
    float acc[VPT][KP];
#pragma unroll
    for (int v = 0; v < VPT; ++v)
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[v][k] = 0.0f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int j = 0; j < CIN / 8; ++j) {
            // this thread's voxels: (h, w0 + tx + kTX * v): consecutive threads read consecutive 16-byte vectors
            int4 xs[VPT][3];                                       // halo columns tx + kTX*v + kw, 8 channels each
#pragma unroll
            for (int v = 0; v < VPT; ++v)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) xs[v][kw] = sx[(j * kHH + ty + kh) * kHW + tx + kTX * v + kw];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float* wc = sw + ((kh * 3 + kw) * CIN + j * 8 + i) * KP;   // warp-uniform address: broadcast
                    float wk[KP];
                    if (KP == 2) {
                        const float2 t = *reinterpret_cast<const float2*>(wc);
                        wk[0] = t.x; wk[1] = t.y;
                    } else {
#pragma unroll
                        for (int q4 = 0; q4 < KP / 4; ++q4) {
                            const float4 t = *reinterpret_cast<const float4*>(wc + 4 * q4);
                            wk[4 * q4] = t.x; wk[4 * q4 + 1] = t.y; wk[4 * q4 + 2] = t.z; wk[4 * q4 + 3] = t.w;
                        }
                    }
#pragma unroll
                    for (int v = 0; v < VPT; ++v)
#pragma unroll
                    {
                        const float f = bf16_lane(xs[v][kw], i);
#pragma unroll
                        for (int k = 0; k < KP; ++k) acc[v][k] = fmaf(f, wk[k], acc[v][k]);
                    }
                }
            }
        }
    }
    const int h = h0 + ty;
    if (h < P.H) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < K) {
                const float b = __ldg(P.bias + k);
                float* out = P.logits + (((int64_t)n * K + k) * P.D + d) * HW + (int64_t)h * P.W + w0 + tx;
#pragma unroll
                for (int v = 0; v < VPT; ++v)
                    if (w0 + tx + kTX * v < P.W) out[kTX * v] = acc[v][k] + b;
            }
        }
    }
}

template <int CIN, int KP>
int launch_head_fwd_kp(const HeadParams& P, cudaStream_t st) {
    constexpr int kTW = kTX * 4, kHW = kTW + 2;
    const int smem = (CIN / 8) * kHH * kHW * 16 + 9 * CIN * KP * 4;
    FPL_CHECK_CUDA(cudaFuncSetAttribute(head_fwd_kernel<CIN, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid((P.W + kTW - 1) / kTW, (P.H + kTH - 1) / kTH, P.N * P.D);
    fpl_launch(head_fwd_kernel<CIN, KP>, grid, kThreadsH, smem, st, P);
    return 0;
}

template <int CIN>
int launch_head_fwd(const HeadParams& P, cudaStream_t st) {
    if (P.classes <= 2) return launch_head_fwd_kp<CIN, 2>(P, st);
    if (P.classes <= 4) return launch_head_fwd_kp<CIN, 4>(P, st);
    return launch_head_fwd_kp<CIN, 8>(P, st);
}

int check_common(const char* who, int n, int d, int h, int w, int cin, int classes) {
    FPL_REQUIRE(cin == 16 || cin == 32, "%s: cin=%d must be 16 or 32", who, cin);
    FPL_REQUIRE(classes >= 1 && classes <= kMaxClasses, "%s: class_num=%d not in [1,%d]", who, classes, kMaxClasses);
    FPL_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && (int64_t)n * d <= 65535, "%s: bad geometry %dx%dx%dx%d", who, n, d, h, w);
    return 0;
}

}  // namespace

bool fpl_head_fwd_tc_eligible(int h, int w, int cin, int classes);
bool fpl_head_dgrad_tc_eligible(int h, int w, int cin, int classes);
int fpl_head_dgrad_tc_launch(const float* dlogits, const float* w, void* g, int g_c8tot, int g_c8off, void* dl8, int dl_c8tot,
                             int dl_c8off, float* dbias, int n, int d, int h, int w_, int cin, int classes, void* stream);
int fpl_head_fwd_tc_launch(const void* x, int x_c8tot, int x_c8off, const float* w, const float* bias, float* logits, int n, int d,
                           int h, int w_, int cin, int classes, void* stream);

extern "C" int fpl_head_dgrad(const float* dlogits, const float* w, void* g, int g_c8tot, int g_c8off, void* dl8, int dl_c8tot,
                              int dl_c8off, float* dbias, int n, int d, int h, int w_, int cin, int classes, void* stream) {
    if (int rc = check_common("fpl_head_dgrad", n, d, h, w_, cin, classes)) return rc;
    FPL_REQUIRE(dlogits != nullptr && w != nullptr && g != nullptr, "fpl_head_dgrad: NULL buffer");
    if (fpl_head_dgrad_tc_eligible(h, w_, cin, classes)) {     // tensor-core form with the operand built in shared memory (head_tc.cu)
        if (int rc = fpl_head_dgrad_tc_launch(dlogits, w, g, g_c8tot, g_c8off, dl8, dl_c8tot, dl_c8off, dbias, n, d, h, w_, cin, classes,
                                              stream)) return rc;
        FPL_LAUNCH_CHECK();
        return 0;
    }
    HeadParams P = {};
    P.w = w; P.dlogits = dlogits; P.g = (bf16x8*)g; P.g_c8tot = g_c8tot; P.g_c8off = g_c8off;
    P.dl8 = (bf16x8*)dl8; P.dl_c8tot = dl_c8tot; P.dl_c8off = dl_c8off; P.dbias = dbias;
    P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.classes = classes;
    if (cin == 16) {
        dim3 grid((w_ + kTX * 4 - 1) / (kTX * 4), (h + kTH - 1) / kTH, n * d);
        fpl_launch(head_dgrad_kernel<16, 4>, grid, kThreadsH, 0, (cudaStream_t)stream, P);
    } else {
        dim3 grid((w_ + kTX * 2 - 1) / (kTX * 2), (h + kTH - 1) / kTH, n * d);
        fpl_launch(head_dgrad_kernel<32, 2>, grid, kThreadsH, 0, (cudaStream_t)stream, P);
    }
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_head_fwd(const void* x, int x_c8tot, int x_c8off, const float* w, const float* bias, float* logits, int n,
                            int d, int h, int w_, int cin, int classes, void* stream) {
    if (int rc = check_common("fpl_head_fwd", n, d, h, w_, cin, classes)) return rc;
    FPL_REQUIRE(x != nullptr && w != nullptr && bias != nullptr && logits != nullptr, "fpl_head_fwd: NULL buffer");
    if (fpl_head_fwd_tc_eligible(h, w_, cin, classes)) {       // tensor-core "scatter" form (head_tc.cu)
        if (int rc = fpl_head_fwd_tc_launch(x, x_c8tot, x_c8off, w, bias, logits, n, d, h, w_, cin, classes, stream)) return rc;
        FPL_LAUNCH_CHECK();
        return 0;
    }
    HeadParams P = {};
    P.x = (const bf16x8*)x; P.x_c8tot = x_c8tot; P.x_c8off = x_c8off; P.w = w; P.bias = bias; P.logits = logits;
    P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.classes = classes;
    int rc = cin == 16 ? launch_head_fwd<16>(P, (cudaStream_t)stream) : launch_head_fwd<32>(P, (cudaStream_t)stream);
    if (rc) return rc;
    FPL_LAUNCH_CHECK();
    return 0;
}
