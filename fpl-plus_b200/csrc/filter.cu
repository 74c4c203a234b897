// Pseudo-label filter and sliding-window stitching kernels (HBM streaming, fp32 NCDHW logits).
//   argmax labels        : net_run_dsbn/agent_seg.py:1049-1050
//   MC-dropout statistics: agent_seg.py:911-929
//   agreement weight     : data/get_pixel_weight.py:21-26 (+ io/nifty_dataset.py:165-168 folding)
//   window accumulate    : net_run_dsbn/infer_func.py:96-112, 202-219
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxK = 8;          // passes held in registers by mc_uncertainty_kernel
constexpr int kMaxKStream = 16;   // passes the streaming variant accepts (= UNet2D5_dsbn.kMaxMcRepeats)

int grid_for(int64_t items) {
    int64_t blocks = (items + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)FPL_NUM_SMS * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// The reference takes the argmax of the fp32 SOFTMAX PROBABILITIES (scipy.special.softmax = exp(x - max) / sum, then
// np.argmax: agent_seg.py:1049-1050), not of the logits: two logits closer than the fp32 resolution of their
// probabilities (|dz| < ~6e-8 around p = 0.5) collapse to a tie that the first index wins.  Same arithmetic here
// (accurate expf, sequential sum, IEEE division), so labels follow the reference through those ties.
template <int C>
__device__ __forceinline__ int argmax_c(const float* z) {
    float m = z[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
    float e[C], s = 0.0f;
#pragma unroll
    for (int c = 0; c < C; ++c) { e[c] = expf(z[c] - m); s += e[c]; }
    int am = 0;
    float best = __fdiv_rn(e[0], s);
#pragma unroll
    for (int c = 1; c < C; ++c) {
        const float p = __fdiv_rn(e[c], s);
        if (p > best) { best = p; am = c; }
    }
    return am;
}

template <int C>
__global__ void __launch_bounds__(kThreads) argmax_label_kernel(const float* __restrict__ logits, uint8_t* label, int B,
                                                               int64_t S4) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int64_t total = (int64_t)B * S4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        int64_t b = g / S4, s4 = g - b * S4;
        float4 zv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) zv[c] = ld_stream_f4(reinterpret_cast<const float4*>(logits) + (b * C + c) * S4 + s4);
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float z[C];
#pragma unroll
            for (int c = 0; c < C; ++c) z[c] = reinterpret_cast<const float*>(&zv[c])[j];
            packed |= (uint32_t)argmax_c<C>(z) << (8 * j);
        }
        reinterpret_cast<uint32_t*>(label)[g] = packed;
    }
}

struct PassPtrs {
    const float* p[kMaxKStream];
};

// out[0] += sum over classes and voxels of the population variance across the K passes of the
// softmax probabilities; out[1] += #voxels with -m*ln(m+1e-6) > 0.01, m = mean class-1 probability.
// KT > 0: the pass count is a compile-time constant (KT = 6 is the reference's hard-coded value, agent_seg.py:898) and,
// when KT * C <= 16, ALL K * C 16-byte loads of a voxel group are issued before the first use: with the runtime-K form
// the `k < K` predicate kept ~2 loads in flight per thread and the kernel ran at 2.0 TB/s, latency bound at 24 %
// occupancy (profiles/r02a_prof_filter_loss_summary.md).
template <int C, int KT>
__global__ void __launch_bounds__(kThreads) mc_uncertainty_kernel(PassPtrs ptrs, int K_rt, int64_t S4, double* out,
                                                                 float* umap) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    constexpr int KK = KT > 0 ? KT : kMaxK;
    constexpr bool kPreload = KT > 0 && KT * C <= 16;
    const int K = KT > 0 ? KT : K_rt;
    float var_acc = 0.0f;
    unsigned int cnt = 0;
    const float invK = 1.0f / (float)K;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < S4; g += (int64_t)gridDim.x * blockDim.x) {
        float prob[KK][C][4];
        float4 zall[kPreload ? KK : 1][C];
        if (kPreload) {
#pragma unroll
            for (int k = 0; k < KK; ++k)
#pragma unroll
                for (int c = 0; c < C; ++c) zall[k][c] = ld_stream_f4(reinterpret_cast<const float4*>(ptrs.p[k]) + c * S4 + g);
        }
#pragma unroll
        for (int k = 0; k < KK; ++k) {
            if (k < K) {
                float4 zv[C];
#pragma unroll
                for (int c = 0; c < C; ++c)
                    zv[c] = kPreload ? zall[kPreload ? k : 0][c] : ld_stream_f4(reinterpret_cast<const float4*>(ptrs.p[k]) + c * S4 + g);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (C == 2) {
                        // two classes: exp(z_max - z_max) is exactly 1, so one expf per voxel gives the same bits as
                        // the general form below (12 expf per voxel at K = 6 were the pacing cost of this kernel)
                        const float z0 = reinterpret_cast<const float*>(&zv[0])[j], z1 = reinterpret_cast<const float*>(&zv[1])[j];
                        const bool first = z0 >= z1;
                        const float e = expf(first ? z1 - z0 : z0 - z1);
                        const float e0 = first ? 1.0f : e, e1 = first ? e : 1.0f;
                        const float s = e0 + e1;
                        prob[k][0][j] = e0 / s;
                        prob[k][1][j] = e1 / s;
                        continue;
                    }
                    float m = reinterpret_cast<const float*>(&zv[0])[j];
#pragma unroll
                    for (int c = 1; c < C; ++c) m = fmaxf(m, reinterpret_cast<const float*>(&zv[c])[j]);
                    float s = 0.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        float e = expf(reinterpret_cast<const float*>(&zv[c])[j] - m);
                        prob[k][c][j] = e;
                        s += e;
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) prob[k][c][j] = prob[k][c][j] / s;
                }
            }
        }
        float4 uo;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float mean1 = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float mean = 0.0f;
#pragma unroll
                for (int k = 0; k < KK; ++k)
                    if (k < K) mean += prob[k][c][j];
                mean *= invK;
                float m2 = 0.0f;
#pragma unroll
                for (int k = 0; k < KK; ++k)
                    if (k < K) { float dlt = prob[k][c][j] - mean; m2 = fmaf(dlt, dlt, m2); }
                var_acc += m2 * invK;
                if (c == 1) mean1 = mean;
            }
            float u = -1.0f * (mean1 * logf(mean1 + 1e-6f));
            cnt += (u > 0.01f) ? 1u : 0u;
            reinterpret_cast<float*>(&uo)[j] = u;
        }
        if (umap != nullptr) reinterpret_cast<float4*>(umap)[g] = uo;
    }
    __shared__ float sv[kThreads / 32];
    __shared__ unsigned int sc[kThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float v = warp_sum(var_acc);
    unsigned int c = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) { sv[wid] = v; sc[wid] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tv = 0.0, tc = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) { tv += (double)sv[k]; tc += (double)sc[k]; }
        atomicAdd(out, tv);
        atomicAdd(out + 1, tc);
    }
}

// K in (kMaxK, kMaxKStream]: the [K][C][4] register tile of the kernel above would spill, so the probabilities are
// recomputed: sweep 1 accumulates the per-class means, sweep 2 re-reads the logits (L1/L2 hits: the same thread read
// them a few hundred cycles earlier) and accumulates the squared deviations.  Same operation order as the register
// kernel (sum over k ascending, * 1/K, fma of squared deviations over k ascending), so both give identical bits.
template <int C>
__device__ __forceinline__ void mc_probs4(const float* base, int64_t S4, int64_t g, float (&prob)[C][4]) {
    float4 zv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) zv[c] = __ldg(reinterpret_cast<const float4*>(base) + c * S4 + g);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float m = reinterpret_cast<const float*>(&zv[0])[j];
#pragma unroll
        for (int c = 1; c < C; ++c) m = fmaxf(m, reinterpret_cast<const float*>(&zv[c])[j]);
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float e = expf(reinterpret_cast<const float*>(&zv[c])[j] - m);
            prob[c][j] = e;
            s += e;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) prob[c][j] = prob[c][j] / s;
    }
}

template <int C>
__global__ void __launch_bounds__(kThreads) mc_uncertainty_stream_kernel(PassPtrs ptrs, int K, int64_t S4, double* out,
                                                                        float* umap) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    float var_acc = 0.0f;
    unsigned int cnt = 0;
    const float invK = 1.0f / (float)K;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < S4; g += (int64_t)gridDim.x * blockDim.x) {
        float mean[C][4], m2[C][4], prob[C][4];
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) { mean[c][j] = 0.0f; m2[c][j] = 0.0f; }
        for (int k = 0; k < K; ++k) {
            mc_probs4<C>(ptrs.p[k], S4, g, prob);
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) mean[c][j] += prob[c][j];
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int j = 0; j < 4; ++j) mean[c][j] *= invK;
        for (int k = 0; k < K; ++k) {
            mc_probs4<C>(ptrs.p[k], S4, g, prob);
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) { float dlt = prob[c][j] - mean[c][j]; m2[c][j] = fmaf(dlt, dlt, m2[c][j]); }
        }
        float4 uo;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int c = 0; c < C; ++c) var_acc += m2[c][j] * invK;
            const float mean1 = mean[1][j];
            float u = -1.0f * (mean1 * logf(mean1 + 1e-6f));
            cnt += (u > 0.01f) ? 1u : 0u;
            reinterpret_cast<float*>(&uo)[j] = u;
        }
        if (umap != nullptr) reinterpret_cast<float4*>(umap)[g] = uo;
    }
    __shared__ float sv[kThreads / 32];
    __shared__ unsigned int sc[kThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float v = warp_sum(var_acc);
    unsigned int c = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) { sv[wid] = v; sc[wid] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tv = 0.0, tc = 0.0;
        for (int k = 0; k < kThreads / 32; ++k) { tv += (double)sv[k]; tc += (double)sc[k]; }
        atomicAdd(out, tv);
        atomicAdd(out + 1, tc);
    }
}

template <int C>
__global__ void __launch_bounds__(kThreads) agree_weight_kernel(const float* __restrict__ lt, const float* __restrict__ ls,
                                                               uint8_t* lab_t, uint8_t* lab_s, float* weight, int fold,
                                                               float image_weight, unsigned long long* out_count,
                                                               int64_t S4) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    unsigned int diff = 0;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < S4; g += (int64_t)gridDim.x * blockDim.x) {
        float4 a[C], b[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            a[c] = ld_stream_f4(reinterpret_cast<const float4*>(lt) + c * S4 + g);
            b[c] = ld_stream_f4(reinterpret_cast<const float4*>(ls) + c * S4 + g);
        }
        uint32_t pa = 0, pb = 0;
        float4 wv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float za[C], zb[C];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                za[c] = reinterpret_cast<const float*>(&a[c])[j];
                zb[c] = reinterpret_cast<const float*>(&b[c])[j];
            }
            int la = argmax_c<C>(za), lb = argmax_c<C>(zb);
            pa |= (uint32_t)la << (8 * j);
            pb |= (uint32_t)lb << (8 * j);
            float w = (la == lb) ? 1.0f : 0.5f;
            diff += (la != lb) ? 1u : 0u;
            if (fold) w = (w < 1.0f ? 0.0f : w) * image_weight;
            reinterpret_cast<float*>(&wv)[j] = w;
        }
        if (lab_t != nullptr) reinterpret_cast<uint32_t*>(lab_t)[g] = pa;
        if (lab_s != nullptr) reinterpret_cast<uint32_t*>(lab_s)[g] = pb;
        if (weight != nullptr) reinterpret_cast<float4*>(weight)[g] = wv;
    }
    if (out_count != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) diff += __shfl_xor_sync(0xffffffffu, diff, o);
        if ((threadIdx.x & 31) == 0 && diff) atomicAdd(out_count, (unsigned long long)diff);
    }
}

__global__ void __launch_bounds__(kThreads) window_accumulate_kernel(const float* __restrict__ patch, float* out,
                                                                    float* count, int BC, int vd, int vh, int vw, int d0,
                                                                    int h0, int w0, int pd, int ph, int pw, int flip_h,
                                                                    int flip_w, float scale) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int64_t total = (int64_t)BC * pd * ph * pw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int k = (int)(i % pw);
        int64_t t = i / pw;
        int j = (int)(t % ph); t /= ph;
        int ii = (int)(t % pd);
        int64_t bc = t / pd;
        int jj = flip_h ? ph - 1 - j : j;
        int kk = flip_w ? pw - 1 - k : k;
        int64_t o = ((bc * vd + d0 + ii) * vh + h0 + jj) * (int64_t)vw + w0 + kk;
        out[o] += scale * __ldg(patch + i);
        if (count != nullptr) count[o] += 1.0f;
    }
}

__global__ void __launch_bounds__(kThreads) window_normalize_kernel(float* out, const float* __restrict__ count,
                                                                   float scale, int64_t numel) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float v = out[i];
        if (count != nullptr) v = v / count[i];
        out[i] = v * scale;
    }
}

}  // namespace

#define FPL_DISPATCH_C(C_, ...)                                                     \
    switch (C_) {                                                                   \
        case 2: { constexpr int CC = 2; __VA_ARGS__; } break;                       \
        case 3: { constexpr int CC = 3; __VA_ARGS__; } break;                       \
        case 4: { constexpr int CC = 4; __VA_ARGS__; } break;                       \
        case 5: { constexpr int CC = 5; __VA_ARGS__; } break;                       \
        case 6: { constexpr int CC = 6; __VA_ARGS__; } break;                       \
        case 7: { constexpr int CC = 7; __VA_ARGS__; } break;                       \
        case 8: { constexpr int CC = 8; __VA_ARGS__; } break;                       \
        default: fpl_set_error("class_num %d not in [2,8]", C_); return 2;          \
    }

extern "C" int fpl_argmax_label(const float* logits, uint8_t* label, int b, int c, int64_t spatial, void* stream) {
    FPL_REQUIRE(spatial % 4 == 0, "fpl_argmax_label: spatial size %lld must be a multiple of 4", (long long)spatial);
    int64_t s4 = spatial / 4;
    FPL_DISPATCH_C(c, (fpl_launch(argmax_label_kernel<CC>, grid_for((int64_t)b * s4), kThreads, 0, (cudaStream_t)stream, logits, label, b, s4)));
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_mc_uncertainty(const float* const* h_logits_k, int k, int c, int64_t spatial, double* out,
                                  float* uncertainty_map, void* stream) {
    FPL_REQUIRE(k >= 1 && k <= kMaxKStream, "fpl_mc_uncertainty: K=%d passes not in [1,%d]", k, kMaxKStream);
    FPL_REQUIRE(spatial % 4 == 0, "fpl_mc_uncertainty: spatial size %lld must be a multiple of 4", (long long)spatial);
    PassPtrs ptrs;
    for (int i = 0; i < kMaxKStream; ++i) ptrs.p[i] = i < k ? h_logits_k[i] : nullptr;
    int64_t s4 = spatial / 4;
    if (k <= kMaxK) {
        if (k == 6) {
            FPL_DISPATCH_C(c, (fpl_launch(mc_uncertainty_kernel<CC, 6>, grid_for(s4), kThreads, 0, (cudaStream_t)stream, ptrs, k, s4, out, uncertainty_map)));
        } else {
            FPL_DISPATCH_C(c, (fpl_launch(mc_uncertainty_kernel<CC, 0>, grid_for(s4), kThreads, 0, (cudaStream_t)stream, ptrs, k, s4, out, uncertainty_map)));
        }
    } else {
        FPL_DISPATCH_C(c, (fpl_launch(mc_uncertainty_stream_kernel<CC>, grid_for(s4), kThreads, 0, (cudaStream_t)stream, ptrs, k, s4, out, uncertainty_map)));
    }
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_mc_uncertainty_max_passes(void) { return kMaxKStream; }

extern "C" int fpl_agree_weight(const float* logits_tgt, const float* logits_src, uint8_t* label_tgt,
                                uint8_t* label_src, float* weight, int fold_image_weight, float image_weight,
                                long long* out_count, int c, int64_t spatial, void* stream) {
    FPL_REQUIRE(spatial % 4 == 0, "fpl_agree_weight: spatial size %lld must be a multiple of 4", (long long)spatial);
    int64_t s4 = spatial / 4;
    FPL_DISPATCH_C(c, (fpl_launch(agree_weight_kernel<CC>, grid_for(s4), kThreads, 0, (cudaStream_t)stream, 
                          logits_tgt, logits_src, label_tgt, label_src, weight, fold_image_weight, image_weight,
                          (unsigned long long*)out_count, s4)));
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_window_accumulate(const float* patch, float* out, float* count, int b, int c, int vd, int vh,
                                     int vw, int d0, int h0, int w0, int pd, int ph, int pw, int flip_h, int flip_w,
                                     float scale, void* stream) {
    FPL_REQUIRE(d0 >= 0 && h0 >= 0 && w0 >= 0 && d0 + pd <= vd && h0 + ph <= vh && w0 + pw <= vw,
                "fpl_window_accumulate: window [%d+%d,%d+%d,%d+%d] outside volume [%d,%d,%d]", d0, pd, h0, ph, w0, pw,
                vd, vh, vw);
    int64_t total = (int64_t)b * c * pd * ph * pw;
    fpl_launch(window_accumulate_kernel, grid_for(total), kThreads, 0, (cudaStream_t)stream, 
        patch, out, count, b * c, vd, vh, vw, d0, h0, w0, pd, ph, pw, flip_h, flip_w, scale);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_window_normalize(float* out, const float* count, float scale, int64_t numel, void* stream) {
    fpl_launch(window_normalize_kernel, grid_for(numel), kThreads, 0, (cudaStream_t)stream, out, count, scale, numel);
    FPL_LAUNCH_CHECK();
    return 0;
}
