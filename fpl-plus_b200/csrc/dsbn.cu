// DSBN BatchNorm3d + PReLU + Dropout + MaxPool, forward and backward, on C8-planar bf16.
// Replaces nn.BatchNorm3d (selected domain, net_run_dsbn/dsbn.py:54-57), nn.PReLU,
// nn.Dropout and nn.MaxPool3d of PyMIC/pymic/net/net3d/unet2d5_dsbn.py:75-81,104-117.
// All kernels are HBM-streaming: one 16-byte vector (8 channels of one voxel) per access,
// consecutive threads on consecutive voxels of one (n, d, channel-group) plane.
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------
// finalize: batch statistics -> scale/shift, running statistics update
// ------------------------------------------------------------------------------------
__global__ void dsbn_finalize_kernel(const double* __restrict__ stats, double inv_count, double unbias,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* running_mean, float* running_var, long long* nbt,
                                     float momentum, float eps, int training,
                                     float* scale, float* shift, float* save_mean, float* save_invstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        float mean, invstd;
        if (training) {
            double m = stats[c] * inv_count;
            double var = stats[C + c] * inv_count - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)eps));
            if (running_mean != nullptr) {
                running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)(var * unbias);
            }
        } else {
            mean = running_mean[c];
            invstd = rsqrtf(running_var[c] + eps);
            invstd = 1.0f / sqrtf(running_var[c] + eps);
        }
        float g = gamma[c] * invstd;
        scale[c] = g;
        shift[c] = beta[c] - mean * g;
        save_mean[c] = mean;
        save_invstd[c] = invstd;
    }
    if (training && nbt != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
}

// ------------------------------------------------------------------------------------
// forward: a = dropout(prelu(y*scale+shift)), optional fused max-pool
// ------------------------------------------------------------------------------------
// optional fused "finalize" (statistics -> scale/shift, running-stat update) executed in the prologue of
// the activation kernels: every block derives the affine map of ITS 8 channels, one block per channel
// group publishes it (for backward) and updates the selected domain's running statistics
struct BnFinalize {
    const double* stats;        // non-NULL (training) or running_mean non-NULL (eval) enables the fused path
    double inv_count, unbias;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* nbt;
    float momentum, eps;
    int training, enabled;
    float* scale;
    float* shift;
    float* save_mean;
    float* save_invstd;
};

struct ActParams {
    BnFinalize fin;
    const bf16x8* y;            // dense [N][D][C8][H][W]
    const float* scale;
    const float* shift;
    const float* slope;
    bf16x8* a;
    int a_c8tot, a_c8off;
    bf16x8* pooled;
    int p_c8tot, p_c8off;
    uint2* pool_idx;            // 8 x uint8 per pooled vector
    int pool_kd;
    float drop_p;
    const uint2* drop_mask;     // 8 x uint8 per vector, dense order
    uint64_t seed, offset;
    const unsigned long long* seed_dev;   // optional device-side addend to the seed (CUDA-graph replays)
    int N, D, C8, H, W;
};

__device__ __forceinline__ void load_affine(const float* scale, const float* shift, int c8, float* sc, float* sh) {
    const float4* s4 = reinterpret_cast<const float4*>(scale + c8 * 8);
    const float4* h4 = reinterpret_cast<const float4*>(shift + c8 * 8);
    float4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(h4), d = __ldg(h4 + 1);
    sc[0] = a.x; sc[1] = a.y; sc[2] = a.z; sc[3] = a.w; sc[4] = b.x; sc[5] = b.y; sc[6] = b.z; sc[7] = b.w;
    sh[0] = c.x; sh[1] = c.y; sh[2] = c.z; sh[3] = c.w; sh[4] = d.x; sh[5] = d.y; sh[6] = d.z; sh[7] = d.w;
}

__device__ __forceinline__ void bn_prologue(const BnFinalize& F, int C, int c8, bool writer, float* sc, float* sh) {
    __shared__ float s_aff[16];
    if (threadIdx.x < 8) {
        const int c = c8 * 8 + threadIdx.x;
        float mean, invstd;
        if (F.training) {
            double m = F.stats[c] * F.inv_count;
            double var = F.stats[C + c] * F.inv_count - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            invstd = (float)(1.0 / sqrt(var + (double)F.eps));
            if (writer && F.running_mean != nullptr) {
                F.running_mean[c] = (1.0f - F.momentum) * F.running_mean[c] + F.momentum * mean;
                F.running_var[c] = (1.0f - F.momentum) * F.running_var[c] + F.momentum * (float)(var * F.unbias);
            }
        } else {
            mean = F.running_mean[c];
            invstd = 1.0f / sqrtf(F.running_var[c] + F.eps);
        }
        const float g = F.gamma[c] * invstd;
        s_aff[threadIdx.x] = g;
        s_aff[8 + threadIdx.x] = F.beta[c] - mean * g;
        if (writer) {
            F.scale[c] = g;
            F.shift[c] = F.beta[c] - mean * g;
            F.save_mean[c] = mean;
            F.save_invstd[c] = invstd;
            if (c == 0 && F.training && F.nbt != nullptr) *F.nbt += 1;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = s_aff[i]; sh[i] = s_aff[8 + i]; }
}

__device__ __forceinline__ uint32_t keep_bits(const uint2* mask, uint64_t seed, uint64_t offset, int64_t vec, float p) {
    if (mask != nullptr) {
        uint2 m = __ldg(mask + vec);
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            bits |= (((m.x >> (8 * i)) & 0xffu) ? 1u : 0u) << i;
            bits |= (((m.y >> (8 * i)) & 0xffu) ? 1u : 0u) << (4 + i);
        }
        return bits;
    }
    return dropout_keep8(seed, offset, (uint64_t)vec, p);
}

// grid: (chunks of H*W, N*D*C8 planes): a block stays inside one channel group, so the affine
// parameters are loaded once and the index math is 32-bit
__global__ void __launch_bounds__(kThreads) dsbn_act_fwd_kernel(ActParams P) {
    const int HW = P.H * P.W;
    const int plane = blockIdx.y;
    const int c8 = plane % P.C8;
    const int nd = plane / P.C8;
    const float slope = __ldg(P.slope);
    const bool drop = P.drop_p > 0.0f;
    const float keep_scale = drop ? 1.0f / (1.0f - P.drop_p) : 1.0f;
    const uint64_t seed = P.seed + (P.seed_dev != nullptr ? (uint64_t)__ldg(P.seed_dev) : 0ull);
    float sc[8], sh[8];
    if (P.fin.enabled) bn_prologue(P.fin, P.C8 * 8, c8, blockIdx.x == 0 && nd == 0, sc, sh);
    else load_affine(P.scale, P.shift, c8, sc, sh);
    const bf16x8* src = P.y + (int64_t)plane * HW;
    bf16x8* dst = P.a + ((int64_t)nd * P.a_c8tot + P.a_c8off + c8) * HW;
    const int stride = gridDim.x * blockDim.x;
    for (int hw = blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += 2 * stride) {
        const int hw2 = hw + stride;
        const bool two = hw2 < HW;
        int4 r0 = ld_stream16(src + hw), r1 = make_int4(0, 0, 0, 0);
        if (two) r1 = ld_stream16(src + hw2);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (k == 1 && !two) break;
            const int h = k == 0 ? hw : hw2;
            float f[8];
            bf16x8_to_float(*reinterpret_cast<bf16x8*>(k == 0 ? &r0 : &r1), f);
            uint32_t keep = drop ? keep_bits(P.drop_mask, seed, P.offset, (int64_t)plane * HW + h, P.drop_p) : 0xffu;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float z = fmaf(f[i], sc[i], sh[i]);
                float a = z > 0.0f ? z : slope * z;
                f[i] = ((keep >> i) & 1u) ? a * keep_scale : 0.0f;
            }
            st_bf16x8(dst + h, f);
        }
    }
}

// one thread per pooled output vector; reads the 2x2x2 (or 1x2x2) window, writes the full
// resolution activations, the pooled max and the 3-bit argmax code (first max wins, scan
// order d,h,w as in torch's max_pool3d).
// grid: (chunks of H2*W2, N*D2*C8 pooled planes)
__global__ void __launch_bounds__(kThreads) dsbn_act_pool_fwd_kernel(ActParams P) {
    const int kd = P.pool_kd;
    const int D2 = P.D / kd, H2 = P.H / 2, W2 = P.W / 2;
    const int HW = P.H * P.W, HW2 = H2 * W2;
    const int pplane = blockIdx.y;
    const int c8 = pplane % P.C8;
    const int nd2 = pplane / P.C8;
    const int d2 = nd2 % D2, n = nd2 / D2;
    const float slope = __ldg(P.slope);
    float sc[8], sh[8];
    if (P.fin.enabled) bn_prologue(P.fin, P.C8 * 8, c8, blockIdx.x == 0 && nd2 == 0, sc, sh);
    else load_affine(P.scale, P.shift, c8, sc, sh);
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < HW2; v += gridDim.x * blockDim.x) {
        const int h2 = v / W2, w2 = v - h2 * W2;
        float best[8];
        uint32_t code[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; code[i] = 0; }
        int4 raw[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {                      // all loads of the 2x2x2 window first
            const int dd = q >> 2, hh = (q >> 1) & 1, ww = q & 1;
            if (dd < kd) {
                const int64_t nd = (int64_t)n * P.D + d2 * kd + dd;
                raw[q] = ld_stream16(P.y + (nd * P.C8 + c8) * HW + (h2 * 2 + hh) * P.W + w2 * 2 + ww);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int dd = q >> 2, hh = (q >> 1) & 1, ww = q & 1;
            if (dd < kd) {
                const int64_t nd = (int64_t)n * P.D + d2 * kd + dd;
                float f[8];
                bf16x8_to_float(*reinterpret_cast<bf16x8*>(&raw[q]), f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float z = fmaf(f[i], sc[i], sh[i]);
                    f[i] = z > 0.0f ? z : slope * z;
                }
                bf16x8 outv = float_to_bf16x8(f);
                st_bf16x8(P.a + (nd * P.a_c8tot + P.a_c8off + c8) * HW + (h2 * 2 + hh) * P.W + w2 * 2 + ww, f);
                // compare on the bf16-rounded values (what backward and the next layer see); first max wins
                bf16x8_to_float(outv, f);
                const uint32_t k = (uint32_t)((dd * 2 + hh) * 2 + ww);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (f[i] > best[i]) { best[i] = f[i]; code[i] = k; }
                }
            }
        }
        st_bf16x8(P.pooled + ((int64_t)nd2 * P.p_c8tot + P.p_c8off + c8) * HW2 + v, best);
        uint2 packed;
        packed.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
        packed.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
        P.pool_idx[(int64_t)pplane * HW2 + v] = packed;
    }
}

// ------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------
struct ActBwdParams {
    const bf16x8* y;
    const bf16x8* g1;
    int g1_c8tot, g1_c8off;
    const bf16x8* g_pool;
    int gp_c8tot, gp_c8off;
    const uint2* pool_idx;
    int pool_kd;
    const float* scale;
    const float* shift;
    const float* mean;
    const float* invstd;
    const float* slope;
    float drop_p;
    const uint2* drop_mask;
    uint64_t seed, offset;
    const unsigned long long* seed_dev;
    double* red;                // [2C+1]: sum dz, sum dz*xhat, dslope
    float* fin_dgamma;          // optional fused finalize of the apply pass (all may be NULL)
    float* fin_dbeta;
    float* fin_dslope;
    float* fin_dbias;
    bf16x8* dy;
    int training;
    double inv_count;
    int N, D, C8, H, W;
};

// grid: (chunks of H*W, N*D*C8 planes) so a block stays inside one channel group.  Two vectors per
// thread are in flight (all loads issued before any arithmetic); xhat = y*k1 + k0.
template <bool APPLY>
__global__ void __launch_bounds__(kThreads) dsbn_act_bwd_kernel(ActBwdParams P) {
    const int HW = P.H * P.W;
    const int plane = blockIdx.y;
    const int c8 = plane % P.C8;
    const int nd = plane / P.C8;
    const float slope = __ldg(P.slope);
    const bool drop = P.drop_p > 0.0f;
    const float keep_scale = drop ? 1.0f / (1.0f - P.drop_p) : 1.0f;
    const uint64_t seed = P.seed + (P.seed_dev != nullptr ? (uint64_t)__ldg(P.seed_dev) : 0ull);
    float sc[8], sh[8], k1[8], k0[8];
    load_affine(P.scale, P.shift, c8, sc, sh);
    load_affine(P.mean, P.invstd, c8, k0, k1);
#pragma unroll
    for (int i = 0; i < 8; ++i) k0[i] = -k0[i] * k1[i];
    float s1[8], s2[8], dsl = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.0f; s2[i] = 0.0f; }
    if (APPLY && blockIdx.x == 0 && nd == 0 && threadIdx.x < 8) {
        // one block per channel group publishes the parameter gradients (what fpl_dsbn_bwd_finalize does)
        const int c = c8 * 8 + threadIdx.x, C = P.C8 * 8;
        if (P.fin_dbeta != nullptr) P.fin_dbeta[c] += (float)P.red[c];
        if (P.fin_dgamma != nullptr) P.fin_dgamma[c] += (float)P.red[C + c];
        if (P.fin_dbias != nullptr && !P.training) P.fin_dbias[c] += __ldg(P.scale + c) * (float)P.red[c];
        if (c == 0 && P.fin_dslope != nullptr) P.fin_dslope[0] += (float)P.red[2 * C];
    }
    if (APPLY && P.training) {
        // fold the batch means into the apply constants: dy = sc*(dz - m1 - xhat*m2)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s1[i] = (float)(P.red[c8 * 8 + i] * P.inv_count);                 // m1
            s2[i] = (float)(P.red[P.C8 * 8 + c8 * 8 + i] * P.inv_count);      // m2
        }
    }
    const bf16x8* yp = P.y + (int64_t)plane * HW;
    const bf16x8* g1p = P.g1 != nullptr ? P.g1 + ((int64_t)nd * P.g1_c8tot + P.g1_c8off + c8) * HW : nullptr;
    // pooled-path pointers (2x2 in-plane window; kd = 1 or 2 planes per pooled plane)
    const int W2 = P.W >> 1, HW2 = (P.H >> 1) * W2;
    const bf16x8* gpp = nullptr;
    const uint2* idxp = nullptr;
    uint32_t code_d = 0;
    if (P.g_pool != nullptr) {
        const int kd = P.pool_kd, D2 = P.D / kd;
        const int n = nd / P.D, d = nd - n * P.D, d2 = d / kd;
        const int64_t pnd = (int64_t)n * D2 + d2;
        gpp = P.g_pool + (pnd * P.gp_c8tot + P.gp_c8off + c8) * HW2;
        idxp = P.pool_idx + (pnd * P.C8 + c8) * HW2;
        code_d = (uint32_t)(d - d2 * kd) * 4u;
    }
    bf16x8* dyp = APPLY ? P.dy + (int64_t)plane * HW : nullptr;
    const int stride = gridDim.x * blockDim.x;
    for (int hw0 = blockIdx.x * blockDim.x + threadIdx.x; hw0 < HW; hw0 += 2 * stride) {
        int hwk[2] = {hw0, hw0 + stride};
        const bool act[2] = {true, hwk[1] < HW};
        int4 ry[2], rg[2], rp[2];
        uint2 rc[2];
        uint32_t mycode[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            ry[k] = rg[k] = rp[k] = make_int4(0, 0, 0, 0);
            rc[k] = make_uint2(0, 0);
            mycode[k] = 0xffu;
            if (act[k]) {
                ry[k] = ld_stream16(yp + hwk[k]);
                if (g1p != nullptr) rg[k] = ld_stream16(g1p + hwk[k]);
                if (gpp != nullptr) {
                    const int h = hwk[k] / P.W, w = hwk[k] - h * P.W;
                    const int pix = (h >> 1) * W2 + (w >> 1);
                    mycode[k] = code_d + (uint32_t)((h & 1) * 2 + (w & 1));
                    rc[k] = __ldg(idxp + pix);
                    rp[k] = __ldg(reinterpret_cast<const int4*>(gpp + pix));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (!act[k]) continue;
            float yf[8], g[8], gp[8];
            bf16x8_to_float(*reinterpret_cast<bf16x8*>(&ry[k]), yf);
            bf16x8_to_float(*reinterpret_cast<bf16x8*>(&rg[k]), g);
            bf16x8_to_float(*reinterpret_cast<bf16x8*>(&rp[k]), gp);
            const uint32_t keep = drop ? keep_bits(P.drop_mask, seed, P.offset, (int64_t)plane * HW + hwk[k], P.drop_p) : 0xffu;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t cd = ((i < 4 ? rc[k].x : rc[k].y) >> (8 * (i & 3))) & 0xffu;
                float gi = g[i] + (cd == mycode[k] ? gp[i] : 0.0f);
                gi = ((keep >> i) & 1u) ? gi * keep_scale : 0.0f;
                const float z = fmaf(yf[i], sc[i], sh[i]);
                const float dz = z > 0.0f ? gi : gi * slope;
                const float xhat = fmaf(yf[i], k1[i], k0[i]);
                if (!APPLY) {
                    dsl += z > 0.0f ? 0.0f : z * gi;
                    s1[i] += dz;
                    s2[i] = fmaf(dz, xhat, s2[i]);
                } else {
                    o[i] = P.training ? sc[i] * (dz - s1[i] - xhat * s2[i]) : sc[i] * dz;
                }
            }
            if (APPLY) st_bf16x8(dyp + hwk[k], o);
        }
    }
    if (!APPLY) {
        __shared__ float sm[kThreads / 32][17];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float a = warp_sum(s1[i]), b = warp_sum(s2[i]);
            if (lane == 0) { sm[wid][i] = a; sm[wid][8 + i] = b; }
        }
        float c = warp_sum(dsl);
        if (lane == 0) sm[wid][16] = c;
        __syncthreads();
        if (threadIdx.x < 17) {
            float t = 0.0f;
#pragma unroll
            for (int k = 0; k < kThreads / 32; ++k) t += sm[k][threadIdx.x];
            int i = threadIdx.x;
            if (i < 8) atomicAdd(P.red + c8 * 8 + i, (double)t);
            else if (i < 16) atomicAdd(P.red + P.C8 * 8 + c8 * 8 + (i - 8), (double)t);
            else atomicAdd(P.red + 2 * P.C8 * 8, (double)t);
        }
    }
}

__global__ void dsbn_bwd_finalize_kernel(const double* __restrict__ red, const float* __restrict__ scale, int training,
                                         float* dgamma, float* dbeta, float* dslope, float* dbias_conv,
                                         const float* __restrict__ invstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        // dz is the gradient wrt the BN output: dbeta = sum dz, dgamma = sum dz*xhat
        if (dbeta != nullptr) dbeta[c] += (float)red[c];
        if (dgamma != nullptr) dgamma[c] += (float)red[C + c];
        // the conv bias feeds straight into BN: in training mode its gradient is identically 0
        // (sum_v dy = 0), in eval mode it is scale * sum dz
        if (dbias_conv != nullptr && !training) dbias_conv[c] += scale[c] * (float)red[c];
    }
    if (c == 0 && dslope != nullptr) dslope[0] += (float)red[2 * C];
}

int grid_for(int64_t work_items, int per_thread) {
    int64_t blocks = (work_items + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
    int64_t cap = (int64_t)FPL_NUM_SMS * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" int fpl_dsbn_finalize(const double* stats, int64_t count, const float* gamma, const float* beta,
                                 float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                 float momentum, float eps, int training, float* scale, float* shift,
                                 float* save_mean, float* save_invstd, int c, void* stream) {
    FPL_REQUIRE(c > 0 && c % 8 == 0, "fpl_dsbn_finalize: channels (%d) must be a positive multiple of 8", c);
    FPL_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean != nullptr && running_var != nullptr),
                "fpl_dsbn_finalize: missing statistics for training=%d", training);
    double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
    dsbn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        stats, 1.0 / (double)(count > 0 ? count : 1), unbias, gamma, beta, running_mean, running_var,
        (long long*)num_batches_tracked, momentum, eps, training, scale, shift, save_mean, save_invstd, c);
    FPL_LAUNCH_CHECK();
    return 0;
}

static int act_fwd_launch(const BnFinalize* fin, const void* y, const float* scale, const float* shift, const float* slope,
                                void* a, int a_c8tot, int a_c8off, void* pooled, int p_c8tot, int p_c8off,
                                uint8_t* pool_idx, int pool_kd, float drop_p, const uint8_t* drop_mask,
                                uint64_t seed, uint64_t offset, const uint64_t* seed_dev, int n, int d, int h, int w, int c,
                                void* stream) {
    FPL_REQUIRE(c > 0 && c % 8 == 0, "fpl_dsbn_act_fwd: channels (%d) must be a multiple of 8", c);
    FPL_REQUIRE(drop_p >= 0.0f && drop_p < 1.0f, "fpl_dsbn_act_fwd: dropout p=%f out of [0,1)", drop_p);
    ActParams P;
    if (fin != nullptr) { P.fin = *fin; P.fin.enabled = 1; } else P.fin.enabled = 0;
    P.y = (const bf16x8*)y; P.scale = scale; P.shift = shift; P.slope = slope;
    P.a = (bf16x8*)a; P.a_c8tot = a_c8tot; P.a_c8off = a_c8off;
    P.pooled = (bf16x8*)pooled; P.p_c8tot = p_c8tot; P.p_c8off = p_c8off; P.pool_idx = (uint2*)pool_idx;
    P.pool_kd = pool_kd; P.drop_p = drop_p; P.drop_mask = (const uint2*)drop_mask; P.seed = seed; P.offset = offset;
    P.seed_dev = (const unsigned long long*)seed_dev;
    P.N = n; P.D = d; P.C8 = c / 8; P.H = h; P.W = w;
    if (pooled != nullptr) {
        FPL_REQUIRE(pool_kd == 1 || pool_kd == 2, "fpl_dsbn_act_fwd: pool_kd must be 1 or 2");
        FPL_REQUIRE(h % 2 == 0 && w % 2 == 0 && d % pool_kd == 0, "fpl_dsbn_act_fwd: pooled dims must be even");
        FPL_REQUIRE(drop_p == 0.0f, "fpl_dsbn_act_fwd: dropout is not combined with pooling");
        FPL_REQUIRE(pool_idx != nullptr, "fpl_dsbn_act_fwd: pool_idx required");
        int64_t planes = (int64_t)n * (d / pool_kd) * (c / 8);
        FPL_REQUIRE(planes <= 65535, "fpl_dsbn_act_fwd: too many planes (%lld)", (long long)planes);
        int hw2 = (h / 2) * (w / 2);
        dim3 grid((hw2 + kThreads - 1) / kThreads, (unsigned)planes);
        dsbn_act_pool_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(P);
    } else {
        int64_t planes = (int64_t)n * d * (c / 8);
        FPL_REQUIRE(planes <= 65535, "fpl_dsbn_act_fwd: too many planes (%lld)", (long long)planes);
        int chunks = (h * w + kThreads * 4 - 1) / (kThreads * 4);
        dim3 grid(chunks < 1 ? 1 : chunks, (unsigned)planes);
        dsbn_act_fwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(P);
    }
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_dsbn_act_fwd(const void* y, const float* scale, const float* shift, const float* slope,
                                void* a, int a_c8tot, int a_c8off, void* pooled, int p_c8tot, int p_c8off,
                                uint8_t* pool_idx, int pool_kd, float drop_p, const uint8_t* drop_mask,
                                uint64_t seed, uint64_t offset, const uint64_t* seed_dev, int n, int d, int h, int w, int c,
                                void* stream) {
    return act_fwd_launch(nullptr, y, scale, shift, slope, a, a_c8tot, a_c8off, pooled, p_c8tot, p_c8off, pool_idx, pool_kd,
                          drop_p, drop_mask, seed, offset, seed_dev, n, d, h, w, c, stream);
}

extern "C" int fpl_dsbn_bn_act_fwd(const void* y, const double* stats, int64_t count, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                                   float eps, int training, float* scale, float* shift, float* save_mean,
                                   float* save_invstd, const float* slope, void* a, int a_c8tot, int a_c8off, void* pooled,
                                   int p_c8tot, int p_c8off, uint8_t* pool_idx, int pool_kd, float drop_p,
                                   const uint8_t* drop_mask, uint64_t seed, uint64_t offset, const uint64_t* seed_dev,
                                   int n, int d, int h, int w, int c, void* stream) {
    FPL_REQUIRE(training ? (stats != nullptr && count > 0) : (running_mean != nullptr && running_var != nullptr),
                "fpl_dsbn_bn_act_fwd: missing statistics for training=%d", training);
    FPL_REQUIRE(scale && shift && save_mean && save_invstd && gamma && beta, "fpl_dsbn_bn_act_fwd: NULL parameter arrays");
    BnFinalize F;
    F.stats = stats; F.inv_count = 1.0 / (double)(count > 0 ? count : 1);
    F.unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
    F.gamma = gamma; F.beta = beta; F.running_mean = running_mean; F.running_var = running_var;
    F.nbt = (long long*)num_batches_tracked; F.momentum = momentum; F.eps = eps; F.training = training; F.enabled = 1;
    F.scale = scale; F.shift = shift; F.save_mean = save_mean; F.save_invstd = save_invstd;
    return act_fwd_launch(&F, y, scale, shift, slope, a, a_c8tot, a_c8off, pooled, p_c8tot, p_c8off, pool_idx, pool_kd,
                          drop_p, drop_mask, seed, offset, seed_dev, n, d, h, w, c, stream);
}

static int fill_bwd(ActBwdParams& P, const void* y, const void* g1, int g1_c8tot, int g1_c8off, const void* g_pool,
                    int gp_c8tot, int gp_c8off, const uint8_t* pool_idx, int pool_kd, const float* scale,
                    const float* shift, const float* save_mean, const float* save_invstd, const float* slope,
                    float drop_p, const uint8_t* drop_mask, uint64_t seed, uint64_t offset, const uint64_t* seed_dev,
                    int n, int d, int h, int w, int c) {
    FPL_REQUIRE(c > 0 && c % 8 == 0, "dsbn bwd: channels (%d) must be a multiple of 8", c);
    FPL_REQUIRE(g1 != nullptr || g_pool != nullptr, "dsbn bwd: no incoming gradient");
    if (g_pool != nullptr) {
        FPL_REQUIRE(pool_idx != nullptr && (pool_kd == 1 || pool_kd == 2), "dsbn bwd: pool_idx/pool_kd invalid");
        FPL_REQUIRE(h % 2 == 0 && w % 2 == 0 && d % pool_kd == 0, "dsbn bwd: pooled dims must be even");
    }
    P.y = (const bf16x8*)y; P.g1 = (const bf16x8*)g1; P.g1_c8tot = g1_c8tot; P.g1_c8off = g1_c8off;
    P.g_pool = (const bf16x8*)g_pool; P.gp_c8tot = gp_c8tot; P.gp_c8off = gp_c8off;
    P.pool_idx = (const uint2*)pool_idx; P.pool_kd = pool_kd > 0 ? pool_kd : 2;
    P.scale = scale; P.shift = shift; P.mean = save_mean; P.invstd = save_invstd; P.slope = slope;
    P.drop_p = drop_p; P.drop_mask = (const uint2*)drop_mask; P.seed = seed; P.offset = offset;
    P.seed_dev = (const unsigned long long*)seed_dev;
    P.red = nullptr; P.dy = nullptr; P.training = 1;
    P.fin_dgamma = P.fin_dbeta = P.fin_dslope = P.fin_dbias = nullptr; P.inv_count = 1.0 / ((double)n * d * h * w);
    P.N = n; P.D = d; P.C8 = c / 8; P.H = h; P.W = w;
    return 0;
}

static dim3 bwd_grid(int n, int d, int h, int w, int c) {
    int64_t hw = (int64_t)h * w;
    int chunks = (int)((hw + kThreads * 4 - 1) / (kThreads * 4));
    if (chunks < 1) chunks = 1;
    return dim3(chunks, n * d * (c / 8));
}

extern "C" int fpl_dsbn_act_bwd_reduce(const void* y, const void* g1, int g1_c8tot, int g1_c8off, const void* g_pool,
                                       int gp_c8tot, int gp_c8off, const uint8_t* pool_idx, int pool_kd,
                                       const float* scale, const float* shift, const float* save_mean,
                                       const float* save_invstd, const float* slope, float drop_p,
                                       const uint8_t* drop_mask, uint64_t seed, uint64_t offset,
                                       const uint64_t* seed_dev, double* red, int n, int d, int h, int w, int c,
                                       void* stream) {
    ActBwdParams P;
    int rc = fill_bwd(P, y, g1, g1_c8tot, g1_c8off, g_pool, gp_c8tot, gp_c8off, pool_idx, pool_kd, scale, shift,
                      save_mean, save_invstd, slope, drop_p, drop_mask, seed, offset, seed_dev, n, d, h, w, c);
    if (rc) return rc;
    FPL_REQUIRE((int64_t)n * d * (c / 8) <= 65535, "dsbn bwd: too many planes (%lld)", (long long)n * d * (c / 8));
    P.red = red;
    dsbn_act_bwd_kernel<false><<<bwd_grid(n, d, h, w, c), kThreads, 0, (cudaStream_t)stream>>>(P);
    FPL_LAUNCH_CHECK();
    return 0;
}

static int act_bwd_apply_launch(float* dgamma, float* dbeta, float* dslope, float* dbias_conv, const void* y, const void* g1, int g1_c8tot, int g1_c8off, const void* g_pool,
                                      int gp_c8tot, int gp_c8off, const uint8_t* pool_idx, int pool_kd,
                                      const float* scale, const float* shift, const float* save_mean,
                                      const float* save_invstd, const float* slope, float drop_p,
                                      const uint8_t* drop_mask, uint64_t seed, uint64_t offset,
                                      const uint64_t* seed_dev, const double* red, int training, void* dy, int n, int d,
                                      int h, int w, int c, void* stream) {
    ActBwdParams P;
    int rc = fill_bwd(P, y, g1, g1_c8tot, g1_c8off, g_pool, gp_c8tot, gp_c8off, pool_idx, pool_kd, scale, shift,
                      save_mean, save_invstd, slope, drop_p, drop_mask, seed, offset, seed_dev, n, d, h, w, c);
    if (rc) return rc;
    FPL_REQUIRE((int64_t)n * d * (c / 8) <= 65535, "dsbn bwd: too many planes (%lld)", (long long)n * d * (c / 8));
    P.red = const_cast<double*>(red);
    P.dy = (bf16x8*)dy;
    P.training = training;
    P.fin_dgamma = dgamma; P.fin_dbeta = dbeta; P.fin_dslope = dslope; P.fin_dbias = dbias_conv;
    dsbn_act_bwd_kernel<true><<<bwd_grid(n, d, h, w, c), kThreads, 0, (cudaStream_t)stream>>>(P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_dsbn_act_bwd_apply(const void* y, const void* g1, int g1_c8tot, int g1_c8off, const void* g_pool,
                                      int gp_c8tot, int gp_c8off, const uint8_t* pool_idx, int pool_kd,
                                      const float* scale, const float* shift, const float* save_mean,
                                      const float* save_invstd, const float* slope, float drop_p,
                                      const uint8_t* drop_mask, uint64_t seed, uint64_t offset,
                                      const uint64_t* seed_dev, const double* red, int training, void* dy, int n, int d,
                                      int h, int w, int c, void* stream) {
    return act_bwd_apply_launch(nullptr, nullptr, nullptr, nullptr, y, g1, g1_c8tot, g1_c8off, g_pool, gp_c8tot, gp_c8off, pool_idx, pool_kd, scale, shift, save_mean, save_invstd, slope, drop_p, drop_mask, seed, offset, seed_dev, red, training, dy, n, d, h, w, c, stream);
}

extern "C" int fpl_dsbn_act_bwd_apply_fin(const void* y, const void* g1, int g1_c8tot, int g1_c8off, const void* g_pool,
                                      int gp_c8tot, int gp_c8off, const uint8_t* pool_idx, int pool_kd,
                                      const float* scale, const float* shift, const float* save_mean,
                                      const float* save_invstd, const float* slope, float drop_p,
                                      const uint8_t* drop_mask, uint64_t seed, uint64_t offset,
                                      const uint64_t* seed_dev, const double* red, int training, void* dy, int n, int d,
                                      int h, int w, int c, void* stream, float* dgamma, float* dbeta, float* dslope, float* dbias_conv) {
    return act_bwd_apply_launch(dgamma, dbeta, dslope, dbias_conv, y, g1, g1_c8tot, g1_c8off, g_pool, gp_c8tot, gp_c8off, pool_idx, pool_kd, scale, shift, save_mean, save_invstd, slope, drop_p, drop_mask, seed, offset, seed_dev, red, training, dy, n, d, h, w, c, stream);
}

extern "C" int fpl_dsbn_bwd_finalize(const double* red, const float* scale, const float* save_invstd, int training,
                                     float* dgamma, float* dbeta, float* dslope, float* dbias_conv, int c,
                                     void* stream) {
    dsbn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(red, scale, training, dgamma, dbeta,
                                                                                dslope, dbias_conv, save_invstd, c);
    FPL_LAUNCH_CHECK();
    return 0;
}
