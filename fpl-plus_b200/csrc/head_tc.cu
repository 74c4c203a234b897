// Tensor-core forward of the network's output conv, nn.Conv3d(ft0 -> classes, k(1,3,3), p(0,1,1))
// (PyMIC/pymic/net/net3d/unet2d5_dsbn.py:268, 306).  With 2..7 classes the direct form is 288 FMAs per voxel on the CUDA
// cores (head_fwd_kernel: FMA-issue bound at 3.5x its HBM time, 53 us at 4x32x128x128) and an implicit GEMM with N = classes
// pays a full 40-cycle MMA per tap for 2 useful columns.  Here the ROW taps are gathered by the tensor pipe and the COLUMN
// taps are scattered over N:
//     Q[v][term, kw, cls] = sum_kh sum_ci x[v + (kh-1) rows][ci] * W_term[cls][ci][kh][kw]     3 * Cin/16 MMAs (A shifted by rows)
//     logits[u][cls]     = bias[cls] + sum_term ( Q[u-1][term,0,cls] + Q[u][term,1,cls] + Q[u+1][term,2,cls] )
//   * the weights stay fp32-accurate: W = W_0 + W_1 + W_2 (three bf16 terms) are three column blocks of B, summed in registers;
//   * input tile = 18 rows x 32 voxels of one plane (one TMA box; rows are exactly four 128-byte core matrices, so any
//     4-row window of the tile is ONE uniform-stride K-major operand); an M tile = 4 rows = 4 warps' worth of TMEM lanes, a
//     WARP = one row: the column taps are two warp shuffles, no shared memory and no block barrier in the epilogue;
//   * output = 16 rows x the 30 interior columns; fp32 NCDHW stores of 30 consecutive floats per warp and class;
//   * warp 0 TMA producer, warp 1 MMA issuer (fully unrolled: 12 / 24 MMAs per tile, two TMEM accumulator sets),
//     warps 2..9 epilogue (two per TMEM lane quarter).
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/fplplus_b200.h"

namespace {

constexpr int kHtOutRows = 16, kHtRows = kHtOutRows + 2, kHtCols = 32;     // output rows; input tile
constexpr int kHtOutCols = kHtCols - 2;
constexpr int kHtPlane = kHtRows * kHtCols * 16;           // bytes of one channel group of the tile: 9216
constexpr int kHtEpiWarps = 8;                            // two per TMEM lane quarter
constexpr int kHtThreads = 64 + 32 * kHtEpiWarps;
constexpr int kHtMaxStages = 6;

struct HtParams {
    const float* w;            // [classes][cin][1][3][3]
    const float* bias;
    float* logits;             // [N][classes][D][H][W]
    int N, D, H, W, cin, classes;
    int x_c8off;
    int nb;                    // UMMA N: 9 * classes rounded up to 16
    int a_bytes, stages;
    int tiles_h, tiles_w, total_tiles;
    int dbg;                   // timing experiments (fpl_debug_set 51): 1 = no stores, 2 = no epilogue work
};

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}

template <int KSTEPS>      // Cin / 16: the MMA loop is fully unrolled (12 / 24 tiny MMAs per tile: loop control per MMA made the
                           // issuing warp the limit, 2.2 k cycles per tile for 0.5 k of MMAs)
__global__ void __launch_bounds__(kHtThreads) head_fwd_tc_kernel(const __grid_constant__ CUtensorMap xmap, HtParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // [A stage ring][B: 3 row taps x cin/8 groups x nb x 16 B][barriers]
    uint8_t* ring = smem;
    uint8_t* b_sm = ring + (size_t)P.stages * P.a_bytes;
    const int b_bytes = 3 * (P.cin / 8) * P.nb * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_sm + ((b_bytes + 127) / 128) * 128);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kHtMaxStages;
    uint64_t* tmem_full = bars + 2 * kHtMaxStages;
    uint64_t* tmem_empty = bars + 2 * kHtMaxStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kHtMaxStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
        for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], kHtEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);              // 2 accumulator sets x 4 M tiles x 64 columns (nb <= 64)
    FPL_PDL_WAIT();
    // B operand (K-major, no swizzle): [kh][k8][n][8 ci] bf16; n = (term * 3 + kw) * classes + cls; K step (kh, j) = 16 channels
    {
        const int groups = P.cin / 8, total = 3 * groups * P.nb * 8;
        __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(b_sm);
        for (int i = threadIdx.x; i < total; i += kHtThreads) {
            int t = i;
            const int kk = t & 7; t >>= 3;
            const int nn = t % P.nb; t /= P.nb;
            const int g = t % groups;
            const int kh = t / groups;
            float v = 0.0f;
            int term = 0;
            if (nn < 9 * P.classes) {
                term = nn / (3 * P.classes);
                const int rem = nn - term * 3 * P.classes;
                const int kw = rem / P.classes, cls = rem - kw * P.classes;
                v = P.w[((int64_t)cls * P.cin + g * 8 + kk) * 9 + kh * 3 + kw];
            }
            const __nv_bfloat16 b0 = __float2bfloat16_rn(v);
            const float r1 = v - __bfloat162float(b0);
            const __nv_bfloat16 b1 = __float2bfloat16_rn(r1);
            const __nv_bfloat16 b2 = __float2bfloat16_rn(r1 - __bfloat162float(b1));
            b[i] = term == 0 ? b0 : (term == 1 ? b1 : b2);
        }
    }
    fence_proxy_async();                                      // generic-proxy writes of B before the tensor pipe reads them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
                int r = t;
                const int tw = r % P.tiles_w; r /= P.tiles_w;
                const int th = r % P.tiles_h; r /= P.tiles_h;     // r = n * D + d
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], (uint32_t)P.a_bytes);
                // map {W*8 bf16, H, c8tot, D, N}; halo origin (h0 - 1, w0 - 1); outside the plane: zero fill = "same" padding
                tma_load_5d(ring + (size_t)stage * P.a_bytes, &xmap, &full_bar[stage], (tw * kHtOutCols - 1) * 8, th * kHtOutRows - 1,
                            P.x_c8off, r % P.D, r / P.D);
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // K-major A (LBO = channel-group plane, SBO = 128: the tile's core matrices are contiguous) and B (LBO = nb*16, SBO = 128)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.nb >> 3) << 17) | (8u << 24);
        const uint64_t a_hi = make_desc(0, kHtPlane, 128), b_hi = make_desc(0, (uint32_t)P.nb * 16, 128);
        const uint64_t b0 = b_hi + (uint64_t)(smem_u32(b_sm) >> 4);
        const uint64_t b_kstep = (uint64_t)(2 * P.nb);            // 16 channels of B, in 16-byte units
        const uint32_t ring_u = smem_u32(ring) >> 4;
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t a0 = a_hi + (uint64_t)(ring_u + (uint32_t)(((size_t)stage * P.a_bytes) >> 4));
            const uint32_t d_acc = tmem_base + (uint32_t)(acc * 256);
#pragma unroll
            for (int m = 0; m < 4; ++m) {                         // M tile m = output rows 4m .. 4m+3
#pragma unroll
                for (int kh = 0; kh < 3; ++kh) {                  // A = input rows 4m+kh .. 4m+kh+3 (tile row 0 = output row -1)
#pragma unroll
                    for (int j = 0; j < KSTEPS; ++j) {
                        const uint64_t ad = a0 + (uint64_t)((4 * m + kh) * (kHtCols * 16 / 16) + j * (2 * kHtPlane / 16));
                        const uint64_t bd = b0 + (uint64_t)(kh * KSTEPS + j) * b_kstep;
                        if (leader) umma_bf16(d_acc + (uint32_t)(m * 64), ad, bd, idesc, (kh | j) ? 1u : 0u);
                    }
                }
            }
            if (leader) { umma_commit(&empty_bar[stage]); umma_commit(&tmem_full[acc]); }
            __syncwarp();
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        FPL_PDL_TRIGGER();
    } else {
        // 8 epilogue warps (a warp alone on its scheduler cannot hide the TMEM latency): warps w and w + 4 share a TMEM lane
        // quarter and split the four M tiles of a tile; a warp = one output row, lane = column of the input tile
        const int quarter = warp & 3, pair = (warp - 2) >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int C = P.classes, nvals = 9 * C;
        const int64_t HW = (int64_t)P.H * P.W;
        int acc = 0; uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
            int r = t;
            const int tw = r % P.tiles_w; r /= P.tiles_w;
            const int th = r % P.tiles_h; r /= P.tiles_h;
            const int dd = r % P.D, n = r / P.D;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int w = tw * kHtOutCols + lane - 1;                       // lanes 0 / 31 are the halo columns
            const bool col_ok = lane >= 1 && lane <= kHtOutCols && w < P.W;
#pragma unroll 1
            for (int m = pair; m < 4 && !(P.dbg & 2); m += 2) {
                const int h = th * kHtOutRows + 4 * m + quarter;
                const uint32_t taddr = lane_base + (uint32_t)(acc * 256 + m * 64);
                float* out = P.logits + (((int64_t)n * C) * P.D + dd) * HW + (int64_t)h * P.W + w;
                if (C == 2) {                                              // the shipped configurations
                    uint32_t v[16], v16, v17;
                    tmem_ld16(taddr, v);
                    tmem_ld2(taddr + 16, v16, v17);
                    tmem_ld_wait();
                    // n = (term * 3 + kw) * 2 + cls
                    float q[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const float t2 = i < 4 ? __uint_as_float(v[12 + i]) : __uint_as_float(i == 4 ? v16 : v17);
                        q[i] = __uint_as_float(v[i]) + __uint_as_float(v[6 + i]) + t2;
                    }
#pragma unroll
                    for (int cls = 0; cls < 2; ++cls) {
                        const float left = __shfl_up_sync(0xffffffffu, q[cls], 1);          // Q[u-1][kw=0]
                        const float right = __shfl_down_sync(0xffffffffu, q[4 + cls], 1);   // Q[u+1][kw=2]
                        const float y = __ldg(P.bias + cls) + left + q[2 + cls] + right;
                        if (col_ok && h < P.H && !(P.dbg & 1)) out[(int64_t)cls * P.D * HW] = y;
                    }
                } else {
                    for (int cls = 0; cls < C; ++cls) {
                        float q[3];
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            q[kw] = 0.0f;
                            for (int term = 0; term < 3; ++term) {
                                const int col = (term * 3 + kw) * C + cls;
                                uint32_t a, b;
                                tmem_ld2(taddr + (uint32_t)(col & ~1), a, b);
                                tmem_ld_wait();
                                q[kw] += __uint_as_float((col & 1) ? b : a);
                            }
                        }
                        const float left = __shfl_up_sync(0xffffffffu, q[0], 1);
                        const float right = __shfl_down_sync(0xffffffffu, q[2], 1);
                        const float y = __ldg(P.bias + cls) + left + q[1] + right;
                        if (col_ok && h < P.H && !(P.dbg & 1)) out[(int64_t)cls * P.D * HW] = y;
                    }
                }
            }
            (void)nvals;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512u);
    }
}

int g_head_tc = 1;      // fpl_debug_set 50: tensor-core head forward on / off
int g_head_tc_dbg = 0;  // fpl_debug_set 51

}  // namespace

void fpl_head_tc_debug_set(int key, long long value) {
    if (key == 50) g_head_tc = (int)value;
    if (key == 51) g_head_tc_dbg = (int)value;
}

bool fpl_head_fwd_tc_eligible(int h, int w, int cin, int classes) {
    return g_head_tc && cin % 16 == 0 && cin >= 16 && cin <= 64 && classes >= 1 && classes <= 7 && h >= 4 && w >= 32 && (cin == 16 || cin == 32);     // 9 * classes <= 64 accumulator columns per M tile
}

int fpl_head_fwd_tc_launch(const void* x, int x_c8tot, int x_c8off, const float* w, const float* bias, float* logits, int n, int d,
                           int h, int w_, int cin, int classes, void* stream) {
    HtParams P;
    P.w = w; P.bias = bias; P.logits = logits; P.N = n; P.D = d; P.H = h; P.W = w_; P.cin = cin; P.classes = classes;
    P.x_c8off = x_c8off; P.dbg = g_head_tc_dbg;
    P.nb = ((9 * classes + 15) / 16) * 16;
    P.a_bytes = (cin / 8) * kHtPlane;
    const int b_bytes = ((3 * (cin / 8) * P.nb * 16 + 127) / 128) * 128;
    P.stages = (160 * 1024 - b_bytes - 512) / P.a_bytes;
    if (P.stages > kHtMaxStages) P.stages = kHtMaxStages;
    FPL_REQUIRE(P.stages >= 2, "fpl_head_fwd: tensor-core tile does not fit shared memory (cin %d, classes %d)", cin, classes);
    const int smem_bytes = P.stages * P.a_bytes + b_bytes + 256 + 1024;
    P.tiles_h = (h + kHtOutRows - 1) / kHtOutRows; P.tiles_w = (w_ + kHtOutCols - 1) / kHtOutCols;
    const int64_t total = (int64_t)P.tiles_h * P.tiles_w * n * d;
    FPL_REQUIRE(total < (1ll << 30), "fpl_head_fwd: too many tiles");
    P.total_tiles = (int)total;
    FPL_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "fpl_head_fwd: x must be 16-byte aligned");
    EncodeTiledFn encode = get_encode_fn();
    FPL_REQUIRE(encode != nullptr, "fpl_head_fwd: cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap xmap;
    cuuint64_t gdim[5] = {(cuuint64_t)w_ * 8, (cuuint64_t)h, (cuuint64_t)x_c8tot, (cuuint64_t)d, (cuuint64_t)n};
    cuuint64_t gstr[4] = {(cuuint64_t)w_ * 16, (cuuint64_t)h * w_ * 16, (cuuint64_t)x_c8tot * h * w_ * 16,
                          (cuuint64_t)d * x_c8tot * h * w_ * 16};
    cuuint32_t box[5] = {(cuuint32_t)kHtCols * 8, (cuuint32_t)kHtRows, (cuuint32_t)(cin / 8), 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FPL_REQUIRE(r == CUDA_SUCCESS, "fpl_head_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
    int grid = FPL_NUM_SMS;
    if (grid > P.total_tiles) grid = P.total_tiles;
    if (cin == 16) {
        FPL_CHECK_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        fpl_launch(head_fwd_tc_kernel<1>, grid, kHtThreads, smem_bytes, (cudaStream_t)stream, xmap, P);
    } else {
        FPL_CHECK_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        fpl_launch(head_fwd_tc_kernel<2>, grid, kHtThreads, smem_bytes, (cudaStream_t)stream, xmap, P);
    }
    return 0;
}
