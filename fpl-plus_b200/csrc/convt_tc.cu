// ConvTranspose3d kernel 2 / stride 2 (PyMIC/pymic/net/net3d/unet2d5_dsbn.py:152,181; (1,2,2) in 2.5-D levels) on the
// 5th-gen tensor cores.  With v a voxel of the LOW-resolution grid and tap = (a,b,c) in {0,1}^3:
//   fwd    y[2v+tap][co]   = bias[co] + sum_ci x[v][ci]      * W[ci][co][tap]      GEMM  M=voxels K=Cin      N=(tap,Cout)
//   dgrad  dx[v][ci]       =            sum_{tap,co} dy[2v+tap][co] * W[ci][co][tap]   GEMM  M=voxels K=(tap,Cout) N=Cin
//   wgrad  dW[ci][co][tap] =            sum_v x[v][ci]       * dy[2v+tap][co]      GEMM  M=Cin    K=voxels   N=Cout per tap
// The stride-2 sub-lattice of the high-resolution tensor that belongs to one tap is fetched by TMA with
// elementStrides = 2 in H and W (4-D tensor map {8 ch, W, H, planes}), so it lands in shared memory as the same
// dense C8-planar tile [C/8][16 rows][8 voxels][8 ch] as a low-resolution tile: K-major core matrices for
// fwd/dgrad, MN-major for wgrad.  fwd scatters its epilogue straight into the second half of the concat buffer.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/fplplus_b200.h"

namespace {

constexpr int kTH = 16, kTW = 8;                    // low-resolution tile: 128 voxels = UMMA M
constexpr int kPlane = kTH * kTW * 16;              // one channel group of a tile: 2048 B
constexpr int kThreadsT = 192;
constexpr int kMaxStagesT = 8;

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// fwd (mode 0) and dgrad (mode 1): D[128 voxels][NB] = sum over K chunks of A[128][KC] * B[KC][NB]
// ---------------------------------------------------------------------------------------------------------
struct CtParams {
    const __nv_bfloat16* image;      // [slice][chunk][KC/8][NB][8]
    const float* bias;               // fwd: [cout]
    bf16x8* out;                     // fwd: high-res y (c8tot/off below); dgrad: low-res dx
    int out_c8tot, out_c8off;
    int a_c8tot, a_c8off;            // A operand channel-slice addressing (fwd: x low-res; dgrad: dy high-res)
    int N, D, H, W;                  // LOW-resolution grid
    int cin, cout, kd2, ntaps;       // convT channels (cin -> cout), kd2 in {1,2}, ntaps = 4*kd2
    int mode;                        // 0 fwd, 1 dgrad
    int kdim, ndim;                  // GEMM K and N sizes (fwd: cin, ntaps*cout; dgrad: ntaps*cout, cin)
    int nb, kc, nslices, nchunks, a_bytes, b_bytes, stages, tmem_cols;
    int tiles_h, tiles_w, total_items;
};

struct CtCfg {
    int nb, kc, nslices, nchunks, a_bytes, b_bytes, stages, tmem_cols, smem_bytes;
};

bool make_ct_cfg(int kdim, int ndim, int kc_div, CtCfg& c) {
    if (kdim % 16 || ndim % 16) return false;
    c.nb = ndim;
    if (c.nb > 128) {
        if (ndim % 128 == 0) c.nb = 128;
        else if (ndim % 64 == 0) c.nb = 64;
        else if (ndim % 32 == 0) c.nb = 32;
        else c.nb = 16;
    }
    c.kc = 64;
    while (c.kc > 16 && (kc_div % c.kc != 0)) c.kc /= 2;      // a K chunk never straddles two taps
    if (kc_div % c.kc != 0) return false;
    c.nslices = ndim / c.nb;
    c.nchunks = kdim / c.kc;
    c.a_bytes = (c.kc / 8) * kPlane;
    c.b_bytes = c.kc * c.nb * 2;
    int stage = c.a_bytes + c.b_bytes;
    c.stages = (96 * 1024) / stage;
    if (c.stages > kMaxStagesT) c.stages = kMaxStagesT;
    if (c.stages < 2) c.stages = 2;
    int cols = 2 * c.nb;
    c.tmem_cols = 32;
    while (c.tmem_cols < cols) c.tmem_cols *= 2;
    c.smem_bytes = c.stages * stage + 1024 + 256 + (ndim > 1024 ? ndim : 1024) * (int)sizeof(float) + 16;   // + the tiled bias [ndim]
    return true;
}

__global__ void __launch_bounds__(kThreadsT) convt_gemm_tc_kernel(const __grid_constant__ CUtensorMap amap, CtParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = P.a_bytes + P.b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.stages * stage_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStagesT;
    uint64_t* tmem_full = bars + 2 * kMaxStagesT;
    uint64_t* tmem_empty = bars + 2 * kMaxStagesT + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStagesT + 4);
    float* bias_sm = reinterpret_cast<float*>(bars + 2 * kMaxStagesT + 6);   // fwd: bias tiled over taps, [ndim]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
        for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    FPL_PDL_WAIT();      // prologue above overlapped the previous kernel's tail; from here on its results are visible
    if (P.mode == 0)
        for (int i = threadIdx.x; i < P.ndim; i += kThreadsT) bias_sm[i] = P.bias != nullptr ? P.bias[i % P.cout] : 0.0f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int Do = P.D * P.kd2;
    const int chunks_per_tap = P.cout / P.kc;       // dgrad only

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int it = blockIdx.x; it < P.total_items; it += gridDim.x) {
                int t = it;
                const int tw = t % P.tiles_w; t /= P.tiles_w;
                const int th = t % P.tiles_h; t /= P.tiles_h;
                const int d = t % P.D; t /= P.D;
                const int n = t % P.N;
                const int slice = t / P.N;
                for (int q = 0; q < P.nchunks; ++q) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* a_dst = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(P.a_bytes + P.b_bytes));
                    if (P.mode == 0) {
                        // low-res x tile, channel chunk q
                        tma_load_4d(a_dst, &amap, &full_bar[stage], 0, tw * kTW, th * kTH,
                                    (n * P.D + d) * P.a_c8tot + P.a_c8off + q * (P.kc / 8));
                    } else {
                        // stride-2 sub-lattice of dy that belongs to tap (a,b,c); channel chunk cq of that tap
                        const int tap = q / chunks_per_tap, cq = q - tap * chunks_per_tap;
                        const int a = P.kd2 == 2 ? tap >> 2 : 0, b = (tap >> 1) & 1, c = tap & 1;
                        tma_load_4d(a_dst, &amap, &full_bar[stage], 0, 2 * tw * kTW + c, 2 * th * kTH + b,
                                    (n * Do + d * P.kd2 + a) * P.a_c8tot + P.a_c8off + cq * (P.kc / 8));
                    }
                    const uint8_t* b_src = reinterpret_cast<const uint8_t*>(P.image) + (size_t)(slice * P.nchunks + q) * P.b_bytes;
                    bulk_load(a_dst + P.a_bytes, b_src, (uint32_t)P.b_bytes, &full_bar[stage]);
                    if (++stage == P.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P.nb >> 3) << 17) | (8u << 24);
        const uint64_t a_hi = make_desc(0, kPlane, kTW * 16), b_hi = make_desc(0, (uint32_t)P.nb * 16, 128);
        const uint32_t smem_u = smem_u32(smem) >> 4;
        const int ksteps = P.kc / 16;
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int it = blockIdx.x; it < P.total_items; it += gridDim.x) {
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P.nb);
            uint32_t accumulate = 0;
            for (int q = 0; q < P.nchunks; ++q) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_base = smem_u + (uint32_t)(((size_t)stage * stage_bytes) >> 4);
                const uint32_t b_base = a_base + (uint32_t)(P.a_bytes >> 4);
                // 64-bit adds of warp-uniform values (the 14-bit start-address field never carries): uniform datapath
                uint64_t adesc = a_hi + (uint64_t)a_base, bdesc = b_hi + (uint64_t)b_base;
                const uint64_t b_step = (uint64_t)(2 * P.nb);
                for (int j = 0; j < ksteps; ++j, adesc += 2 * kPlane / 16, bdesc += b_step) {
                    if (leader) umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                    accumulate = 1;
                }
                if (leader) umma_commit(&empty_bar[stage]);
                __syncwarp();
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
            if (leader) umma_commit(&tmem_full[acc]);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        FPL_PDL_TRIGGER();   // this CTA has issued its last tile: the next kernel of the stream may be scheduled as SMs drain
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int hl = row / kTW, wl = row % kTW;
        int acc = 0; uint32_t acc_phase = 0;
        const int Ho = P.H * 2, Wo = P.W * 2;
        for (int it = blockIdx.x; it < P.total_items; it += gridDim.x) {
            int t = it;
            const int tw = t % P.tiles_w; t /= P.tiles_w;
            const int th = t % P.tiles_h; t /= P.tiles_h;
            const int d = t % P.D; t /= P.D;
            const int n = t % P.N;
            const int slice = t / P.N;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int h = th * kTH + hl, w = tw * kTW + wl;
            const bool valid = h < P.H && w < P.W;
            const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * P.nb);
            for (int c0 = 0; c0 < P.nb; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(t_row + (uint32_t)c0, r);
                tmem_ld_wait();
                float v[16];
                const int col = slice * P.nb + c0;          // GEMM N index of v[0]
                if (P.mode == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bias_sm[col + i];
                    if (valid) {
                        const int tap = col / P.cout, co = col - tap * P.cout;
                        const int a = P.kd2 == 2 ? tap >> 2 : 0, b = (tap >> 1) & 1, c = tap & 1;
                        const int64_t HWo = (int64_t)Ho * Wo;
                        bf16x8* dst = P.out + (((int64_t)n * Do + d * P.kd2 + a) * P.out_c8tot + P.out_c8off + co / 8) * HWo +
                                      (int64_t)(2 * h + b) * Wo + (2 * w + c);
                        st_bf16x8(dst, v);
                        st_bf16x8(dst + HWo, v + 8);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
                    if (valid) {
                        const int64_t HW = (int64_t)P.H * P.W;
                        bf16x8* dst = P.out + (((int64_t)n * P.D + d) * P.out_c8tot + P.out_c8off + col / 8) * HW + (int64_t)h * P.W + w;
                        st_bf16x8(dst, v);
                        st_bf16x8(dst + HW, v + 8);
                    }
                }
            }
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
        tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

// weight staging: fp32 W[cin][cout][ntaps] -> bf16 image [slice][chunk][KC/8][NB][8] of the GEMM B operand
//   mode 0 (fwd):   B[k = ci][n = tap*cout + co]          mode 1 (dgrad): B[k = tap*cout + co][n = ci]
__global__ void convt_prep_kernel(const float* __restrict__ w, __nv_bfloat16* image, int cin, int cout, int ntaps, int mode,
                                  int nb, int kc, int nchunks, int total) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int t = i;
        const int el = t % 8; t /= 8;
        const int nrow = t % nb; t /= nb;
        const int k8 = t % (kc / 8); t /= (kc / 8);
        const int q = t % nchunks;
        const int sl = t / nchunks;
        const int k = q * kc + k8 * 8 + el, nn = sl * nb + nrow;
        int ci, co, tap;
        if (mode == 0) { ci = k; tap = nn / cout; co = nn - tap * cout; }
        else { tap = k / cout; co = k - tap * cout; ci = nn; }
        image[i] = __float2bfloat16_rn(w[((int64_t)ci * cout + co) * ntaps + tap]);
    }
}

CUresult encode_4d(EncodeTiledFn encode, CUtensorMap* map, const void* base, int64_t planes, int h, int w, int box_c8,
                   int stride) {
    cuuint64_t gdim[4] = {8, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)planes};
    cuuint64_t gstr[3] = {16, (cuuint64_t)w * 16, (cuuint64_t)h * w * 16};
    cuuint32_t box[4] = {8, (cuuint32_t)(kTW * stride), (cuuint32_t)(kTH * stride), (cuuint32_t)box_c8};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// ---------------------------------------------------------------------------------------------------------
// wgrad: dW[ci][co][tap] += sum_v x[v][ci] * dy[2v+tap][co]   (both operands MN-major, one TMEM accumulator per tap)
// ---------------------------------------------------------------------------------------------------------
struct CtwParams {
    float* dw;
    int N, D, H, W, cin, cout, kd2, ntaps;
    int x_c8tot, x_c8off, dy_c8tot, dy_c8off;
    int mgroups, m, nb, mtiles, nchunks;
    int x_bytes, dy_tile_bytes, stage_bytes, stages, tmem_cols;
    int tiles_h, tiles_w, tiles_total, split;
    int tapmajor;          // dW written as scratch [tap][cout][cin]
};

__global__ void __launch_bounds__(kThreadsT) convt_wgrad_tc_kernel(const __grid_constant__ CUtensorMap xmap,
                                                                  const __grid_constant__ CUtensorMap dymap, CtwParams P) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStagesT;
    uint64_t* done_bar = bars + 2 * kMaxStagesT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStagesT + 1);
    uint8_t* ring = smem + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int b = blockIdx.x;
    const int slice = b % P.split; b /= P.split;
    const int nc = b % P.nchunks;
    const int mt = b / P.nchunks;
    const int tile_begin = (int)(((int64_t)P.tiles_total * slice) / P.split);
    const int tile_end = (int)(((int64_t)P.tiles_total * (slice + 1)) / P.split);
    const int Do = P.D * P.kd2;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&dymap) : "memory");
        for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
    FPL_PDL_WAIT();      // prologue above overlapped the previous kernel's tail; from here on its results are visible
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int t = tile_begin; t < tile_end; ++t) {
                int r = t;
                const int tw = r % P.tiles_w; r /= P.tiles_w;
                const int th = r % P.tiles_h; r /= P.tiles_h;
                const int d = r % P.D;
                const int n = r / P.D;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* x_dst = ring + (size_t)stage * P.stage_bytes;
                mbar_expect_tx(&full_bar[stage], (uint32_t)(P.x_bytes + P.ntaps * P.dy_tile_bytes));
                tma_load_4d(x_dst, &xmap, &full_bar[stage], 0, tw * kTW, th * kTH,
                            (n * P.D + d) * P.x_c8tot + P.x_c8off + mt * P.mgroups);
                for (int tap = 0; tap < P.ntaps; ++tap) {
                    const int a = P.kd2 == 2 ? tap >> 2 : 0, bb = (tap >> 1) & 1, c = tap & 1;
                    tma_load_4d(x_dst + P.x_bytes + (size_t)tap * P.dy_tile_bytes, &dymap, &full_bar[stage], 0,
                                2 * tw * kTW + c, 2 * th * kTH + bb,
                                (n * Do + d * P.kd2 + a) * P.dy_c8tot + P.dy_c8off + nc * (P.nb / 8));
                }
                if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(P.nb >> 3) << 17) | ((uint32_t)(P.m >> 4) << 24);
        // MN-major no-swizzle: LBO = stride between the two 8-voxel K groups (next tile row), SBO = channel-group plane
        const uint64_t hi = make_desc(0, kTW * 16, kPlane);
        const uint32_t ring_u = smem_u32(ring);
        const bool leader = elect_one();
        int stage = 0; uint32_t phase = 0;
        uint32_t accumulate = 0;
        for (int t = tile_begin; t < tile_end; ++t) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t x_base = (ring_u + (uint32_t)stage * (uint32_t)P.stage_bytes) >> 4;
            const uint32_t dy_base = x_base + ((uint32_t)P.x_bytes >> 4);
            // 64-bit adds of warp-uniform values (the 14-bit start-address field never carries): uniform datapath
            uint64_t adesc = hi + (uint64_t)x_base, b_hp = hi + (uint64_t)dy_base;
            const uint64_t dy_tile16 = (uint64_t)((uint32_t)P.dy_tile_bytes >> 4);
            for (int hp = 0; hp < kTH / 2; ++hp, adesc += 2 * kTW * 16 / 16, b_hp += 2 * kTW * 16 / 16) {
                uint64_t bdesc = b_hp;
                uint32_t d_tap = tmem_base;
                for (int tap = 0; tap < P.ntaps; ++tap, bdesc += dy_tile16, d_tap += (uint32_t)P.nb) {
                    if (leader) umma_bf16(d_tap, adesc, bdesc, idesc, accumulate);
                }
                accumulate = 1;
            }
            if (leader) umma_commit(&empty_bar[stage]);
            __syncwarp();
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(done_bar);
        __syncwarp();
        FPL_PDL_TRIGGER();   // this CTA has issued its last tile: the next kernel of the stream may be scheduled as SMs drain
    } else if (tile_end > tile_begin) {
        const int quarter = warp & 3;
        mbar_wait(done_bar, 0);
        tc_fence_after();
        const int tlane = quarter * 32 + lane;
        int row = P.m == 128 ? tlane : ((lane < 16) ? quarter * 16 + lane : -1);     // M=64: 16 rows per lane quadrant
        const bool valid = row >= 0 && row < P.mgroups * 8 && (mt * P.mgroups * 8 + row) < P.cin;
        const int ci = mt * P.mgroups * 8 + row;
        for (int tap = 0; tap < P.ntaps; ++tap) {
            for (int c0 = 0; c0 < P.nb; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tap * P.nb + c0), r);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int co = nc * P.nb + c0 + i;
                        if (P.tapmajor) atomicAdd(P.dw + ((int64_t)tap * P.cout + co) * P.cin + ci, __uint_as_float(r[i]));   // lanes = ci: coalesced
                        else atomicAdd(P.dw + ((int64_t)ci * P.cout + co) * P.ntaps + tap, __uint_as_float(r[i]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
    }
}

}  // namespace

extern "C" int64_t fpl_convt_weight_image_bytes(int cin, int cout, int kd2) {
    return (int64_t)cin * cout * 4 * kd2 * 2;
}

// mode 0: forward image, mode 1: dgrad image (w fp32 [cin][cout][kd2][2][2])
extern "C" int fpl_convt_prep_weight(const float* w, int cin, int cout, int kd2, int mode, void* image, void* stream) {
    FPL_REQUIRE(kd2 == 1 || kd2 == 2, "fpl_convt_prep_weight: kd2=%d must be 1 or 2", kd2);
    const int ntaps = 4 * kd2;
    CtCfg c;
    const int kdim = mode == 0 ? cin : ntaps * cout, ndim = mode == 0 ? ntaps * cout : cin;
    FPL_REQUIRE(make_ct_cfg(kdim, ndim, mode == 0 ? cin : cout, c), "fpl_convt_prep_weight: unsupported channels (%d -> %d)", cin, cout);
    const int total = kdim * ndim;
    int blocks = (total + 255) / 256;
    if (blocks > FPL_NUM_SMS * 4) blocks = FPL_NUM_SMS * 4;
    fpl_launch(convt_prep_kernel, blocks, 256, 0, (cudaStream_t)stream, w, (__nv_bfloat16*)image, cin, cout, ntaps, mode, c.nb, c.kc,
                                                               c.nchunks, total);
    FPL_LAUNCH_CHECK();
    return 0;
}

namespace {
constexpr int kMaxCtBatch = 16;
struct CtPrepBatch {
    const float* w[kMaxCtBatch];
    __nv_bfloat16* image[kMaxCtBatch];
    int cin[kMaxCtBatch], cout[kMaxCtBatch], ntaps[kMaxCtBatch], mode[kMaxCtBatch], nb[kMaxCtBatch], kc[kMaxCtBatch],
        nchunks[kMaxCtBatch], total[kMaxCtBatch];
};
__global__ void convt_prep_batch_kernel(const __grid_constant__ CtPrepBatch B) {
    FPL_PDL_WAIT();      // everything below may read what earlier kernels of the stream wrote
    const int e = blockIdx.y;
    const float* __restrict__ w = B.w[e];
    __nv_bfloat16* image = B.image[e];
    const int cout = B.cout[e], ntaps = B.ntaps[e], mode = B.mode[e], nb = B.nb[e], kc = B.kc[e], nchunks = B.nchunks[e];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B.total[e]; i += gridDim.x * blockDim.x) {
        int t = i;
        const int el = t % 8; t /= 8;
        const int nrow = t % nb; t /= nb;
        const int k8 = t % (kc / 8); t /= (kc / 8);
        const int q = t % nchunks;
        const int sl = t / nchunks;
        const int k = q * kc + k8 * 8 + el, nn = sl * nb + nrow;
        int ci, co, tap;
        if (mode == 0) { ci = k; tap = nn / cout; co = nn - tap * cout; }
        else { tap = k / cout; co = k - tap * cout; ci = nn; }
        image[i] = __float2bfloat16_rn(w[((int64_t)ci * cout + co) * ntaps + tap]);
    }
}
}  // namespace

extern "C" int fpl_convt_prep_weight_batch(int count, const float* const* h_w, const int* h_cin, const int* h_cout,
                                           const int* h_kd2, const int* h_mode, void* const* h_images, void* stream) {
    FPL_REQUIRE(count >= 0 && count <= kMaxCtBatch, "fpl_convt_prep_weight_batch: count %d not in [0,%d]", count, kMaxCtBatch);
    if (count == 0) return 0;
    CtPrepBatch B;
    int max_total = 0;
    for (int e = 0; e < count; ++e) {
        const int ntaps = 4 * h_kd2[e], mode = h_mode[e], cin = h_cin[e], cout = h_cout[e];
        const int kdim = mode == 0 ? cin : ntaps * cout, ndim = mode == 0 ? ntaps * cout : cin;
        CtCfg c;
        FPL_REQUIRE(make_ct_cfg(kdim, ndim, mode == 0 ? cin : cout, c), "fpl_convt_prep_weight_batch: unsupported channels (%d -> %d)", cin, cout);
        B.w[e] = h_w[e]; B.image[e] = (__nv_bfloat16*)h_images[e]; B.cin[e] = cin; B.cout[e] = cout; B.ntaps[e] = ntaps;
        B.mode[e] = mode; B.nb[e] = c.nb; B.kc[e] = c.kc; B.nchunks[e] = c.nchunks; B.total[e] = kdim * ndim;
        if (B.total[e] > max_total) max_total = B.total[e];
    }
    int bx = (max_total + 255) / 256;
    if (bx > 64) bx = 64;
    fpl_launch(convt_prep_batch_kernel, dim3(bx, count), 256, 0, (cudaStream_t)stream, B);
    FPL_LAUNCH_CHECK();
    return 0;
}

static int convt_gemm_launch(int mode, const void* a, int a_c8tot, int a_c8off, const void* image, const float* bias, void* out,
                             int out_c8tot, int out_c8off, int n, int d, int h, int w, int cin, int cout, int kd2,
                             void* stream) {
    FPL_REQUIRE(kd2 == 1 || kd2 == 2, "fpl_convt_k2s2_tc: kd2=%d must be 1 or 2", kd2);
    const int ntaps = 4 * kd2;
    const int kdim = mode == 0 ? cin : ntaps * cout, ndim = mode == 0 ? ntaps * cout : cin;
    CtCfg c;
    FPL_REQUIRE(make_ct_cfg(kdim, ndim, mode == 0 ? cin : cout, c), "fpl_convt_k2s2_tc: unsupported channels (%d -> %d)", cin, cout);
    FPL_REQUIRE(ndim <= 8192, "fpl_convt_k2s2_tc: GEMM N = %d too large", ndim);      // 8 taps x 1024 output channels
    EncodeTiledFn encode = get_encode_fn();
    FPL_REQUIRE(encode != nullptr, "fpl_convt_k2s2_tc: cuTensorMapEncodeTiled not available from the driver");
    CUtensorMap amap;
    CUresult r;
    if (mode == 0) r = encode_4d(encode, &amap, a, (int64_t)n * d * a_c8tot, h, w, c.kc / 8, 1);
    else r = encode_4d(encode, &amap, a, (int64_t)n * d * kd2 * a_c8tot, 2 * h, 2 * w, c.kc / 8, 2);
    FPL_REQUIRE(r == CUDA_SUCCESS, "fpl_convt_k2s2_tc: tensor map failed (%d)", (int)r);
    CtParams P;
    P.image = (const __nv_bfloat16*)image; P.bias = bias; P.out = (bf16x8*)out; P.out_c8tot = out_c8tot; P.out_c8off = out_c8off;
    P.a_c8tot = a_c8tot; P.a_c8off = a_c8off; P.N = n; P.D = d; P.H = h; P.W = w; P.cin = cin; P.cout = cout; P.kd2 = kd2;
    P.ntaps = ntaps; P.mode = mode; P.kdim = kdim; P.ndim = ndim;
    P.nb = c.nb; P.kc = c.kc; P.nslices = c.nslices; P.nchunks = c.nchunks; P.a_bytes = c.a_bytes; P.b_bytes = c.b_bytes;
    P.stages = c.stages; P.tmem_cols = c.tmem_cols;
    P.tiles_h = (h + kTH - 1) / kTH; P.tiles_w = (w + kTW - 1) / kTW;
    int64_t total = (int64_t)P.tiles_h * P.tiles_w * d * n * c.nslices;
    FPL_REQUIRE(total < (1ll << 30), "fpl_convt_k2s2_tc: too many tiles");
    P.total_items = (int)total;
    FPL_CHECK_CUDA(cudaFuncSetAttribute(convt_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c.smem_bytes));
    int grid = FPL_NUM_SMS * 2;
    if (grid > P.total_items) grid = P.total_items;
    fpl_launch(convt_gemm_tc_kernel, grid, kThreadsT, c.smem_bytes, (cudaStream_t)stream, amap, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_convt_k2s2_fwd_tc(const void* x, int x_c8tot, int x_c8off, const void* image, const float* bias, void* y,
                                     int y_c8tot, int y_c8off, int n, int d, int h, int w, int cin, int cout, int kd2,
                                     void* stream) {
    return convt_gemm_launch(0, x, x_c8tot, x_c8off, image, bias, y, y_c8tot, y_c8off, n, d, h, w, cin, cout, kd2, stream);
}

extern "C" int fpl_convt_k2s2_dgrad_tc(const void* dy, int dy_c8tot, int dy_c8off, const void* image_t, void* dx, int dx_c8tot,
                                       int dx_c8off, int n, int d, int h, int w, int cin, int cout, int kd2, void* stream) {
    return convt_gemm_launch(1, dy, dy_c8tot, dy_c8off, image_t, nullptr, dx, dx_c8tot, dx_c8off, n, d, h, w, cin, cout, kd2,
                             stream);
}

static int convt_wgrad_launch(const void* x, int x_c8tot, int x_c8off, const void* dy, int dy_c8tot, int dy_c8off,
                              float* dw, int n, int d, int h, int w, int cin, int cout, int kd2, void* stream, int tapmajor) {
    FPL_REQUIRE(kd2 == 1 || kd2 == 2, "fpl_convt_k2s2_wgrad_tc: kd2=%d must be 1 or 2", kd2);
    FPL_REQUIRE(cin % 8 == 0 && cout % 16 == 0, "fpl_convt_k2s2_wgrad_tc: unsupported channels (%d -> %d)", cin, cout);
    EncodeTiledFn encode = get_encode_fn();
    FPL_REQUIRE(encode != nullptr, "fpl_convt_k2s2_wgrad_tc: cuTensorMapEncodeTiled not available from the driver");
    CtwParams P;
    P.dw = dw; P.N = n; P.D = d; P.H = h; P.W = w; P.cin = cin; P.cout = cout; P.kd2 = kd2; P.ntaps = 4 * kd2;
    P.x_c8tot = x_c8tot; P.x_c8off = x_c8off; P.dy_c8tot = dy_c8tot; P.dy_c8off = dy_c8off; P.tapmajor = tapmajor;
    const int g_all = cin / 8;
    P.mgroups = g_all < 16 ? g_all : 16;
    FPL_REQUIRE(g_all % P.mgroups == 0, "fpl_convt_k2s2_wgrad_tc: cin=%d not tileable", cin);
    P.mtiles = g_all / P.mgroups;
    P.m = P.mgroups <= 8 ? 64 : 128;
    P.nb = cout % 32 == 0 ? 32 : 16;
    P.x_bytes = P.mgroups * kPlane;
    if (2 * (P.x_bytes + P.ntaps * (P.nb / 8) * kPlane) > 200 * 1024 - (P.m / 8) * kPlane) P.nb = 16;   // keep two stages
    P.nchunks = cout / P.nb;
    P.dy_tile_bytes = (P.nb / 8) * kPlane;
    P.stage_bytes = P.x_bytes + P.ntaps * P.dy_tile_bytes;
    // the M=64/128 operand reads 8/16 planes from the x tile start: with fewer valid groups it runs into the dy tiles of
    // the same stage (harmless garbage rows); keep one extra stage worth of slack behind the ring
    P.stages = (200 * 1024 - (P.m / 8) * kPlane) / P.stage_bytes;
    if (P.stages > kMaxStagesT) P.stages = kMaxStagesT;
    FPL_REQUIRE(P.stages >= 2, "fpl_convt_k2s2_wgrad_tc: stage of %d bytes does not fit twice", P.stage_bytes);
    int cols = P.ntaps * P.nb;
    P.tmem_cols = 32;
    while (P.tmem_cols < cols) P.tmem_cols *= 2;
    CUtensorMap xmap, dymap;
    CUresult r = encode_4d(encode, &xmap, x, (int64_t)n * d * x_c8tot, h, w, P.mgroups, 1);
    FPL_REQUIRE(r == CUDA_SUCCESS, "fpl_convt_k2s2_wgrad_tc: tensor map (x) failed (%d)", (int)r);
    r = encode_4d(encode, &dymap, dy, (int64_t)n * d * kd2 * dy_c8tot, 2 * h, 2 * w, P.nb / 8, 2);
    FPL_REQUIRE(r == CUDA_SUCCESS, "fpl_convt_k2s2_wgrad_tc: tensor map (dy) failed (%d)", (int)r);
    P.tiles_h = (h + kTH - 1) / kTH; P.tiles_w = (w + kTW - 1) / kTW;
    int64_t tiles = (int64_t)P.tiles_h * P.tiles_w * d * n;
    FPL_REQUIRE(tiles < (1ll << 30), "fpl_convt_k2s2_wgrad_tc: too many tiles");
    P.tiles_total = (int)tiles;
    const int pairs = P.mtiles * P.nchunks;
    int split = FPL_NUM_SMS / pairs;
    if (split > P.tiles_total / 4) split = P.tiles_total / 4;      // amortise the per-CTA atomics epilogue
    if (split < 1) split = 1;
    P.split = split;
    const int smem_bytes = P.stages * P.stage_bytes + (P.m / 8) * kPlane + 1024 + 256;
    FPL_CHECK_CUDA(cudaFuncSetAttribute(convt_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    fpl_launch(convt_wgrad_tc_kernel, pairs * split, kThreadsT, smem_bytes, (cudaStream_t)stream, xmap, dymap, P);
    FPL_LAUNCH_CHECK();
    return 0;
}

extern "C" int fpl_convt_k2s2_wgrad_tc(const void* x, int x_c8tot, int x_c8off, const void* dy, int dy_c8tot, int dy_c8off,
                                       float* dw, int n, int d, int h, int w, int cin, int cout, int kd2, void* stream) {
    return convt_wgrad_launch(x, x_c8tot, x_c8off, dy, dy_c8tot, dy_c8off, dw, n, d, h, w, cin, cout, kd2, stream, 0);
}

/* Tap-major variant: scratch S[tap][cout][cin] (fp32, ACCUMULATED into; coalesced epilogue atomics); folded into the
 * nn.ConvTranspose3d layout [cin][cout][tap] by fpl_wgrad_tapmajor_to_dw_batch with a NEGATIVE tap count. */
extern "C" int fpl_convt_k2s2_wgrad_tc_tapmajor(const void* x, int x_c8tot, int x_c8off, const void* dy, int dy_c8tot,
                                                int dy_c8off, float* scratch, int n, int d, int h, int w, int cin, int cout,
                                                int kd2, void* stream) {
    return convt_wgrad_launch(x, x_c8tot, x_c8off, dy, dy_c8tot, dy_c8off, scratch, n, d, h, w, cin, cout, kd2, stream, 1);
}
