// Does cuTensorMapEncodeTiled accept a tensor map whose dimensions are NOT ordered by stride (C8-planar activation
// [N][D][C8][H][W][16 B] presented as (w, c8, d, h, n)), and does the box land as [h][d][c8][w]?  Development check.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "../fpl-plus_b200/csrc/tc_ptx.cuh"

__global__ void k(const __grid_constant__ CUtensorMap map, uint64_t* out, int bytes, int c0, int c1, int c2, int c3, int c4) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)bytes);
        tma_load_5d(smem, &map, &bar, c0, c1, c2, c3, c4);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bytes / 8; i += blockDim.x) out[i] = reinterpret_cast<uint64_t*>(smem)[i];
}

int main() {
    const int N = 2, D = 5, C8 = 3, H = 12, W = 24;
    const size_t elems = (size_t)N * D * C8 * H * W * 2;   // u64 elements (8 B = 4 channels)
    std::vector<uint64_t> h(elems);
    for (size_t i = 0; i < elems; ++i) h[i] = i;           // value = linear u64 index
    uint64_t* d_x; cudaMalloc(&d_x, elems * 8); cudaMemcpy(d_x, h.data(), elems * 8, cudaMemcpyHostToDevice);
    EncodeTiledFn encode = get_encode_fn();
    CUtensorMap map;
    cuuint64_t gdim[5] = {(cuuint64_t)W * 2, (cuuint64_t)C8, (cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)H * W * 16, (cuuint64_t)C8 * H * W * 16, (cuuint64_t)W * 16, (cuuint64_t)D * C8 * H * W * 16};
    const int bw = 10, bc = 2, bd = 4, bh = 6;
    cuuint32_t box[5] = {(cuuint32_t)bw * 2, (cuuint32_t)bc, (cuuint32_t)bd, (cuuint32_t)bh, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, d_x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode result %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    const int bytes = bw * bc * bd * bh * 16;
    uint64_t* d_out; cudaMalloc(&d_out, bytes);
    // box origin: w0 = 3 voxels, c8 = 1, d = -1 (one plane outside: zero fill), h = 2, n = 1
    const int w0 = 3, c0 = 1, d0 = -1, h0 = 2, n0 = 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024);
    k<<<1, 128, bytes>>>(map, d_out, bytes, w0 * 2, c0, d0, h0, n0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<uint64_t> o(bytes / 8);
    cudaMemcpy(o.data(), d_out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int hh = 0; hh < bh; ++hh) for (int dd = 0; dd < bd; ++dd) for (int cc = 0; cc < bc; ++cc) for (int ww = 0; ww < bw * 2; ++ww) {
        const size_t si = (((size_t)hh * bd + dd) * bc + cc) * bw * 2 + ww;
        const int d = d0 + dd;
        uint64_t want = 0;
        if (d >= 0 && d < D) want = ((((size_t)n0 * D + d) * C8 + c0 + cc) * H + h0 + hh) * W * 2 + w0 * 2 + ww;
        if (o[si] != want) { if (bad < 5) printf("mismatch at h%d d%d c%d w%d: got %llu want %llu\n", hh, dd, cc, ww, (unsigned long long)o[si], (unsigned long long)want); ++bad; }
    }
    printf("box lands as [h][d][c8][w]: %s (%d mismatches)\n", bad == 0 ? "YES" : "NO", bad);
    return 0;
}
