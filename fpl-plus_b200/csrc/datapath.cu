// Device data path of the training inputs: RandomCrop + RandomFlip of PyMIC's loader
// (PyMIC/pymic/transform/crop.py:170-244, flip.py:14-62) as ONE gather kernel over volumes that stay resident in HBM.
// The host draws the crop origin and the flip axes per sample (same decisions, same order as the reference transforms);
// this kernel cuts the image patch (fp32), the label patch (uint8) and the agreement-code patch (uint8) and applies the
// flips on the fly.  LabelToProbability (one-hot) and NiftyDataset.set_weight_ happen inside the loss kernels
// (fpl_dice_ce_*_ex), so no fp32 one-hot / weight tensor is ever materialised and nothing crosses PCIe per step.
#include "common.cuh"
#include "../../include/fplplus_b200.h"

namespace {

struct __align__(16) PatchRow {      // one sample of the batch (64 bytes, built by the host)
    const float* image;              // [C][D][H][W] fp32 (normalised, padded to >= patch size)
    const uint8_t* label;            // [D][H][W] or NULL
    const uint8_t* code;             // [D][H][W] agreement code (0/1/2) or NULL
    int D, H, W;
    int d0, h0, w0;                  // crop origin
    int flip;                        // bit 0: depth, bit 1: height, bit 2: width
    int pad[3];
};
static_assert(sizeof(PatchRow) == 64, "PatchRow layout is part of the C ABI (fpl_gather_patches)");

__global__ void __launch_bounds__(256) gather_patches_kernel(const PatchRow* __restrict__ rows, int C, int pd, int ph, int pw,
                                                             float* __restrict__ out_img, uint8_t* __restrict__ out_lab,
                                                             uint8_t* __restrict__ out_code) {
    FPL_PDL_WAIT();
    const int n = blockIdx.y;
    const PatchRow r = rows[n];
    const int64_t pvox = (int64_t)pd * ph * pw;
    const int64_t svox = (int64_t)r.D * r.H * r.W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pvox; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % pw);
        const int64_t t = i / pw;
        const int h = (int)(t % ph), d = (int)(t / ph);
        const int sd = r.d0 + ((r.flip & 1) ? pd - 1 - d : d);
        const int sh = r.h0 + ((r.flip & 2) ? ph - 1 - h : h);
        const int sw = r.w0 + ((r.flip & 4) ? pw - 1 - w : w);
        const int64_t s = ((int64_t)sd * r.H + sh) * r.W + sw;
        for (int c = 0; c < C; ++c) out_img[((int64_t)n * C + c) * pvox + i] = __ldg(r.image + (int64_t)c * svox + s);
        if (out_lab != nullptr) out_lab[(int64_t)n * pvox + i] = r.label != nullptr ? __ldg(r.label + s) : (uint8_t)0;
        if (out_code != nullptr) out_code[(int64_t)n * pvox + i] = r.code != nullptr ? __ldg(r.code + s) : (uint8_t)2;
    }
}

}  // namespace

extern "C" int fpl_gather_patches(const void* d_rows, int n, int c, int pd, int ph, int pw, float* out_image,
                                  uint8_t* out_label, uint8_t* out_code, void* stream) {
    FPL_REQUIRE(n >= 1 && n <= 65535 && c >= 1 && pd >= 1 && ph >= 1 && pw >= 1, "fpl_gather_patches: bad sizes");
    FPL_REQUIRE(d_rows != nullptr && out_image != nullptr, "fpl_gather_patches: NULL argument");
    FPL_REQUIRE((reinterpret_cast<uintptr_t>(d_rows) & 15) == 0, "fpl_gather_patches: table must be 16-byte aligned");
    const int64_t pvox = (int64_t)pd * ph * pw;
    int bx = (int)((pvox + 256 * 4 - 1) / (256 * 4));
    if (bx > FPL_NUM_SMS * 4) bx = FPL_NUM_SMS * 4;
    if (bx < 1) bx = 1;
    fpl_launch(gather_patches_kernel, dim3(bx, n), 256, 0, (cudaStream_t)stream, reinterpret_cast<const PatchRow*>(d_rows), c, pd,
               ph, pw, out_image, out_label, out_code);
    FPL_LAUNCH_CHECK();
    return 0;
}
