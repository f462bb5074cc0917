// Host build of the per-row / per-pixel bodies of the lab8 panorama kernels
// (coursecomputationalphotography_b200/csrc/gsb_pano_body.h) for tests/test_pano_host.py.  Test infrastructure only.
// Each function applies the body exactly as the corresponding kernel does (run finder per row, then the copy).
#include <vector>
#include "../../coursecomputationalphotography_b200/csrc/gsb_pano_body.h"

extern "C" void host_pano_mask_image(const unsigned char *src, const unsigned char *mask, int W, int H,
                                     unsigned char *out) {
    for (int64_t p = 0; p < (int64_t)W * H; ++p) pano_mask_image_at(src, mask, p, out);
}

extern "C" void host_pano_gradients_masked(const unsigned char *img, const unsigned char *mask, int W, int H, float *gx,
                                           float *gy) {
    std::vector<int> first((size_t)H);
    for (int y = 0; y < H; ++y) first[(size_t)y] = pano_first_nonzero(mask + (size_t)y * W, W);
    for (int64_t p = 0; p < (int64_t)W * H; ++p) pano_gradients_masked_at(img, first.data(), W, H, p, gx, gy);
}
extern "C" void host_pano_gradients(const unsigned char *img, int W, int H, float *gx, float *gy) {
    for (int64_t p = 0; p < (int64_t)W * H; ++p) pano_gradients_at(img, W, H, p, gx, gy);
}

extern "C" void host_pano_merge2_f32(float *target, const float *src, const unsigned char *target_mask,
                                     const unsigned char *outer, const unsigned char *inner, int W, int H) {
    for (int i = 0; i < H; ++i) {
        const int64_t o = (int64_t)i * W;
        const PanoRun r = pano_merge2_run(target_mask + o, outer + o, inner + o, W);
        for (int64_t j = 0; j < (int64_t)3 * r.count; ++j) target[(o + r.start) * 3 + j] = src[(o + r.start) * 3 + j];
    }
}

extern "C" void host_pano_merge_u8(unsigned char *target, const unsigned char *src, const unsigned char *target_mask,
                                   const unsigned char *src_mask, int channel, double skip, int W, int H) {
    // runs first, copies afterwards: the kernels do the same, which is what makes target == target_mask safe
    PanoRun *runs = new PanoRun[H];
    for (int i = 0; i < H; ++i) runs[i] = pano_merge_run(target_mask + (int64_t)i * W, src_mask + (int64_t)i * W, W, skip);
    for (int i = 0; i < H; ++i) {
        const int64_t o = (int64_t)i * W + runs[i].start;
        for (int64_t j = 0; j < (int64_t)channel * runs[i].count; ++j) target[o * channel + j] = src[o * channel + j];
    }
    delete[] runs;
}

extern "C" void host_pano_enforce_gradient_bound(float *dx, float *dy, const unsigned char *src,
                                                 const unsigned char *mask, int W, int H) {
    for (int64_t p = 0; p < (int64_t)W * H; ++p) pano_enforce_bound_at(src, mask, W, H, p, dx, dy);
}
