// Host build of the per-pixel bodies of the gradient-domain-fusion kernels
// (coursecomputationalphotography_b200/csrc/gsb_gdf_body.h), so that tests/test_gdf_host.py can check their
// indexing against the oracle without a GPU.  Test infrastructure only.
#include "../../coursecomputationalphotography_b200/csrc/gsb_gdf_body.h"

extern "C" int host_gdf_gradients(const unsigned char *images, int n_images, const unsigned char *labels, int W,
                                  int H, float *gx, float *gy) {
    int bad = 0;
    for (int64_t p = 0; p < (int64_t)W * H; ++p) bad |= gdf_gradient_at(images, n_images, labels, W, H, p, gx, gy);
    return bad;
}

extern "C" int host_gdf_composite(const unsigned char *images, int n_images, const unsigned char *labels, int W,
                                  int H, double *x0) {
    int bad = 0;
    for (int64_t p = 0; p < (int64_t)W * H; ++p) bad |= gdf_composite_at(images, n_images, labels, W, H, p, x0);
    return bad;
}

extern "C" void host_gdf_writeback(const double *x, int64_t n, unsigned char *out) {
    for (int64_t p = 0; p < n; ++p) gdf_writeback_at(x, n, p, out);
}
