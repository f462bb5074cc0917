// gsb_gdf_body.h -- per-pixel bodies of the gradient-domain-fusion kernels (gsb_gdf.cu).
//
// Plain C++ on purpose: the same functions are compiled by nvcc into the kernels and by g++ into the CPU
// check of tests/test_gdf_host.py (GSB_HD expands to `inline` there), so the indexing of the interleaved
// cv::Mat layouts is verified on a machine without a GPU as well.  Not an alternative compute path: nothing
// in the library calls these on the host.
//
// Layouts (reference: project/src/PhotoMontage/PhotoMontage.cpp)
//   images  n_images x (H x W x 3) bytes, channel-interleaved (cv::Mat CV_8UC3, continuous), one after another
//   labels  H x W bytes (ResultLabel.at<uchar>(y, x))
//   gx, gy  3 planes of H x W float32: plane c = component c of the Vec3f the reference stores (:419-425)
//   x       3 planes of H x W doubles (one solve per channel, :428-433)
#pragma once
#include <stdint.h>

#ifndef GSB_HD
#ifdef __CUDACC__
#define GSB_HD __host__ __device__ __forceinline__
#else
#define GSB_HD inline
#endif
#endif

// GradientAt of the labelled source image (PhotoMontage.cpp:399-408, called at :419-425 for x < W-1, y < H-1):
//   grad_x = Image(y, x+1) - Image(y, x), grad_y = Image(y+1, x) - Image(y, x), per channel, Vec3i arithmetic
// stored as float.  The reference leaves the last row and column of the gradient images unset and never reads
// them (SolveChannel loops to H-1 / W-1); they are written as 0 here.  Returns 1 when the label is out of range.
GSB_HD int gdf_gradient_at(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                           int64_t p, float *gx, float *gy) {
    const int64_t n = (int64_t)W * H;
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    const int l = labels[p];
    if (x >= W - 1 || y >= H - 1) {
        for (int c = 0; c < 3; ++c) {
            gx[c * n + p] = 0.0f;
            gy[c * n + p] = 0.0f;
        }
        return l >= n_images ? 1 : 0;
    }
    if (l >= n_images) return 1;
    const unsigned char *img = images + (int64_t)l * n * 3;
    for (int c = 0; c < 3; ++c) {
        const int c1 = img[p * 3 + c], c2 = img[(p + 1) * 3 + c], c3 = img[(p + W) * 3 + c];
        gx[c * n + p] = (float)(c2 - c1);
        gy[c * n + p] = (float)(c3 - c1);
    }
    return 0;
}

// fast_init_value (PhotoMontage.cpp:599-610): init[y*W + x] = Images[label(y, x)](y, x)[channel]
GSB_HD int gdf_composite_at(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                            int64_t p, double *x0) {
    const int64_t n = (int64_t)W * H;
    const int l = labels[p];
    if (l >= n_images) return 1;
    const unsigned char *img = images + (int64_t)l * n * 3;
    for (int c = 0; c < 3; ++c) x0[c * n + p] = (double)img[p * 3 + c];
    return 0;
}

// write-back of the three solved channels into the interleaved result (PhotoMontage.cpp:617-626):
//   output(y, x)[channel] = uchar(max(min(solution, 255.0), 0.0)) -- truncation
GSB_HD void gdf_writeback_at(const double *x, int64_t n, int64_t p, unsigned char *out) {
    for (int c = 0; c < 3; ++c) {
        double v = x[c * n + p];
        if (v > 255.0) v = 255.0;
        if (!(v > 0.0)) v = 0.0;
        out[p * 3 + c] = (unsigned char)(int)v;
    }
}
