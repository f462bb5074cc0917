// gsb_pano_body.h -- per-row / per-pixel bodies of the lab8 panorama kernels (gsb_pano.cu), "next" row N3:
// the producers of the gradient-domain right-hand side in labs/lab8/src/OpenCVHW1/hw8_pa.cc.
//
// Plain C++ on purpose (see gsb_gdf_body.h): nvcc compiles these into the kernels, g++ compiles them into the CPU
// check of tests/test_pano_host.py.  Nothing in the library calls them on the host.
//
// The reference's merge functions are written as pointer walks along a row that are not all bounded by the row
// length; on a continuous cv::Mat such a walk reads the next row's bytes.  Every such walk ends in "copy nothing"
// (the copy loop is bounded by the column count), so the bodies below use the bounded form, which gives the same
// result for every input and never leaves the row.
//
// Layouts: images H x W x 3 bytes, gradients H x W x 3 floats (CV_32FC3), masks H x W bytes -- all continuous.
#pragma once
#include <stdint.h>

#ifndef GSB_HD
#ifdef __CUDACC__
#define GSB_HD __host__ __device__ __forceinline__
#else
#define GSB_HD inline
#endif
#endif

struct PanoRun {
    int start; // first column copied
    int count; // columns copied
};

// MergeImage2<T> row scan (hw8_pa.cc:338-385): skip the columns outside the source's outer mask; skip on while the
// source's inner mask is 0 and the target is already covered; copy the rest of that outer-mask run.
GSB_HD PanoRun pano_merge2_run(const unsigned char *target_mask_row, const unsigned char *outer_row,
                               const unsigned char *inner_row, int cols) {
    int k = 0;
    while (k < cols && outer_row[k] == 0) ++k;
    while (k < cols && inner_row[k] == 0 && target_mask_row[k] != 0) ++k;
    PanoRun r;
    r.start = k;
    r.count = 0;
    while (k < cols && outer_row[k]) {
        ++k;
        ++r.count;
    }
    return r;
}

// MergeImage<T, channel> row scan (hw8_pa.cc:387-441): skip the columns outside the source mask; if the target is
// already covered there, skip `skip_how_many` more columns; copy the rest of the source-mask run.
GSB_HD PanoRun pano_merge_run(const unsigned char *target_mask_row, const unsigned char *src_mask_row, int cols,
                              double skip_how_many) {
    PanoRun r;
    int k = 0;
    while (k < cols && src_mask_row[k] == 0) ++k;
    if (k < cols && target_mask_row[k] != 0 && skip_how_many > 0) {
        int c = 0;
        while (c < skip_how_many) {
            ++k;
            ++c;
        }
    }
    r.start = k;
    r.count = 0;
    while (k < cols && src_mask_row[k]) {
        ++k;
        ++r.count;
    }
    if (r.count == 0) r.start = 0; // nothing to copy: keep the start inside the row
    return r;
}

// MaskImage (hw8_pa.cc:443-466)
GSB_HD void pano_mask_image_at(const unsigned char *src, const unsigned char *mask, int64_t p, unsigned char *out) {
    const bool keep = mask[p] != 0;
    for (int c = 0; c < 3; ++c) out[p * 3 + c] = keep ? src[p * 3 + c] : (unsigned char)0;
}

// GradientAt (hw8_pa.cc:314-323) at flat pixel p = y*W + x, interleaved output; the caller guarantees that
// p + 1 and p + W are inside the image buffer.
GSB_HD void pano_gradient_write(const unsigned char *img, int W, int64_t p, float *gx, float *gy) {
    for (int c = 0; c < 3; ++c) {
        const int color1 = img[p * 3 + c], color2 = img[(p + 1) * 3 + c], color3 = img[(p + W) * 3 + c];
        gx[p * 3 + c] = (float)(color2 - color1);
        gy[p * 3 + c] = (float)(color3 - color1);
    }
}

// struct Gradients, first constructor (hw8_pa.cc:604-636): GradientAt for y < H-1, x < W-1, 0 elsewhere (the
// reference leaves those entries unset)
GSB_HD void pano_gradients_at(const unsigned char *img, int W, int H, int64_t p, float *gx, float *gy) {
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    if (x < W - 1 && y < H - 1) {
        pano_gradient_write(img, W, p, gx, gy);
    } else {
        for (int c = 0; c < 3; ++c) gx[p * 3 + c] = gy[p * 3 + c] = 0.0f;
    }
}

// struct Gradients, second constructor (hw8_pa.cc:638-676), mask-driven.  Per image row y < H-1 the reference finds
// the first non-zero mask pixel x0 (its walk does not stop at the row end, but then x >= W and nothing else happens),
// writes ZeroGradientAt (:325-334: grad_x = pixel (y, x0), grad_y = pixel (y+1, x0-1), NOT differences) at column
// x0 - 1 when x0 < W-2, and GradientAt at every x in [x0, W-2] -- the loop tests the mask pointer it stopped at, so it
// runs to the row end whatever the mask holds after x0.  Everything else stays 0.  Mat::at(y, -1) is the previous
// row's last pixel on a continuous image, so x0 == 0 writes entry (y-1, W-1); for y == 0 that lies outside the image
// and is dropped (zero guard rows, see pano_enforce_bound_at).  first[y] = x0, or W when the row has no mask pixel.
// Bounded per-pixel form: every entry is written by at most one of the three cases.
GSB_HD int pano_first_nonzero(const unsigned char *mask_row, int W) {
    int x = 0;
    while (x < W && mask_row[x] == 0) ++x;
    return x;
}
GSB_HD void pano_gradients_masked_at(const unsigned char *img, const int *first, int W, int H, int64_t p, float *gx,
                                     float *gy) {
    const int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    float ox[3] = {0.f, 0.f, 0.f}, oy[3] = {0.f, 0.f, 0.f};
    const int x0 = y < H - 1 ? first[y] : W;
    if (x0 < W && x >= x0 && x <= W - 2) { // GradientAt
        for (int c = 0; c < 3; ++c) {
            const int color1 = img[p * 3 + c], color2 = img[(p + 1) * 3 + c], color3 = img[(p + W) * 3 + c];
            ox[c] = (float)(color2 - color1);
            oy[c] = (float)(color3 - color1);
        }
    } else if (x0 < W - 2 && x0 >= 1 && x == x0 - 1) { // ZeroGradientAt(m, x0 - 1, y)
        for (int c = 0; c < 3; ++c) {
            ox[c] = (float)img[(p + 1) * 3 + c];
            oy[c] = (float)img[(p + W) * 3 + c];
        }
    } else if (x == W - 1 && y + 1 < H - 1 && first[y + 1] == 0 && 0 < W - 2) { // row y+1's ZeroGradientAt at column -1
        for (int c = 0; c < 3; ++c) {
            ox[c] = (float)img[(p + 1) * 3 + c];
            oy[c] = (float)img[(p + W) * 3 + c];
        }
    }
    for (int c = 0; c < 3; ++c) {
        gx[p * 3 + c] = ox[c];
        gy[p * 3 + c] = oy[c];
    }
}

// EnforceGradientBound (hw8_pa.cc:468-498) for mask pixel p = i*W + j: GradientAt(src) into rows i, i-1, i+1 of
// dx / dy (Mat::at on a continuous Mat: column W-1 reads the next row's first pixel).  The walk leaves the images at
// the first and last row (rows -1 and H are written, row H is read) -- undefined upstream.  Semantics here, pinned to
// the compiled reference running over images with zero guard rows (see tests/test_pano_vs_reference.py): pixels outside the
// image read as 0, writes outside it are dropped.  Concurrent pixels may write the same entry; they write the same
// value.
GSB_HD void pano_enforce_bound_at(const unsigned char *src, const unsigned char *mask, int W, int H, int64_t p,
                                  float *dx, float *dy) {
    if (!mask[p]) return;
    const int i = (int)(p / W);
    const int64_t j = p - (int64_t)i * W, n = (int64_t)W * H;
    for (int t = 0; t < 3; ++t) {
        const int r = t == 0 ? i : (t == 1 ? i - 1 : i + 1);
        if (r < 0 || r > H - 1) continue;
        const int64_t q = (int64_t)r * W + j;
        for (int c = 0; c < 3; ++c) {
            const int color1 = src[q * 3 + c];
            const int color2 = q + 1 < n ? src[(q + 1) * 3 + c] : 0, color3 = q + W < n ? src[(q + W) * 3 + c] : 0;
            dx[q * 3 + c] = (float)(color2 - color1);
            dy[q * 3 + c] = (float)(color3 - color1);
        }
    }
}
