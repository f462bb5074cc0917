// gsb_pano.cu -- "next" row N3 of SURVEY 8f: the producers of the gradient-domain right-hand side in the lab8
// panorama (labs/lab8/src/OpenCVHW1/hw8_pa.cc), on the device:
//   MaskImage                      :443-466      Gradients (first constructor)   :604-636 (GradientAt :314-323)
//   MergeImage2<float>             :338-385      MergeImage<uchar, channel>      :387-441
//   EnforceGradientBound           :468-498      one iteration of the stitch loop :740-768
// The warps and erosions around them are OpenCV calls and stay with the caller.  Row scans run one thread per image
// row (a run finder), the copies they decide on run one block per row; none of this is hot -- it runs once per
// source image, against thousands of sweeps -- so the kernels are kept simple.  Per-row / per-pixel arithmetic:
// gsb_pano_body.h.
#include "gsb_internal.cuh"
#include "gsb_pano_body.h"

__global__ void __launch_bounds__(256) pano_mask_image_kernel(const unsigned char *__restrict__ src,
                                                              const unsigned char *__restrict__ mask, int64_t n,
                                                              unsigned char *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        pano_mask_image_at(src, mask, p, out);
}

__global__ void __launch_bounds__(256) pano_gradients_kernel(const unsigned char *__restrict__ img, int W, int H,
                                                             float *__restrict__ gx, float *__restrict__ gy) {
    const int64_t n = (int64_t)W * H;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        pano_gradients_at(img, W, H, p, gx, gy);
}

__global__ void __launch_bounds__(128) pano_runs_merge2_kernel(const unsigned char *__restrict__ target_mask,
                                                               const unsigned char *__restrict__ outer,
                                                               const unsigned char *__restrict__ inner, int W, int H,
                                                               PanoRun *__restrict__ runs) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= H) return;
    const int64_t o = (int64_t)i * W;
    runs[i] = pano_merge2_run(target_mask + o, outer + o, inner + o, W);
}

__global__ void __launch_bounds__(128) pano_runs_merge_kernel(const unsigned char *__restrict__ target_mask,
                                                              const unsigned char *__restrict__ src_mask, int W, int H,
                                                              double skip, PanoRun *__restrict__ runs) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= H) return;
    const int64_t o = (int64_t)i * W;
    runs[i] = pano_merge_run(target_mask + o, src_mask + o, W, skip);
}

// memcpy(dp, sp, sizeof(T) * channel * count) of every row's run; one block per row
template <typename T>
__global__ void __launch_bounds__(256) pano_copy_runs_kernel(T *target, const T *__restrict__ src,
                                                             const PanoRun *__restrict__ runs, int W, int channel) {
    const int i = blockIdx.x;
    const PanoRun r = runs[i];
    const int64_t o = ((int64_t)i * W + r.start) * channel;
    const int64_t cnt = (int64_t)r.count * channel;
    for (int64_t j = threadIdx.x; j < cnt; j += 256) target[o + j] = src[o + j];
}

__global__ void __launch_bounds__(256) pano_enforce_kernel(const unsigned char *__restrict__ src,
                                                           const unsigned char *__restrict__ mask, int W, int H,
                                                           float *dx, float *dy) {
    const int64_t n = (int64_t)W * H;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        pano_enforce_bound_at(src, mask, W, H, p, dx, dy);
}

// CV_32FC3 (interleaved) -> three planes, the layout gsb_poisson_rhs / gsb_gdf_solve take
__global__ void __launch_bounds__(256) pano_split_kernel(const float *__restrict__ in, int64_t n,
                                                         float *__restrict__ planes) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        for (int c = 0; c < 3; ++c) planes[c * n + p] = in[p * 3 + c];
}

// ---------------------------------------------------------------------------------------------
// device-pointer steps
// ---------------------------------------------------------------------------------------------
static int pano_check(const char *who, int W, int H) {
    if (W < 1 || H < 1 || (int64_t)W * H > INT32_MAX - 1) {
        gsb_set_error("%s: bad image size %d x %d", who, W, H);
        return GSB_ERR_ARG;
    }
    return gsb_ensure_device();
}
static inline int pix_blocks(int64_t n) { return gsb_blocks_for(n, 256, gsb_sm_count() * 16); }

static int merge2_dev(float *target, const float *src, const unsigned char *target_mask, const unsigned char *outer,
                      const unsigned char *inner, int W, int H, PanoRun *runs, cudaStream_t st) {
    pano_runs_merge2_kernel<<<(H + 127) / 128, 128, 0, st>>>(target_mask, outer, inner, W, H, runs);
    GSB_KERNEL_CHECK();
    pano_copy_runs_kernel<float><<<H, 256, 0, st>>>(target, src, runs, W, 3);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

// target may alias target_mask (the mask merge): the runs are complete before the first byte is copied
static int merge_u8_dev(unsigned char *target, const unsigned char *src, const unsigned char *target_mask,
                        const unsigned char *src_mask, int channel, double skip, int W, int H, PanoRun *runs,
                        cudaStream_t st) {
    pano_runs_merge_kernel<<<(H + 127) / 128, 128, 0, st>>>(target_mask, src_mask, W, H, skip, runs);
    GSB_KERNEL_CHECK();
    pano_copy_runs_kernel<unsigned char><<<H, 256, 0, st>>>(target, src, runs, W, channel);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI (host pointers)
// ---------------------------------------------------------------------------------------------
template <typename T>
static int up(DevBuf<T> &d, const T *h, int64_t count, cudaStream_t st) {
    GSB_TRY(d.alloc(count));
    GSB_CUDA(cudaMemcpyAsync(d.p, h, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice, st));
    return GSB_OK;
}
template <typename T>
static int down(T *h, const DevBuf<T> &d, int64_t count, cudaStream_t st) {
    GSB_CUDA(cudaMemcpyAsync(h, d.p, sizeof(T) * (size_t)count, cudaMemcpyDeviceToHost, st));
    return GSB_OK;
}

extern "C" int gsb_pano_mask_image(const unsigned char *src, const unsigned char *mask, int W, int H,
                                   unsigned char *out) {
    if (!src || !mask || !out) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_mask_image", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<unsigned char> ds, dm, dout;
    GSB_TRY(up(ds, src, 3 * n, st));
    GSB_TRY(up(dm, mask, n, st));
    GSB_TRY(dout.alloc(3 * n));
    pano_mask_image_kernel<<<pix_blocks(n), 256, 0, st>>>(ds.p, dm.p, n, dout.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(down(out, dout, 3 * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_pano_gradients(const unsigned char *img, int W, int H, float *gx, float *gy) {
    if (!img || !gx || !gy) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_gradients", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<unsigned char> di;
    DevBuf<float> dgx, dgy;
    GSB_TRY(up(di, img, 3 * n, st));
    GSB_TRY(dgx.alloc(3 * n));
    GSB_TRY(dgy.alloc(3 * n));
    pano_gradients_kernel<<<pix_blocks(n), 256, 0, st>>>(di.p, W, H, dgx.p, dgy.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(down(gx, dgx, 3 * n, st));
    GSB_TRY(down(gy, dgy, 3 * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

// struct Gradients, second (mask-driven) constructor: first mask pixel per row, then one pass over the pixels
__global__ void __launch_bounds__(128) pano_first_nonzero_kernel(const unsigned char *__restrict__ mask, int W, int H,
                                                                 int *__restrict__ first) {
    const int y = blockIdx.x * 128 + threadIdx.x;
    if (y < H) first[y] = pano_first_nonzero(mask + (size_t)y * W, W);
}
__global__ void __launch_bounds__(256) pano_gradients_masked_kernel(const unsigned char *__restrict__ img,
                                                                    const int *__restrict__ first, int W, int H,
                                                                    float *__restrict__ gx, float *__restrict__ gy) {
    const int64_t n = (int64_t)W * H;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        pano_gradients_masked_at(img, first, W, H, p, gx, gy);
}

extern "C" int gsb_pano_gradients_masked(const unsigned char *img, const unsigned char *mask, int W, int H, float *gx,
                                         float *gy) {
    if (!img || !mask || !gx || !gy) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_gradients_masked", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<unsigned char> di, dm;
    DevBuf<float> dgx, dgy;
    DevBuf<int> first;
    GSB_TRY(up(di, img, 3 * n, st));
    GSB_TRY(up(dm, mask, n, st));
    GSB_TRY(dgx.alloc(3 * n));
    GSB_TRY(dgy.alloc(3 * n));
    GSB_TRY(first.alloc(H));
    pano_first_nonzero_kernel<<<(H + 127) / 128, 128, 0, st>>>(dm.p, W, H, first.p);
    GSB_KERNEL_CHECK();
    pano_gradients_masked_kernel<<<pix_blocks(n), 256, 0, st>>>(di.p, first.p, W, H, dgx.p, dgy.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(down(gx, dgx, 3 * n, st));
    GSB_TRY(down(gy, dgy, 3 * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_pano_merge2_f32(float *target, const float *src, const unsigned char *target_mask,
                                   const unsigned char *src_outer_mask, const unsigned char *src_inner_mask, int W,
                                   int H) {
    if (!target || !src || !target_mask || !src_outer_mask || !src_inner_mask) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_merge2_f32", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<float> dt, ds;
    DevBuf<unsigned char> dtm, dso, dsi;
    DevBuf<PanoRun> runs;
    GSB_TRY(up(dt, (const float *)target, 3 * n, st));
    GSB_TRY(up(ds, src, 3 * n, st));
    GSB_TRY(up(dtm, target_mask, n, st));
    GSB_TRY(up(dso, src_outer_mask, n, st));
    GSB_TRY(up(dsi, src_inner_mask, n, st));
    GSB_TRY(runs.alloc(H));
    GSB_TRY(merge2_dev(dt.p, ds.p, dtm.p, dso.p, dsi.p, W, H, runs.p, st));
    GSB_TRY(down(target, dt, 3 * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_pano_merge_u8(unsigned char *target, const unsigned char *src, const unsigned char *target_mask,
                                 const unsigned char *src_mask, int channel, double skip_how_many, int W, int H) {
    if (!target || !src || !target_mask || !src_mask || (channel != 1 && channel != 3)) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_merge_u8", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<unsigned char> dt, ds, dtm, dsm;
    DevBuf<PanoRun> runs;
    GSB_TRY(up(dt, (const unsigned char *)target, channel * n, st));
    GSB_TRY(up(ds, src, channel * n, st));
    GSB_TRY(up(dtm, target_mask, n, st)); // a copy: also correct when the caller passes target == target_mask
    GSB_TRY(up(dsm, src_mask, n, st));
    GSB_TRY(runs.alloc(H));
    GSB_TRY(merge_u8_dev(dt.p, ds.p, dtm.p, dsm.p, channel, skip_how_many, W, H, runs.p, st));
    GSB_TRY(down(target, dt, channel * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_pano_enforce_gradient_bound(float *dx, float *dy, const unsigned char *src,
                                               const unsigned char *mask, int W, int H) {
    if (!dx || !dy || !src || !mask) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_enforce_gradient_bound", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<float> ddx, ddy;
    DevBuf<unsigned char> ds, dm;
    GSB_TRY(up(ddx, (const float *)dx, 3 * n, st));
    GSB_TRY(up(ddy, (const float *)dy, 3 * n, st));
    GSB_TRY(up(ds, src, 3 * n, st));
    GSB_TRY(up(dm, mask, n, st));
    pano_enforce_kernel<<<pix_blocks(n), 256, 0, st>>>(ds.p, dm.p, W, H, ddx.p, ddy.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(down(dx, ddx, 3 * n, st));
    GSB_TRY(down(dy, ddy, 3 * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

// One iteration of the stitch loop (hw8_pa.cc:740-768) with the warps already done by the caller:
//   tmp_masked = MaskImage(warped, erode_mask2);  grad = Gradients(tmp_masked);
//   MergeImage2<float>(dx, grad.x, mask, erode_mask2, erode_mask);  same for dy;
//   MergeImage<uchar>(raw, warped, mask, erode_mask, 1);  MergeImage<uchar, 1>(mask, erode_mask, mask, erode_mask, 0)
// raw, dx, dy, mask are updated in place; everything between the upload and the download stays on the device.
extern "C" int gsb_pano_merge_step(unsigned char *raw, float *dx, float *dy, unsigned char *mask,
                                   const unsigned char *warped, const unsigned char *erode_mask,
                                   const unsigned char *erode_mask2, int W, int H) {
    if (!raw || !dx || !dy || !mask || !warped || !erode_mask || !erode_mask2) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_merge_step", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<unsigned char> draw, dmask, dwarp, de1, de2, dmasked;
    DevBuf<float> ddx, ddy, dgx, dgy;
    DevBuf<PanoRun> runs;
    GSB_TRY(up(draw, (const unsigned char *)raw, 3 * n, st));
    GSB_TRY(up(ddx, (const float *)dx, 3 * n, st));
    GSB_TRY(up(ddy, (const float *)dy, 3 * n, st));
    GSB_TRY(up(dmask, (const unsigned char *)mask, n, st));
    GSB_TRY(up(dwarp, warped, 3 * n, st));
    GSB_TRY(up(de1, erode_mask, n, st));
    GSB_TRY(up(de2, erode_mask2, n, st));
    GSB_TRY(dmasked.alloc(3 * n));
    GSB_TRY(dgx.alloc(3 * n));
    GSB_TRY(dgy.alloc(3 * n));
    GSB_TRY(runs.alloc(H));
    pano_mask_image_kernel<<<pix_blocks(n), 256, 0, st>>>(dwarp.p, de2.p, n, dmasked.p);
    GSB_KERNEL_CHECK();
    pano_gradients_kernel<<<pix_blocks(n), 256, 0, st>>>(dmasked.p, W, H, dgx.p, dgy.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(merge2_dev(ddx.p, dgx.p, dmask.p, de2.p, de1.p, W, H, runs.p, st));
    GSB_TRY(merge2_dev(ddy.p, dgy.p, dmask.p, de2.p, de1.p, W, H, runs.p, st));
    GSB_TRY(merge_u8_dev(draw.p, dwarp.p, dmask.p, de1.p, 3, 1.0, W, H, runs.p, st));
    GSB_TRY(merge_u8_dev(dmask.p, de1.p, dmask.p, de1.p, 1, 0.0, W, H, runs.p, st));
    GSB_TRY(down(raw, draw, 3 * n, st));
    GSB_TRY(down(dx, ddx, 3 * n, st));
    GSB_TRY(down(dy, ddy, 3 * n, st));
    GSB_TRY(down(mask, dmask, n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_pano_split_planes_f32(const float *interleaved, int W, int H, float *planes) {
    if (!interleaved || !planes) return GSB_ERR_ARG;
    GSB_TRY(pano_check("pano_split_planes_f32", W, H));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<float> din, dout;
    GSB_TRY(up(din, interleaved, 3 * n, st));
    GSB_TRY(dout.alloc(3 * n));
    pano_split_kernel<<<pix_blocks(n), 256, 0, st>>>(din.p, n, dout.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(down(planes, dout, 3 * n, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}
