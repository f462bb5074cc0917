// gsb_gdf.cu -- the gradient-domain-fusion driver around the solve ("next" rows N2 + N4 of SURVEY 8f):
//   GradientAt + label-driven pick       project/src/PhotoMontage/PhotoMontage.cpp:399-408, :419-425
//   initial guess from the composite     PhotoMontage.cpp:599-610
//   SolveChannel x 3                     PhotoMontage.cpp:535-628 == labs/lab8/src/OpenCVHW1/hw8_pa.cc:902-986
//   BuildSolveGradientFusion             PhotoMontage.cpp:410-433
// Everything between the caller's host images and the fused host image runs on the device: gradients, A^T b,
// the Poisson matrix (closed form, gsb_poisson.cu), the solve (three channels fused for Gauss-Seidel), clamp.
// The per-pixel arithmetic lives in gsb_gdf_body.h.
#include "gsb_internal.cuh"
#include "gsb_gdf_body.h"

#include <chrono>
#include <new>

__global__ void __launch_bounds__(256) gdf_gradients_kernel(const unsigned char *__restrict__ images, int n_images,
                                                            const unsigned char *__restrict__ labels, int W, int H,
                                                            float *__restrict__ gx, float *__restrict__ gy,
                                                            int *__restrict__ bad) {
    const int64_t n = (int64_t)W * H;
    int any = 0;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        any |= gdf_gradient_at(images, n_images, labels, W, H, p, gx, gy);
    if (any) *bad = 1;
}

__global__ void __launch_bounds__(256) gdf_composite_kernel(const unsigned char *__restrict__ images, int n_images,
                                                            const unsigned char *__restrict__ labels, int W, int H,
                                                            double *__restrict__ x0, int *__restrict__ bad) {
    const int64_t n = (int64_t)W * H;
    int any = 0;
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        any |= gdf_composite_at(images, n_images, labels, W, H, p, x0);
    if (any) *bad = 1;
}

__global__ void __launch_bounds__(256) gdf_writeback_kernel(const double *__restrict__ x, int64_t n,
                                                            unsigned char *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < n; p += (int64_t)gridDim.x * 256)
        gdf_writeback_at(x, n, p, out);
}

static int gdf_check_shape(const char *who, int W, int H) {
    if (W < 1 || H < 1) {
        gsb_set_error("%s: bad image size %d x %d", who, W, H);
        return GSB_ERR_ARG;
    }
    const int64_t n = (int64_t)W * H;
    const int64_t nnz_bound = (n - 1) + 4 * (int64_t)(W - 1) * (H - 1) + 1;
    if (n > INT32_MAX - 1 || nnz_bound > INT32_MAX - 1) {
        gsb_set_error("%s: %d x %d exceeds the int32 index range", who, W, H);
        return GSB_ERR_OVERFLOW;
    }
    return GSB_OK;
}

// images / labels (device) -> gx, gy (device, 3 planes each).  Errors if a label is >= n_images.
static int gdf_gradients_dev(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                             float *gx, float *gy, cudaStream_t st) {
    const int64_t n = (int64_t)W * H;
    DevBuf<int> bad;
    GSB_TRY(bad.alloc(1));
    GSB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    gdf_gradients_kernel<<<gsb_blocks_for(n, 256, gsb_sm_count() * 16), 256, 0, st>>>(images, n_images, labels, W, H, gx,
                                                                                    gy, bad.p);
    GSB_KERNEL_CHECK();
    int h = 0;
    GSB_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h) {
        gsb_set_error("gdf: the label map refers to an image >= n_images (%d)", n_images);
        return GSB_ERR_ARG;
    }
    return GSB_OK;
}

static int gdf_composite_dev(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                             double *x0, cudaStream_t st) {
    const int64_t n = (int64_t)W * H;
    DevBuf<int> bad;
    GSB_TRY(bad.alloc(1));
    GSB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    gdf_composite_kernel<<<gsb_blocks_for(n, 256, gsb_sm_count() * 16), 256, 0, st>>>(images, n_images, labels, W, H, x0,
                                                                                    bad.p);
    GSB_KERNEL_CHECK();
    int h = 0;
    GSB_CUDA(cudaMemcpyAsync(&h, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h) {
        gsb_set_error("gdf: the label map refers to an image >= n_images (%d)", n_images);
        return GSB_ERR_ARG;
    }
    return GSB_OK;
}

// host images + labels -> device copies
struct GdfInputs {
    DevBuf<unsigned char> images, labels;
};
static int gdf_upload(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                      GdfInputs *in, cudaStream_t st) {
    const int64_t n = (int64_t)W * H;
    GSB_TRY(in->images.alloc(n * 3 * n_images));
    GSB_TRY(in->labels.alloc(n));
    GSB_CUDA(cudaMemcpyAsync(in->images.p, images, (size_t)(n * 3 * n_images), cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(in->labels.p, labels, (size_t)n, cudaMemcpyHostToDevice, st));
    return GSB_OK;
}

static int gdf_check_images(const char *who, const void *images, int n_images, const void *labels, int W, int H,
                            const void *out) {
    if (!images || !labels || !out || n_images < 1 || n_images > 256) {
        gsb_set_error("%s: bad argument (1..256 images, non-null buffers)", who);
        return GSB_ERR_ARG;
    }
    GSB_TRY(gdf_check_shape(who, W, H));
    return gsb_ensure_device();
}

extern "C" void gsb_gdf_default_options(gsb_gdf_options *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->solver = GSB_GDF_GS;
    o->max_iteration = 1000; // v2 :350
    o->epsilon = 1e-6;       // v2 :350
    gsb_gs_default_options(&o->gs);
}

extern "C" int gsb_gdf_gradients(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                                 float *gx, float *gy) {
    GSB_TRY(gdf_check_images("gdf_gradients", images, n_images, labels, W, H, gx));
    if (!gy) return GSB_ERR_ARG;
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    GdfInputs in;
    GSB_TRY(gdf_upload(images, n_images, labels, W, H, &in, st));
    DevBuf<float> dgx, dgy;
    GSB_TRY(dgx.alloc(3 * n));
    GSB_TRY(dgy.alloc(3 * n));
    GSB_TRY(gdf_gradients_dev(in.images.p, n_images, in.labels.p, W, H, dgx.p, dgy.p, st));
    GSB_CUDA(cudaMemcpyAsync(gx, dgx.p, sizeof(float) * (size_t)(3 * n), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaMemcpyAsync(gy, dgy.p, sizeof(float) * (size_t)(3 * n), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_gdf_composite(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                                 double *x0) {
    GSB_TRY(gdf_check_images("gdf_composite", images, n_images, labels, W, H, x0));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    GdfInputs in;
    GSB_TRY(gdf_upload(images, n_images, labels, W, H, &in, st));
    DevBuf<double> dx;
    GSB_TRY(dx.alloc(3 * n));
    GSB_TRY(gdf_composite_dev(in.images.p, n_images, in.labels.p, W, H, dx.p, st));
    GSB_CUDA(cudaMemcpyAsync(x0, dx.p, sizeof(double) * (size_t)(3 * n), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// the Poisson matrix the driver keeps between calls (one per process: the reference builds the same
// matrix three times per image, PhotoMontage.cpp:428-433; here it is built once per image size)
// ---------------------------------------------------------------------------------------------
static gsb_matrix *g_gdf_matrix = nullptr;
static int g_gdf_W = 0, g_gdf_H = 0, g_gdf_device = -1;

extern "C" int gsb_gdf_release(void) {
    if (g_gdf_matrix) gsb_matrix_destroy(g_gdf_matrix);
    g_gdf_matrix = nullptr;
    g_gdf_W = g_gdf_H = 0;
    g_gdf_device = -1;
    return GSB_OK;
}

static int gdf_matrix_for(int W, int H, gsb_matrix **out) {
    const int dev = gsb_current_device();
    if (g_gdf_matrix && (g_gdf_W != W || g_gdf_H != H || g_gdf_device != dev)) gsb_gdf_release();
    if (!g_gdf_matrix) {
        gsb_matrix *m = nullptr;
        GSB_TRY(gsb_matrix_create(&m, GSB_F64));
        int s = gsb_poisson_matrix(m, W, H);
        if (s != GSB_OK) {
            gsb_matrix_destroy(m);
            return s;
        }
        g_gdf_matrix = m;
        g_gdf_W = W;
        g_gdf_H = H;
        g_gdf_device = dev;
    }
    *out = g_gdf_matrix;
    return GSB_OK;
}

// gx, gy (device, 3 planes), constraint[3], init (device, 3 planes, or null) -> out (device, H*W*3 interleaved)
static int gdf_core(int W, int H, const float *gx_dev, const float *gy_dev, const double *constraint,
                    const double *init_dev, const gsb_gdf_options *opts_in, unsigned char *out_dev,
                    gsb_gdf_stats *stats) {
    gsb_gdf_options opts;
    gsb_gdf_default_options(&opts);
    if (opts_in) opts = *opts_in;
    if (opts.solver != GSB_GDF_GS && opts.solver != GSB_GDF_CG) {
        gsb_set_error("gdf: unknown solver %d", opts.solver);
        return GSB_ERR_ARG;
    }
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    gsb_matrix *m = nullptr;
    GSB_TRY(gdf_matrix_for(W, H, &m));
    DevBuf<double> b, x;
    GSB_TRY(b.alloc(3 * n));
    GSB_TRY(x.alloc(3 * n));
    GSB_TRY(gsb_poisson_rhs_dev(W, H, 3, gx_dev, gy_dev, constraint, b.p));
    gsb_gdf_stats s;
    memset(&s, 0, sizeof(s));
    if (opts.solver == GSB_GDF_GS) {
        gsb_gs_stats gs;
        memset(&gs, 0, sizeof(gs));
        GSB_TRY(gsb_gs_solve_device_x0(m, b.p, init_dev, 3, opts.epsilon, opts.max_iteration, &opts.gs, x.p, &gs));
        for (int c = 0; c < 3; ++c) {
            s.iterations[c] = gs.sweeps;
            s.last_eps[c] = gs.last_eps[c];
        }
        s.solve_ms = gs.solve_ms;
    } else {
        // the reference's own call: one conjugateGradient(b, eps, iterations, init) per channel
        cudaEvent_t e0, e1;
        GSB_CUDA(cudaEventCreate(&e0));
        GSB_CUDA(cudaEventCreate(&e1));
        GSB_CUDA(cudaEventRecord(e0, st));
        // the three channels share every pass over the matrix (one SpMV for all of them); each stops on its own
        int status = gsb_cg_solve_device_multi(m, b.p, init_dev, 3, opts.epsilon, opts.max_iteration, x.p, s.iterations);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (status != GSB_OK) return status;
        s.solve_ms = ms;
    }
    for (int c = 0; c < 3; ++c) GSB_TRY(gsb_residual_l2_dev(m, b.p + c * n, x.p + c * n, &s.residual_l2[c]));
    gdf_writeback_kernel<<<gsb_blocks_for(n, 256, gsb_sm_count() * 16), 256, 0, st>>>(x.p, n, out_dev);
    GSB_KERNEL_CHECK();
    if (stats) *stats = s;
    return GSB_OK;
}

extern "C" int gsb_gdf_solve(int W, int H, const float *gx, const float *gy, const double *constraint,
                             const double *init, const gsb_gdf_options *opts, unsigned char *out, gsb_gdf_stats *stats) {
    if (!gx || !gy || !constraint || !out) {
        gsb_set_error("gdf_solve: bad argument");
        return GSB_ERR_ARG;
    }
    GSB_TRY(gdf_check_shape("gdf_solve", W, H));
    GSB_TRY(gsb_ensure_device());
    const auto t0 = std::chrono::steady_clock::now();
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    DevBuf<float> dgx, dgy;
    DevBuf<double> dinit;
    DevBuf<unsigned char> dout;
    GSB_TRY(dgx.alloc(3 * n));
    GSB_TRY(dgy.alloc(3 * n));
    GSB_TRY(dout.alloc(3 * n));
    GSB_CUDA(cudaMemcpyAsync(dgx.p, gx, sizeof(float) * (size_t)(3 * n), cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(dgy.p, gy, sizeof(float) * (size_t)(3 * n), cudaMemcpyHostToDevice, st));
    if (init) {
        GSB_TRY(dinit.alloc(3 * n));
        GSB_CUDA(cudaMemcpyAsync(dinit.p, init, sizeof(double) * (size_t)(3 * n), cudaMemcpyHostToDevice, st));
    }
    GSB_TRY(gdf_core(W, H, dgx.p, dgy.p, constraint, init ? dinit.p : nullptr, opts, dout.p, stats));
    GSB_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)(3 * n), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (stats)
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return GSB_OK;
}

extern "C" int gsb_gdf_fuse(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                            int fast_init, const gsb_gdf_options *opts, unsigned char *out, gsb_gdf_stats *stats) {
    GSB_TRY(gdf_check_images("gdf_fuse", images, n_images, labels, W, H, out));
    const auto t0 = std::chrono::steady_clock::now();
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = (int64_t)W * H;
    GdfInputs in;
    GSB_TRY(gdf_upload(images, n_images, labels, W, H, &in, st));
    DevBuf<float> dgx, dgy;
    DevBuf<double> dinit;
    DevBuf<unsigned char> dout;
    GSB_TRY(dgx.alloc(3 * n));
    GSB_TRY(dgy.alloc(3 * n));
    GSB_TRY(dout.alloc(3 * n));
    GSB_TRY(gdf_gradients_dev(in.images.p, n_images, in.labels.p, W, H, dgx.p, dgy.p, st));
    if (fast_init) {
        GSB_TRY(dinit.alloc(3 * n));
        GSB_TRY(gdf_composite_dev(in.images.p, n_images, in.labels.p, W, H, dinit.p, st));
    }
    // Vec3b color0 = Images[0].at<Vec3b>(0,0) -- image 0 whatever the label of pixel 0 (PhotoMontage.cpp:428)
    const double constraint[3] = {(double)images[0], (double)images[1], (double)images[2]};
    GSB_TRY(gdf_core(W, H, dgx.p, dgy.p, constraint, fast_init ? dinit.p : nullptr, opts, dout.p, stats));
    GSB_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)(3 * n), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (stats)
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return GSB_OK;
}
