// gsb_poisson.cu -- the Poisson system the reference's SolveChannel builds,
// generated on the device in closed form instead of through Eigen's A^T*A.
//
// Reference: project/src/PhotoMontage/PhotoMontage.cpp:541-597 == labs/lab8/src/OpenCVHW1/hw8_pa.cc:911-967.
// The over-determined system has, for every pixel with x < W-1 and y < H-1, a row
// v(x+1,y)-v(x,y) = gx(x,y) and a row v(x,y+1)-v(x,y) = gy(x,y), plus the pin v(0,0) = constraint.
// A^T*A is the graph Laplacian of exactly those edges (+1 at pixel 0); pixel (W-1,H-1) has no
// edge and therefore an empty row.  A^T*b sums the incident gradients with sign.
#include "gsb_internal.cuh"

struct PoissonEdges {
    bool l, r, u, d;
};

__device__ __forceinline__ PoissonEdges poisson_edges(int x, int y, int W, int H) {
    PoissonEdges e;
    e.l = x >= 1 && y < H - 1;     // edge (x-1,y)-(x,y) exists iff its left end has y < H-1
    e.r = x < W - 1 && y < H - 1;
    e.u = y >= 1 && x < W - 1;     // edge (x,y-1)-(x,y) exists iff its upper end has x < W-1
    e.d = x < W - 1 && y < H - 1;
    return e;
}

// row lengths for rows [p0, p1) of the full grid; len[p1-p0] = 0 (scan sentinel)
__global__ void __launch_bounds__(256) poisson_row_len(int W, int H, int64_t p0, int64_t p1, int *__restrict__ len) {
    int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    int64_t p = p0 + q;
    if (p > p1) return;
    if (p == p1) {
        len[q] = 0;
        return;
    }
    int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    PoissonEdges e = poisson_edges(x, y, W, H);
    int deg = (int)e.l + (int)e.r + (int)e.u + (int)e.d + (p == 0 ? 1 : 0);
    len[q] = deg ? deg - (p == 0 ? 1 : 0) + 1 : 0;
}

// fills compressed CSR rows (ascending columns: up, left, diag, right, down).
// col_base is subtracted from every column (0 for a whole matrix; strips keep global columns).
__global__ void __launch_bounds__(256) poisson_fill(int W, int H, int64_t p0, int64_t p1, const int *__restrict__ rp,
                                                    int *__restrict__ ci, double *__restrict__ va) {
    int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    int64_t p = p0 + q;
    if (p >= p1) return;
    int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    PoissonEdges e = poisson_edges(x, y, W, H);
    int deg = (int)e.l + (int)e.r + (int)e.u + (int)e.d + (p == 0 ? 1 : 0);
    int k = rp[q];
    if (e.u) { ci[k] = (int)(p - W); va[k++] = -1.0; }
    if (e.l) { ci[k] = (int)(p - 1); va[k++] = -1.0; }
    if (deg) { ci[k] = (int)p; va[k++] = (double)deg; }
    if (e.r) { ci[k] = (int)(p + 1); va[k++] = -1.0; }
    if (e.d) { ci[k] = (int)(p + W); va[k++] = -1.0; }
}

// launchers for other translation units (the strip generator in gsb_dist.cu)
int gsb_poisson_launch_row_len(int W, int H, int64_t p0, int64_t p1, int *len, cudaStream_t st) {
    poisson_row_len<<<(unsigned)((p1 - p0 + 1 + 255) / 256), 256, 0, st>>>(W, H, p0, p1, len);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}
int gsb_poisson_launch_fill(int W, int H, int64_t p0, int64_t p1, const int *rp, int *ci, double *va,
                            cudaStream_t st) {
    if (p1 <= p0) return GSB_OK;
    poisson_fill<<<(unsigned)((p1 - p0 + 255) / 256), 256, 0, st>>>(W, H, p0, p1, rp, ci, va);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

// row arrays exactly as initializeFromEigenRowMajor leaves them for a compressed matrix:
// trailing empty rows whose offset equals n_values are decremented to stay in bounds (v2 :608-614)
__global__ void __launch_bounds__(256) rowlen_from_rp(const int *__restrict__ rp, int n, int total,
                                                      int *__restrict__ begin, int *__restrict__ nnz,
                                                      int *__restrict__ left) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        int o = rp[i];
        nnz[i] = rp[i + 1] - o;
        left[i] = 0;
        begin[i] = (o == total) ? o - 1 : o;
    }
}

extern "C" int gsb_poisson_matrix(gsb_matrix *m, int W, int H) {
    if (!m || W < 1 || H < 1) return GSB_ERR_ARG;
    if (m->vtype != GSB_F64) {
        gsb_set_error("poisson_matrix: needs a GSB_F64 matrix");
        return GSB_ERR_ARG;
    }
    const int64_t n = (int64_t)W * H;
    const int64_t nnz_bound = (n - 1) + 4 * (int64_t)(W - 1) * (H - 1) + 1;
    if (n > INT32_MAX - 1 || nnz_bound > INT32_MAX - 1) {
        gsb_set_error("poisson_matrix: %d x %d exceeds the int32 index range", W, H);
        return GSB_ERR_OVERFLOW;
    }
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    m->drop_analysis();
    DevBuf<int> rp;
    GSB_TRY(rp.alloc(n + 1));
    poisson_row_len<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(W, H, 0, n, rp.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(gsb_exclusive_scan_i32(rp.p, rp.p, n + 1, nullptr, st));
    int total = 0; // == (n-1) + 4(W-1)(H-1) when W,H >= 2
    GSB_CUDA(cudaMemcpyAsync(&total, rp.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    const int64_t nnz = total;
    GSB_TRY(m->values_raw.alloc(nnz * 8));
    GSB_TRY(m->cols.alloc(nnz));
    GSB_TRY(m->row_begin.alloc(n));
    GSB_TRY(m->row_nnz.alloc(n));
    GSB_TRY(m->row_left.alloc(n));
    m->store = nnz;
    m->n_rows = (int)n;
    m->n_cols = (int)n;
    poisson_fill<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, H, 0, n, rp.p, m->cols.p, (double *)m->values_raw.p);
    GSB_KERNEL_CHECK();
    rowlen_from_rp<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rp.p, (int)n, total, m->row_begin.p, m->row_nnz.p,
                                                               m->row_left.p);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaStreamSynchronize(st));
    return gsb_matrix_finish_layout(m);
}

// A^T*b, summed in ascending equation order (gy above, gx left, gx here, gy here, pin) with
// separate roundings -- the order a row-major sparse product visits column p of A.
// Rows [y0,y1) of the image; gx/gy hold image rows [ybase, y1) (ybase = max(y0-1,0)), one plane per channel.
__global__ void __launch_bounds__(256) poisson_rhs_kernel(int W, int H, int y0, int y1, int ybase, int nch,
                                                          const float *__restrict__ gx, const float *__restrict__ gy,
                                                          double c0, double c1, double c2, double c3,
                                                          double *__restrict__ b) {
    const int64_t n_local = (int64_t)W * (y1 - y0);
    const int64_t plane = (int64_t)W * (y1 - ybase);
    int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (q >= n_local) return;
    const int64_t p = (int64_t)y0 * W + q;               // global pixel
    const int64_t g = p - (int64_t)ybase * W;            // index inside the gradient planes
    int y = (int)(p / W), x = (int)(p - (int64_t)y * W);
    PoissonEdges e = poisson_edges(x, y, W, H);
    for (int ch = 0; ch < nch; ++ch) {
        const float *GX = gx + ch * plane, *GY = gy + ch * plane;
        double s = 0.0;
        if (e.u) s = __dadd_rn(s, (double)GY[g - W]);
        if (e.l) s = __dadd_rn(s, (double)GX[g - 1]);
        if (e.r) s = __dsub_rn(s, (double)GX[g]);
        if (e.d) s = __dsub_rn(s, (double)GY[g]);
        if (p == 0) s = __dadd_rn(s, ch == 0 ? c0 : ch == 1 ? c1 : ch == 2 ? c2 : c3);
        b[ch * n_local + q] = s;
    }
}

extern "C" int gsb_poisson_rhs_rows_dev(int W, int H, int y0, int y1, int nch, const float *gx_rows_dev,
                                        const float *gy_rows_dev, const double *constraint, double *b_dev) {
    if (W < 1 || H < 1 || y0 < 0 || y1 > H || y0 >= y1 || nch < 1 || nch > 4 || !gx_rows_dev || !gy_rows_dev ||
        !constraint || !b_dev)
        return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    cudaStream_t st = gsb_cur_stream();
    const int64_t n_local = (int64_t)W * (y1 - y0);
    const int ybase = y0 > 0 ? y0 - 1 : 0;
    double c[4] = {0, 0, 0, 0};
    for (int i = 0; i < nch; ++i) c[i] = constraint[i];
    poisson_rhs_kernel<<<(unsigned)((n_local + 255) / 256), 256, 0, st>>>(W, H, y0, y1, ybase, nch, gx_rows_dev,
                                                                         gy_rows_dev, c[0], c[1], c[2], c[3], b_dev);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

extern "C" int gsb_poisson_rhs_rows(int W, int H, int y0, int y1, int nch, const float *gx_rows, const float *gy_rows,
                                    const double *constraint, double *b_out) {
    if (W < 1 || H < 1 || y0 < 0 || y1 > H || y0 >= y1 || nch < 1 || nch > 4 || !gx_rows || !gy_rows || !constraint ||
        !b_out)
        return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    cudaStream_t st = gsb_cur_stream();
    const int ybase = y0 > 0 ? y0 - 1 : 0;
    const int64_t ng = (int64_t)W * (y1 - ybase) * nch, nb = (int64_t)W * (y1 - y0) * nch;
    DevBuf<float> dgx, dgy;
    DevBuf<double> db;
    GSB_TRY(dgx.alloc(ng));
    GSB_TRY(dgy.alloc(ng));
    GSB_TRY(db.alloc(nb));
    GSB_CUDA(cudaMemcpyAsync(dgx.p, gx_rows, sizeof(float) * (size_t)ng, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(dgy.p, gy_rows, sizeof(float) * (size_t)ng, cudaMemcpyHostToDevice, st));
    GSB_TRY(gsb_poisson_rhs_rows_dev(W, H, y0, y1, nch, dgx.p, dgy.p, constraint, db.p));
    GSB_CUDA(cudaMemcpyAsync(b_out, db.p, sizeof(double) * (size_t)nb, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_poisson_rhs_dev(int W, int H, int nch, const float *gx_dev, const float *gy_dev,
                                   const double *constraint, double *b_dev) {
    return gsb_poisson_rhs_rows_dev(W, H, 0, H, nch, gx_dev, gy_dev, constraint, b_dev);
}

extern "C" int gsb_poisson_rhs(int W, int H, int nch, const float *gx, const float *gy, const double *constraint,
                               double *b_out) {
    return gsb_poisson_rhs_rows(W, H, 0, H, nch, gx, gy, constraint, b_out);
}

// A10: uchar(max(min(v, 255), 0)), truncating (PhotoMontage.cpp:622)
__global__ void __launch_bounds__(256) writeback_kernel(const double *__restrict__ x, int64_t n,
                                                        unsigned char *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double v = x[i];
        if (v > 255.0) v = 255.0;
        if (!(v > 0.0)) v = 0.0;
        out[i] = (unsigned char)(int)v;
    }
}

extern "C" int gsb_writeback_u8_dev(const double *x_dev, int64_t n, unsigned char *out_dev) {
    if (!x_dev || !out_dev || n < 0) return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    if (n == 0) return GSB_OK;
    writeback_kernel<<<gsb_blocks_for(n, 256 * 4, gsb_sm_count() * 16), 256, 0, gsb_cur_stream()>>>(x_dev, n, out_dev);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

extern "C" int gsb_writeback_u8(const double *x, int64_t n, unsigned char *out) {
    if (!x || !out || n < 0) return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    cudaStream_t st = gsb_cur_stream();
    DevBuf<double> dx;
    DevBuf<unsigned char> dout;
    GSB_TRY(dx.alloc(n));
    GSB_TRY(dout.alloc(n));
    GSB_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    GSB_TRY(gsb_writeback_u8_dev(dx.p, n, dout.p));
    GSB_CUDA(cudaMemcpyAsync(out, dout.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}
