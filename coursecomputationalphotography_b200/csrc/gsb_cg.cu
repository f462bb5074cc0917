// gsb_cg.cu -- "next" row N1: the conjugate-gradient solvers the reference's Poisson drivers
// actually call (hw8_pa.cc:972, PhotoMontage.cpp:613), on the device.
//   conjugateGradient(b, eps, max_iter, initialize)   v2 :396-434
//   conjugateGradientEigen(b, eps, max_iter)          v2 :472-535  (Jacobi-preconditioned)
// Same recurrences, same unfused a + s*b vector updates, same SpMV (storage order).  Dot products
// are tree reductions (the reference's transform_reduce leaves the order unspecified), so the
// iterates agree with the reference to rounding, not bit for bit.
#include "gsb_internal.cuh"

#include <math.h>

__global__ void __launch_bounds__(256) cg_axpy(const double *a, const double *b, double s,
                                               int64_t n, double *out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = __dadd_rn(a[i], __dmul_rn(s, b[i]));
}

__global__ void __launch_bounds__(256) cg_sub(const double *a, const double *b, int64_t n,
                                              double *out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = __dsub_rn(a[i], b[i]);
}

__global__ void __launch_bounds__(256) cg_mul(const double *a, const double *b, int64_t n,
                                              double *out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
        out[i] = __dmul_rn(a[i], b[i]);
}

// extractDiagnolColInv (v2 :472-491): 1/a_ii where the diagonal is stored and nonzero, else 1
__global__ void __launch_bounds__(256) cg_inv_diag(const double *__restrict__ vals, const int *__restrict__ cols,
                                                   const int *__restrict__ row_begin,
                                                   const int *__restrict__ row_nnz, int n_rows, int n_cols,
                                                   double *__restrict__ inv) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_cols) return;
    double r = 1.0;
    if (i < n_rows) {
        int k = row_begin[i];
        const int e = k + row_nnz[i];
        for (; k < e; ++k)
            if (cols[k] == i) {
                if (vals[k] != 0.0) r = 1.0 / vals[k];
                break;
            }
    }
    inv[i] = r;
}

struct CgWork {
    cudaStream_t st;
    int64_t n;
    int nb;
    DevBuf<double> scal;
    int dot(const double *a, const double *b, double *out) {
        GSB_TRY(gsb_dot_dev(a, b, n, scal.p, st));
        GSB_CUDA(cudaMemcpyAsync(out, scal.p, sizeof(double), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        return GSB_OK;
    }
    int axpy(const double *a, const double *b, double s, double *out) {
        cg_axpy<<<nb, 256, 0, st>>>(a, b, s, n, out);
        GSB_KERNEL_CHECK();
        return GSB_OK;
    }
};

static int cg_common_check(gsb_matrix *m, const double *b, double *x) {
    if (!m || !b || !x) return GSB_ERR_ARG;
    if (!m->has_layout) {
        gsb_set_error("conjugate_gradient: matrix holds no layout yet");
        return GSB_ERR_STATE;
    }
    if (m->n_rows != m->n_cols) {
        gsb_set_error("conjugate_gradient: matrix must be square");
        return GSB_ERR_SHAPE;
    }
    return gsb_set_device(m->device);
}

// device-pointer core of conjugateGradient (v2 :396-434): b_dev, x_dev natural order; x0_dev may be null (zero start)
int gsb_cg_solve_device(gsb_matrix *m, const double *b_dev, const double *x0_dev, double epsilon, int max_iteration,
                        double *x_dev, int *iters) {
    CgWork w;
    w.st = gsb_cur_stream();
    w.n = m->n_rows;
    w.nb = gsb_blocks_for(w.n, 256 * 4, gsb_sm_count() * 16);
    GSB_TRY(w.scal.alloc(1));
    const int64_t n = w.n;
    const size_t bytes = sizeof(double) * (size_t)n;
    DevBuf<double> x, r, r1, p, Ap;
    GSB_TRY(x.alloc(n));
    GSB_TRY(r.alloc(n));
    GSB_TRY(r1.alloc(n));
    GSB_TRY(p.alloc(n));
    GSB_TRY(Ap.alloc(n));
    if (x0_dev)
        GSB_CUDA(cudaMemcpyAsync(x.p, x0_dev, bytes, cudaMemcpyDeviceToDevice, w.st));
    else
        GSB_CUDA(cudaMemsetAsync(x.p, 0, bytes, w.st));
    GSB_TRY(gsb_spmv_dev(m, x.p, r.p)); // r0 = b - A x   (:405-407)
    cg_sub<<<w.nb, 256, 0, w.st>>>(b_dev, r.p, n, r.p);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaMemcpyAsync(p.p, r.p, bytes, cudaMemcpyDeviceToDevice, w.st));
    int cnt = 0;
    while (cnt < max_iteration) {
        double rlen, pAp, r1len;
        GSB_TRY(w.dot(r.p, r.p, &rlen));
        GSB_TRY(gsb_spmv_dev(m, p.p, Ap.p));
        GSB_TRY(w.dot(p.p, Ap.p, &pAp));
        double alpha = rlen / pAp;
        GSB_TRY(w.axpy(x.p, p.p, alpha, x.p));
        GSB_TRY(w.axpy(r.p, Ap.p, -alpha, r1.p));
        GSB_TRY(w.dot(r1.p, r1.p, &r1len));
        if (sqrt(r1len) < epsilon) break; // :425 (cnt is not incremented on the break)
        double beta = r1len / rlen;
        GSB_TRY(w.axpy(r1.p, p.p, beta, p.p));
        r.swap(r1);
        ++cnt;
    }
    if (iters) *iters = cnt;
    GSB_CUDA(cudaMemcpyAsync(x_dev, x.p, bytes, cudaMemcpyDeviceToDevice, w.st));
    GSB_CUDA(cudaStreamSynchronize(w.st));
    return GSB_OK;
}

extern "C" int gsb_conjugate_gradient(gsb_matrix *m, const double *b, double epsilon, int max_iteration,
                                      const double *x0, double *x_out, int *iters) {
    GSB_TRY(cg_common_check(m, b, x_out));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = m->n_rows;
    const size_t bytes = sizeof(double) * (size_t)n;
    DevBuf<double> db, dx;
    GSB_TRY(db.alloc(n));
    GSB_TRY(dx.alloc(n));
    GSB_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    if (x0) GSB_CUDA(cudaMemcpyAsync(dx.p, x0, bytes, cudaMemcpyHostToDevice, st));
    GSB_TRY(gsb_cg_solve_device(m, db.p, x0 ? dx.p : nullptr, epsilon, max_iteration, dx.p, iters));
    GSB_CUDA(cudaMemcpyAsync(x_out, dx.p, bytes, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_conjugate_gradient_jacobi(gsb_matrix *m, const double *b, double epsilon, int max_iteration,
                                             double *x_out, int *iters) {
    GSB_TRY(cg_common_check(m, b, x_out));
    CgWork w;
    w.st = gsb_cur_stream();
    w.n = m->n_rows;
    w.nb = gsb_blocks_for(w.n, 256 * 4, gsb_sm_count() * 16);
    GSB_TRY(w.scal.alloc(1));
    const int64_t n = w.n;
    const size_t bytes = sizeof(double) * (size_t)n;
    DevBuf<double> db, x, r, z, p, Ap, inv;
    GSB_TRY(db.alloc(n));
    GSB_TRY(x.alloc(n));
    GSB_TRY(r.alloc(n));
    GSB_TRY(z.alloc(n));
    GSB_TRY(p.alloc(n));
    GSB_TRY(Ap.alloc(n));
    GSB_TRY(inv.alloc(n));
    GSB_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, w.st));
    GSB_CUDA(cudaMemsetAsync(x.p, 0, bytes, w.st));
    cg_inv_diag<<<(int)((n + 255) / 256), 256, 0, w.st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p,
                                                         m->n_rows, m->n_cols, inv.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(gsb_spmv_dev(m, x.p, r.p));
    cg_sub<<<w.nb, 256, 0, w.st>>>(db.p, r.p, n, r.p);
    GSB_KERNEL_CHECK();
    cg_mul<<<w.nb, 256, 0, w.st>>>(r.p, inv.p, n, p.p); // p0 = M^-1 r0   (:505)
    GSB_KERNEL_CHECK();
    double olddist;
    GSB_TRY(w.dot(p.p, r.p, &olddist));
    int cnt = 0;
    while (cnt < max_iteration) {
        double pAp, err, newdist;
        GSB_TRY(gsb_spmv_dev(m, p.p, Ap.p));
        GSB_TRY(w.dot(p.p, Ap.p, &pAp));
        double alpha = olddist / pAp;
        GSB_TRY(w.axpy(x.p, p.p, alpha, x.p));
        GSB_TRY(w.axpy(r.p, Ap.p, -alpha, r.p));
        GSB_TRY(w.dot(r.p, r.p, &err));
        if (sqrt(err) < epsilon) break; // :524
        cg_mul<<<w.nb, 256, 0, w.st>>>(r.p, inv.p, n, z.p);
        GSB_KERNEL_CHECK();
        GSB_TRY(w.dot(z.p, r.p, &newdist));
        double beta = newdist / olddist;
        olddist = newdist;
        GSB_TRY(w.axpy(z.p, p.p, beta, p.p));
        ++cnt;
    }
    if (iters) *iters = cnt;
    GSB_CUDA(cudaMemcpyAsync(x_out, x.p, bytes, cudaMemcpyDeviceToHost, w.st));
    GSB_CUDA(cudaStreamSynchronize(w.st));
    return GSB_OK;
}
