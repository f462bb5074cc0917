// gsb_assembly.cu -- the reference container's bulk entry points on the device:
//   initializeFromVector          -> gsb_matrix_assemble_sorted_coo   (COO -> slack CSR, scan based)
//   initializeFromEigenRowMajor   -> gsb_matrix_import_csr
//   at / coeff                    -> gsb_matrix_at                    (batched)
// plus raw upload/download of the five layout arrays.  Layout parity with the reference is
// bit-exact; tests/test_assembly_gpu.py checks it against the CPU checker and the compiled reference.
#include "gsb_internal.cuh"

#include <new>

static inline size_t elem_size(int vtype) { return vtype == GSB_I32 ? 4 : 8; }

gsb_matrix::~gsb_matrix() {
    drop_analysis();
    if (group) gsb_dist_group_finalize(group);
    group = nullptr;
    delete plan;
    plan = nullptr;
    if (ctl_host) cudaFreeHost(ctl_host);
    ctl_host = nullptr;
    if (cg_state_host) cudaFreeHost(cg_state_host);
    cg_state_host = nullptr;
    if (b_ready_event) cudaEventDestroy((cudaEvent_t)b_ready_event);
    b_ready_event = nullptr;
    if (ev_t0) cudaEventDestroy((cudaEvent_t)ev_t0);
    if (ev_t1) cudaEventDestroy((cudaEvent_t)ev_t1);
    ev_t0 = ev_t1 = nullptr;
}

extern "C" int gsb_matrix_create(gsb_matrix **out, int vtype) {
    if (!out || (vtype != GSB_F64 && vtype != GSB_I32)) {
        gsb_set_error("gsb_matrix_create: bad argument");
        return GSB_ERR_ARG;
    }
    GSB_TRY(gsb_ensure_device());
    gsb_matrix *m = new (std::nothrow) gsb_matrix();
    if (!m) return GSB_ERR_ALLOC;
    m->vtype = vtype;
    m->device = gsb_current_device();
    *out = m;
    return GSB_OK;
}

extern "C" int gsb_matrix_destroy(gsb_matrix *m) {
    if (!m) return GSB_OK;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(gsb_cur_stream());
    delete m;
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// layout bookkeeping shared by all assembly paths
// ---------------------------------------------------------------------------------------------
__global__ void i32_to_f64(const int *__restrict__ in, double *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (double)in[i];
}

__global__ void sum_i32_to_i64(const int *__restrict__ in, int64_t n, unsigned long long *out) {
    long long s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += in[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, (unsigned long long)s);
}

int gsb_matrix_finish_layout(gsb_matrix *m) {
    cudaStream_t st = gsb_cur_stream();
    m->drop_analysis();
    GSB_TRY(gsb_tiny_alloc(m));
    unsigned long long *tot = reinterpret_cast<unsigned long long *>(m->tiny.p + GSB_TINY_TOT64);
    GSB_CUDA(cudaMemsetAsync(tot, 0, sizeof(unsigned long long), st));
    if (m->n_rows > 0) {
        sum_i32_to_i64<<<gsb_blocks_for(m->n_rows, 256 * 4, gsb_sm_count() * 8), 256, 0, st>>>(m->row_nnz.p,
                                                                                              m->n_rows, tot);
        GSB_KERNEL_CHECK();
    }
    unsigned long long h = 0;
    GSB_CUDA(cudaMemcpyAsync(&h, tot, sizeof(h), cudaMemcpyDeviceToHost, st));
    if (m->vtype == GSB_I32) {
        GSB_TRY(m->values_f64.alloc(m->store));
        if (m->store > 0) {
            i32_to_f64<<<gsb_blocks_for(m->store, 256 * 4, gsb_sm_count() * 16), 256, 0, st>>>(
                (const int *)m->values_raw.p, m->values_f64.p, m->store);
            GSB_KERNEL_CHECK();
        }
    }
    GSB_CUDA(cudaStreamSynchronize(st));
    m->nnz = (int64_t)h;
    m->has_layout = true;
    return GSB_OK;
}

static int alloc_layout(gsb_matrix *m, int64_t store, int n_rows) {
    GSB_TRY(m->values_raw.alloc(store * (int64_t)elem_size(m->vtype)));
    GSB_TRY(m->cols.alloc(store));
    GSB_TRY(m->row_begin.alloc(n_rows));
    GSB_TRY(m->row_nnz.alloc(n_rows));
    GSB_TRY(m->row_left.alloc(n_rows));
    m->store = store;
    m->n_rows = n_rows;
    m->has_layout = false;
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// A1: sorted COO -> slack CSR
// ---------------------------------------------------------------------------------------------
// status word: bit0 rows decrease, bit1 negative row/col
template <typename T>
__global__ void __launch_bounds__(256) coo_bounds_flags(const int *__restrict__ rows, const int *__restrict__ cols,
                                                        const T *__restrict__ vals, int64_t n,
                                                        int n_rows, int *__restrict__ row_begin,
                                                        int *__restrict__ nzflag, int *__restrict__ status) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i > n) return;
    if (i == n) { // sentinel so that the exclusive scan also yields the grand total at [n]
        nzflag[n] = 0;
        return;
    }
    int r = rows[i];
    int prev = i > 0 ? rows[i - 1] : -1;
    if (r < prev || r >= n_rows) atomicOr(status, 1); // r >= n_rows: some row exceeds rows[n-1]
    if (r < 0 || cols[i] < 0) atomicOr(status, 2);
    nzflag[i] = (vals[i] != T(0)) ? 1 : 0;
    // rows prev+1 .. r begin at entry i (rows without entries share their successor's begin,
    // which is what the reference's running sum over nnz+slack produces, v2 :311-318)
    if (r > prev && r >= 0 && r < n_rows)
        for (int q = (prev < -1 ? -1 : prev) + 1; q <= r; ++q) row_begin[q] = (int)i;
    if (i == n - 1 && r >= 0) // rows after the last entry (only when the caller fixed n_rows): empty, begin at n
        for (int q = r + 1; q < n_rows; ++q) row_begin[q] = (int)n;
}

__global__ void __launch_bounds__(256) coo_row_counts(const int *__restrict__ row_begin,
                                                      const int *__restrict__ nzscan, int n_rows, int64_t n,
                                                      int *__restrict__ row_nnz, int *__restrict__ row_left) {
    int r = blockIdx.x * 256 + threadIdx.x;
    if (r >= n_rows) return;
    int b = row_begin[r];
    int e = (r + 1 < n_rows) ? row_begin[r + 1] : (int)n;
    int nz = nzscan[e] - nzscan[b];
    row_nnz[r] = nz;
    row_left[r] = (e - b) - nz;
}

template <typename T>
__global__ void __launch_bounds__(256) coo_compact(const int *__restrict__ rows, const int *__restrict__ cols_in,
                                                   const T *__restrict__ vals_in, int64_t n,
                                                   const int *__restrict__ row_begin,
                                                   const int *__restrict__ row_nnz,
                                                   const int *__restrict__ nzscan, int *__restrict__ cols_out,
                                                   T *__restrict__ vals_out) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int r = rows[i];
    int b = row_begin[r];
    int nz = row_nnz[r];
    T v = vals_in[i];
    int c = cols_in[i];
    int s = nzscan[i];
    if (nzscan[i + 1] != s) { // a nonzero: goes to the front of its row in input order (v2 :299-304)
        int dst = b + (s - nzscan[b]);
        vals_out[dst] = v;
        cols_out[dst] = c;
    }
    if ((int)i - b >= nz) { // slack slot: the reference's in-place compaction never touches it
        vals_out[i] = v;
        cols_out[i] = c;
    }
}

// Device core shared by the sorted path (A1) and, after its radix sort, the triplet path.
// n_cols_override < 0: n_cols = max(col)+1 as the reference estimates it (v2 :271-275).
template <typename T>
int gsb_assemble_sorted_device(gsb_matrix *m, const int *d_rows, const int *d_cols_in, const T *d_vals_in, int64_t n,
                               int n_rows, int n_cols_override) {
    cudaStream_t st = gsb_cur_stream();
    DevBuf<int> d_scan, d_misc;
    GSB_TRY(d_scan.alloc(n + 1));
    GSB_TRY(d_misc.alloc(2));
    GSB_TRY(alloc_layout(m, n, n_rows));
    GSB_CUDA(cudaMemsetAsync(d_misc.p, 0, 2 * sizeof(int), st));
    GSB_CUDA(cudaMemsetAsync(m->row_begin.p, 0, sizeof(int) * (size_t)n_rows, st));
    if (n == 0) { // only reachable from the triplet path: every row empty
        GSB_CUDA(cudaMemsetAsync(m->row_nnz.p, 0, sizeof(int) * (size_t)n_rows, st));
        GSB_CUDA(cudaMemsetAsync(m->row_left.p, 0, sizeof(int) * (size_t)n_rows, st));
        m->n_cols = n_cols_override < 0 ? 1 : n_cols_override;
        return gsb_matrix_finish_layout(m);
    }
    unsigned nb = (unsigned)((n + 1 + 255) / 256);
    coo_bounds_flags<T><<<nb, 256, 0, st>>>(d_rows, d_cols_in, d_vals_in, n, n_rows, m->row_begin.p, d_scan.p,
                                           d_misc.p);
    GSB_KERNEL_CHECK();
    int h_status = 0;
    GSB_CUDA(cudaMemcpyAsync(&h_status, d_misc.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h_status & 2) {
        gsb_set_error("assemble_sorted_coo: negative row or column index");
        return GSB_ERR_ARG;
    }
    if (h_status & 1) {
        gsb_set_error("assemble_sorted_coo: rows are not non-decreasing (input must be sorted by row, col)");
        return GSB_ERR_UNSORTED;
    }
    GSB_TRY(gsb_exclusive_scan_i32(d_scan.p, d_scan.p, n + 1, nullptr, st));
    coo_row_counts<<<(n_rows + 255) / 256, 256, 0, st>>>(m->row_begin.p, d_scan.p, n_rows, n, m->row_nnz.p,
                                                        m->row_left.p);
    GSB_KERNEL_CHECK();
    coo_compact<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_rows, d_cols_in, d_vals_in, n, m->row_begin.p,
                                                               m->row_nnz.p, d_scan.p, m->cols.p,
                                                               (T *)m->values_raw.p);
    GSB_KERNEL_CHECK();
    if (n_cols_override < 0) {
        GSB_TRY(gsb_reduce_max_i32(d_cols_in, n, d_misc.p + 1, st)); // v2 :271-275
        int h_maxcol = 0;
        GSB_CUDA(cudaMemcpyAsync(&h_maxcol, d_misc.p + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        m->n_cols = (h_maxcol < 0 ? 0 : h_maxcol) + 1;
    } else {
        m->n_cols = n_cols_override;
    }
    return gsb_matrix_finish_layout(m);
}
template int gsb_assemble_sorted_device<int>(gsb_matrix *, const int *, const int *, const int *, int64_t, int, int);
template int gsb_assemble_sorted_device<double>(gsb_matrix *, const int *, const int *, const double *, int64_t, int,
                                                int);

template <typename T>
static int assemble_sorted_t(gsb_matrix *m, const int *rows, const int *cols, const T *vals, int64_t n) {
    cudaStream_t st = gsb_cur_stream();
    if (n > (int64_t)INT32_MAX - 1) {
        gsb_set_error("assemble_sorted_coo: %lld entries exceed the int32 index range", (long long)n);
        return GSB_ERR_OVERFLOW;
    }
    int n_rows = rows[n - 1] + 1; // v2 :270 (host read, exactly as the reference does)
    if (n_rows <= 0) {
        gsb_set_error("assemble_sorted_coo: last row index %d is negative", n_rows - 1);
        return GSB_ERR_ARG;
    }
    DevBuf<int> d_rows, d_cols_in;
    DevBuf<T> d_vals_in;
    GSB_TRY(d_rows.alloc(n));
    GSB_TRY(d_cols_in.alloc(n));
    GSB_TRY(d_vals_in.alloc(n));
    GSB_CUDA(cudaMemcpyAsync(d_rows.p, rows, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(d_cols_in.p, cols, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(d_vals_in.p, vals, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st));
    return gsb_assemble_sorted_device<T>(m, d_rows.p, d_cols_in.p, d_vals_in.p, n, n_rows, -1);
}

extern "C" int gsb_matrix_assemble_sorted_coo(gsb_matrix *m, const int *rows, const int *cols, const void *vals,
                                              int64_t n) {
    if (!m || !rows || !cols || !vals || n <= 0) {
        gsb_set_error("assemble_sorted_coo: null pointer or empty input (the reference reads rows.back())");
        return GSB_ERR_ARG;
    }
    GSB_TRY(gsb_set_device(m->device));
    return m->vtype == GSB_I32 ? assemble_sorted_t<int>(m, rows, cols, (const int *)vals, n)
                               : assemble_sorted_t<double>(m, rows, cols, (const double *)vals, n);
}

// ---------------------------------------------------------------------------------------------
// A3: CSR import
// ---------------------------------------------------------------------------------------------
// first row index whose offset equals n_values (search range [0, limit))
__global__ void __launch_bounds__(256) csr_first_full(const int *__restrict__ rb, int limit, int n_values,
                                                      int *__restrict__ first) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < limit && rb[i] == n_values) atomicMin(first, i);
}

// compressed case, v2 :592-618
__global__ void __launch_bounds__(256) csr_import_compressed(const int *__restrict__ off_in, int nr, int n_values,
                                                             const int *__restrict__ first_p,
                                                             int *__restrict__ rb, int *__restrict__ nnz,
                                                             int *__restrict__ left) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nr) return;
    int istar = *first_p; // first i in [0, nr-1) with off[i]==n_values, else nr-1
    if (istar > nr - 1) istar = nr - 1;
    int o = off_in[i];
    bool tail_full = off_in[istar] == n_values;
    int z = 0;
    if (i < istar)
        z = off_in[i + 1] - o;
    else if (i == istar) {
        if (istar < nr - 1)
            z = off_in[i + 1] - o; // assigned before the break test (:598-601)
        else
            z = tail_full ? 0 : n_values - o; // :616
    }
    nnz[i] = z;
    rb[i] = (tail_full && i >= istar) ? o - 1 : o; // :608-614
    left[i] = 0;
}

// uncompressed case (per-row counts given), v2 :560-589
__global__ void __launch_bounds__(256) csr_import_counts(const int *__restrict__ off_in,
                                                         const int *__restrict__ nnz_in, int nr, int n_values,
                                                         const int *__restrict__ first_p, int *__restrict__ rb,
                                                         int *__restrict__ left) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nr) return;
    int istar = *first_p; // first i in [0, nr) with off[i]==n_values, else nr
    if (istar > nr) istar = nr;
    int last = istar > 0 ? off_in[istar - 1] + nnz_in[istar - 1] : 0;
    auto RB = [&](int k) { return k >= istar ? last : off_in[k]; };
    rb[i] = RB(i);
    if (i < nr - 1)
        left[i] = RB(i + 1) - RB(i) - nnz_in[i];
    else
        left[i] = n_values - RB(nr - 2) - nnz_in[nr - 1]; // (sic) :588
}

extern "C" int gsb_matrix_import_csr(gsb_matrix *m, const void *values, int n_values, const int *row_off,
                                     int n_row_off, const int *col_idx, int n_col_off, const int *nnz_per_row,
                                     int n_nnz_per_row) {
    if (!m || n_values < 0 || n_row_off <= 0 || n_col_off < 0 || !row_off || (n_values > 0 && (!values || !col_idx))) {
        gsb_set_error("import_csr: bad argument");
        return GSB_ERR_ARG;
    }
    if (nnz_per_row && (n_row_off < 2 || n_nnz_per_row < n_row_off)) {
        gsb_set_error("import_csr: per-row counts need >= 2 rows and n_non_zeros >= rows (reference indexes row n-2)");
        return GSB_ERR_SHAPE;
    }
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    const int nr = n_row_off;
    GSB_TRY(alloc_layout(m, n_values, nr));
    m->n_cols = n_col_off;
    DevBuf<int> &off_in = m->scratch_rows;
    GSB_TRY(off_in.alloc((int64_t)nr + 1));
    GSB_TRY(gsb_tiny_alloc(m));
    struct { int *p; } first = {m->tiny.p + GSB_TINY_FIRST};
    size_t es = elem_size(m->vtype);
    if (n_values > 0) {
        GSB_CUDA(cudaMemcpyAsync(m->values_raw.p, values, es * (size_t)n_values, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaMemcpyAsync(m->cols.p, col_idx, sizeof(int) * (size_t)n_values, cudaMemcpyHostToDevice, st));
    }
    GSB_CUDA(cudaMemcpyAsync(off_in.p, row_off, sizeof(int) * (size_t)nr, cudaMemcpyHostToDevice, st));
    int big = INT32_MAX;
    GSB_CUDA(cudaMemcpyAsync(first.p, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    const int nb = (nr + 255) / 256;
    if (!nnz_per_row) {
        if (nr > 1) {
            csr_first_full<<<(nr - 1 + 255) / 256, 256, 0, st>>>(off_in.p, nr - 1, n_values, first.p);
            GSB_KERNEL_CHECK();
        }
        csr_import_compressed<<<nb, 256, 0, st>>>(off_in.p, nr, n_values, first.p, m->row_begin.p, m->row_nnz.p,
                                                  m->row_left.p);
        GSB_KERNEL_CHECK();
    } else {
        GSB_CUDA(cudaMemcpyAsync(m->row_nnz.p, nnz_per_row, sizeof(int) * (size_t)nr, cudaMemcpyHostToDevice, st));
        csr_first_full<<<nb, 256, 0, st>>>(off_in.p, nr, n_values, first.p);
        GSB_KERNEL_CHECK();
        csr_import_counts<<<nb, 256, 0, st>>>(off_in.p, m->row_nnz.p, nr, n_values, first.p, m->row_begin.p,
                                              m->row_left.p);
        GSB_KERNEL_CHECK();
    }
    GSB_CUDA(cudaStreamSynchronize(st));
    return gsb_matrix_finish_layout(m);
}

// ---------------------------------------------------------------------------------------------
// raw upload / download / shape
// ---------------------------------------------------------------------------------------------
extern "C" int gsb_matrix_upload(gsb_matrix *m, const void *values, const int *cols, int64_t store,
                                 const int *row_begin, const int *row_nnz, const int *row_left, int n_rows,
                                 int n_cols) {
    if (!m || store < 0 || n_rows <= 0 || n_cols < 0 || !row_begin || !row_nnz || (store > 0 && (!values || !cols))) {
        gsb_set_error("matrix_upload: bad argument");
        return GSB_ERR_ARG;
    }
    if (store > INT32_MAX) return GSB_ERR_OVERFLOW;
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    GSB_TRY(alloc_layout(m, store, n_rows));
    m->n_cols = n_cols;
    size_t es = elem_size(m->vtype);
    if (store > 0) {
        GSB_CUDA(cudaMemcpyAsync(m->values_raw.p, values, es * (size_t)store, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaMemcpyAsync(m->cols.p, cols, sizeof(int) * (size_t)store, cudaMemcpyHostToDevice, st));
    }
    GSB_CUDA(cudaMemcpyAsync(m->row_begin.p, row_begin, sizeof(int) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(m->row_nnz.p, row_nnz, sizeof(int) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
    if (row_left)
        GSB_CUDA(cudaMemcpyAsync(m->row_left.p, row_left, sizeof(int) * (size_t)n_rows, cudaMemcpyHostToDevice, st));
    else
        GSB_CUDA(cudaMemsetAsync(m->row_left.p, 0, sizeof(int) * (size_t)n_rows, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return gsb_matrix_finish_layout(m);
}

extern "C" int gsb_matrix_shape(const gsb_matrix *m, int64_t *store, int *n_rows, int *n_cols, int64_t *nnz) {
    if (!m) return GSB_ERR_ARG;
    if (store) *store = m->store;
    if (n_rows) *n_rows = m->n_rows;
    if (n_cols) *n_cols = m->n_cols;
    if (nnz) *nnz = m->nnz;
    return GSB_OK;
}

extern "C" int gsb_matrix_download(const gsb_matrix *m, void *values, int *cols, int *row_begin, int *row_nnz,
                                   int *row_left) {
    if (!m) return GSB_ERR_ARG;
    if (!m->has_layout) {
        gsb_set_error("matrix_download: matrix holds no layout yet");
        return GSB_ERR_STATE;
    }
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    size_t es = elem_size(m->vtype);
    if (values && m->store)
        GSB_CUDA(cudaMemcpyAsync(values, m->values_raw.p, es * (size_t)m->store, cudaMemcpyDeviceToHost, st));
    if (cols && m->store)
        GSB_CUDA(cudaMemcpyAsync(cols, m->cols.p, sizeof(int) * (size_t)m->store, cudaMemcpyDeviceToHost, st));
    size_t rb = sizeof(int) * (size_t)m->n_rows;
    if (row_begin) GSB_CUDA(cudaMemcpyAsync(row_begin, m->row_begin.p, rb, cudaMemcpyDeviceToHost, st));
    if (row_nnz) GSB_CUDA(cudaMemcpyAsync(row_nnz, m->row_nnz.p, rb, cudaMemcpyDeviceToHost, st));
    if (row_left) GSB_CUDA(cudaMemcpyAsync(row_left, m->row_left.p, rb, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_csr_from_sorted_coo(int vtype, const int *rows, const int *cols, const void *vals, int64_t n,
                                       void *values_out, int *cols_out, int *row_begin, int *row_nnz,
                                       int *row_left, int *n_rows, int *n_cols) {
    gsb_matrix *m = nullptr;
    GSB_TRY(gsb_matrix_create(&m, vtype));
    int s = gsb_matrix_assemble_sorted_coo(m, rows, cols, vals, n);
    if (s == GSB_OK) s = gsb_matrix_download(m, values_out, cols_out, row_begin, row_nnz, row_left);
    if (s == GSB_OK) {
        if (n_rows) *n_rows = m->n_rows;
        if (n_cols) *n_cols = m->n_cols;
    }
    gsb_matrix_destroy(m);
    return s;
}

extern "C" int gsb_csr_import(int vtype, const void *values, int n_values, const int *row_off, int n_row_off,
                              const int *col_idx, int n_col_off, const int *nnz_per_row, int n_nnz_per_row,
                              void *values_out, int *cols_out, int *row_begin, int *row_nnz, int *row_left) {
    gsb_matrix *m = nullptr;
    GSB_TRY(gsb_matrix_create(&m, vtype));
    int s = gsb_matrix_import_csr(m, values, n_values, row_off, n_row_off, col_idx, n_col_off, nnz_per_row,
                                  n_nnz_per_row);
    if (s == GSB_OK) s = gsb_matrix_download(m, values_out, cols_out, row_begin, row_nnz, row_left);
    gsb_matrix_destroy(m);
    return s;
}

// ---------------------------------------------------------------------------------------------
// A4: batched at() -- the reference's clamped lower-bound search (v2 :627-645) per query
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) at_batch(const double *__restrict__ vals, const int *__restrict__ cols,
                                                const int *__restrict__ row_begin,
                                                const int *__restrict__ row_nnz, int n_rows,
                                                const int *__restrict__ qr, const int *__restrict__ qc,
                                                int64_t count, double *__restrict__ out) {
    int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (q >= count) return;
    int r = qr[q], c = qc[q];
    double res = 0.0;
    if (r >= 0 && r < n_rows && row_nnz[r] > 0) {
        int lo = row_begin[r];
        int hi = lo + row_nnz[r] - 1;
        if (cols[lo] != c) {
            while (hi > lo) {
                int mid = (hi + lo) / 2;
                if (cols[mid] < c)
                    lo = mid + 1;
                else
                    hi = mid;
            }
        }
        if (cols[lo] == c) res = vals[lo];
    }
    out[q] = res;
}

extern "C" int gsb_matrix_at(const gsb_matrix *m, const int *rows, const int *cols, int64_t count, double *out) {
    if (!m || !rows || !cols || !out || count < 0) return GSB_ERR_ARG;
    if (!m->has_layout) {
        gsb_set_error("matrix_at: matrix holds no layout yet");
        return GSB_ERR_STATE;
    }
    if (count == 0) return GSB_OK;
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    DevBuf<int> qr, qc;
    DevBuf<double> res;
    GSB_TRY(qr.alloc(count));
    GSB_TRY(qc.alloc(count));
    GSB_TRY(res.alloc(count));
    GSB_CUDA(cudaMemcpyAsync(qr.p, rows, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(qc.p, cols, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, st));
    at_batch<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p,
                                                             m->n_rows, qr.p, qc.p, count, res.p);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaMemcpyAsync(out, res.p, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}
