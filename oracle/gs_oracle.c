/* TEST INFRASTRUCTURE ONLY -- see gs_oracle.h for the rules and the parity status.
 *
 * Plain-C restatement of the reference's slack-CSR container, its Gauss-Seidel /
 * SpMV / CG solvers and the Poisson system it is fed.  Built with
 *   gcc -std=c11 -O2 -ffp-contract=off
 * so that a*b+c is NOT contracted to an FMA: the reference's MSVC x64 /O2 build (and
 * g++ -O2 without -march) round the product and the sum separately, and the GPU
 * kernels do the same (__dmul_rn/__dadd_rn), which is what makes sweep-by-sweep
 * bit-exact comparison possible.
 *
 * "v1" = /root/reference/labs/lab3/src/OpenCVHW1/sparse-matrix.h
 * "v2" = /root/reference/labs/lab8/src/OpenCVHW1/sparse-matrix.h
 */
#include "gs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

orc_matrix *orc_new(void) { return (orc_matrix *)calloc(1, sizeof(orc_matrix)); }

static void orc_clear(orc_matrix *m) {
    free(m->values);
    free(m->cols);
    free(m->row_begin);
    free(m->row_nnz);
    free(m->row_left);
    memset(m, 0, sizeof(*m));
}

void orc_free(orc_matrix *m) {
    if (!m) return;
    orc_clear(m);
    free(m);
}

static void orc_alloc(orc_matrix *m, int64_t store, int n_rows) {
    m->store = store;
    m->cap = store > 0 ? store : 1;
    m->values = (double *)malloc(sizeof(double) * (size_t)m->cap);
    m->cols = (int *)malloc(sizeof(int) * (size_t)m->cap);
    size_t nr = n_rows > 0 ? (size_t)n_rows : 1;
    m->row_begin = (int *)calloc(nr, sizeof(int));
    m->row_nnz = (int *)calloc(nr, sizeof(int));
    m->row_left = (int *)calloc(nr, sizeof(int));
    m->n_rows = n_rows;
}

/* A1.  v2 :265-319.  Input sorted by (row, col), explicit zeros allowed.
 * Layout contract (what the reference's in-place forward compaction leaves behind):
 *   row_begin[r] = number of input entries whose row < r
 *   row_nnz[r]   = nonzeros of r,  row_left[r] = explicit zeros of r
 *   slot row_begin[r]+j holds the j-th nonzero of r for j < row_nnz[r], and otherwise
 *   still holds input entry row_begin[r]+j (stale). */
int orc_init_from_vector(orc_matrix *m, const int *rows, const int *cols, const double *vals, int64_t n) {
    if (n <= 0) return 1;
    for (int64_t i = 1; i < n; ++i)
        if (rows[i] < rows[i - 1]) return 2; /* contract violation: reference gives garbage here */
    orc_clear(m);
    int n_rows = rows[n - 1] + 1; /* v2 :270 */
    int n_cols = 0;
    for (int64_t i = 0; i < n; ++i)
        if (cols[i] > n_cols) n_cols = cols[i]; /* v2 :271-275 */
    n_cols += 1;
    orc_alloc(m, n, n_rows);
    m->n_cols = n_cols;
    memcpy(m->values, vals, sizeof(double) * (size_t)n);
    memcpy(m->cols, cols, sizeof(int) * (size_t)n);

    int64_t seg = 0;
    while (seg < n) {
        int r = rows[seg];
        int64_t end = seg;
        while (end < n && rows[end] == r) ++end;
        int64_t w = seg;
        for (int64_t i = seg; i < end; ++i) {
            if (m->values[i] == 0) continue; /* v2 :295 */
            m->values[w] = m->values[i];
            m->cols[w] = m->cols[i];
            ++w;
        }
        m->row_begin[r] = (int)seg; /* provisional; rows without entries fixed below */
        m->row_nnz[r] = (int)(w - seg);
        m->row_left[r] = (int)((end - seg) - (w - seg));
        seg = end;
    }
    /* v2 :311-318: row_begin = running sum of (nnz + slack), also for rows with no entries */
    int sum = 0;
    for (int r = 0; r < n_rows; ++r) {
        m->row_begin[r] = sum;
        sum += m->row_nnz[r] + m->row_left[r];
    }
    return 0;
}

/* A2.  v2 :332-347: dense row-major list, every entry (zeros included) becomes a COO item */
int orc_init_dense(orc_matrix *m, int n_rows, int n_cols, const double *dense) {
    int64_t n = (int64_t)n_rows * n_cols;
    int *r = (int *)malloc(sizeof(int) * (size_t)n), *c = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n_rows; ++i)
        for (int j = 0; j < n_cols; ++j) {
            r[(int64_t)i * n_cols + j] = i;
            c[(int64_t)i * n_cols + j] = j;
        }
    int rc = orc_init_from_vector(m, r, c, dense, n);
    free(r);
    free(c);
    return rc;
}

/* A3.  v2 :537-620.  n_rows = n_row_off, n_cols = n_col_off (the caller passes Eigen's
 * outerSize()/innerSize(), hw8_pa.cc:889-899). */
int orc_import_csr(orc_matrix *m, const double *values, int n_values, const int *row_off, int n_row_off,
                   const int *col_idx, int n_col_off, const int *nnz_per_row) {
    if (n_row_off <= 0 || n_values < 0) return 1;
    orc_clear(m);
    orc_alloc(m, n_values, n_row_off);
    m->n_cols = n_col_off;
    int nr = n_row_off;
    memcpy(m->values, values, sizeof(double) * (size_t)n_values);
    memcpy(m->cols, col_idx, sizeof(int) * (size_t)n_values);
    memcpy(m->row_begin, row_off, sizeof(int) * (size_t)nr);
    int *rb = m->row_begin;

    if (nnz_per_row) { /* v2 :560-589: uncompressed Eigen matrix, counts given */
        if (nr < 2) return 3; /* reference indexes row n_rows-2 unconditionally (:588) */
        memcpy(m->row_nnz, nnz_per_row, sizeof(int) * (size_t)nr);
        int i = 0;
        while (i < nr && rb[i] != n_values) ++i;
        int last = 0;
        if (i > 0) last = rb[i - 1] + m->row_nnz[i - 1];
        for (; i < nr; ++i) rb[i] = last;
        for (i = 0; i < nr - 1; ++i) m->row_left[i] = rb[i + 1] - rb[i] - m->row_nnz[i];
        m->row_left[nr - 1] = n_values - rb[nr - 2] - m->row_nnz[nr - 1]; /* (sic) :588 */
    } else { /* v2 :592-618: compressed, counts derived from consecutive offsets */
        int i = 0;
        for (; i < nr - 1; ++i) {
            m->row_nnz[i] = rb[i + 1] - rb[i];
            if (rb[i] == n_values) break;
        }
        if (rb[i] == n_values) {
            for (; i < nr; ++i) --rb[i]; /* trailing empty rows kept in bounds */
        } else {
            m->row_nnz[i] = n_values - rb[i];
        }
    }
    return 0;
}

/* v2 :627-645: lower bound inside the live range, clamped to the last live slot */
static int orc_nearest(const orc_matrix *m, int row, int col) {
    int lo = m->row_begin[row];
    int hi = lo + m->row_nnz[row] - 1;
    if (m->cols[lo] == col) return lo;
    while (hi > lo) {
        int mid = (hi + lo) / 2;
        if (m->cols[mid] < col)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

/* A4.  v2 :162-173 */
double orc_at(const orc_matrix *m, int row, int col) {
    if (m->row_nnz[row] == 0) return 0.0;
    int k = orc_nearest(m, row, col);
    return m->cols[k] == col ? m->values[k] : 0.0;
}

/* A5.  v2 :183-247.  Semantics are the dense-mirror ones the reference's own tests check
 * (main6.cc:19-33, T1-T5 at :213-231).  Three reference defects are deliberately not
 * restated: the column memmove of insertZero uses sizeof(T)
 * (:198), insertNoneZero places a column beyond the last live one *before* it (:207,:218),
 * and its slack-path memmove moves one element too few (:219). */
void orc_insert(orc_matrix *m, double val, int row, int col) {
    int rb = m->row_begin[row];
    int nz = m->row_nnz[row];
    if (val == 0) {
        if (nz == 0) return; /* T1 */
        int k = orc_nearest(m, row, col);
        if (m->cols[k] != col) return;
        int tail = (rb + nz - 1) - k; /* live entries after k */
        memmove(m->values + k, m->values + k + 1, sizeof(double) * (size_t)tail);
        memmove(m->cols + k, m->cols + k + 1, sizeof(int) * (size_t)tail);
        m->row_nnz[row] = nz - 1; /* T2 */
        m->row_left[row] += 1;
        return;
    }
    int k = rb;
    if (nz) {
        k = orc_nearest(m, row, col);
        if (m->cols[k] == col) { /* T3 */
            m->values[k] = val;
            return;
        }
        if (m->cols[k] < col) ++k; /* col is beyond every live column */
    }
    int tail = (rb + nz) - k;
    if (m->row_left[row]) { /* T4: shift right inside the row's own slack */
        memmove(m->values + k + 1, m->values + k, sizeof(double) * (size_t)tail);
        memmove(m->cols + k + 1, m->cols + k, sizeof(int) * (size_t)tail);
        m->row_left[row] -= 1;
    } else { /* T5: grow the store by one slot and bump every later row */
        if (m->store + 1 > m->cap) {
            m->cap = m->cap * 2 + 16;
            m->values = (double *)realloc(m->values, sizeof(double) * (size_t)m->cap);
            m->cols = (int *)realloc(m->cols, sizeof(int) * (size_t)m->cap);
        }
        memmove(m->values + k + 1, m->values + k, sizeof(double) * (size_t)(m->store - k));
        memmove(m->cols + k + 1, m->cols + k, sizeof(int) * (size_t)(m->store - k));
        m->store += 1;
        for (int r = row + 1; r < m->n_rows; ++r) m->row_begin[r] += 1;
    }
    m->values[k] = val;
    m->cols[k] = col;
    m->row_nnz[row] = nz + 1;
}

/* A8.  v2 :45-49 */
double orc_l1_dist(const double *a, const double *b, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s = s + fabs(a[i] - b[i]);
    return s;
}

double orc_dot(const double *a, const double *b, int64_t n) {
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s = s + a[i] * b[i];
    return s;
}

/* A6.  v2 :350-380.  x0 = 1.0 everywhere; lexicographic sweep; a zero (or absent)
 * diagonal skips the row; stop on the L1 norm of the sweep's update. */
int orc_gauss_seidel(const orc_matrix *m, const double *b, int64_t n, double epsilon, int max_iteration,
                     double *x, int *sweeps, double *last_eps) {
    double *prev = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) x[i] = 1.0;
    double eps = 10;
    int cnt = 0;
    while (eps > epsilon && cnt < max_iteration) {
        memcpy(prev, x, sizeof(double) * (size_t)n);
        for (int i = 0; i < m->n_rows; ++i) {
            double a_ii = orc_at(m, i, i);
            if (a_ii == 0) continue;
            double sigma = 0;
            int k = m->row_begin[i];
            for (int j = 0; j < m->row_nnz[i]; ++j, ++k) {
                int c = m->cols[k];
                if (c != i) sigma = sigma + m->values[k] * x[c];
            }
            x[i] = (b[i] - sigma) / a_ii;
        }
        eps = orc_l1_dist(x, prev, n);
        ++cnt;
    }
    free(prev);
    if (sweeps) *sweeps = cnt;
    if (last_eps) *last_eps = eps;
    return 0;
}

/* A6 with a trace (test infrastructure for the large goldens): the same loop as orc_gauss_seidel, run ONCE for a
 * descending list of thresholds.  The iterate of the sweep on which `eps <= eps_list[k]` first holds is exactly what
 * gaussSeidel(b, eps_list[k], max_iteration) returns (v2 :350-380: the loop tests eps before each sweep), so it is
 * copied to x_snap[k] and its sweep count to sweeps_at[k]; eps_hist[s] = the L1 update norm of sweep s+1. */
int orc_gauss_seidel_trace(const orc_matrix *m, const double *b, int64_t n, const double *eps_list, int n_eps,
                           int max_iteration, double *x_snap, int *sweeps_at, double *eps_hist) {
    double *prev = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *x = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) x[i] = 1.0;
    for (int k = 0; k < n_eps; ++k) sweeps_at[k] = -1;
    double eps = 10;
    int cnt = 0, next = 0;
    while (next < n_eps && cnt < max_iteration) {
        memcpy(prev, x, sizeof(double) * (size_t)n);
        for (int i = 0; i < m->n_rows; ++i) {
            double a_ii = orc_at(m, i, i);
            if (a_ii == 0) continue;
            double sigma = 0;
            int k = m->row_begin[i];
            for (int j = 0; j < m->row_nnz[i]; ++j, ++k) {
                int c = m->cols[k];
                if (c != i) sigma = sigma + m->values[k] * x[c];
            }
            x[i] = (b[i] - sigma) / a_ii;
        }
        eps = orc_l1_dist(x, prev, n);
        if (eps_hist) eps_hist[cnt] = eps;
        ++cnt;
        while (next < n_eps && !(eps > eps_list[next])) {
            memcpy(x_snap + (size_t)next * (size_t)n, x, sizeof(double) * (size_t)n);
            sweeps_at[next++] = cnt;
        }
    }
    /* thresholds never reached: the iterate at max_iteration, as the reference would return it */
    for (int k = next; k < n_eps; ++k) {
        memcpy(x_snap + (size_t)k * (size_t)n, x, sizeof(double) * (size_t)n);
        sweeps_at[k] = cnt;
    }
    free(prev);
    free(x);
    return 0;
}

/* A7.  v2 :382-393 */
void orc_spmv(const orc_matrix *m, const double *in, double *out) {
    for (int i = 0; i < m->n_rows; ++i) {
        double s = 0;
        int k = m->row_begin[i];
        for (int j = 0; j < m->row_nnz[i]; ++j, ++k) s = s + m->values[k] * in[m->cols[k]];
        out[i] = s;
    }
}

/* N1.  v2 :396-434 */
int orc_cg(const orc_matrix *m, const double *b, int64_t n, double epsilon, int max_iteration, const double *x0,
           double *x, int *iters) {
    size_t bytes = sizeof(double) * (size_t)(n > 0 ? n : 1);
    double *r = (double *)malloc(bytes), *r1 = (double *)malloc(bytes), *p = (double *)malloc(bytes),
           *Ap = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    for (int64_t i = 0; i < n; ++i) x[i] = x0 ? x0[i] : 0.0;
    for (int64_t i = 0; i < n; ++i) r[i] = 0.0;
    orc_spmv(m, x, r);
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
    memcpy(p, r, bytes);
    int cnt = 0;
    while (cnt < max_iteration) {
        double rlen = orc_dot(r, r, n);
        orc_spmv(m, p, Ap);
        double alpha = rlen / orc_dot(p, Ap, n);
        for (int64_t i = 0; i < n; ++i) x[i] = x[i] + alpha * p[i];
        double nalpha = -alpha;
        for (int64_t i = 0; i < n; ++i) r1[i] = r[i] + nalpha * Ap[i];
        double r1len = orc_dot(r1, r1, n);
        if (sqrt(r1len) < epsilon) break;
        double beta = r1len / rlen;
        for (int64_t i = 0; i < n; ++i) p[i] = r1[i] + beta * p[i];
        double *t = r1;
        r1 = r;
        r = t;
        ++cnt;
    }
    if (iters) *iters = cnt;
    free(r);
    free(r1);
    free(p);
    free(Ap);
    return 0;
}

/* N1.  v2 :472-535 (extractDiagnolColInv + conjugateGradientEigen) */
int orc_pcg(const orc_matrix *m, const double *b, int64_t n, double epsilon, int max_iteration, double *x,
            int *iters) {
    size_t cnt_n = (size_t)(n > 0 ? n : 1);
    size_t bytes = sizeof(double) * cnt_n;
    double *r = (double *)calloc(cnt_n, sizeof(double)), *z = (double *)malloc(bytes), *p = (double *)malloc(bytes),
           *Ap = (double *)calloc(cnt_n, sizeof(double)), *inv = (double *)malloc(bytes);
    for (int64_t i = 0; i < n; ++i) inv[i] = 1.0;
    for (int i = 0; i < m->n_rows; ++i) {
        int k = m->row_begin[i];
        for (int j = 0; j < m->row_nnz[i]; ++j, ++k)
            if (m->cols[k] == i) {
                if (m->values[k] != 0) inv[i] = 1.0 / m->values[k];
                break;
            }
    }
    for (int64_t i = 0; i < n; ++i) x[i] = 0.0;
    orc_spmv(m, x, r);
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
    for (int64_t i = 0; i < n; ++i) p[i] = r[i] * inv[i];
    double olddist = orc_dot(p, r, n);
    int cnt = 0;
    while (cnt < max_iteration) {
        orc_spmv(m, p, Ap);
        double alpha = olddist / orc_dot(p, Ap, n);
        for (int64_t i = 0; i < n; ++i) x[i] = x[i] + alpha * p[i];
        double nalpha = -alpha;
        for (int64_t i = 0; i < n; ++i) r[i] = r[i] + nalpha * Ap[i];
        double err = orc_dot(r, r, n);
        if (sqrt(err) < epsilon) break;
        for (int64_t i = 0; i < n; ++i) z[i] = r[i] * inv[i];
        double newdist = orc_dot(z, r, n);
        double beta = newdist / olddist;
        olddist = newdist;
        for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
        ++cnt;
    }
    if (iters) *iters = cnt;
    free(r);
    free(z);
    free(p);
    free(Ap);
    free(inv);
    return 0;
}

/* A9.  PhotoMontage.cpp:551-592.  The over-determined system has, for every pixel with
 * x < W-1 and y < H-1, one row v(x+1,y)-v(x,y)=gx and one row v(x,y+1)-v(x,y)=gy, plus the
 * pin v(0,0)=constraint.  A^T*A is therefore the graph Laplacian of the grid whose edges
 * are exactly those pairs (so the last image row has no horizontal edges and the last
 * column no vertical ones), +1 on (0,0).  Pixel (W-1,H-1) has no edge: an empty row. */
static inline int pe_left(int x, int y, int W, int H) { (void)W; return x >= 1 && y < H - 1; }
static inline int pe_right(int x, int y, int W, int H) { return x < W - 1 && y < H - 1; }
static inline int pe_up(int x, int y, int W, int H) { (void)H; return y >= 1 && x < W - 1; }
static inline int pe_down(int x, int y, int W, int H) { return x < W - 1 && y < H - 1; }

int64_t orc_poisson_nnz(int W, int H) {
    int64_t n = (int64_t)W * H;
    return (n - 1) + 4 * (int64_t)(W - 1) * (H - 1);
}

void orc_poisson_csr(int W, int H, int *row_off, int *col_idx, double *values) {
    int64_t k = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int p = y * W + x;
            row_off[p] = (int)k;
            int l = pe_left(x, y, W, H), r = pe_right(x, y, W, H), u = pe_up(x, y, W, H), d = pe_down(x, y, W, H);
            int deg = l + r + u + d + (p == 0);
            if (u) { col_idx[k] = p - W; values[k++] = -1.0; }
            if (l) { col_idx[k] = p - 1; values[k++] = -1.0; }
            if (deg) { col_idx[k] = p; values[k++] = (double)deg; }
            if (r) { col_idx[k] = p + 1; values[k++] = -1.0; }
            if (d) { col_idx[k] = p + W; values[k++] = -1.0; }
        }
    row_off[(int64_t)W * H] = (int)k;
}

void orc_poisson_rhs(int W, int H, const float *gx, const float *gy, double constraint, double *b) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int64_t p = (int64_t)y * W + x;
            /* column p of A in ascending equation index: gy(x,y-1), gx(x-1,y), gx(x,y), gy(x,y), pin */
            double s = 0.0;
            if (pe_up(x, y, W, H)) s = s + (double)gy[p - W];
            if (pe_left(x, y, W, H)) s = s + (double)gx[p - 1];
            if (pe_right(x, y, W, H)) s = s - (double)gx[p];
            if (pe_down(x, y, W, H)) s = s - (double)gy[p];
            if (p == 0) s = s + constraint;
            b[p] = s;
        }
}

/* A10.  PhotoMontage.cpp:622: uchar(max(min(v,255.0),0.0)) -- truncation toward zero */
void orc_writeback_u8(const double *x, int64_t n, unsigned char *out) {
    for (int64_t i = 0; i < n; ++i) {
        double v = x[i];
        if (v > 255.0) v = 255.0;
        if (!(v > 0.0)) v = 0.0;
        out[i] = (unsigned char)v;
    }
}

/* ---- gradient-domain-fusion driver ("next" rows N2 + N4) ------------------------------------------------
 * Restated from project/src/PhotoMontage/PhotoMontage.cpp; the arithmetic is integer differences of 8-bit
 * values stored as float, so any correct restatement is exact.  Parity unpinned against a reference RUN
 * (the driver needs OpenCV + Eigen, absent here); tests/test_gdf_host.py checks these loops against an
 * independent numpy formulation instead.
 * Layouts: images n x (H x W x 3) bytes interleaved (cv::Mat CV_8UC3), labels H x W bytes,
 * gx / gy 3 planes of H x W float, x 3 planes of H x W double, out H x W x 3 bytes interleaved. */

/* GradientAt (:399-408) of Images[ResultLabel(y,x)] for y < H-1, x < W-1 (:419-425); the rest 0.
 * Returns the number of pixels whose label is >= n_images (0 = ok). */
int64_t orc_gdf_gradients(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                          float *gx, float *gy) {
    const int64_t n = (int64_t)W * H;
    int64_t bad = 0;
    for (int64_t i = 0; i < 3 * n; ++i) gx[i] = gy[i] = 0.0f;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int64_t p = (int64_t)y * W + x;
            const int l = labels[p];
            if (l >= n_images) {
                ++bad;
                continue;
            }
            if (y >= H - 1 || x >= W - 1) continue;
            const unsigned char *img = images + (int64_t)l * n * 3;
            for (int c = 0; c < 3; ++c) {
                const int color1 = img[p * 3 + c];       /* Image.at<Vec3b>(y, x)     */
                const int color2 = img[(p + 1) * 3 + c]; /* Image.at<Vec3b>(y, x + 1) */
                const int color3 = img[(p + W) * 3 + c]; /* Image.at<Vec3b>(y + 1, x) */
                gx[c * n + p] = (float)(color2 - color1);
                gy[c * n + p] = (float)(color3 - color1);
            }
        }
    return bad;
}

/* fast_init_value (:599-610): init[y*W+x] = Images[label(y,x)].at<Vec3b>(y,x)[channel] */
int64_t orc_gdf_composite(const unsigned char *images, int n_images, const unsigned char *labels, int W, int H,
                          double *x0) {
    const int64_t n = (int64_t)W * H;
    int64_t bad = 0;
    for (int64_t p = 0; p < n; ++p) {
        const int l = labels[p];
        if (l >= n_images) {
            ++bad;
            continue;
        }
        for (int c = 0; c < 3; ++c) x0[c * n + p] = (double)images[((int64_t)l * n + p) * 3 + c];
    }
    return bad;
}

/* write-back of three solved channels into the interleaved image (:617-626) */
void orc_gdf_writeback(const double *x, int64_t n, unsigned char *out) {
    for (int c = 0; c < 3; ++c)
        for (int64_t p = 0; p < n; ++p) {
            double v = x[c * n + p];
            if (v > 255.0) v = 255.0;
            if (!(v > 0.0)) v = 0.0;
            out[p * 3 + c] = (unsigned char)v;
        }
}

/* ---- lab8 panorama: producers of the right-hand side ("next" row N3) ----------------------------------------
 * Restated from labs/lab8/src/OpenCVHW1/hw8_pa.cc as the loops are written there, over flat (continuous cv::Mat)
 * buffers.  Where the reference walks a pointer past the end of a row it reads the next row's data, as a
 * continuous Mat does; where it would read or write outside the whole buffer (undefined behaviour upstream) the
 * access is skipped -- those cases never change a result inside the buffer.  Parity unpinned against a reference
 * RUN (the translation unit needs OpenCV); tests/test_pano_host.py checks these loops against the bounded,
 * per-row formulation the device kernels use.
 * Layouts: images H x W x 3 bytes, gradients H x W x 3 floats (CV_32FC3), masks H x W bytes. */

/* MaskImage (:443-466) */
void orc_pano_mask_image(const unsigned char *src, const unsigned char *mask, int W, int H, unsigned char *out) {
    for (int64_t p = 0; p < (int64_t)W * H; ++p)
        for (int c = 0; c < 3; ++c) out[p * 3 + c] = mask[p] == 0 ? 0 : src[p * 3 + c];
}

/* struct Gradients, first constructor (:604-636): GradientAt (:314-323) for y < H-1, x < W-1.  The reference
 * leaves the last row and column of the CV_32FC3 images unset; 0 here. */
void orc_pano_gradients(const unsigned char *img, int W, int H, float *gx, float *gy) {
    const int64_t n = (int64_t)W * H;
    for (int64_t i = 0; i < 3 * n; ++i) gx[i] = gy[i] = 0.0f;
    for (int y = 0; y < H - 1; ++y)
        for (int x = 0; x < W - 1; ++x) {
            const int64_t p = (int64_t)y * W + x;
            for (int c = 0; c < 3; ++c) {
                const int color1 = img[p * 3 + c], color2 = img[(p + 1) * 3 + c], color3 = img[(p + W) * 3 + c];
                gx[p * 3 + c] = (float)(color2 - color1);
                gy[p * 3 + c] = (float)(color3 - color1);
            }
        }
}

/* struct Gradients, second constructor (:638-676), literal: per row the pointer walk to the first non-zero mask byte
 * (not bounded by the row: it reads on into the following rows; `end` guards the buffer, where the reference's
 * stand-in images carry a terminating sentinel), ZeroGradientAt (:325-334) at column x-1 when x < W-2, then
 * GradientAt while x < W-1 and the byte the walk stopped at is non-zero (the loop never advances the pointer).
 * Mat::at(y, -1) addresses the previous row's last pixel on a continuous image; outside the image: dropped. */
void orc_pano_gradients_masked(const unsigned char *img, const unsigned char *mask, int W, int H, float *gx, float *gy) {
    const int64_t n = (int64_t)W * H;
    for (int64_t i = 0; i < 3 * n; ++i) gx[i] = gy[i] = 0.0f;
    for (int y = 0; y < H - 1; ++y) {
        int64_t q = (int64_t)y * W; /* ptr */
        int64_t x = 0;
        while (q < n && mask[q] == 0) {
            ++q;
            ++x;
        }
        const int stopped_on = q < n ? mask[q] : 1; /* the sentinel */
        if (x < W - 2) { /* ZeroGradientAt(m, x-1, y) */
            const int64_t p = (int64_t)y * W + (x - 1);
            if (p >= 0)
                for (int c = 0; c < 3; ++c) {
                    gx[p * 3 + c] = (float)img[(p + 1) * 3 + c];
                    gy[p * 3 + c] = (float)img[(p + W) * 3 + c];
                }
        }
        for (; x < W - 1 && stopped_on; ++x) {
            const int64_t p = (int64_t)y * W + x;
            for (int c = 0; c < 3; ++c) {
                const int color1 = img[p * 3 + c], color2 = img[(p + 1) * 3 + c], color3 = img[(p + W) * 3 + c];
                gx[p * 3 + c] = (float)(color2 - color1);
                gy[p * 3 + c] = (float)(color3 - color1);
            }
        }
    }
}

/* MergeImage2<float> (:338-385): per row, skip to the source's outer mask, skip on while the source's inner mask
 * is 0 and the target is already covered, then copy the rest of the outer-mask run.  `end` = H*W guards the
 * reads the reference makes past the buffer. */
void orc_pano_merge2_f32(float *target, const float *src, const unsigned char *target_mask,
                         const unsigned char *src_outer_mask, const unsigned char *src_inner_mask, int W, int H) {
    const int64_t end = (int64_t)W * H;
    for (int i = 0; i < H; ++i) {
        int64_t q = (int64_t)i * W; /* flat pixel index of dp / sp / tgp / sop / sip */
        int k = 0;
        while (k < W && src_outer_mask[q] == 0) {
            ++q;
            ++k;
        }
        while (q < end && src_inner_mask[q] == 0 && target_mask[q] != 0) { /* no bound on k upstream */
            ++q;
            ++k;
        }
        int c = 0;
        int64_t s = q;
        while (s < end && src_outer_mask[s] && k < W) {
            ++s;
            ++c;
            ++k;
        }
        for (int64_t j = 0; j < (int64_t)3 * c; ++j) target[q * 3 + j] = src[q * 3 + j]; /* memcpy(dp, sp, ...) */
    }
}

/* MergeImage<uchar, channel> (:387-441), channel = 3 (image) or 1 (mask; target and target_mask may alias) */
void orc_pano_merge_u8(unsigned char *target, const unsigned char *src, const unsigned char *target_mask,
                       const unsigned char *src_mask, int channel, double skip_how_many, int W, int H) {
    const int64_t end = (int64_t)W * H;
    for (int i = 0; i < H; ++i) {
        int64_t q = (int64_t)i * W;  /* dp / sp / smp */
        int64_t tq = (int64_t)i * W; /* tgp: advanced by the first loop only */
        int k = 0;
        while (k < W && src_mask[q] == 0) {
            ++q;
            ++tq;
            ++k;
        }
        int c = 0;
        if (tq < end && target_mask[tq] != 0 && skip_how_many > 0) {
            while (c < skip_how_many) {
                ++q;
                ++k;
                ++c;
            }
        }
        c = 0;
        int64_t s = q;
        while (s < end && src_mask[s] && k < W) {
            ++s;
            ++c;
            ++k;
        }
        for (int64_t j = 0; j < (int64_t)channel * c; ++j) target[q * channel + j] = src[q * channel + j];
    }
}

/* EnforceGradientBound (:468-498): where the mask is set, GradientAt(src) into rows i, i-1, i+1 of dx / dy.
 * Flat indexing as Mat::at on a continuous Mat (x + 1 == W reads the next row's first pixel).  At the first / last
 * row the reference walks out of the images (undefined upstream); pinned here to the compiled reference running over
 * images with zero guard rows (ref_pano_shim.cc): outside pixels read as 0, outside writes are dropped. */
void orc_pano_enforce_gradient_bound(float *dx, float *dy, const unsigned char *src, const unsigned char *mask, int W,
                                     int H) {
    const int64_t n = (int64_t)W * H;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            if (!mask[(int64_t)i * W + j]) continue;
            const int rows[3] = {i, i - 1, i + 1};
            for (int t = 0; t < 3; ++t) {
                const int r = rows[t];
                if (r < 0 || r > H - 1) continue;
                const int64_t p = (int64_t)r * W + j;
                for (int c = 0; c < 3; ++c) {
                    const int color1 = src[p * 3 + c];
                    const int color2 = p + 1 < n ? src[(p + 1) * 3 + c] : 0;
                    const int color3 = p + W < n ? src[(p + W) * 3 + c] : 0;
                    dx[p * 3 + c] = (float)(color2 - color1);
                    dy[p * 3 + c] = (float)(color3 - color1);
                }
            }
        }
}
