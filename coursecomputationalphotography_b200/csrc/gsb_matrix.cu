// gsb_matrix.cu -- ordering analysis of an uploaded matrix and construction of the solver format.
//
// The reference sweeps rows lexicographically (v2 :359-374), which is a serial dependency chain.
// Here rows are coloured so that no row reads an unknown of its own colour, and the system is
// permuted to colour-major order: a colour phase is then a contiguous, dependency-free row range
// and multicolour GS is *exactly* lexicographic GS on P*A*P^T (tests check that bit for bit
// against the reference run on the permuted matrix).
//   - 5-point grid probe (all off-diagonal offsets in {+-1, +-W}) -> red-black by pixel parity
//   - otherwise greedy multicolour (Jones-Plassmann order, speculative + conflict repair)
//   - or a caller-supplied colouring (verified)
// Solver format: rp[n+1], ci[nnz] (permuted columns, ascending per row), va[nnz], perm/iperm.
#include "gsb_internal.cuh"

#define MAX_COLORS 64

// Invalidates the solver format.  The device buffers stay with the handle (DevBuf::alloc reuses them when the next
// analysis needs the same sizes): a caller that re-imports a matrix of the same shape -- one Poisson system per
// image, as the reference's SolveChannel does -- pays no cudaFree / cudaMalloc of GB-sized buffers.  They are freed
// when the handle is destroyed or a differently sized matrix arrives.
void gsb_matrix::drop_analysis() {
    analyzed = false;
    n_colors = 0;
    ordering_used = 0;
    grid_width = 0;
    ws_nrhs = 0;
    if (graph_exec) {
        cudaGraphExecDestroy((cudaGraphExec_t)graph_exec);
        graph_exec = nullptr;
    }
    memset(graph_key, 0, sizeof(graph_key));
    if (plan) plan->valid = false; // (its tile tables are rebuilt, into the same allocations, by gsb_plan_build)
    group_built = false;           // the strips of a multi-device solve belong to the previous matrix
}

// ---------------------------------------------------------------------------------------------
// structure probes
// ---------------------------------------------------------------------------------------------
// info[0] = max |col-row| over live off-diagonal entries, info[1] = max col, info[2] = min col,
// info[3] = max row_nnz
__global__ void __launch_bounds__(256) probe_offsets(const int *__restrict__ cols, const int *__restrict__ row_begin,
                                                     const int *__restrict__ row_nnz, int n_rows,
                                                     int *__restrict__ info) {
    int i = blockIdx.x * 256 + threadIdx.x;
    int maxoff = 0, maxcol = -1, mincol = INT32_MAX, maxlen = 0;
    if (i < n_rows) {
        int b = row_begin[i], len = row_nnz[i];
        maxlen = len;
        for (int k = 0; k < len; ++k) {
            int c = cols[b + k];
            int d = c > i ? c - i : i - c;
            maxoff = max(maxoff, d);
            maxcol = max(maxcol, c);
            mincol = min(mincol, c);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        maxoff = max(maxoff, __shfl_down_sync(0xffffffffu, maxoff, d));
        maxcol = max(maxcol, __shfl_down_sync(0xffffffffu, maxcol, d));
        mincol = min(mincol, __shfl_down_sync(0xffffffffu, mincol, d));
        maxlen = max(maxlen, __shfl_down_sync(0xffffffffu, maxlen, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&info[0], maxoff);
        atomicMax(&info[1], maxcol);
        atomicMin(&info[2], mincol);
        atomicMax(&info[3], maxlen);
    }
}

// bad[0] |= 1 if some off-diagonal offset is outside {+-1, +-W}
__global__ void __launch_bounds__(256) probe_grid(const int *__restrict__ cols, const int *__restrict__ row_begin,
                                                  const int *__restrict__ row_nnz, int n_rows, int W,
                                                  int *__restrict__ bad) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int b = row_begin[i], len = row_nnz[i];
    bool ok = true;
    for (int k = 0; k < len; ++k) {
        int d = cols[b + k] - i;
        if (d < 0) d = -d;
        ok = ok && (d == 0 || d == 1 || d == W);
    }
    if (!ok) atomicOr(bad, 1);
}

__global__ void __launch_bounds__(256) color_parity(int n_rows, int W, int *__restrict__ colors) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_rows) colors[i] = ((i % W) + (i / W)) & 1;
}

// A colouring is usable iff no row reads an unknown of its own colour (the directed condition is
// exactly what makes a colour phase race-free).  bad[0] |= 1 on violation, |= 2 on range error.
__global__ void __launch_bounds__(256) check_coloring(const int *__restrict__ cols, const int *__restrict__ row_begin,
                                                      const int *__restrict__ row_nnz, int n_rows,
                                                      const int *__restrict__ colors, int n_colors,
                                                      int *__restrict__ bad) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int ci = colors[i];
    if (ci < 0 || ci >= n_colors) {
        atomicOr(bad, 2);
        return;
    }
    int b = row_begin[i], len = row_nnz[i];
    bool ok = true;
    for (int k = 0; k < len; ++k) {
        int c = cols[b + k];
        if (c != i && colors[c] == ci) ok = false;
    }
    if (!ok) atomicOr(bad, 1);
}

// ---------------------------------------------------------------------------------------------
// greedy multicolour: Jones-Plassmann rounds on hashed priorities, first-fit colour choice, on the SYMMETRISED
// pattern: rows i and j conflict when a_ij OR a_ji is stored (either makes one of them read the other inside a
// colour phase).  A row's own entries are its out-neighbours; the rows that read it (in-neighbours) come from a
// transposed adjacency built once per analysis (histogram -> scan -> scatter).  With both directions visible two
// adjacent uncoloured rows never colour themselves in the same round -- the lower priority waits -- so the result
// is a proper colouring by construction and the rounds terminate in O(log n).
//   (Round 1 coloured speculatively on the rows' own entries and repaired clashes afterwards.  On structurally
//    unsymmetric patterns the repairs chase each other: a re-coloured row clashes with a row that reads it, which
//    re-colours, ...  -- a livelock that ran into the round cap on a 400 000-row matrix, found in round 2.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned prio_hash(unsigned x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

__device__ __forceinline__ bool prio_greater(int a, int b) { // does a outrank b?
    unsigned ha = prio_hash((unsigned)a), hb = prio_hash((unsigned)b);
    return ha > hb || (ha == hb && a > b);
}

// transposed adjacency: in_cnt[j] = number of rows i != j that store column j
__global__ void __launch_bounds__(256) jp_count_in(const int *__restrict__ cols, const int *__restrict__ row_begin,
                                                   const int *__restrict__ row_nnz, int n_rows, int *__restrict__ in_cnt) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int b = row_begin[i], len = row_nnz[i];
    for (int k = 0; k < len; ++k) {
        int j = cols[b + k];
        if (j != i) atomicAdd(&in_cnt[j], 1);
    }
}
// (the order inside a row's in-list depends on the atomics; the colouring only uses the list as a set)
__global__ void __launch_bounds__(256) jp_fill_in(const int *__restrict__ cols, const int *__restrict__ row_begin,
                                                  const int *__restrict__ row_nnz, int n_rows,
                                                  const int *__restrict__ in_ptr, int *__restrict__ cursor,
                                                  int *__restrict__ in_idx) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int b = row_begin[i], len = row_nnz[i];
    for (int k = 0; k < len; ++k) {
        int j = cols[b + k];
        if (j != i) in_idx[in_ptr[j] + atomicAdd(&cursor[j], 1)] = i;
    }
}

// counters[0] = rows still uncoloured after this round, counters[1] |= 1 if > MAX_COLORS needed
__global__ void __launch_bounds__(256) jp_round(const int *__restrict__ cols, const int *__restrict__ row_begin,
                                                const int *__restrict__ row_nnz, const int *__restrict__ in_ptr,
                                                const int *__restrict__ in_idx, int n_rows,
                                                const int *__restrict__ cin, int *__restrict__ cout,
                                                int *__restrict__ counters) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int cur = cin[i];
    if (cur >= 0) {
        cout[i] = cur;
        return;
    }
    unsigned long long forbidden = 0ull;
    bool is_max = true;
    auto visit = [&](int j) {
        if (j == i) return;
        int cj = cin[j];
        if (cj < 0) {
            if (prio_greater(j, i)) is_max = false;
        } else {
            forbidden |= 1ull << cj;
        }
    };
    const int b = row_begin[i], len = row_nnz[i];
    for (int k = 0; k < len && is_max; ++k) visit(cols[b + k]);
    for (int k = in_ptr[i]; k < in_ptr[i + 1] && is_max; ++k) visit(in_idx[k]);
    if (!is_max) {
        cout[i] = -1;
        atomicAdd(&counters[0], 1);
        return;
    }
    if (forbidden == ~0ull) {
        atomicOr(&counters[1], 1);
        cout[i] = -1;
        return;
    }
    cout[i] = __ffsll((long long)~forbidden) - 1;
}

// ---------------------------------------------------------------------------------------------
// permutation + solver format
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) color_flag(const int *__restrict__ colors, int n_rows, int c,
                                                  int *__restrict__ flag) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_rows) flag[i] = colors[i] == c ? 1 : 0;
}

__global__ void __launch_bounds__(256) color_place(const int *__restrict__ colors, const int *__restrict__ rank,
                                                   int n_rows, int c, int base, int *__restrict__ perm,
                                                   int *__restrict__ iperm) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_rows && colors[i] == c) {
        int p = base + rank[i];
        perm[p] = i;
        iperm[i] = p;
    }
}

// off-diagonal length of every row in permuted order (the diagonal goes to its own array)
__global__ void __launch_bounds__(256) perm_row_len(const int *__restrict__ perm, const int *__restrict__ row_begin,
                                                    const int *__restrict__ row_nnz, const int *__restrict__ cols,
                                                    int n_rows, int *__restrict__ len) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p < n_rows) {
        const int old = perm[p];
        const int src = row_begin[old], cnt = row_nnz[old];
        int off = 0;
        for (int k = 0; k < cnt; ++k) off += cols[src + k] != old;
        len[p] = off;
    }
    if (p == n_rows) len[p] = 0;
}

__global__ void __launch_bounds__(128) perm_fill_rows(const int *__restrict__ perm, const int *__restrict__ iperm,
                                                      const int *__restrict__ row_begin,
                                                      const int *__restrict__ row_nnz,
                                                      const int *__restrict__ cols, const double *__restrict__ vals,
                                                      int n_rows, const int *__restrict__ rp, int *__restrict__ ci,
                                                      double *__restrict__ va, double *__restrict__ dg) {
    int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_rows) return;
    int old = perm[p];
    int src = row_begin[old], len = row_nnz[old], dst = rp[p];
    // insertion sort by permuted column while copying (rows are short and nearly sorted); the diagonal entry
    // (what at(i,i) finds, v2 :360) is split off into dg
    double d = 0.0;
    int w = 0;
    for (int k = 0; k < len; ++k) {
        int co = cols[src + k];
        double v = vals[src + k];
        if (co == old) {
            d = v;
            continue;
        }
        int c = iperm[co];
        int q = dst + w;
        while (q > dst && ci[q - 1] > c) {
            ci[q] = ci[q - 1];
            va[q] = va[q - 1];
            --q;
        }
        ci[q] = c;
        va[q] = v;
        ++w;
    }
    dg[p] = d;
}

static int build_solver_format(gsb_matrix *m, cudaStream_t st) {
    const int n = m->n_rows;
    const int nb = (n + 255) / 256;
    GSB_TRY(m->perm.alloc(n));
    GSB_TRY(m->iperm.alloc(n));
    DevBuf<int> &flag = m->scratch_rows;
    GSB_TRY(flag.alloc((int64_t)n + 1));
    GSB_TRY(gsb_tiny_alloc(m));
    GSB_TRY(m->scan_scratch.alloc(gsb_scan_scratch_ints((int64_t)n + 1)));
    struct { int *p; } tot = {m->tiny.p + GSB_TINY_COLOR_TOT};
    int base = 0;
    for (int c = 0; c < m->n_colors; ++c) {
        color_flag<<<nb, 256, 0, st>>>(m->colors.p, n, c, flag.p);
        GSB_KERNEL_CHECK();
        GSB_TRY(gsb_exclusive_scan_i32(flag.p, flag.p, n, tot.p, st, m->scan_scratch.p, m->scan_scratch.n));
        int h = 0;
        GSB_CUDA(cudaMemcpyAsync(&h, tot.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        color_place<<<nb, 256, 0, st>>>(m->colors.p, flag.p, n, c, base, m->perm.p, m->iperm.p);
        GSB_KERNEL_CHECK();
        m->color_start[c] = base;
        base += h;
    }
    m->color_start[m->n_colors] = base;
    if (base != n) {
        gsb_set_error("internal: colour classes cover %d of %d rows", base, n);
        return GSB_ERR_COLORING;
    }
    GSB_TRY(m->rp.alloc((int64_t)n + 1 + 8)); // +8: aligned bulk copies may over-read
    perm_row_len<<<(n + 1 + 255) / 256, 256, 0, st>>>(m->perm.p, m->row_begin.p, m->row_nnz.p, m->cols.p, n, m->rp.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(gsb_exclusive_scan_i32(m->rp.p, m->rp.p, (int64_t)n + 1, nullptr, st, m->scan_scratch.p, m->scan_scratch.n));
    int nnz_off = 0;
    GSB_CUDA(cudaMemcpyAsync(&nnz_off, m->rp.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    m->nnz_off = nnz_off;
    GSB_TRY(m->ci.alloc(m->nnz_off + 8)); // +8: the staged kernels' 16-byte aligned bulk copies may over-read
    GSB_TRY(m->va.alloc(m->nnz_off + 8));
    GSB_TRY(m->dg.alloc((int64_t)n + 8));
    perm_fill_rows<<<(n + 127) / 128, 128, 0, st>>>(m->perm.p, m->iperm.p, m->row_begin.p, m->row_nnz.p, m->cols.p,
                                                   m->vals(), n, m->rp.p, m->ci.p, m->va.p, m->dg.p);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

static int run_check(gsb_matrix *m, int *bad_dev, cudaStream_t st, int *h_bad) {
    GSB_CUDA(cudaMemsetAsync(bad_dev, 0, sizeof(int), st));
    check_coloring<<<(m->n_rows + 255) / 256, 256, 0, st>>>(m->cols.p, m->row_begin.p, m->row_nnz.p, m->n_rows,
                                                          m->colors.p, m->n_colors, bad_dev);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaMemcpyAsync(h_bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

static int try_redblack(gsb_matrix *m, int maxoff, int *bad_dev, cudaStream_t st, bool *ok) {
    *ok = false;
    const int n = m->n_rows;
    int W = maxoff; // the largest offset is the image width (1 for a 1-D chain, 0 for a diagonal matrix)
    if (W <= 0) W = 1;
    int h_bad = 0;
    GSB_CUDA(cudaMemsetAsync(bad_dev, 0, sizeof(int), st));
    probe_grid<<<(n + 255) / 256, 256, 0, st>>>(m->cols.p, m->row_begin.p, m->row_nnz.p, n, W, bad_dev);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaMemcpyAsync(&h_bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h_bad) return GSB_OK;
    color_parity<<<(n + 255) / 256, 256, 0, st>>>(n, W, m->colors.p);
    GSB_KERNEL_CHECK();
    m->n_colors = n > 1 ? 2 : 1;
    GSB_TRY(run_check(m, bad_dev, st, &h_bad));
    if (h_bad) return GSB_OK; // e.g. wrap-around edges between image rows
    m->grid_width = W;
    *ok = true;
    return GSB_OK;
}

static int run_multicolor(gsb_matrix *m, cudaStream_t st) {
    const int n = m->n_rows;
    const int nb = (n + 255) / 256;
    DevBuf<int> other, counters, in_ptr, in_idx, cursor;
    GSB_TRY(other.alloc(n));
    GSB_TRY(counters.alloc(4));
    // transposed adjacency of the stored pattern
    GSB_TRY(in_ptr.alloc((int64_t)n + 1));
    GSB_TRY(cursor.alloc(n));
    GSB_CUDA(cudaMemsetAsync(in_ptr.p, 0, sizeof(int) * (size_t)(n + 1), st));
    GSB_CUDA(cudaMemsetAsync(cursor.p, 0, sizeof(int) * (size_t)n, st));
    jp_count_in<<<nb, 256, 0, st>>>(m->cols.p, m->row_begin.p, m->row_nnz.p, n, in_ptr.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(gsb_exclusive_scan_i32(in_ptr.p, in_ptr.p, (int64_t)n + 1, nullptr, st));
    int n_in = 0;
    GSB_CUDA(cudaMemcpyAsync(&n_in, in_ptr.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    GSB_TRY(in_idx.alloc(n_in));
    jp_fill_in<<<nb, 256, 0, st>>>(m->cols.p, m->row_begin.p, m->row_nnz.p, n, in_ptr.p, cursor.p, in_idx.p);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaMemsetAsync(m->colors.p, 0xff, sizeof(int) * (size_t)n, st)); // all -1
    int *cin = m->colors.p, *cout = other.p;
    int h[4] = {1, 0, 0, 0};
    int round = 0;
    for (; round < 4096 && h[0] != 0; ++round) {
        GSB_CUDA(cudaMemsetAsync(counters.p, 0, 4 * sizeof(int), st));
        jp_round<<<nb, 256, 0, st>>>(m->cols.p, m->row_begin.p, m->row_nnz.p, in_ptr.p, in_idx.p, n, cin, cout, counters.p);
        GSB_KERNEL_CHECK();
        GSB_CUDA(cudaMemcpyAsync(h, counters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        if (h[1]) {
            gsb_set_error("multicolour ordering needs more than %d colours (row degree too high)", MAX_COLORS);
            return GSB_ERR_COLORING;
        }
        int *t = cin;
        cin = cout;
        cout = t;
    }
    if (h[0] != 0) { // cannot happen: every round colours at least the highest-priority uncoloured row
        gsb_set_error("internal: multicolour ordering did not finish in %d rounds (%d rows left)", round, h[0]);
        return GSB_ERR_COLORING;
    }
    if (cin != m->colors.p)
        GSB_CUDA(cudaMemcpyAsync(m->colors.p, cin, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    int *mx = counters.p;
    GSB_TRY(gsb_reduce_max_i32(m->colors.p, n, mx, st));
    int hmax = 0;
    GSB_CUDA(cudaMemcpyAsync(&hmax, mx, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    m->n_colors = hmax + 1;
    return GSB_OK;
}

extern "C" int gsb_matrix_analyze(gsb_matrix *m, int ordering, const int *user_colors) {
    if (!m) return GSB_ERR_ARG;
    if (!m->has_layout) {
        gsb_set_error("matrix_analyze: matrix holds no layout yet");
        return GSB_ERR_STATE;
    }
    if (m->n_rows != m->n_cols) {
        gsb_set_error("matrix_analyze: Gauss-Seidel needs a square matrix (have %d x %d)", m->n_rows, m->n_cols);
        return GSB_ERR_SHAPE;
    }
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    m->drop_analysis();
    const int n = m->n_rows;
    GSB_TRY(gsb_tiny_alloc(m));
    struct { int *p; } info = {m->tiny.p + GSB_TINY_INFO};
    int init[4] = {0, -1, INT32_MAX, 0};
    GSB_CUDA(cudaMemcpyAsync(info.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    probe_offsets<<<(n + 255) / 256, 256, 0, st>>>(m->cols.p, m->row_begin.p, m->row_nnz.p, n, info.p);
    GSB_KERNEL_CHECK();
    int h[4];
    GSB_CUDA(cudaMemcpyAsync(h, info.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (m->nnz > 0 && (h[1] >= n || h[2] < 0)) {
        gsb_set_error("matrix_analyze: column index out of range [0,%d): min %d max %d", n, h[2], h[1]);
        return GSB_ERR_SHAPE;
    }
    m->max_row_nnz = h[3];
    GSB_TRY(m->colors.alloc(n));
    int *bad = info.p + 4;

    if (ordering == GSB_ORDER_USER) {
        if (!user_colors) {
            gsb_set_error("matrix_analyze: GSB_ORDER_USER needs a colour array");
            return GSB_ERR_ARG;
        }
        GSB_CUDA(cudaMemcpyAsync(m->colors.p, user_colors, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
        int *mx = info.p + 5;
        GSB_TRY(gsb_reduce_max_i32(m->colors.p, n, mx, st));
        int hmax = 0;
        GSB_CUDA(cudaMemcpyAsync(&hmax, mx, sizeof(int), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        if (hmax < 0 || hmax >= MAX_COLORS) {
            gsb_set_error("matrix_analyze: user colours must lie in [0,%d)", MAX_COLORS);
            return GSB_ERR_COLORING;
        }
        m->n_colors = hmax + 1;
        int h_bad = 0;
        GSB_TRY(run_check(m, bad, st, &h_bad));
        if (h_bad) {
            gsb_set_error("matrix_analyze: user colouring is not proper (a row reads an unknown of its own colour)");
            return GSB_ERR_COLORING;
        }
        m->ordering_used = GSB_ORDER_USER;
    } else {
        bool done = false;
        if (ordering == GSB_ORDER_AUTO || ordering == GSB_ORDER_REDBLACK) {
            GSB_TRY(try_redblack(m, h[0], bad, st, &done));
            if (done) m->ordering_used = GSB_ORDER_REDBLACK;
            if (!done && ordering == GSB_ORDER_REDBLACK) {
                gsb_set_error("matrix_analyze: matrix is not a 5-point grid operator (red-black by parity is improper)");
                return GSB_ERR_COLORING;
            }
        }
        if (!done) {
            if (ordering != GSB_ORDER_AUTO && ordering != GSB_ORDER_MULTICOLOR) {
                gsb_set_error("matrix_analyze: unknown ordering %d", ordering);
                return GSB_ERR_ARG;
            }
            GSB_TRY(run_multicolor(m, st));
            int h_bad = 0;
            GSB_TRY(run_check(m, bad, st, &h_bad));
            if (h_bad) {
                gsb_set_error("internal: multicolour ordering failed verification (%d)", h_bad);
                return GSB_ERR_COLORING;
            }
            m->ordering_used = GSB_ORDER_MULTICOLOR;
        }
    }
    GSB_TRY(build_solver_format(m, st));
    m->analyzed = true;
    return GSB_OK;
}

extern "C" int gsb_matrix_coloring(const gsb_matrix *m, int *n_colors, int *ordering_used, int *grid_width) {
    if (!m) return GSB_ERR_ARG;
    if (!m->analyzed) {
        gsb_set_error("matrix_coloring: matrix has not been analysed");
        return GSB_ERR_STATE;
    }
    if (n_colors) *n_colors = m->n_colors;
    if (ordering_used) *ordering_used = m->ordering_used;
    if (grid_width) *grid_width = m->grid_width;
    return GSB_OK;
}

extern "C" int gsb_matrix_ordering(const gsb_matrix *m, int *perm, int *colors) {
    if (!m) return GSB_ERR_ARG;
    if (!m->analyzed) {
        gsb_set_error("matrix_ordering: matrix has not been analysed");
        return GSB_ERR_STATE;
    }
    GSB_TRY(gsb_set_device(m->device));
    cudaStream_t st = gsb_cur_stream();
    size_t bytes = sizeof(int) * (size_t)m->n_rows;
    if (perm) GSB_CUDA(cudaMemcpyAsync(perm, m->perm.p, bytes, cudaMemcpyDeviceToHost, st));
    if (colors) GSB_CUDA(cudaMemcpyAsync(colors, m->colors.p, bytes, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}
