// gsb_cg.cu -- row N1: the conjugate-gradient solvers the reference's Poisson drivers actually call
// (hw8_pa.cc:972, PhotoMontage.cpp:613), as a DEVICE loop.
//   conjugateGradient(b, eps, max_iter, initialize)   v2 :396-434
//   conjugateGradientEigen(b, eps, max_iter)          v2 :472-535  (Jacobi-preconditioned)
// Same recurrences, same unfused a + s*b vector updates, same SpMV (storage order, bit-exact with applyToVector).
//
// One iteration = three launches and no host round trip:
//   cg_spmv_dot   Ap = A p for up to 4 right-hand sides in ONE pass over the CSR (the colour channels share the
//                 matrix, PhotoMontage.cpp:428-433), fused with the partial sums of p.Ap; the last CTA to retire folds
//                 the partials in a fixed order and leaves alpha = (r.r) / (p.Ap) in device memory
//   cg_update_xr  x += alpha p, r -= alpha Ap (and z = M^-1 r for the Jacobi variant), fused with the partial sums of
//                 r.r (and z.r); the last CTA folds, takes the stop decision sqrt(r.r) < eps (v2 :425 / :524: the
//                 loop breaks BEFORE p is updated and before the counter is bumped) and leaves beta
//   cg_update_p   p = z + beta p
// alpha, beta, the norms, the per-RHS iteration counters and `done` flags live in a CgState block on the device; a
// right-hand side that has stopped is frozen while the others go on.  The host enqueues a batch of iterations and
// reads the block once per batch (kernels launched after the last RHS stopped return at their first instruction).
// Dot products are fixed-order tree reductions (the reference's transform_reduce leaves the order unspecified), so
// iterates agree with the reference to rounding, not bit for bit; run to run they are identical.
// Workspaces stay with the matrix handle: repeated solves (three channels, every frame) allocate nothing.
#include "gsb_ring.cuh"

#include <math.h>
#include <stdlib.h>

#define CG_THREADS 256
#define CG_UNROLL 6      // rows up to this many entries take the gather-prefetch path
#define CG_SPAN_CAP 2048 // CSR entries of a 256-row tile staged in shared memory (24 KB): rows up to 8 entries on average

struct CgState {
    double rho[GSB_MAX_RHS];   // r.r (CG) / z.r (Jacobi) of the current residual: numerator of alpha, denominator of beta
    double alpha[GSB_MAX_RHS], beta[GSB_MAX_RHS];
    double err[GSB_MAX_RHS];   // r.r after the last update
    int done[GSB_MAX_RHS];     // 1 = stopped (converged, or max_iteration reached)
    int cnt[GSB_MAX_RHS];      // the reference's `cnt` at exit
    int all_done;
    int max_iter;
    double epsilon;
    unsigned ticket[2];        // "last CTA folds" counters of the two reducing kernels
};

// fixed-order fold of `np` per-CTA partials (NV values each) by the calling CTA -> sum[0..NV) in shared memory
template <int NV>
__device__ __forceinline__ void cg_fold(const double *partials, int np, double (&sum)[NV], double (*ws)[CG_THREADS / 32]) {
    double s[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) s[v] = 0.0;
    for (int i = threadIdx.x; i < np; i += CG_THREADS) {
#pragma unroll
        for (int v = 0; v < NV; ++v) s[v] += __ldcg(partials + (size_t)i * NV + v);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double t = s[v];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
        if (lane == 0) ws[v][wid] = t;
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double t = 0.0;
        for (int w = 0; w < CG_THREADS / 32; ++w) t += ws[v][w];
        sum[v] = t;
    }
}

// per-CTA partial (NV values per thread) -> partials[blockIdx.x]; returns true in the last CTA to retire
template <int NV>
__device__ __forceinline__ bool cg_block_partial(double (&v)[NV], double *partials, unsigned *ticket,
                                                 double (*ws)[CG_THREADS / 32]) {
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double t = v[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
        if (lane == 0) ws[q][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
        for (int w = 0; w < CG_THREADS / 32; ++w) t += ws[threadIdx.x][w];
        partials[(size_t)blockIdx.x * NV + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        __threadfence();
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) *ticket = 0;
    return s_last != 0;
}

// out_r = A * in_r (storage order, product and sum rounded separately: bit-exact with applyToVector, v2 :382-393) for
// NRHS vectors in one pass over the CSR.  A tile of 256 rows is one contiguous span of the slack CSR; it is staged
// through shared memory with coalesced loads when it fits, else read row per thread.  DOT: also sum in_r . out_r.
template <int NRHS, bool DOT>
__global__ void __launch_bounds__(CG_THREADS) cg_spmv_dot(const double *__restrict__ vals, const int *__restrict__ cols,
                                                          const int *__restrict__ row_begin,
                                                          const int *__restrict__ row_nnz, int n_rows, int64_t store,
                                                          const double *__restrict__ in, double *__restrict__ out,
                                                          int64_t ld, CgState *st, double *__restrict__ partials) {
    __shared__ double v_s[CG_SPAN_CAP];
    __shared__ int c_s[CG_SPAN_CAP];
    __shared__ double ws[NRHS][CG_THREADS / 32];
    if (DOT && *(volatile int *)&st->all_done) return;
    bool live[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) live[r] = !DOT || !st->done[r];
    double dot[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) dot[r] = 0.0;
    const int ntiles = (n_rows + CG_THREADS - 1) / CG_THREADS;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int r0 = t * CG_THREADS, r1 = min(r0 + CG_THREADS, n_rows);
        const int64_t k0 = row_begin[r0], k1 = r1 < n_rows ? (int64_t)row_begin[r1] : store;
        const bool staged = k1 - k0 <= CG_SPAN_CAP;
        __syncthreads(); // the previous tile's rows are done with the stage
        if (staged) {
            for (int k = threadIdx.x; k < (int)(k1 - k0); k += CG_THREADS) {
                v_s[k] = vals[k0 + k];
                c_s[k] = cols[k0 + k];
            }
        }
        __syncthreads();
        const int i = r0 + threadIdx.x;
        if (i < r1) {
            const int64_t kb = row_begin[i];
            const int len = row_nnz[i];
            double s[NRHS];
#pragma unroll
            for (int r = 0; r < NRHS; ++r) s[r] = 0.0;
            if (len <= CG_UNROLL) {
                // short rows (a 5-point row has 5 entries): fetch the row, issue ALL gathers, then accumulate in storage
                // order -- the gathers' latencies overlap instead of adding up; padded slots load column 0 and are
                // not accumulated
                int cc[CG_UNROLL];
                double vv[CG_UNROLL], xg[CG_UNROLL][NRHS];
#pragma unroll
                for (int j = 0; j < CG_UNROLL; ++j) {
                    const bool in_row = j < len;
                    cc[j] = in_row ? (staged ? c_s[kb - k0 + j] : cols[kb + j]) : 0;
                    vv[j] = in_row ? (staged ? v_s[kb - k0 + j] : vals[kb + j]) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < CG_UNROLL; ++j)
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) xg[j][r] = live[r] ? in[r * ld + cc[j]] : 0.0;
#pragma unroll
                for (int j = 0; j < CG_UNROLL; ++j)
                    if (j < len) {
#pragma unroll
                        for (int r = 0; r < NRHS; ++r) s[r] = __dadd_rn(s[r], __dmul_rn(vv[j], xg[j][r]));
                    }
            } else {
                for (int j = 0; j < len; ++j) {
                    const double v = staged ? v_s[kb - k0 + j] : vals[kb + j];
                    const int c = staged ? c_s[kb - k0 + j] : cols[kb + j];
#pragma unroll
                    for (int r = 0; r < NRHS; ++r)
                        if (live[r]) s[r] = __dadd_rn(s[r], __dmul_rn(v, in[r * ld + c]));
                }
            }
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                if (live[r]) {
                    out[r * ld + i] = s[r];
                    if (DOT) dot[r] += in[r * ld + i] * s[r];
                }
        }
    }
    if (DOT) {
        if (cg_block_partial<NRHS>(dot, partials, &st->ticket[0], ws)) {
            double sum[NRHS];
            cg_fold<NRHS>(partials, gridDim.x, sum, ws);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
                    if (!st->done[r]) st->alpha[r] = st->rho[r] / sum[r]; // v2 :420-421 / :517-518
            }
        }
    }
}

// The same product as a PERSISTENT kernel with a two-stage ring of bulk copies (the structure of the colour-phase
// kernels, gsb_phase.cu): cg_spmv_dot loads a tile's span, waits, computes, waits -- every tile pays two exposed
// round trips to HBM and the kernel reaches half of the copy bandwidth at 4096^2 (ncu launch list, round 2 call 20:
// 605 us for 1.95 GB).  Here thread 0 issues tile j+2's four spans (values, columns, row_begin, row_nnz: each one
// contiguous piece of the slack CSR, 16-byte aligned by rounding its ends outwards) as cp.async.bulk copies on the
// stage's mbarrier while the CTA computes tile j; the tile boundaries of the tile after that are fetched into registers an
// iteration ahead, so that the issue never waits for a load.  Tiles that do not fit the stage (long rows) and the
// last tile (its aligned ends could leave the arrays) are read row per thread as before.  Same arithmetic, same
// storage order: bit-exact with applyToVector and with cg_spmv_dot.
struct CgStage {
    int v_off, c_off, rb_off, rn_off, hdr_off, bytes;
};
__host__ __device__ inline CgStage cg_stage_layout() {
    CgStage L;
    L.v_off = 0;
    L.c_off = L.v_off + (CG_SPAN_CAP + 2) * 8;
    L.rb_off = L.c_off + (CG_SPAN_CAP + 8) * 4;
    L.rn_off = L.rb_off + CG_THREADS * 4;
    L.hdr_off = L.rn_off + CG_THREADS * 4;
    L.bytes = L.hdr_off + 16;
    return L;
}

template <int NRHS, bool DOT>
__global__ void __launch_bounds__(CG_THREADS, 4) cg_spmv_ring(const double *__restrict__ vals, const int *__restrict__ cols,
                                                              const int *__restrict__ row_begin,
                                                              const int *__restrict__ row_nnz, int n_rows, int64_t store,
                                                              const double *__restrict__ in, double *__restrict__ out,
                                                              int64_t ld, CgState *st, double *__restrict__ partials) {
    extern __shared__ __align__(128) unsigned char cg_smem[];
    __shared__ double ws[NRHS][CG_THREADS / 32];
    const CgStage L = cg_stage_layout();
    uint64_t *full = reinterpret_cast<uint64_t *>(cg_smem);
    unsigned char *stage0 = cg_smem + 64;
    const int tid = threadIdx.x, bid = blockIdx.x, gsz = gridDim.x;
    if (DOT && *(volatile int *)&st->all_done) return;
    bool live[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) live[r] = !DOT || !st->done[r];
    double dot[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) dot[r] = 0.0;
    const int ntiles = (n_rows + CG_THREADS - 1) / CG_THREADS;
    const int my_tiles = bid < ntiles ? (ntiles - bid + gsz - 1) / gsz : 0;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
    }
    __syncthreads();
    // thread 0: CSR offsets at the two ends of the tile it will issue next (fetched an iteration ahead)
    int64_t nk0 = 0, nk1 = 0;
    auto fetch_bounds = [&](int j) {
        if (j >= my_tiles) return;
        const int r0 = (bid + j * gsz) * CG_THREADS, r1 = r0 + CG_THREADS;
        nk0 = row_begin[r0];
        nk1 = r1 < n_rows ? (int64_t)row_begin[r1] : store;
    };
    auto issue = [&](int j, int s) { // thread 0
        const int t = bid + j * gsz, r0 = t * CG_THREADS;
        unsigned char *sp = stage0 + (size_t)s * L.bytes;
        int *hdr = reinterpret_cast<int *>(sp + L.hdr_off);
        const int64_t k0 = nk0, k1 = nk1;
        const int64_t kv0 = k0 & ~(int64_t)1, kv1 = (k1 + 1) & ~(int64_t)1;
        const int64_t kc0 = k0 & ~(int64_t)3, kc1 = (k1 + 3) & ~(int64_t)3;
        const bool staged = t != ntiles - 1 && k1 - k0 <= CG_SPAN_CAP && kc1 <= store && k1 >= k0;
        hdr[0] = staged ? 1 : 0;
        if (staged) {
            const uint32_t bytes_v = (uint32_t)(kv1 - kv0) * 8u, bytes_c = (uint32_t)(kc1 - kc0) * 4u;
            const uint32_t bytes_r = CG_THREADS * 4u;
            mbar_expect_tx(&full[s], bytes_v + bytes_c + 2u * bytes_r);
            if (bytes_v) bulk_g2s(sp + L.v_off, vals + kv0, bytes_v, &full[s]);
            if (bytes_c) bulk_g2s(sp + L.c_off, cols + kc0, bytes_c, &full[s]);
            bulk_g2s(sp + L.rb_off, row_begin + r0, bytes_r, &full[s]);
            bulk_g2s(sp + L.rn_off, row_nnz + r0, bytes_r, &full[s]);
        } else {
            mbar_expect_tx(&full[s], 0u); // nothing to wait for: the rows read global memory
        }
    };
    if (tid == 0) {
        fetch_bounds(0);
        if (my_tiles > 0) issue(0, 0);
        fetch_bounds(1);
        if (my_tiles > 1) issue(1, 1);
        fetch_bounds(2);
    }
    for (int j = 0; j < my_tiles; ++j) {
        const int s = j & 1;
        const unsigned char *sp = stage0 + (size_t)s * L.bytes;
        mbar_wait(&full[s], (uint32_t)(j >> 1) & 1u);
        const bool staged = reinterpret_cast<const int *>(sp + L.hdr_off)[0] != 0;
        const int t = bid + j * gsz, r0 = t * CG_THREADS, r1 = min(r0 + CG_THREADS, n_rows);
        const int i = r0 + tid;
        if (i < r1) {
            const int64_t kb = staged ? (int64_t)reinterpret_cast<const int *>(sp + L.rb_off)[tid] : (int64_t)row_begin[i];
            const int len = staged ? reinterpret_cast<const int *>(sp + L.rn_off)[tid] : row_nnz[i];
            // the stage holds vals[kv0 ..) and cols[kc0 ..): index them by the global offset k
            const int64_t k0 = staged ? (int64_t)reinterpret_cast<const int *>(sp + L.rb_off)[0] : 0;
            const double *v_s = reinterpret_cast<const double *>(sp + L.v_off) - (k0 & ~(int64_t)1);
            const int *c_s = reinterpret_cast<const int *>(sp + L.c_off) - (k0 & ~(int64_t)3);
            double sum[NRHS];
#pragma unroll
            for (int r = 0; r < NRHS; ++r) sum[r] = 0.0;
            if (len <= CG_UNROLL) {
                int cc[CG_UNROLL];
                double vv[CG_UNROLL], xg[CG_UNROLL][NRHS];
#pragma unroll
                for (int q = 0; q < CG_UNROLL; ++q) {
                    const bool in_row = q < len;
                    cc[q] = in_row ? (staged ? c_s[kb + q] : cols[kb + q]) : 0;
                    vv[q] = in_row ? (staged ? v_s[kb + q] : vals[kb + q]) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < CG_UNROLL; ++q)
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) xg[q][r] = live[r] ? in[r * ld + cc[q]] : 0.0;
#pragma unroll
                for (int q = 0; q < CG_UNROLL; ++q)
                    if (q < len) {
#pragma unroll
                        for (int r = 0; r < NRHS; ++r) sum[r] = __dadd_rn(sum[r], __dmul_rn(vv[q], xg[q][r]));
                    }
            } else {
                for (int q = 0; q < len; ++q) {
                    const double v = staged ? v_s[kb + q] : vals[kb + q];
                    const int c = staged ? c_s[kb + q] : cols[kb + q];
#pragma unroll
                    for (int r = 0; r < NRHS; ++r)
                        if (live[r]) sum[r] = __dadd_rn(sum[r], __dmul_rn(v, in[r * ld + c]));
                }
            }
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                if (live[r]) {
                    out[r * ld + i] = sum[r];
                    if (DOT) dot[r] += in[r * ld + i] * sum[r];
                }
        }
        __syncthreads(); // every row of the tile has read its stage
        if (tid == 0) {
            if (j + 2 < my_tiles) issue(j + 2, s);
            fetch_bounds(j + 3);
        }
    }
    if (DOT) {
        if (cg_block_partial<NRHS>(dot, partials, &st->ticket[0], ws)) {
            double sum[NRHS];
            cg_fold<NRHS>(partials, gridDim.x, sum, ws);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int r = 0; r < NRHS; ++r)
                    if (!st->done[r]) st->alpha[r] = st->rho[r] / sum[r]; // v2 :420-421 / :517-518
            }
        }
    }
}

// start-up: r = b - A x0 (Ax0 in `ax`, or absent: r = b), z = M^-1 r (inv != null), p = z, rho = z.r
template <int NRHS>
__global__ void __launch_bounds__(CG_THREADS) cg_start(const double *__restrict__ b, const double *__restrict__ ax,
                                                       const double *__restrict__ inv, int64_t n, int64_t ld,
                                                       double *__restrict__ r, double *__restrict__ p, CgState *st,
                                                       double *__restrict__ partials) {
    __shared__ double ws[NRHS][CG_THREADS / 32];
    double dot[NRHS];
#pragma unroll
    for (int q = 0; q < NRHS; ++q) dot[q] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS) {
#pragma unroll
        for (int q = 0; q < NRHS; ++q) {
            const double rv = ax ? __dsub_rn(b[q * ld + i], ax[q * ld + i]) : b[q * ld + i]; // v2 :405-407
            const double zv = inv ? __dmul_rn(rv, inv[i]) : rv;                            // v2 :505
            r[q * ld + i] = rv;
            p[q * ld + i] = zv;
            dot[q] += zv * rv;
        }
    }
    if (cg_block_partial<NRHS>(dot, partials, &st->ticket[1], ws)) {
        double sum[NRHS];
        cg_fold<NRHS>(partials, gridDim.x, sum, ws);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < NRHS; ++q) st->rho[q] = sum[q];
        }
    }
}

// x += alpha p; r -= alpha Ap; (z = M^-1 r;) err = r.r, rho' = z.r; stop decision and beta by the last CTA
template <int NRHS, bool JACOBI>
__global__ void __launch_bounds__(CG_THREADS) cg_update_xr(double *__restrict__ x, double *__restrict__ r,
                                                           const double *__restrict__ p, const double *__restrict__ ap,
                                                           const double *__restrict__ inv, double *__restrict__ z,
                                                           int64_t n, int64_t ld, CgState *st,
                                                           double *__restrict__ partials) {
    constexpr int NV = JACOBI ? 2 * NRHS : NRHS;
    __shared__ double ws[NV][CG_THREADS / 32];
    if (*(volatile int *)&st->all_done) return;
    double alpha[NRHS];
    bool live[NRHS];
#pragma unroll
    for (int q = 0; q < NRHS; ++q) {
        live[q] = !st->done[q];
        alpha[q] = st->alpha[q];
    }
    double dot[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) dot[q] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS) {
#pragma unroll
        for (int q = 0; q < NRHS; ++q)
            if (live[q]) {
                x[q * ld + i] = __dadd_rn(x[q * ld + i], __dmul_rn(alpha[q], p[q * ld + i]));          // v2 :422 / :519
                const double rv = __dadd_rn(r[q * ld + i], __dmul_rn(-alpha[q], ap[q * ld + i]));     // v2 :423 / :520
                r[q * ld + i] = rv;
                dot[q] += rv * rv;
                if (JACOBI) {
                    const double zv = __dmul_rn(rv, inv[i]); // v2 :526
                    z[q * ld + i] = zv;
                    dot[NRHS + q] += zv * rv;
                }
            }
    }
    if (cg_block_partial<NV>(dot, partials, &st->ticket[1], ws)) {
        double sum[NV];
        cg_fold<NV>(partials, gridDim.x, sum, ws);
        if (threadIdx.x == 0) {
            int all = 1;
#pragma unroll
            for (int q = 0; q < NRHS; ++q) {
                if (!st->done[q]) {
                    st->err[q] = sum[q];
                    if (sqrt(sum[q]) < st->epsilon) {
                        st->done[q] = 1; // break: p is not updated, cnt is not bumped (v2 :425 / :524)
                    } else {
                        const double rho_new = JACOBI ? sum[NRHS + q] : sum[q];
                        st->beta[q] = rho_new / st->rho[q]; // v2 :426 / :528
                        st->rho[q] = rho_new;
                        st->cnt[q] += 1;
                        // the reference updates p, bumps cnt and only then tests cnt < max_iteration: a right-hand side
                        // that has used up its iterations stops here -- its last p update is skipped, which nothing
                        // observes (x and cnt are final)
                        if (st->cnt[q] >= st->max_iter) st->done[q] = 1;
                    }
                }
                all = all && st->done[q];
            }
            st->all_done = all;
        }
    }
}

// p = z + beta p (z = r for plain CG); then the loop condition cnt < max_iteration
template <int NRHS>
__global__ void __launch_bounds__(CG_THREADS) cg_update_p(double *__restrict__ p, const double *__restrict__ z, int64_t n,
                                                          int64_t ld, CgState *st) {
    if (*(volatile int *)&st->all_done) return;
    double beta[NRHS];
    bool live[NRHS];
#pragma unroll
    for (int q = 0; q < NRHS; ++q) {
        live[q] = !st->done[q];
        beta[q] = st->beta[q];
    }
    for (int64_t i = (int64_t)blockIdx.x * CG_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * CG_THREADS) {
#pragma unroll
        for (int q = 0; q < NRHS; ++q)
            if (live[q]) p[q * ld + i] = __dadd_rn(z[q * ld + i], __dmul_rn(beta[q], p[q * ld + i])); // v2 :427 / :530
    }
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

static int cg_common_check(gsb_matrix *m, const double *b, double *x, int nrhs) {
    if (!m || !b || !x || nrhs < 1 || nrhs > GSB_MAX_RHS) return GSB_ERR_ARG;
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

// A p (+ p.Ap): the ring kernel when the rows are short enough for its stage to be of use (GSB_CG_RING=0: never)
template <int NRHS, bool DOT>
static int cg_launch_spmv(gsb_matrix *m, const double *in, double *out, int64_t ld, CgState *ds, int grid_s, int *grid_ring,
                          cudaStream_t st) {
    static const int ring_env = [] {
        const char *e = getenv("GSB_CG_RING");
        return e ? atoi(e) : 1;
    }();
    const int64_t n = m->n_rows;
    if (ring_env && n >= 4 * CG_THREADS && m->store <= (int64_t)n * (CG_SPAN_CAP / CG_THREADS)) {
        auto kern = cg_spmv_ring<NRHS, DOT>;
        const int smem = 64 + 2 * cg_stage_layout().bytes;
        int per_sm = 1;
        GSB_TRY(gsb_kernel_occupancy((const void *)kern, smem, &per_sm, CG_THREADS));
        const int ntiles = (int)((n + CG_THREADS - 1) / CG_THREADS);
        int grid = gsb_sm_count() * per_sm;
        if (grid > ntiles) grid = ntiles;
        if (grid > *grid_ring) grid = *grid_ring; // (the partials buffer was sized for this many CTAs)
        kern<<<grid, CG_THREADS, smem, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p, m->n_rows, m->store, in, out, ld,
                                            ds, m->cg_partials.p);
        return GSB_OK;
    }
    cg_spmv_dot<NRHS, DOT><<<grid_s, CG_THREADS, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p, m->n_rows, m->store,
                                                         in, out, ld, ds, m->cg_partials.p);
    return GSB_OK;
}

template <int NRHS>
static int cg_run(gsb_matrix *m, const double *b_dev, const double *x0_dev, bool jacobi, double epsilon,
                  int max_iteration, double *x_dev, int *iters) {
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = m->n_rows;
    const int64_t ld = n; // callers hand over nrhs vectors of n doubles, one after another
    const int grid_v = gsb_blocks_for(n, CG_THREADS * 4, gsb_sm_count() * 8);
    const int ntiles = (int)((n + CG_THREADS - 1) / CG_THREADS);
    const int grid_s = ntiles < gsb_sm_count() * 8 ? ntiles : gsb_sm_count() * 8;
    int grid_ring = grid_v > grid_s ? grid_v : grid_s; // upper bound of the ring kernel's grid (slots in cg_partials)
    // workspace with the handle: r, p, Ap (, z, inv) + partials + state
    const int nvec = jacobi ? 4 : 3;
    GSB_TRY(m->cg_ws.alloc((int64_t)nvec * NRHS * n + (jacobi ? n : 0)));
    GSB_TRY(m->cg_partials.alloc((int64_t)(grid_v > grid_s ? grid_v : grid_s) * 2 * NRHS + 64));
    GSB_TRY(m->cg_state.alloc(sizeof(CgState)));
    if (!m->cg_state_host) GSB_CUDA(cudaHostAlloc(&m->cg_state_host, sizeof(CgState), cudaHostAllocDefault));
    double *r = m->cg_ws.p, *p = r + NRHS * n, *ap = p + NRHS * n, *z = jacobi ? ap + NRHS * n : r;
    double *inv = jacobi ? z + NRHS * n : nullptr;
    CgState *ds = (CgState *)m->cg_state.p, *hs = (CgState *)m->cg_state_host;
    memset(hs, 0, sizeof(CgState));
    hs->max_iter = max_iteration;
    hs->epsilon = epsilon;
    for (int q = NRHS; q < GSB_MAX_RHS; ++q) hs->done[q] = 1;
    if (max_iteration <= 0) { // while (cnt < max_iteration) never runs: x = x0
        for (int q = 0; q < NRHS; ++q) hs->done[q] = 1;
        hs->all_done = 1;
    }
    GSB_CUDA(cudaMemcpyAsync(ds, hs, sizeof(CgState), cudaMemcpyHostToDevice, st));
    const size_t bytes = sizeof(double) * (size_t)(NRHS * n);
    if (x0_dev) {
        if (x0_dev != x_dev) GSB_CUDA(cudaMemcpyAsync(x_dev, x0_dev, bytes, cudaMemcpyDeviceToDevice, st));
        GSB_TRY((cg_launch_spmv<NRHS, false>(m, x_dev, ap, ld, ds, grid_s, &grid_ring, st)));
        GSB_KERNEL_CHECK();
    } else {
        GSB_CUDA(cudaMemsetAsync(x_dev, 0, bytes, st)); // v2 :398-403: zero start
    }
    if (jacobi) {
        cg_inv_diag<<<(int)((n + 255) / 256), 256, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p, m->n_rows,
                                                           m->n_cols, inv);
        GSB_KERNEL_CHECK();
    }
    cg_start<NRHS><<<grid_v, CG_THREADS, 0, st>>>(b_dev, x0_dev ? ap : nullptr, inv, n, ld, r, p, ds, m->cg_partials.p);
    GSB_KERNEL_CHECK();
    int issued = 0;
    const int batch = 16;
    while (!hs->all_done) {
        for (int k = 0; k < batch; ++k) {
            GSB_TRY((cg_launch_spmv<NRHS, true>(m, p, ap, ld, ds, grid_s, &grid_ring, st)));
            if (jacobi)
                cg_update_xr<NRHS, true><<<grid_v, CG_THREADS, 0, st>>>(x_dev, r, p, ap, inv, z, n, ld, ds, m->cg_partials.p);
            else
                cg_update_xr<NRHS, false><<<grid_v, CG_THREADS, 0, st>>>(x_dev, r, p, ap, nullptr, nullptr, n, ld, ds,
                                                                        m->cg_partials.p);
            cg_update_p<NRHS><<<grid_v, CG_THREADS, 0, st>>>(p, z, n, ld, ds);
        }
        GSB_KERNEL_CHECK();
        issued += batch;
        GSB_CUDA(cudaMemcpyAsync(hs, ds, sizeof(CgState), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        if (issued > max_iteration + batch) break; // cannot happen (cg_update_xr stops a right-hand side at max_iteration)
    }
    GSB_CUDA(cudaStreamSynchronize(st));
    if (iters)
        for (int q = 0; q < NRHS; ++q) iters[q] = hs->cnt[q];
    return GSB_OK;
}

static int cg_dispatch(gsb_matrix *m, const double *b_dev, const double *x0_dev, int nrhs, bool jacobi, double epsilon,
                       int max_iteration, double *x_dev, int *iters) {
    switch (nrhs) {
        case 1: return cg_run<1>(m, b_dev, x0_dev, jacobi, epsilon, max_iteration, x_dev, iters);
        case 2: return cg_run<2>(m, b_dev, x0_dev, jacobi, epsilon, max_iteration, x_dev, iters);
        case 3: return cg_run<3>(m, b_dev, x0_dev, jacobi, epsilon, max_iteration, x_dev, iters);
        case 4: return cg_run<4>(m, b_dev, x0_dev, jacobi, epsilon, max_iteration, x_dev, iters);
    }
    return GSB_ERR_ARG;
}

// device-pointer core of conjugateGradient (v2 :396-434): b_dev, x_dev natural order; x0_dev may be null (zero start)
int gsb_cg_solve_device(gsb_matrix *m, const double *b_dev, const double *x0_dev, double epsilon, int max_iteration,
                        double *x_dev, int *iters) {
    return cg_dispatch(m, b_dev, x0_dev, 1, false, epsilon, max_iteration, x_dev, iters);
}
int gsb_cg_solve_device_multi(gsb_matrix *m, const double *b_dev, const double *x0_dev, int nrhs, double epsilon,
                              int max_iteration, double *x_dev, int *iters) {
    return cg_dispatch(m, b_dev, x0_dev, nrhs, false, epsilon, max_iteration, x_dev, iters);
}

static int cg_host(gsb_matrix *m, const double *b, int nrhs, bool jacobi, double epsilon, int max_iteration,
                   const double *x0, double *x_out, int *iters) {
    GSB_TRY(cg_common_check(m, b, x_out, nrhs));
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = m->n_rows;
    const size_t bytes = sizeof(double) * (size_t)(n * nrhs);
    DevBuf<double> &db = m->stage_b, &dx = m->stage_x; // device staging of the caller's host vectors, kept with the handle
    GSB_TRY(db.alloc(n * nrhs));
    GSB_TRY(dx.alloc(n * nrhs));
    GSB_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    if (x0) GSB_CUDA(cudaMemcpyAsync(dx.p, x0, bytes, cudaMemcpyHostToDevice, st));
    GSB_TRY(cg_dispatch(m, db.p, x0 ? dx.p : nullptr, nrhs, jacobi, epsilon, max_iteration, dx.p, iters));
    GSB_CUDA(cudaMemcpyAsync(x_out, dx.p, bytes, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_conjugate_gradient(gsb_matrix *m, const double *b, double epsilon, int max_iteration,
                                      const double *x0, double *x_out, int *iters) {
    return cg_host(m, b, 1, false, epsilon, max_iteration, x0, x_out, iters);
}

// EXTENSION: up to 4 right-hand sides (the colour channels) in one call; every SpMV reads the CSR once for all of them
extern "C" int gsb_conjugate_gradient_multi(gsb_matrix *m, const double *b, int nrhs, double epsilon, int max_iteration,
                                            const double *x0, double *x_out, int *iters) {
    return cg_host(m, b, nrhs, false, epsilon, max_iteration, x0, x_out, iters);
}

extern "C" int gsb_conjugate_gradient_jacobi(gsb_matrix *m, const double *b, double epsilon, int max_iteration,
                                             double *x_out, int *iters) {
    return cg_host(m, b, 1, true, epsilon, max_iteration, nullptr, x_out, iters);
}
