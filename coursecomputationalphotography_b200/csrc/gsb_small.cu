// gsb_small.cu -- kernel 6: small systems, the whole solve in ONE persistent launch.
//
// BASELINE configs[0] (lab3: n = 1e4, ~5 entries per row, 18 sweeps, 5 colours) is launch bound: as a replayed CUDA
// graph it is 17 x (5 colour phases + end of sweep) = ~100 kernel nodes of ~7 us each for ~1 us of work per node
// (profiles/README.md, c1: 764 us per solve).  Here one grid of resident CTAs keeps running: a colour phase is a
// grid-stride pass over the colour's rows (row per thread, the row body of kernel 1), colours are separated by a
// grid barrier, and the stop rule is part of the loop -- every CTA leaves its partial of the sweep's L1 update norm,
// CTA 0 folds them in a fixed order, bumps the counter and decides, a second barrier publishes the decision.  One
// launch, (colours + 2) barriers per sweep, no host round trip until the solve has stopped.
// Arithmetic: gs_row_sigma, unfused, storage order -> x is bit-identical to kernels 1-5 after every sweep.
// All CTAs must be resident (the barrier spins): grid = min(ceil(largest colour / 256), SMs x occupancy).
#include "gsb_ring.cuh"

#include <stdlib.h>

#define GS_SMALL_MAX_COLORS 64

struct GsbSmallArgs {
    int n_colors;
    int color_start[GS_SMALL_MAX_COLORS + 1];
    unsigned *bar; // [0] arrival counter, [1] generation
    int max_sweeps; // per launch
};

// Barrier over the whole grid.  One CTA (the smallest systems): __syncthreads -- the CTA's own global writes are
// visible to its threads.  Several CTAs: a counting barrier in global memory; `gen` is the number of barriers this
// CTA has passed (all CTAs pass the same sequence, so it needs no load), the last arrival resets the counter and
// publishes gen + 1 with a release store, the others spin on it with acquire loads (which also drop the SM's stale L1
// lines).  Bounded: a CTA that never arrives (it cannot, all are resident) raises ctl->error instead of hanging.
__device__ __forceinline__ bool small_grid_barrier(unsigned *bar, unsigned &gen, GsCtl *ctl) {
    __syncthreads();
    if (gridDim.x == 1) return true;
    __shared__ int ok_s;
    if (threadIdx.x == 0) {
        int ok = 1;
        unsigned ticket;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(bar) : "memory");
        if (ticket == gridDim.x - 1) {
            asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(0u) : "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
        } else {
            unsigned cur;
            int spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar + 1) : "memory");
                if (cur != gen) break;
                if (++spins > (1 << 24)) {
                    ctl->error = 3;
                    ok = 0;
                    break;
                }
            }
        }
        ok_s = ok;
    }
    ++gen;
    __syncthreads();
    return ok_s != 0;
}

template <int NRHS, int THREADS>
__global__ void __launch_bounds__(THREADS) gs_small_persistent(const int *__restrict__ rp, const int *__restrict__ ci,
                                                               const double *__restrict__ va,
                                                               const double *__restrict__ dg,
                                                               const double *__restrict__ b, double *x, int64_t n,
                                                               GsCtl *ctl, double *partials, const GsbSmallArgs a) {
    __shared__ double ws[NRHS][THREADS / 32];
    __shared__ double tot_s[NRHS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool single = gridDim.x == 1; // one CTA: its own writes are visible through L1, no need to go to L2
    unsigned gen = 0;                   // (the host zeroes the barrier words before every solve)
    if (*(volatile int *)&ctl->done) return;
    for (int sweep = 0; sweep < a.max_sweeps; ++sweep) {
        double acc[NRHS];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
        for (int c = 0; c < a.n_colors; ++c) {
            const int row0 = a.color_start[c], row1 = a.color_start[c + 1];
            for (int i = row0 + blockIdx.x * THREADS + tid; i < row1; i += gridDim.x * THREADS) {
                const int k0 = rp[i], k1 = rp[i + 1];
                const double d = dg[i];
                double sig[NRHS];
                // x of the other colours was written before the last barrier, by other CTAs unless there is only one
                if (single)
                    gs_row_sigma<NRHS>(ci + k0, va + k0, k1 - k0, [&](int col, int r) { return x[r * n + col]; }, sig);
                else
                    gs_row_sigma<NRHS>(ci + k0, va + k0, k1 - k0, [&](int col, int r) { return __ldcg(x + r * n + col); }, sig);
                if (d != 0.0) { // zero or absent diagonal: row skipped, x_i unchanged (v2 :360-363)
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) {
                        const double xn = __ddiv_rn(__dsub_rn(b[r * n + i], sig[r]), d);
                        acc[r] += fabs(xn - (single ? x[r * n + i] : __ldcg(x + r * n + i)));
                        x[r * n + i] = xn;
                    }
                }
            }
            if (c + 1 < a.n_colors && !small_grid_barrier(a.bar, gen, ctl)) return;
        }
        // this CTA's share of the sweep's L1 update norm (fixed order) -> partial slot blockIdx.x
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            double t = acc[r];
#pragma unroll
            for (int d2 = 16; d2 > 0; d2 >>= 1) t += __shfl_down_sync(0xffffffffu, t, d2);
            if (lane == 0) ws[r][wid] = t;
        }
        __syncthreads();
        if (tid < NRHS) {
            double t = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) t += ws[tid][w];
            partials[(size_t)blockIdx.x * NRHS + tid] = t;
        }
        if (!small_grid_barrier(a.bar, gen, ctl)) return;
        if (blockIdx.x == 0) { // fold in slot order, decide (v2 :356, :376-377)
            double tot = 0.0;
            if (tid < NRHS)
                for (unsigned q = 0; q < gridDim.x; ++q) tot += __ldcg(partials + (size_t)q * NRHS + tid);
            if (tid < NRHS) tot_s[tid] = tot;
            __syncthreads();
            if (tid == 0) {
                bool all_ok = true;
#pragma unroll
                for (int r = 0; r < NRHS; ++r) {
                    ctl->eps_last[r] = tot_s[r];
                    if (tot_s[r] > ctl->epsilon) all_ok = false;
                }
                const int cnt = ctl->sweeps + 1;
                ctl->sweeps = cnt;
                if (all_ok || cnt >= ctl->max_iter) ctl->done = 1;
            }
        }
        if (!small_grid_barrier(a.bar, gen, ctl)) return;
        if (*(volatile int *)&ctl->done) return;
    }
}

// Is the system small enough for kernel 6 to be the better path?  (auto policy; GSB_SMALL_PERSISTENT=0|1 forces)
bool gsb_small_auto(int64_t n_rows, int n_colors, int check_every) {
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("GSB_SMALL_PERSISTENT");
        env = e ? atoi(e) : 2;
    }
    if (env == 0 || check_every != 1 || n_colors < 1 || n_colors > GS_SMALL_MAX_COLORS) return false;
    if (env == 1) return true;
    return n_rows * (int64_t)n_colors < ((int64_t)1 << 20); // where the CUDA-graph path used to be chosen
}

#define GS_SMALL_SINGLE_ROWS 16384 // up to this many rows: ONE CTA of 1024 threads, barriers are __syncthreads

template <int NRHS>
static int launch_small_t(const int *rp, const int *ci, const double *va, const double *dg, const double *b, double *x,
                          int64_t ld, const int *color_start, int n_colors, GsCtl *ctl, double *partials, unsigned *bar,
                          int max_sweeps, cudaStream_t st, int *slots) {
    GsbSmallArgs a;
    memset(&a, 0, sizeof(a));
    a.n_colors = n_colors;
    for (int c = 0; c <= n_colors; ++c) a.color_start[c] = color_start[c];
    a.bar = bar;
    a.max_sweeps = max_sweeps;
    int largest = 1;
    for (int c = 0; c < n_colors; ++c) largest = largest > color_start[c + 1] - color_start[c] ? largest : color_start[c + 1] - color_start[c];
    if (color_start[n_colors] - color_start[0] <= GS_SMALL_SINGLE_ROWS) {
        gs_small_persistent<NRHS, 1024><<<1, 1024, 0, st>>>(rp, ci, va, dg, b, x, ld, ctl, partials, a);
        GSB_KERNEL_CHECK();
        if (slots) *slots = 1;
        return GSB_OK;
    }
    auto kern = gs_small_persistent<NRHS, GS_THREADS>;
    int per_sm = 1;
    GSB_TRY(gsb_kernel_occupancy((const void *)kern, 0, &per_sm, GS_THREADS));
    int grid = (largest + GS_THREADS - 1) / GS_THREADS;
    const int resident = gsb_sm_count() * per_sm;
    if (grid > resident) grid = resident;
    if (grid > GSB_RING_SLOTS_MAX) grid = GSB_RING_SLOTS_MAX;
    kern<<<grid, GS_THREADS, 0, st>>>(rp, ci, va, dg, b, x, ld, ctl, partials, a);
    GSB_KERNEL_CHECK();
    if (slots) *slots = grid;
    return GSB_OK;
}

int gsb_launch_small_persistent(const int *rp, const int *ci, const double *va, const double *dg, const double *b, double *x,
                                int64_t ld, int nrhs, const int *color_start, int n_colors, GsCtl *ctl, double *partials,
                                unsigned *bar, int max_sweeps, cudaStream_t st, int *slots) {
    switch (nrhs) {
        case 1: return launch_small_t<1>(rp, ci, va, dg, b, x, ld, color_start, n_colors, ctl, partials, bar, max_sweeps, st, slots);
        case 2: return launch_small_t<2>(rp, ci, va, dg, b, x, ld, color_start, n_colors, ctl, partials, bar, max_sweeps, st, slots);
        case 3: return launch_small_t<3>(rp, ci, va, dg, b, x, ld, color_start, n_colors, ctl, partials, bar, max_sweeps, st, slots);
        case 4: return launch_small_t<4>(rp, ci, va, dg, b, x, ld, color_start, n_colors, ctl, partials, bar, max_sweeps, st, slots);
    }
    gsb_set_error("nrhs must be 1..%d", GSB_MAX_RHS);
    return GSB_ERR_ARG;
}
