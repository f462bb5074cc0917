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

#include <cooperative_groups.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

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

// ---------------------------------------------------------------------------------------------
// One CTA (<= GS_SMALL_SINGLE_ROWS rows, BASELINE configs[0]): the solve is a chain of (colour, pass) steps separated
// by __syncthreads, and the time of a step is the chain of dependent loads inside it -- row offsets -> entries ->
// x gathers -> divide (measured: 3.7 us per colour of 1250 rows, 2 passes).  The matrix does not change and the
// schedule is static (a thread meets the same rows in the same order every sweep), so everything that does not depend
// on x is loaded ahead: the row offsets two steps ahead, the entries, the diagonal, b (and the row's own old x) one
// step ahead, while the current step's gathers are in flight.  What stays on the critical path of a step is one
// gather round trip, the multiply-add chain, the divide and the barrier -- and with XS the gathers come out of shared
// memory: x (all right-hand sides) lives there for the whole solve when it fits and is written back once.
// CSIZE > 1: the same loop on a thread-block CLUSTER of CSIZE CTAs (one SM each).  One SM is what limits the one-CTA
// version (measured: 1.4 us per step of 1024 rows -- ~200 instructions per row, 45 of them the IEEE divide, through
// one SM's issue slots).  Every CTA keeps a complete copy of x in its shared memory; a thread that has computed x_i
// stores it into all CSIZE copies (distributed shared memory), so that every gather stays a local shared-memory read,
// and the colours are separated by the hardware cluster barrier (release / acquire at cluster scope orders the
// remote stores) instead of __syncthreads.  The stop rule's partials meet in CTA 0's shared memory, which folds
// them in rank order and writes its decision into every CTA.
// Arithmetic and its order are those of gs_row_sigma / gs_small_persistent: the same bits.
struct SmallRowA {
    int i, k0, k1; // i < 0: this thread has no row in the step.  (The row length is k1 - k0, formed where it is used:
                   // an in-order warp stalls at the first instruction that consumes a load.)
    __device__ __forceinline__ int len() const { return k1 - k0; }
};
template <int NRHS>
struct SmallRowB {
    int cc[GS_UNROLL];
    double vv[GS_UNROLL];
    double d, bb[NRHS], xo[NRHS];
};

template <int NRHS, int THREADS, bool XS, int CSIZE>
__global__ void __launch_bounds__(THREADS) gs_small_one_cta(const int *__restrict__ rp, const int *__restrict__ ci,
                                                            const double *__restrict__ va, const double *__restrict__ dg,
                                                            const double *__restrict__ b, double *x, int64_t n, GsCtl *ctl,
                                                            double *partials, const GsbSmallArgs a) {
    extern __shared__ __align__(16) double xs[]; // XS: NRHS planes of `nrows` doubles
    __shared__ double ws[NRHS][THREADS / 32];
    __shared__ int cs[GS_SMALL_MAX_COLORS + 1];
    __shared__ int done_s;
    __shared__ double part_s[CSIZE][NRHS]; // CTA 0: every CTA's share of the sweep's update norm
    static_assert(CSIZE == 1 || XS, "the cluster version keeps x in (distributed) shared memory");
    constexpr int STRIDE = CSIZE * THREADS; // rows per step
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int rank = 0;
    double *xs_of[CSIZE]; // every CTA's copy of x ([0] = this CTA's for CSIZE == 1)
    xs_of[0] = xs;
    if constexpr (CSIZE > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        rank = (int)cluster.block_rank();
#pragma unroll
        for (int q = 0; q < CSIZE; ++q) xs_of[q] = cluster.map_shared_rank(xs, q);
    }
    auto barrier = [&]() {
        if constexpr (CSIZE > 1)
            cg::this_cluster().sync();
        else
            __syncthreads();
    };
    if (*(volatile int *)&ctl->done) return; // (the same value in every CTA: nobody is left waiting)
    const int nc = a.n_colors;
    for (int c = tid; c <= nc; c += THREADS) cs[c] = a.color_start[c];
    const int nrows = a.color_start[nc];
    if (XS) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
            for (int i = tid; i < nrows; i += THREADS) xs[r * nrows + i] = x[r * n + i];
    }
    barrier(); // (cluster: no CTA stores into a copy that is still being filled)
    auto xread = [&](int col, int r) -> double { return XS ? xs[r * nrows + col] : x[r * n + col]; };
    // the step after (c, p) in sweep order; true when it belongs to the next sweep
    auto advance = [&](int &c, int &p) -> bool {
        if (cs[c] + (p + 1) * STRIDE < cs[c + 1]) {
            ++p;
            return false;
        }
        p = 0;
        if (++c == nc) {
            c = 0;
            return true;
        }
        return false;
    };
    auto load_a = [&](int c, int p) -> SmallRowA {
        SmallRowA A;
        const int i = cs[c] + p * STRIDE + rank * THREADS + tid;
        if (i < cs[c + 1]) {
            A.i = i;
            A.k0 = rp[i];
            A.k1 = rp[i + 1];
        } else {
            A.i = -1;
            A.k0 = 0;
            A.k1 = 0;
        }
        return A;
    };
    auto load_b = [&](const SmallRowA &A) -> SmallRowB<NRHS> {
        SmallRowB<NRHS> B;
        const bool row = A.i >= 0, shortrow = row && A.len() <= GS_UNROLL;
#pragma unroll
        for (int j = 0; j < GS_UNROLL; ++j) {
            B.cc[j] = (shortrow && j < A.len()) ? ci[A.k0 + j] : 0;
            B.vv[j] = (shortrow && j < A.len()) ? va[A.k0 + j] : 0.0;
        }
        B.d = row ? dg[A.i] : 0.0;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            B.bb[r] = row ? b[r * n + A.i] : 0.0;
            B.xo[r] = (row && !XS) ? x[r * n + A.i] : 0.0; // the row's own old value: nobody else writes it
        }
        return B;
    };
    double acc[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
    auto compute = [&](const SmallRowA &A, const SmallRowB<NRHS> &B) {
        if (A.i < 0) return;
        double sig[NRHS];
        if (A.len() <= GS_UNROLL) {
            double xg[GS_UNROLL][NRHS];
#pragma unroll
            for (int j = 0; j < GS_UNROLL; ++j)
#pragma unroll
                for (int r = 0; r < NRHS; ++r) xg[j][r] = xread(B.cc[j], r); // padded positions read index 0 (valid)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) sig[r] = 0.0;
#pragma unroll
            for (int j = 0; j < GS_UNROLL; ++j)
                if (j < A.len()) {
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) sig[r] = __dadd_rn(sig[r], __dmul_rn(B.vv[j], xg[j][r]));
                }
        } else {
            gs_row_sigma<NRHS>(ci + A.k0, va + A.k0, A.len(), xread, sig);
        }
        if (B.d != 0.0) { // zero or absent diagonal: row skipped, x_i unchanged (v2 :360-363)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                const double xn = __ddiv_rn(__dsub_rn(B.bb[r], sig[r]), B.d);
                if (XS) {
                    acc[r] += fabs(xn - xs[r * nrows + A.i]);
#pragma unroll
                    for (int q = 0; q < CSIZE; ++q) xs_of[q][r * nrows + A.i] = xn;
                } else {
                    acc[r] += fabs(xn - B.xo[r]);
                    x[r * n + A.i] = xn;
                }
            }
        }
    };

    int c0 = 0, p0 = 0, c1 = 0, p1 = 0, sweep = 0;
    SmallRowA A0 = load_a(0, 0);
    SmallRowB<NRHS> B0 = load_b(A0);
    bool wrap1 = advance(c1, p1);
    SmallRowA A1 = load_a(c1, p1);
    for (;;) {
        int c2 = c1, p2 = p1;
        const bool wrap2 = advance(c2, p2);
        const SmallRowA A2 = load_a(c2, p2);      // two steps ahead: row offsets
        const SmallRowB<NRHS> B1 = load_b(A1);    // one step ahead: entries, diagonal, b
        compute(A0, B0);
        if (wrap1) { // the sweep is complete: fold the update norm in a fixed order, decide (v2 :356, :376-377)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                double t = acc[r];
#pragma unroll
                for (int d2 = 16; d2 > 0; d2 >>= 1) t += __shfl_down_sync(0xffffffffu, t, d2);
                if (lane == 0) ws[r][wid] = t;
                acc[r] = 0.0;
            }
            __syncthreads();
            if (tid < NRHS) { // this CTA's share, warps in order -> slot `rank` of CTA 0
                double t = 0.0;
                for (int w = 0; w < THREADS / 32; ++w) t += ws[tid][w];
                if constexpr (CSIZE > 1)
                    cg::this_cluster().map_shared_rank(&part_s[0][0], 0)[rank * NRHS + tid] = t;
                else
                    part_s[0][tid] = t;
            }
            barrier();
            if (rank == 0 && tid == 0) {
                bool all_ok = true;
#pragma unroll
                for (int r = 0; r < NRHS; ++r) {
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < CSIZE; ++q) {
                        partials[q * NRHS + r] = part_s[q][r];
                        tot += part_s[q][r];
                    }
                    ctl->eps_last[r] = tot;
                    if (tot > ctl->epsilon) all_ok = false;
                }
                const int cnt = ctl->sweeps + 1;
                ctl->sweeps = cnt;
                const int d = (all_ok || cnt >= ctl->max_iter) ? 1 : 0;
                if (d) ctl->done = 1;
                if constexpr (CSIZE > 1) {
#pragma unroll
                    for (int q = 0; q < CSIZE; ++q) *cg::this_cluster().map_shared_rank(&done_s, q) = d;
                } else {
                    done_s = d;
                }
            }
            barrier();
            if (done_s || ++sweep >= a.max_sweeps) break;
        } else if (c1 != c0) {
            barrier();
        }
        A0 = A1;
        B0 = B1;
        A1 = A2;
        c0 = c1;
        p0 = p1;
        c1 = c2;
        p1 = p2;
        wrap1 = wrap2;
    }
    if (XS) { // all copies are equal after the last barrier: every CTA writes a share back
#pragma unroll
        for (int r = 0; r < NRHS; ++r)
            for (int i = rank * THREADS + tid; i < nrows; i += STRIDE) x[r * n + i] = xs[r * nrows + i];
    }
}

// ---------------------------------------------------------------------------------------------
// Cluster version, second generation: no fence and no cluster barrier inside the loop.
//
// ncu of gs_small_one_cta<.., CSIZE = 8> on configs[0] (profiles/r02_small_cluster_v1.txt): 3600 cycles per colour,
// of which the row arithmetic is ~400 -- cluster.sync() is MEMBAR.ALL.GPU + ERRBAR + UCGABAR_ARV / _WAIT + CCTL.IVALL
// (48 % of the stall samples), and the per-step bookkeeping is executed by 16 warps per SM of which 5 have rows.
// Here the colour barrier is the data itself: a thread sends x_i to every CTA's copy with st.async (SASS STAS: an
// asynchronous store into a peer's shared memory that completes transaction bytes on a peer mbarrier), every CTA
// arms one mbarrier per colour phase with the bytes that phase delivers -- (rows of the colour) x NRHS x 8, the same
// in every CTA -- and its threads wait on that mbarrier before they gather.  The hardware counts the bytes; nothing
// is flushed, nobody waits for anybody's outstanding loads.  Two mbarriers alternate (a CTA can be at most one phase
// ahead of another: it passes phase k only with every CTA's phase-k values, which a CTA sends only after it has
// passed phase k-1), and one __syncthreads per phase keeps the threads of a CTA within a phase of each other, so that
// the one-bit phase parity cannot be misread.  Write after read needs no extra step: the bytes of a thread's row
// arrive only after that thread's gathers (its value depends on them), so when a CTA has seen phase k complete no
// thread anywhere still reads the old values that phase k+1 overwrites.  Rows with a zero diagonal send their
// unchanged value (the byte count stays exact).  The stop rule travels the same way: every CTA's partial to CTA 0's
// `red` mbarrier, CTA 0's decision to every CTA's `done` mbarrier.
// Arithmetic and its order are those of gs_row_sigma / gs_small_persistent: the same bits.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_b64(uint32_t addr, unsigned long long v, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr), "l"(v), "r"(mbar)
                 : "memory");
}
// bounded: a byte that never arrives (it cannot, short of a fault) raises ctl->error instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity, GsCtl *ctl) {
    uint32_t done = 0;
    for (int spins = 0; spins < (1 << 24); ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return true;
    }
    ctl->error = 3;
    return false;
}

#define GS_SMALL_CLUSTER_THREADS_MAX 512
#define GS_SMALL_MAX_STEPS 192 // (colour, pass) steps per sweep the step table holds

// One (colour, pass) step of the sweep as the cluster sees it: rows [base, end) are dealt out `threads` per CTA;
// flags: 1 = the step opens its colour's phase, 2 = it closes it (wait for the phase's bytes), 4 = it closes the sweep
struct __align__(16) SmallStep {
    int base, end, flags, phase_bytes;
};

template <int NRHS, int CSIZE>
__global__ void __launch_bounds__(GS_SMALL_CLUSTER_THREADS_MAX) gs_small_cluster(const int *__restrict__ rp, const int *__restrict__ ci,
                                                            const double *__restrict__ va, const double *__restrict__ dg,
                                                            const double *__restrict__ b, double *x, int64_t n, GsCtl *ctl,
                                                            double *partials, const GsbSmallArgs a) {
    extern __shared__ __align__(16) double xs[]; // this CTA's copy of x: NRHS planes of `nrows` doubles
    __shared__ double ws[NRHS][GS_SMALL_CLUSTER_THREADS_MAX / 32];
    __shared__ SmallStep steps[GS_SMALL_MAX_STEPS];
    __shared__ int nsteps_s;
    __shared__ __align__(8) double part_s[CSIZE][NRHS]; // CTA 0: every CTA's share of the sweep's update norm
    __shared__ __align__(8) unsigned long long done_s;
    __shared__ __align__(8) uint64_t bar_x[2], bar_red, bar_done;
    const int threads = (int)blockDim.x, stride = CSIZE * threads; // rows per step
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = threads >> 5;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    if (*(volatile int *)&ctl->done) return; // (the same value in every CTA: nobody is left waiting)
    const int nc = a.n_colors;
    const int nrows = a.color_start[nc];
    const int ldx = (nrows + 1) & ~1; // plane stride of the copies: even, so that a row's parity decides 16-byte alignment
    if (tid == 0) { // the sweep's schedule, once (the host has checked that it fits)
        int ns = 0;
        for (int c = 0; c < nc; ++c) {
            const int c0 = a.color_start[c], c1 = a.color_start[c + 1];
            int base = c0;
            do {
                SmallStep sp;
                sp.base = base;
                sp.end = c1;
                sp.phase_bytes = (c1 - c0) * NRHS * 8; // every row of the colour arrives here, this CTA's own included
                sp.flags = (base == c0 ? 1 : 0) | (base + stride >= c1 ? 2 : 0) | ((base + stride >= c1 && c == nc - 1) ? 4 : 0);
                steps[ns++] = sp;
                base += stride;
            } while (base < c1);
        }
        nsteps_s = ns;
        mbar_init(&bar_x[0], 1);
        mbar_init(&bar_x[1], 1);
        mbar_init(&bar_red, 1);
        mbar_init(&bar_done, 1);
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r)
        for (int i = tid; i < nrows; i += threads) xs[r * ldx + i] = x[r * n + i];
    cluster.sync(); // copies filled, schedule and mbarriers in place everywhere before the first remote store
    const int nsteps = nsteps_s;
    uint32_t xs_of[CSIZE], bx_of[CSIZE][2]; // shared::cluster addresses of every CTA's copy and phase mbarriers
#pragma unroll
    for (int q = 0; q < CSIZE; ++q) {
        xs_of[q] = mapa_u32(smem_u32(xs), q);
        bx_of[q][0] = mapa_u32(smem_u32(&bar_x[0]), q);
        bx_of[q][1] = mapa_u32(smem_u32(&bar_x[1]), q);
    }
    auto xread = [&](int col, int r) -> double { return xs[r * ldx + col]; };
    const int my_off = rank * threads + tid;
    auto load_a = [&](int sidx) -> SmallRowA {
        SmallRowA A;
        const int i = steps[sidx].base + my_off;
        if (i < steps[sidx].end) {
            A.i = i;
            A.k0 = rp[i];
            A.k1 = rp[i + 1];
        } else {
            A.i = -1;
            A.k0 = 0;
            A.k1 = 0;
        }
        return A;
    };
    auto load_b = [&](const SmallRowA &A) -> SmallRowB<NRHS> {
        SmallRowB<NRHS> B;
        const bool row = A.i >= 0, shortrow = row && A.len() <= GS_UNROLL;
#pragma unroll
        for (int j = 0; j < GS_UNROLL; ++j) {
            B.cc[j] = (shortrow && j < A.len()) ? ci[A.k0 + j] : 0;
            B.vv[j] = (shortrow && j < A.len()) ? va[A.k0 + j] : 0.0;
        }
        B.d = row ? dg[A.i] : 0.0;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            B.bb[r] = row ? b[r * n + A.i] : 0.0;
            B.xo[r] = 0.0;
        }
        return B;
    };
    double acc[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
    int phase = 0; // colour phases since the start of the launch: mbarrier bar_x[phase & 1], parity (phase >> 1) & 1
    // (One bulk copy per warp and peer -- shared memory to shared::cluster, 256 bytes instead of 32 st.async packets --
    // was measured too: 220 us against 193 us on configs[0], profiles/r02_call24_summary.txt.  A bulk copy's fixed
    // latency is on the critical path of every step; the single stores leave as soon as the row is computed.)
    auto compute = [&](const SmallRowA &A, const SmallRowB<NRHS> &B) {
        if (A.i < 0) return;
        double sig[NRHS];
        if (A.len() <= GS_UNROLL) {
            double xg[GS_UNROLL][NRHS];
#pragma unroll
            for (int j = 0; j < GS_UNROLL; ++j)
#pragma unroll
                for (int r = 0; r < NRHS; ++r) xg[j][r] = xread(B.cc[j], r); // padded positions read index 0 (valid)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) sig[r] = 0.0;
#pragma unroll
            for (int j = 0; j < GS_UNROLL; ++j)
                if (j < A.len()) {
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) sig[r] = __dadd_rn(sig[r], __dmul_rn(B.vv[j], xg[j][r]));
                }
        } else {
            gs_row_sigma<NRHS>(ci + A.k0, va + A.k0, A.len(), xread, sig);
        }
        const uint32_t pb = (uint32_t)phase & 1u;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            const double xo = xs[r * ldx + A.i];
            double xn = xo; // zero or absent diagonal: row skipped, x_i unchanged (v2 :360-363) -- and sent as it is
            if (B.d != 0.0) {
                xn = __ddiv_rn(__dsub_rn(B.bb[r], sig[r]), B.d);
                acc[r] += fabs(xn - xo);
            }
            const uint32_t off = (uint32_t)(r * ldx + A.i) * 8u;
#pragma unroll
            for (int q = 0; q < CSIZE; ++q)
                st_async_b64(xs_of[q] + off, (unsigned long long)__double_as_longlong(xn), pb ? bx_of[q][1] : bx_of[q][0]);
        }
    };
    auto next = [&](int sidx) -> int { return sidx + 1 == nsteps ? 0 : sidx + 1; };

    int s0 = 0, s1 = next(0), sweep = 0;
    SmallRowA A0 = load_a(0);
    SmallRowB<NRHS> B0 = load_b(A0);
    SmallRowA A1 = load_a(s1);
    for (;;) {
        const int flags = steps[s0].flags;
        if (flags & 1) { // a colour phase opens: every thread of the CTA has passed the previous phase's wait
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&bar_x[phase & 1], (uint32_t)steps[s0].phase_bytes);
                if (s0 == 0) { // ... and a sweep: the stop rule's two mbarriers
                    if (rank == 0) mbar_expect_tx(&bar_red, (uint32_t)(CSIZE * NRHS * 8));
                    mbar_expect_tx(&bar_done, 8u);
                }
            }
        }
        const int s2 = next(s1);
        const SmallRowA A2 = load_a(s2);       // two steps ahead: row offsets
        const SmallRowB<NRHS> B1 = load_b(A1); // one step ahead: entries, diagonal, b
        compute(A0, B0);
        if (flags & 2) { // the colour is complete once its bytes are: every CTA's values of it are in this copy
            if (!mbar_wait_bounded(&bar_x[phase & 1], (uint32_t)(phase >> 1) & 1u, ctl)) return;
            ++phase;
        }
        if (flags & 4) { // the sweep is complete: fold the update norm in a fixed order, decide (v2 :356, :376-377)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                double t = acc[r];
#pragma unroll
                for (int d2 = 16; d2 > 0; d2 >>= 1) t += __shfl_down_sync(0xffffffffu, t, d2);
                if (lane == 0) ws[r][wid] = t;
                acc[r] = 0.0;
            }
            __syncthreads();
            if (tid < NRHS) { // this CTA's share, warps in order -> slot `rank` of CTA 0
                double t = 0.0;
                for (int w = 0; w < nwarps; ++w) t += ws[tid][w];
                st_async_b64(mapa_u32(smem_u32(&part_s[rank][tid]), 0), (unsigned long long)__double_as_longlong(t),
                             mapa_u32(smem_u32(&bar_red), 0));
            }
            const uint32_t sp = (uint32_t)sweep & 1u;
            if (rank == 0 && tid == 0) {
                if (!mbar_wait_bounded(&bar_red, sp, ctl)) return;
                bool all_ok = true;
#pragma unroll
                for (int r = 0; r < NRHS; ++r) {
                    double tot = 0.0;
#pragma unroll
                    for (int q = 0; q < CSIZE; ++q) {
                        partials[q * NRHS + r] = part_s[q][r];
                        tot += part_s[q][r];
                    }
                    ctl->eps_last[r] = tot;
                    if (tot > ctl->epsilon) all_ok = false;
                }
                const int cnt = ctl->sweeps + 1;
                ctl->sweeps = cnt;
                const int d = (all_ok || cnt >= ctl->max_iter) ? 1 : 0;
                if (d) ctl->done = 1;
#pragma unroll
                for (int q = 0; q < CSIZE; ++q)
                    st_async_b64(mapa_u32(smem_u32(&done_s), q), (unsigned long long)d, mapa_u32(smem_u32(&bar_done), q));
            }
            if (!mbar_wait_bounded(&bar_done, sp, ctl)) return;
            if (done_s != 0ull || ++sweep >= a.max_sweeps) break;
        }
        A0 = A1;
        B0 = B1;
        A1 = A2;
        s0 = s1;
        s1 = s2;
    }
    // all copies are equal (every phase has been waited for): every CTA writes a share back
#pragma unroll
    for (int r = 0; r < NRHS; ++r)
        for (int i = my_off; i < nrows; i += stride) x[r * n + i] = xs[r * ldx + i];
    cluster.sync(); // no CTA leaves while a peer could still be sending to it
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

#define GS_SMALL_SINGLE_ROWS 16384 // up to this many rows: ONE CTA, barriers are __syncthreads
#define GS_SMALL_XS_BYTES_MAX (200 * 1024) // x of all right-hand sides in shared memory up to this size
#define GS_SMALL_CLUSTER 8                 // CTAs of the cluster version (the portable maximum)
#define GS_SMALL_CLUSTER_MIN_ROWS 1024     // largest colour: below this one CTA has the rows in a single pass anyway

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
        // GSB_SMALL_PIPE: 0 = the plain one-CTA loop, 1 = loads issued ahead (x in global memory), 2 = and x in shared
        // memory when all right-hand sides fit, 3 (default) = and on a cluster of GS_SMALL_CLUSTER CTAs when a colour
        // has enough rows to occupy it (gs_small_cluster: st.async + mbarriers); 4 = the cluster.sync() version of that
        static const int pipe = [] {
            const char *e = getenv("GSB_SMALL_PIPE");
            return e ? atoi(e) : 3;
        }();
        // (planes padded to an even length: the cluster version aligns its bulk copies on that)
        const size_t xs_bytes = sizeof(double) * (size_t)NRHS * (size_t)((color_start[n_colors] + 1) & ~1);
        const int dev = gsb_current_device();
        const bool dev_ok = dev >= 0 && dev < 64;
        if (pipe >= 1 && color_start[0] == 0) {
            constexpr int T = NRHS == 1 ? 1024 : 512; // registers: the rows in flight of k right-hand sides
            constexpr int CS = GS_SMALL_CLUSTER; // cluster version: CS CTAs
            // threads per CTA: enough for the largest colour in ONE pass when that is possible (a pass costs a fixed
            // latency, not throughput), a multiple of 32 in [128, 512]
            int TC = ((largest + CS - 1) / CS + 31) / 32 * 32;
            if (TC < 128) TC = 128;
            if (TC > GS_SMALL_CLUSTER_THREADS_MAX) TC = GS_SMALL_CLUSTER_THREADS_MAX;
            int cluster_steps = 0;
            for (int c = 0; c < n_colors; ++c) {
                const int rows_c = color_start[c + 1] - color_start[c];
                cluster_steps += rows_c > 0 ? (rows_c + CS * TC - 1) / (CS * TC) : 1;
            }
            static int cluster_ok[64] = {0};                // per device: 0 unknown, 1 usable, -1 not
            bool launched = false;
            if (pipe >= 3 && xs_bytes <= GS_SMALL_XS_BYTES_MAX && largest >= GS_SMALL_CLUSTER_MIN_ROWS &&
                cluster_steps <= GS_SMALL_MAX_STEPS && !(dev_ok && cluster_ok[dev] < 0)) {
                if (pipe == 4) TC = 256;
                auto kern = pipe == 4 ? gs_small_one_cta<NRHS, 256, true, CS> : gs_small_cluster<NRHS, CS>;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(CS);
                cfg.blockDim = dim3(TC);
                cfg.dynamicSmemBytes = xs_bytes;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = CS;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                if (!dev_ok || cluster_ok[dev] == 0) { // once per device: shared-memory limit, can a cluster be resident?
                    int nclu = 0;
                    cudaError_t e = cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         GS_SMALL_XS_BYTES_MAX);
                    cudaLaunchConfig_t probe = cfg;
                    probe.dynamicSmemBytes = GS_SMALL_XS_BYTES_MAX;
                    if (pipe != 4) probe.blockDim = dim3(GS_SMALL_CLUSTER_THREADS_MAX);
                    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&nclu, (const void *)kern, &probe);
                    if (e != cudaSuccess) cudaGetLastError();
                    if (dev_ok) cluster_ok[dev] = (e == cudaSuccess && nclu >= 1) ? 1 : -1;
                    if (e != cudaSuccess || nclu < 1) cfg.numAttrs = 0; // -> the one-CTA version below
                }
                if (cfg.numAttrs == 1) {
                    GSB_CUDA(cudaLaunchKernelEx(&cfg, kern, rp, ci, va, dg, b, x, ld, ctl, partials, a));
                    launched = true;
                    if (slots) *slots = CS;
                }
            }
            if (launched) {
                GSB_KERNEL_CHECK();
                return GSB_OK;
            }
            if (pipe >= 2 && xs_bytes <= GS_SMALL_XS_BYTES_MAX) {
                auto kern = gs_small_one_cta<NRHS, T, true, 1>;
                static bool attr_set[64] = {false}; // per device
                if (!dev_ok || !attr_set[dev]) {
                    GSB_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  GS_SMALL_XS_BYTES_MAX));
                    if (dev_ok) attr_set[dev] = true;
                }
                kern<<<1, T, xs_bytes, st>>>(rp, ci, va, dg, b, x, ld, ctl, partials, a);
            } else {
                gs_small_one_cta<NRHS, T, false, 1><<<1, T, 0, st>>>(rp, ci, va, dg, b, x, ld, ctl, partials, a);
            }
        } else {
            gs_small_persistent<NRHS, 1024><<<1, 1024, 0, st>>>(rp, ci, va, dg, b, x, ld, ctl, partials, a);
        }
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
