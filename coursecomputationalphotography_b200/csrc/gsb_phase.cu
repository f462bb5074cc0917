// gsb_phase.cu -- the Gauss-Seidel colour-phase kernels and their launch plan.
//
// A colour phase updates the contiguous row range [row0,row1) of a colour-major CSR; rows of one
// colour never read each other, so every row is independent inside the phase:
//     x_i <- (b_i - sum_{j != i} a_ij x_j) / a_ii        (v2 :359-374, zero diagonal -> row skipped)
// Products and sums are rounded separately, in storage order (no FMA), as the reference's build does.
//
// Two kernels:
//   gs_phase_direct  row per thread, CSR read straight from global memory.  Simple; latency bound
//                    (three dependent load levels, 40-byte-stride accesses to values/columns).
//   gs_phase_staged  one CTA per tile of <= 256 rows.  The tile's slice of `values` and `columns` is
//                    one contiguous span of the CSR arrays, so an elected thread fetches it with two
//                    1-D bulk async copies (cp.async.bulk, the TMA engine, completion on an mbarrier)
//                    into shared memory while all threads issue their coalesced row-pointer / b / x_old
//                    loads.  Rows are then walked out of shared memory and only the x gathers touch
//                    global memory; for short rows all gathers of a row are issued before the first is
//                    consumed.  HBM sees long contiguous bursts instead of 32-byte sectors.
// The stop rule's L1 update norm is accumulated per block into `partials` (fixed order, no atomics).
//
// Solver format: the colour-major CSR holds the OFF-DIAGONAL entries only (rp/ci/va) and the diagonal lives in
// its own array dg[row] (0 when the row has no diagonal -> the row is skipped).  The row body then has no
// per-entry "is this the diagonal" test, and a 5-point row has at most 4 entries.
#include "gsb_ring.cuh"

#include <stdlib.h>

#include <mutex>

// ---------------------------------------------------------------------------------------------
// kernel 1: row per thread, direct global loads
// ---------------------------------------------------------------------------------------------
template <int NRHS, bool CHECK>
__global__ void __launch_bounds__(GS_THREADS)
    gs_phase_direct(const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                    const double *__restrict__ dg, const double *__restrict__ b, double *x, int64_t n, int row0,
                    int row1, const GsCtl *__restrict__ ctl, double *__restrict__ partials) {
    if (*(volatile const int *)&ctl->done) return;
    const int i = row0 + blockIdx.x * GS_THREADS + threadIdx.x;
    double diff[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) diff[r] = 0.0;
    if (i < row1) {
        const int k0 = rp[i], k1 = rp[i + 1];
        const double d = dg[i];
        double sig[NRHS];
        gs_row_sigma<NRHS>(ci + k0, va + k0, k1 - k0, [&](int c, int r) { return x[r * n + c]; }, sig);
        if (d != 0.0) { // zero or absent diagonal: row skipped, x_i unchanged (v2 :360-363)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                const double xn = __ddiv_rn(__dsub_rn(b[r * n + i], sig[r]), d);
                if (CHECK) diff[r] = fabs(xn - x[r * n + i]);
                x[r * n + i] = xn;
            }
        }
    }
    if (CHECK) gsb_block_reduce_store<NRHS, GS_THREADS>(diff, partials + (size_t)blockIdx.x * NRHS);
}

// ---------------------------------------------------------------------------------------------
// kernel 2: bulk-copy staged CSR tiles
// ---------------------------------------------------------------------------------------------
// HINT: L2 residency control for matrices whose x (all right-hand sides) fits in L2 while the matrix does not
// (BASELINE configs[4]: n = 1e7 -> x is 80 MB, the CSR 3.2 GB per sweep).  With random columns every gather touches its
// own 32-byte sector, and when the matrix stream has pushed x out of L2 each of them goes to DRAM -- several times the
// algorithmic traffic.  The streamed arrays (values, columns, row offsets, diagonal, b) are fetched evict-first and
// the x lines (gathers, the row's own old value, the store) evict-last, so that x stays resident across the sweep.
__device__ __forceinline__ double ld_f64_l2hint(const double *p, uint64_t policy) {
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy) : "memory");
    return v;
}
__device__ __forceinline__ void st_f64_l2hint(double *p, double v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}

template <int NRHS, bool CHECK, bool HINT>
__global__ void __launch_bounds__(GS_THREADS, 4)
    gs_phase_staged(const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                    const double *__restrict__ dg, const double *__restrict__ b, double *x, int64_t n, int row0,
                    int row1, int tile_rows, const int *__restrict__ tile_k, int cap, const GsCtl *__restrict__ ctl,
                    double *__restrict__ partials) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);
    double *va_s = reinterpret_cast<double *>(smem_raw + 16);
    int *ci_s = reinterpret_cast<int *>(smem_raw + 16 + (size_t)cap * 8);
    if (*(volatile const int *)&ctl->done) return;

    const int t = blockIdx.x;
    const int r_begin = row0 + t * tile_rows;
    const int r_end = min(r_begin + tile_rows, row1);
    const int k0 = tile_k[t], k1 = tile_k[t + 1];
    const int kv0 = k0 & ~1, kc0 = k0 & ~3; // 16-byte aligned starts of the two spans
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes_v = (uint32_t)(((k1 + 1) & ~1) - kv0) * 8u;
        const uint32_t bytes_c = (uint32_t)(((k1 + 3) & ~3) - kc0) * 4u;
        mbar_expect_tx(mbar, bytes_v + bytes_c);
        if (HINT) {
            const uint64_t pol_stream = l2_policy_evict_first();
            if (bytes_v) bulk_g2s_hint(va_s, va + kv0, bytes_v, mbar, pol_stream);
            if (bytes_c) bulk_g2s_hint(ci_s, ci + kc0, bytes_c, mbar, pol_stream);
        } else {
            if (bytes_v) bulk_g2s(va_s, va + kv0, bytes_v, mbar);
            if (bytes_c) bulk_g2s(ci_s, ci + kc0, bytes_c, mbar);
        }
    }
    // coalesced per-row loads overlap the bulk copies
    const int i = r_begin + threadIdx.x;
    const bool valid = threadIdx.x < tile_rows && i < r_end;
    const uint64_t pol_keep = HINT ? l2_policy_evict_last() : 0;
    int rs = 0, re = 0;
    double d = 0.0, bb[NRHS], xo[NRHS], diff[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) bb[r] = xo[r] = diff[r] = 0.0;
    if (valid) {
        rs = HINT ? __ldcs(rp + i) : rp[i];
        re = HINT ? __ldcs(rp + i + 1) : rp[i + 1];
        d = HINT ? __ldcs(dg + i) : dg[i];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            bb[r] = HINT ? __ldcs(b + r * n + i) : b[r * n + i];
            if (CHECK) xo[r] = HINT ? ld_f64_l2hint(x + r * n + i, pol_keep) : x[r * n + i];
        }
    }
    mbar_wait(mbar, 0);

    if (valid) {
        double sig[NRHS];
        if (HINT)
            gs_row_sigma<NRHS>(ci_s + (rs - kc0), va_s + (rs - kv0), re - rs,
                               [&](int c, int r) { return ld_f64_l2hint(x + r * n + c, pol_keep); }, sig);
        else
            gs_row_sigma<NRHS>(ci_s + (rs - kc0), va_s + (rs - kv0), re - rs, [&](int c, int r) { return x[r * n + c]; }, sig);
        if (d != 0.0) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                const double xn = __ddiv_rn(__dsub_rn(bb[r], sig[r]), d);
                if (CHECK) diff[r] = fabs(xn - xo[r]);
                if (HINT)
                    st_f64_l2hint(x + r * n + i, xn, pol_keep);
                else
                    x[r * n + i] = xn;
            }
        }
    }
    if (CHECK) gsb_block_reduce_store<NRHS, GS_THREADS>(diff, partials + (size_t)blockIdx.x * NRHS);
}

// ---------------------------------------------------------------------------------------------
// kernels 3 and 4: persistent CTAs, ring of bulk-copy stages
//
// gs_phase_staged still exposes one bulk-copy latency plus one gather latency per CTA lifetime.  Here
// a CTA stays resident, walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... and keeps STAGES tiles
// in flight: while tile k is computed, the bulk copies of tile k+1 (values, columns, row offsets, and
// the b / x_old planes -- every per-tile input is a contiguous span) are already landing in the next
// stage.
//   kernel 3 (WIN = false): the x gathers still go to global memory (read-only path).
//   kernel 4 (WIN = true):  the columns a tile gathers from lie in a few contiguous windows of x (for a
//       5-point grid in colour-major order: the neighbour colour's rows above, beside and below the
//       tile).  The plan records up to 4 windows per tile (64-column granules); they are bulk-copied
//       into the stage too, so the whole phase runs out of shared memory and HBM only sees TMA bursts
//       plus the coalesced x stores.  Tiles whose gathers do not fit windows fall back to global gathers.
// ---------------------------------------------------------------------------------------------
// The halo CTAs' synchronisation with the neighbour GPUs.  Kept out of line on purpose: with the spin loop, the
// system-scope fences and the atomic inlined, ptxas stops using the uniform datapath for the whole kernel and the
// interior tiles run 8 % slower (measured); as calls on a cold path they cost the interior nothing.
#define GS_HALO_ABORT_EPOCH 0x7fffffff // a rank that gave up raises its neighbours' flags to this: nobody waits on it
__device__ __noinline__ void halo_wait_flags(const int *f0, const int *f1, int epoch, GsCtl *ctl, int *peer_flag0,
                                             int *peer_flag1) {
    const int *f[2] = {f0, f1};
    bool gave_up = false;
#pragma unroll
    for (int pr = 0; pr < 2; ++pr)
        if (f[pr]) {
            int v;
            long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f[pr]) : "memory");
                if (v >= epoch) break;
                // bounded like the stop-rule exchange (~10 s): a neighbour that crashed or never launched must not
                // wedge this GPU; also leave as soon as another CTA of this rank has given up
                if (++spins > (1ll << 25) || ((spins & 1023) == 0 && *(volatile int *)&ctl->error)) {
                    gave_up = true;
                    break;
                }
                if (spins > 64) __nanosleep(100);
            }
        }
    if (gave_up) {
        ctl->error = 1;
        ctl->done = 1;
        // publish the abort: neighbours waiting on this rank's flags fall through instead of spinning forever (their
        // own stop-rule exchange then times out against this rank and reports the error)
        if (peer_flag0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flag0), "r"(GS_HALO_ABORT_EPOCH) : "memory");
        if (peer_flag1) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peer_flag1), "r"(GS_HALO_ABORT_EPOCH) : "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __noinline__ void halo_fence_system() { __threadfence_system(); }
// one halo tile of the phase is complete; the last one raises the neighbours' flags (release at system scope
// orders this GPU's peer stores before the flag)
__device__ __noinline__ void halo_tile_done(int *counter, int n_halo_tiles, int *flag0, int *flag1, int epoch) {
    const int done_tiles = atomicAdd(counter, 1) + 1;
    if (done_tiles != n_halo_tiles) return;
    *counter = 0;
    __threadfence_system();
    if (flag0) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag0), "r"(epoch) : "memory");
    if (flag1) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag1), "r"(epoch) : "memory");
}

// End of a sweep, ONE implementation for every path: the separate kernels (gs_end_sweep, gs_end_sweep_peer) and the
// fused end (GsbEndArgs: the last CTA of the sweep's last colour phase to retire).  Always run by GS_THREADS threads
// of one CTA, so the fold order -- and with it the stop norm, bit for bit -- does not depend on the path.
//   mode 0: fold the partials, (strip solver: exchange the sums with the peers,) bump the counter, decide
//   mode 1: fold into ctl->eps_local only (strip solver on the ncclAllReduce path, before the all-reduce)
//   mode 2: bump the counter and decide from ctl->eps_last (after the all-reduce)
// Peer exchange (GsbEpsExchange): this rank's k sums go into slot `rank` of every rank's box (NVLink peer stores)
// with the slot's flag raised to `epoch` (release, system scope); the rank then waits for the `world` flags of its
// own box (bounded: a missing rank raises ctl->error instead of hanging the GPU) and adds the slots in rank order --
// the same bits, hence the same decision, on every rank.
// Out of line: it runs once per sweep and must not weigh on the tile loop's register allocation.
template <int NRHS>
__device__ __noinline__ void gs_end_of_sweep_body(GsCtl *ctl, const double *partials, int n_partials, int checked,
                                                  int mode, int exchange, const GsbEpsExchange ex) {
    __shared__ double ws[NRHS][GS_THREADS / 32];
    __shared__ double tot[NRHS];
    __shared__ double all[GSB_DIST_MAX_WORLD][NRHS];
    __shared__ int timed_out;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (mode == 2 || !checked) {
        if (tid == 0) {
            bool all_ok = checked != 0;
            if (checked)
                for (int r = 0; r < NRHS; ++r)
                    if (ctl->eps_last[r] > ctl->epsilon) all_ok = false;
            const int cnt = ctl->sweeps + 1;
            ctl->sweeps = cnt;
            if (all_ok || cnt >= ctl->max_iter) ctl->done = 1;
        }
        return;
    }
    double s[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) s[r] = 0.0;
    for (int i = tid; i < n_partials; i += GS_THREADS) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) s[r] += __ldcg(partials + (size_t)i * NRHS + r); // L2: written by other CTAs
    }
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
        double t = s[r];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
        if (lane == 0) ws[r][wid] = t;
    }
    if (tid == 0) timed_out = 0;
    __syncthreads();
    if (tid < NRHS) {
        double t = 0.0;
        for (int w = 0; w < GS_THREADS / 32; ++w) t += ws[tid][w];
        tot[tid] = t;
    }
    __syncthreads();
    if (mode == 1) { // this rank's share; the all-reduce writes eps_last (idempotent after the stop)
        if (tid < NRHS) ctl->eps_local[tid] = tot[tid];
        return;
    }
    if (exchange) {
        const int par = ex.epoch & 1;
        const size_t flag_off = (size_t)2 * ex.world * GSB_MAX_RHS; // in doubles
        if (tid < ex.world) { // push this rank's sums into slot `rank` of rank q's box
            const int q = tid;
            double *dst = ex.box[q] + (size_t)(par * ex.world + ex.rank) * GSB_MAX_RHS;
#pragma unroll
            for (int r = 0; r < NRHS; ++r)
                asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + r), "d"(tot[r]) : "memory");
            int *flag = reinterpret_cast<int *>(ex.box[q] + flag_off) + par * ex.world + ex.rank;
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(ex.epoch) : "memory");
        }
        if (tid < ex.world) { // collect slot q of the own box
            const int q = tid;
            const int *flag = reinterpret_cast<const int *>(ex.box[ex.rank] + flag_off) + par * ex.world + q;
            int v = 0;
            long long spins = 0;
            for (;;) {
                asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (v >= ex.epoch) break;
                if (++spins > (1ll << 25)) { // ~10 s: a rank is missing
                    timed_out = 1;
                    break;
                }
                __nanosleep(spins < 64 ? 20 : 200);
            }
            const double *src = ex.box[ex.rank] + (size_t)(par * ex.world + q) * GSB_MAX_RHS;
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                double t;
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(t) : "l"(src + r) : "memory");
                all[q][r] = t;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        bool all_ok = true;
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            double t = tot[r];
            if (exchange) {
                t = 0.0;
                for (int q = 0; q < ex.world; ++q) t += all[q][r];
            }
            ctl->eps_last[r] = t;
            if (t > ctl->epsilon) all_ok = false; // v2 :356: the loop continues while eps > epsilon
        }
        const int cnt = ctl->sweeps + 1;
        ctl->sweeps = cnt;
        if (timed_out) ctl->error = 1;
        if (all_ok || cnt >= ctl->max_iter || timed_out) ctl->done = 1;
    }
}

// resident CTAs per SM the register allocation must allow: 4 wherever shared memory lets 4 stages-pairs fit
// (k = 3 with windows is limited to 2-3 by its 49 KB stages)
//
// HALO (strip solver): the grid is [n_halo_tiles one-tile CTAs | the persistent ring over the interior tiles].
// CTA b < n_halo_tiles takes the b-th halo tile -- a tile that reads ghost unknowns and/or owns rows a neighbour
// GPU reads -- and runs the same code as a ring of exactly one tile, plus the flag wait, the peer stores and the
// completion count.  The other CTAs run the ring over the interior tiles [interior_base, interior_base + n_interior),
// numbered arithmetically, so the interior path is instruction-for-instruction the single-GPU kernel (a table
// look-up per tile, or halo code in the tile body, costs 8-13 % -- measured, profiles/README.md).
template <int NRHS, bool CHECK, int STAGES, bool WIN, bool HALO, bool FEND>
__global__ void __launch_bounds__(GS_THREADS, (NRHS >= 3 && WIN ? 3 : 4))
    gs_phase_ring(const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                  const double *__restrict__ dg, const double *__restrict__ b, double *x, int64_t n, int row0_in,
                  int row1_in, int ntiles_in,
                  const int *__restrict__ tile_k_in, const int *__restrict__ tile_win_in, int cap, int wcap,
                  const GsCtl *__restrict__ ctl, double *__restrict__ partials, const GsbHaloArgs halo,
                  const GsbEndArgs end) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RingLayout L = ring_layout(cap, NRHS, CHECK, WIN ? wcap : 0);
    // kernel 3 with window descriptors: the x spans a tile will gather from can be pulled into L2 when the tile's
    // stage is issued (one tile ahead); off by default (GSB_X_PREFETCH)
    const bool xpf = !WIN && tile_win_in != nullptr;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned char *stage0 = smem_raw + 64;
    const int tid = threadIdx.x;

    // the slice of the colour phase this CTA walks: tiles t = bid, bid + gsz, ... < ntiles of rows [row0, row1)
    int bid = blockIdx.x, gsz = gridDim.x, row0 = row0_in, row1 = row1_in, ntiles = ntiles_in;
    const int *tile_k = tile_k_in, *tile_win = tile_win_in;
    int hinfo = 0; // halo CTA: bit0 the tile reads ghosts, bit1 it owns rows a neighbour reads
    if (HALO) {
        const int nh = halo.n_halo_tiles;
        int first, count;
        if (bid < nh) {
            // halo tiles are the prefix [0, interior_base) and the suffix [interior_base + n_interior, tiles) of the
            // colour's tiles; every one of them is treated as "reads ghosts and pushes" (a tile that does only one
            // of the two waits on a flag that is raised anyway / finds no row to push) -- no table, all uniform
            first = bid < halo.interior_base ? bid : bid + halo.n_interior;
            count = 1;
            hinfo = 3;
            bid = 0;
            gsz = 1;
        } else {
            first = halo.interior_base;
            count = halo.n_interior;
            bid -= nh;
            gsz -= nh;
        }
        row0 = row0_in + first * GS_THREADS;
        row1 = min(row1_in, row0 + count * GS_THREADS);
        ntiles = count;
        tile_k = tile_k_in + first;
        if (tile_win_in) tile_win = tile_win_in + (size_t)first * GS_WIN_DESC;
    }

    // tile descriptor: fetched by thread 0 one iteration before it is needed, so that its L2 latency hides
    // behind the compute of the current tile
    struct TileDesc {
        int k0, k1;
        int4 w0, w1, w2; // {nwin, lo0, lo1, lo2}, {lo3, len0, len1, len2}, {len3, -, -, -}
    };
    auto load_desc = [&](int t) -> TileDesc {
        TileDesc d;
        d.k0 = tile_k[t];
        d.k1 = tile_k[t + 1];
        if (WIN || xpf) {
            const int4 *wd = reinterpret_cast<const int4 *>(tile_win + (size_t)t * GS_WIN_DESC);
            d.w0 = wd[0];
            d.w1 = wd[1];
            d.w2 = wd[2];
        } else {
            d.w0 = d.w1 = d.w2 = make_int4(0, 0, 0, 0);
        }
        return d;
    };
    auto issue = [&](const TileDesc &td, int t, int s) { // thread 0 only
        unsigned char *st = stage0 + (size_t)s * L.stage_bytes;
        int *hdr = reinterpret_cast<int *>(st + L.hdr_off);
        const int r_begin = row0 + t * GS_THREADS;
        const int rows = min(GS_THREADS, row1 - r_begin);
        const int k0 = td.k0, k1 = td.k1;
        const int kv0 = k0 & ~1, kc0 = k0 & ~3;
        const uint32_t bytes_v = (uint32_t)(((k1 + 1) & ~1) - kv0) * 8u;
        const uint32_t bytes_c = (uint32_t)(((k1 + 3) & ~3) - kc0) * 4u;
        const int ra = r_begin & ~3;
        const uint32_t bytes_r = (uint32_t)(((r_begin + rows + 1 + 3) & ~3) - ra) * 4u;
        const int ea = r_begin & ~1; // n (the leading dimension) is even: the same alignment for every plane
        const uint32_t bytes_p = (uint32_t)(((r_begin + rows + 1) & ~1) - ea) * 8u;
        uint32_t total = bytes_v + bytes_c + bytes_r + bytes_p + (uint32_t)NRHS * bytes_p * (CHECK ? 2u : 1u);
        int nwin = 0, lo[GS_WIN_MAX], len[GS_WIN_MAX];
        if (WIN || xpf) {
            nwin = td.w0.x;
            lo[0] = td.w0.y; lo[1] = td.w0.z; lo[2] = td.w0.w; lo[3] = td.w1.x;
            len[0] = td.w1.y; len[1] = td.w1.z; len[2] = td.w1.w; len[3] = td.w2.x;
        }
        if (WIN) {
#pragma unroll
            for (int w = 0; w < GS_WIN_MAX; ++w) total += (uint32_t)NRHS * (uint32_t)len[w] * 8u;
        }
        hdr[0] = k0;
        hdr[1] = nwin;
#pragma unroll
        for (int w = 0; w < GS_WIN_MAX; ++w) {
            hdr[2 + w] = WIN ? lo[w] : 0;
            hdr[2 + GS_WIN_MAX + w] = WIN ? len[w] : 0;
        }
        mbar_expect_tx(&full[s], total);
        if (bytes_v) bulk_g2s(st + L.va_off, va + kv0, bytes_v, &full[s]);
        if (bytes_c) bulk_g2s(st + L.ci_off, ci + kc0, bytes_c, &full[s]);
        bulk_g2s(st + L.rp_off, rp + ra, bytes_r, &full[s]);
        bulk_g2s(st + L.dg_off, dg + ea, bytes_p, &full[s]);
#pragma unroll
        for (int r = 0; r < NRHS; ++r) {
            bulk_g2s(st + L.b_off + r * L.plane * 8, b + r * n + ea, bytes_p, &full[s]);
            if (CHECK) bulk_g2s(st + L.xo_off + r * L.plane * 8, x + r * n + ea, bytes_p, &full[s]);
        }
        if (WIN) {
            int base = 0;
#pragma unroll
            for (int w = 0; w < GS_WIN_MAX; ++w) {
                if (len[w] > 0) {
#pragma unroll
                    for (int r = 0; r < NRHS; ++r)
                        bulk_g2s(st + L.xw_off + ((size_t)r * wcap + base) * 8, x + r * n + lo[w], (uint32_t)len[w] * 8u,
                                 &full[s]);
                }
                base += len[w];
            }
        } else if (xpf) {
            // wcap carries the prefetch mode for kernel 3: 1 = every window, 2 = only the last (highest) window --
            // with tiles walked in ascending order the lower windows were requested by earlier tiles
#pragma unroll
            for (int w = 0; w < GS_WIN_MAX; ++w)
                if (len[w] > 0 && (wcap != 2 || w == nwin - 1)) {
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) bulk_prefetch_l2(x + r * n + lo[w], (uint32_t)len[w] * 8u);
                }
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    }
    __syncthreads();
    // Programmatic dependent launch: the next kernel of the stream may take the SM slots this grid frees as its
    // CTAs retire, and this grid may itself have started while the previous colour phase was still draining.
    // Everything the prologue stages is either constant (CSR, diagonal, b) or this colour's own x (x_old), which
    // the predecessor does not write (and the kernel that did write it has completed, see below); the
    // predecessor's x (the gathers, the window copies) and the control block are only touched after pdl_wait().
    // The dependents are released only AFTER this kernel's own wait: the kernel after this one may then overlap
    // this one, but never the one before it (whose x -- this colour's x_old two phases later -- it stages early).
    // prologue before the predecessor has finished -- not for halo CTAs, which must know that the solve is still
    // running before they wait for a neighbour (after the stop decision no flag is raised any more)
    const bool early = !WIN && halo.pdl_early && !(HALO && hinfo);
    if (!early) {
        pdl_wait();
        pdl_launch_dependents();
        if (*(volatile const int *)&ctl->done) return; // nothing staged yet
    }
    // halo CTA: its tile reads ghost unknowns -- the neighbours' values of the other colour must have landed before
    // anything of the tile is staged or gathered.  (The halo calls sit before and after the tile loop: a call
    // inside it makes ptxas give up the uniform datapath for the whole loop.)
    if (HALO && hinfo && tid == 0 && halo.wait_epoch > 0)
        halo_wait_flags(halo.has_peer[0] ? halo.wait_flag[0] : nullptr, halo.has_peer[1] ? halo.wait_flag[1] : nullptr,
                        halo.wait_epoch, const_cast<GsCtl *>(ctl), halo.has_peer[0] ? halo.peer_flag[0] : nullptr,
                        halo.has_peer[1] ? halo.peer_flag[1] : nullptr);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const int t = bid + s * gsz;
            if (t < ntiles) issue(load_desc(t), t, s);
        }
    }
    if (early) {
        pdl_wait();
        pdl_launch_dependents();
        if (*(volatile const int *)&ctl->done) { // written only by gs_end_sweep, i.e. constant from here on
            // the prologue's bulk copies must land before the shared memory is released
#pragma unroll
            for (int s = 0; s < STAGES; ++s)
                if (bid + s * gsz < ntiles) mbar_wait(&full[s], 0);
            return;
        }
    }

    // stop-rule partial: accumulated per thread over the CTA's tiles (static schedule -> fixed order) and folded
    // once at the end into partial slot blockIdx.x; the phase's other slots are zeroed
    double acc[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
    int k = 0;
    for (int t = bid; t < ntiles; t += gsz, ++k) {
        const int s = k % STAGES;
        const uint32_t parity = (uint32_t)(k / STAGES) & 1u;
        unsigned char *st = stage0 + (size_t)s * L.stage_bytes;
        const int tn = t + STAGES * gsz; // the tile that will reuse this stage
        TileDesc next_desc;
        if (tid == 0 && tn < ntiles) next_desc = load_desc(tn);
        const int r_begin = row0 + t * GS_THREADS;
        const int rows = min(GS_THREADS, row1 - r_begin);
        mbar_wait(&full[s], parity);

        const int *hdr = reinterpret_cast<const int *>(st + L.hdr_off);
        const int k0 = hdr[0];
        const double *va_s = reinterpret_cast<const double *>(st + L.va_off) - (k0 & ~1);
        const int *ci_s = reinterpret_cast<const int *>(st + L.ci_off) - (k0 & ~3);
        const int *rp_s = reinterpret_cast<const int *>(st + L.rp_off) + (r_begin & 3);
        const double *xw_s = reinterpret_cast<const double *>(st + L.xw_off);
        // kernel 4: in a window tile the staged index array holds shared-memory slots precomputed by
        // plan_tile_slots; otherwise it holds column numbers and the gathers go to global memory
        const bool use_win = WIN && hdr[1] > 0;
        const int i = r_begin + tid;
        if (tid < rows) {
            const int rs = rp_s[tid], len = rp_s[tid + 1] - rs;
            const int po = (r_begin & 1) + tid;
            const double d = reinterpret_cast<const double *>(st + L.dg_off)[po];
            double sig[NRHS];
            if (use_win)
                gs_row_sigma<NRHS>(ci_s + rs, va_s + rs, len, [&](int c, int r) { return xw_s[(size_t)r * wcap + c]; }, sig);
            else
                gs_row_sigma<NRHS>(ci_s + rs, va_s + rs, len, [&](int c, int r) { return x[r * n + c]; }, sig);
            if (d != 0.0) { // zero or absent diagonal: row skipped, x_i unchanged (v2 :360-363)
                double xn[NRHS];
#pragma unroll
                for (int r = 0; r < NRHS; ++r) { // the k divisions are independent: keep them in one straight-line block
                    const double bb = reinterpret_cast<const double *>(st + L.b_off)[r * L.plane + po];
                    xn[r] = __ddiv_rn(__dsub_rn(bb, sig[r]), d);
                    if (CHECK) acc[r] += fabs(xn[r] - reinterpret_cast<const double *>(st + L.xo_off)[r * L.plane + po]);
                    x[r * n + i] = xn[r];
                }
                if (HALO && (hinfo & 2)) { // a neighbour GPU reads rows of this tile: store them into its ghost slots too
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr)
                        if (halo.has_peer[pr]) {
                            const int slot = halo.push_map[pr][i];
                            if (slot >= 0) {
#pragma unroll
                                for (int r = 0; r < NRHS; ++r)
                                    halo.peer_x[pr][r * halo.peer_ld[pr] + halo.peer_gs[pr] + slot] = xn[r];
                            }
                        }
                }
            }
        }
        __syncthreads(); // every thread is done with stage s
        if (tid == 0 && tn < ntiles) issue(next_desc, tn, s);
    }
    if (HALO && hinfo) {
        // halo CTA, its one tile done: peer stores visible (system scope) before the tile is counted; the last
        // halo tile of the phase raises the neighbours' flags
        halo_fence_system();
        __syncthreads();
        if (tid == 0)
            halo_tile_done(halo.counter, halo.n_halo_tiles, halo.has_peer[0] ? halo.peer_flag[0] : nullptr,
                           halo.has_peer[1] ? halo.peer_flag[1] : nullptr, halo.signal_epoch);
    }
    if (CHECK) {
        gsb_block_reduce_store<NRHS, GS_THREADS>(acc, partials + (size_t)blockIdx.x * NRHS);
        // the phase owns min(tiles, GS_RING_SLOTS_MAX) partial slots (gsb_plan_partial_slots); unused ones are zero
        const int nslots = min(ntiles_in, GS_RING_SLOTS_MAX);
        for (int t2 = blockIdx.x + gridDim.x; t2 < nslots; t2 += gridDim.x)
            if (tid < NRHS) partials[(size_t)t2 * NRHS + tid] = 0.0;
    }
    if (FEND && end.enabled) {
        // fused end of sweep (its own instantiation: the tail's call and by-value arguments change the register
        // allocation of the tile loop, so the default path is compiled without it).  This CTA's partials are
        // visible device-wide before it takes its ticket; the CTA that draws the last ticket sees every other
        // CTA's partials (fence + atomic on both sides) and ends the sweep
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            s_last = atomicAdd(&end.ctl->ticket, 1) == (int)gridDim.x - 1;
            __threadfence();
        }
        __syncthreads();
        if (s_last) {
            if (tid == 0) end.ctl->ticket = 0;
            gs_end_of_sweep_body<NRHS>(end.ctl, end.partials, end.n_partials, end.checked, 0, end.exchange, end.ex);
        }
    }
}

// Per-tile gather windows (kernel 4).  One CTA per tile: the 64-column granules the tile's off-diagonal
// columns touch are collected in a small hash set, sorted, and runs of consecutive granules become windows.
// desc = {nwin, lo[4], len[4]}; nwin = 0 marks a tile that keeps global gathers.  stats[0] = max total
// window length, stats[1] = tiles without windows.
__global__ void __launch_bounds__(GS_THREADS) plan_tile_windows(const int *__restrict__ rp, const int *__restrict__ ci,
                                                                int row0, int row1, int *__restrict__ desc,
                                                                int *__restrict__ stats) {
    constexpr int TABLE = 512, MAXG = 40;
    __shared__ int table[TABLE];
    __shared__ int overflow;
    const int t = blockIdx.x;
    for (int q = threadIdx.x; q < TABLE; q += GS_THREADS) table[q] = -1;
    if (threadIdx.x == 0) overflow = 0;
    __syncthreads();
    const int i = row0 + t * GS_THREADS + threadIdx.x;
    if (i < row1) {
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int c = ci[k];
            const int g = c / GS_WIN_GRANULE;
            unsigned h = ((unsigned)g * 2654435761u) % TABLE;
            for (int probe = 0; probe < TABLE; ++probe) {
                int old = atomicCAS(&table[h], -1, g);
                if (old == -1 || old == g) break;
                h = (h + 1) % TABLE;
                if (probe == TABLE - 1) overflow = 1;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int *d = desc + (size_t)t * GS_WIN_DESC;
        for (int q = 0; q < GS_WIN_DESC; ++q) d[q] = 0;
        int g[MAXG], cnt = 0;
        bool ok = !overflow;
        for (int q = 0; q < TABLE && ok; ++q)
            if (table[q] >= 0) {
                if (cnt == MAXG) {
                    ok = false;
                    break;
                }
                int v = table[q], p = cnt++;
                while (p > 0 && g[p - 1] > v) {
                    g[p] = g[p - 1];
                    --p;
                }
                g[p] = v;
            }
        int nwin = 0, total = 0;
        if (ok && cnt > 0) {
            int lo = g[0], prev = g[0];
            for (int q = 1; q <= cnt && ok; ++q) {
                if (q < cnt && g[q] == prev + 1) {
                    prev = g[q];
                    continue;
                }
                if (nwin == GS_WIN_MAX) {
                    ok = false;
                    break;
                }
                d[1 + nwin] = lo * GS_WIN_GRANULE;
                d[1 + GS_WIN_MAX + nwin] = (prev - lo + 1) * GS_WIN_GRANULE;
                total += (prev - lo + 1) * GS_WIN_GRANULE;
                ++nwin;
                if (q < cnt) lo = prev = g[q];
            }
        }
        if (!ok || total > GS_WIN_CAP_MAX) {
            for (int q = 0; q < GS_WIN_DESC; ++q) d[q] = 0;
            nwin = 0;
            total = 0;
            if (cnt > 0 || !ok) atomicAdd(&stats[1], 1);
        }
        d[0] = nwin;
        atomicMax(&stats[0], total);
    }
}

// Index array of kernel 4: for a window tile, the shared-memory slot of every (off-diagonal) entry; for a
// fallback tile, the column number unchanged.
__global__ void __launch_bounds__(GS_THREADS) plan_tile_slots(const int *__restrict__ rp, const int *__restrict__ ci,
                                                              int row0, int row1, const int *__restrict__ desc,
                                                              int *__restrict__ ci_slot) {
    const int t = blockIdx.x;
    const int *d = desc + (size_t)t * GS_WIN_DESC;
    const int nwin = d[0];
    int lo[GS_WIN_MAX], len[GS_WIN_MAX];
#pragma unroll
    for (int w = 0; w < GS_WIN_MAX; ++w) {
        lo[w] = d[1 + w];
        len[w] = d[1 + GS_WIN_MAX + w];
    }
    const int i = row0 + t * GS_THREADS + threadIdx.x;
    if (i >= row1) return;
    for (int k = rp[i]; k < rp[i + 1]; ++k) {
        const int c = ci[k];
        int out = c;
        if (nwin > 0) {
            int base = 0;
#pragma unroll
            for (int w = 0; w < GS_WIN_MAX; ++w) {
                const unsigned dd = (unsigned)(c - lo[w]);
                if (dd < (unsigned)len[w]) out = base + (int)dd;
                base += len[w];
            }
        }
        ci_slot[k] = out;
    }
}

// ---------------------------------------------------------------------------------------------
// end of sweep: fold the partials (fixed order), update the control block
//   mode 0: fold, bump the counter, decide (single GPU)
//   mode 1: fold into ctl->eps_last only (strip solver, before the all-reduce)
//   mode 2: bump the counter and decide from ctl->eps_last (after the all-reduce)
// ---------------------------------------------------------------------------------------------
template <int NRHS>
__global__ void __launch_bounds__(GS_THREADS) gs_end_sweep(GsCtl *ctl, const double *__restrict__ partials, int n_partials,
                                                           int checked, int mode) {
    // the predecessor (the sweep's last colour phase, or the all-reduce) must be complete before anything is read;
    // the dependents are released only after that, so that the next sweep's first kernel can overlap this one but
    // never the colour phase before it
    pdl_wait();
    pdl_launch_dependents();
    if (*(volatile const int *)&ctl->done) return;
    GsbEpsExchange none;
    none.world = none.rank = none.epoch = 0;
    gs_end_of_sweep_body<NRHS>(ctl, partials, n_partials, checked, mode, 0, none);
}

// end of a checked sweep of the strip solver, stop-rule all-reduce fused in (see GsbEpsExchange)
template <int NRHS>
__global__ void __launch_bounds__(GS_THREADS) gs_end_sweep_peer(GsCtl *ctl, const double *partials, int n_partials,
                                                                const GsbEpsExchange ex) {
    pdl_wait();
    pdl_launch_dependents();
    if (*(volatile const int *)&ctl->done) return;
    gs_end_of_sweep_body<NRHS>(ctl, partials, n_partials, 1, 0, 1, ex);
}

int gsb_launch_end_sweep_peer(GsCtl *ctl, const double *partials, int n_partials, int nrhs, const GsbEpsExchange *ex,
                              cudaStream_t st) {
    if (n_partials > 8192) {
        gsb_set_error("end_sweep_peer: %d partials (the ring kernels write at most %d per colour)", n_partials,
                      GS_RING_SLOTS_MAX);
        return GSB_ERR_ARG;
    }
    void (*kern)(GsCtl *, const double *, int, const GsbEpsExchange) = nullptr;
    switch (nrhs) {
        case 1: kern = gs_end_sweep_peer<1>; break;
        case 2: kern = gs_end_sweep_peer<2>; break;
        case 3: kern = gs_end_sweep_peer<3>; break;
        case 4: kern = gs_end_sweep_peer<4>; break;
        default: return GSB_ERR_ARG;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(GS_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gsb_pdl_enabled() ? 1 : 0;
    GSB_CUDA(cudaLaunchKernelEx(&cfg, kern, ctl, partials, n_partials, (const GsbEpsExchange)*ex));
    return GSB_OK;
}

// first level of the fold for many partials: GS_FOLD_BLOCKS blocks, each folds one contiguous chunk in a
// fixed order -> out[b * NRHS + r]; gs_end_sweep then folds those.  The grouping depends only on n_partials.
#define GS_FOLD_BLOCKS 64
template <int NRHS>
__global__ void __launch_bounds__(256) gs_fold_partials(const double *__restrict__ partials, int n_partials,
                                                        double *__restrict__ out) {
    const int chunk = (n_partials + GS_FOLD_BLOCKS - 1) / GS_FOLD_BLOCKS;
    const int lo = blockIdx.x * chunk, hi = min(n_partials, lo + chunk);
    double s[NRHS];
#pragma unroll
    for (int r = 0; r < NRHS; ++r) s[r] = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += 256) {
#pragma unroll
        for (int r = 0; r < NRHS; ++r) s[r] += partials[(size_t)i * NRHS + r];
    }
    gsb_block_reduce_store<NRHS, 256>(s, out + (size_t)blockIdx.x * NRHS);
}

// Programmatic dependent launch is used for the ring kernels and gs_end_sweep unless GSB_PDL=0 or the caller is
// capturing a CUDA graph (gsb_pdl_suppress).
// Measured on B200 (profiles/README.md): at 4096^2 (8.4 M rows per phase, 0.22 ms) the overlap costs 3 %, at
// 1024^2 (0.5 M rows, 12 us) it gains 8 %; a strip of an 8-GPU solve is on the small side.  Policy: unset = auto
// (on for phases of at most GS_PDL_AUTO_ROWS rows), GSB_PDL=0 off, 1 on, 2 on without the early prologue.
#define GS_PDL_AUTO_ROWS (3 << 20)
#define GS_STAGED_HINT_AUTO 1                 // measured (r02_call26): configs[4] n = 1e7 58.5 -> 51.7 ms, n = 5e6 21.0 -> 16.7 ms per 20 sweeps
#define GS_STAGED_HINT_X_BYTES (96.0 * 1048576.0) // x of all right-hand sides that L2 (126 MB) can keep
static thread_local int g_pdl_suppress = 0;
static thread_local int g_pdl_last = 0; // decision of the most recent ring launch: gs_end_sweep follows it
void gsb_pdl_suppress(int on) { g_pdl_suppress = on; }
int gsb_pdl_mode(int64_t phase_rows) {
    static int env = -2;
    if (env == -2) {
        const char *e = getenv("GSB_PDL");
        env = e ? atoi(e) : -1;
    }
    int mode = env >= 0 ? env : (phase_rows <= GS_PDL_AUTO_ROWS ? 1 : 0);
    if (g_pdl_suppress) mode = 0;
    g_pdl_last = mode;
    return mode;
}
bool gsb_pdl_enabled() { return g_pdl_last != 0 && !g_pdl_suppress; }

// partials must have room for GS_FOLD_BLOCKS * nrhs more doubles after the n_partials * nrhs used ones
int gsb_launch_end_sweep(GsCtl *ctl, const double *partials, int n_partials, int nrhs, int checked, int mode,
                         cudaStream_t st) {
    if (checked && mode != 2 && n_partials > 8192) {
        double *scratch = const_cast<double *>(partials) + (size_t)n_partials * nrhs;
        switch (nrhs) {
            case 1: gs_fold_partials<1><<<GS_FOLD_BLOCKS, 256, 0, st>>>(partials, n_partials, scratch); break;
            case 2: gs_fold_partials<2><<<GS_FOLD_BLOCKS, 256, 0, st>>>(partials, n_partials, scratch); break;
            case 3: gs_fold_partials<3><<<GS_FOLD_BLOCKS, 256, 0, st>>>(partials, n_partials, scratch); break;
            case 4: gs_fold_partials<4><<<GS_FOLD_BLOCKS, 256, 0, st>>>(partials, n_partials, scratch); break;
            default: return GSB_ERR_ARG;
        }
        GSB_KERNEL_CHECK();
        partials = scratch;
        n_partials = GS_FOLD_BLOCKS;
    }
    void (*kern)(GsCtl *, const double *, int, int, int) = nullptr;
    switch (nrhs) {
        case 1: kern = gs_end_sweep<1>; break;
        case 2: kern = gs_end_sweep<2>; break;
        case 3: kern = gs_end_sweep<3>; break;
        case 4: kern = gs_end_sweep<4>; break;
        default: return GSB_ERR_ARG;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(GS_THREADS);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gsb_pdl_enabled() ? 1 : 0;
    GSB_CUDA(cudaLaunchKernelEx(&cfg, kern, ctl, partials, n_partials, checked, mode));
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// launch plan
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plan_tile_k(const int *__restrict__ rp, int row0, int row1, int tile_rows,
                                                   int ntiles, int *__restrict__ out, int *__restrict__ max_nnz) {
    int t = blockIdx.x * 256 + threadIdx.x;
    if (t > ntiles) return;
    int r = min(row0 + t * tile_rows, row1);
    int k = rp[r];
    out[t] = k;
    if (t < ntiles) {
        int r2 = min(row0 + (t + 1) * tile_rows, row1);
        atomicMax(max_nnz, rp[r2] - k);
    }
}

// Kernel that actually runs for `nrhs` right-hand sides.  plan->kernel == 4 means "gather windows available".
// Measured on B200 (profiles/README.md): one right-hand side is fastest with the windows (4); with several
// fused right-hand sides a window stage (49 KB at k = 3) leaves room for two CTAs per SM only, and the ring
// with global gathers (3), which fits three, wins.  Kernel 4 stays selectable explicitly.
// Kernel 5 (both colours in one launch, gsb_fused.cu) pays when the reuse it creates cannot happen otherwise -- the
// sweep's working set must exceed L2 -- and when a tile carries enough bytes to amortise its flag traffic.  Measured on
// B200 (profiles/README.md, 4096^2): k = 3: 0.85 of the copy bandwidth against 0.80 for kernel 3; k = 1: 0.82 against
// 0.93 for the window kernel (4); 1024^2 (L2-resident): 0.45 against 0.70.  GSB_FUSED_SWEEP=0 never, =2 whenever possible.
// (the 5.03 M-row masked blend of the time-to-tolerance leg, k = 3: kernel 3 1206 / 1216 ms against kernel 5 1239 / 1243 /
// 1283 ms over calls 15 and 25 -- at a third of the 4096^2 grid a sweep is short enough for the ring kernels' dependent
// launches to hide what kernel 5 saves in traffic; hence 8 M rows, not the 3 M at which the working set leaves L2)
#define GS_FUSED_AUTO_MIN_ROWS (8 << 20)
static bool fused_sweep_auto(const GsbPlan *p, int nrhs) {
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("GSB_FUSED_SWEEP");
        env = e ? atoi(e) : 1;
    }
    if (env == 0) return false;
    if (env == 2) return true;
    return nrhs >= 2 && p->color_start[p->n_colors] - p->color_start[0] >= GS_FUSED_AUTO_MIN_ROWS;
}
int gsb_plan_effective_kernel(const GsbPlan *p, int nrhs) {
    if (p->fused_ok && (p->requested == 5 || (p->requested == 0 && fused_sweep_auto(p, nrhs)))) return 5;
    if (p->kernel != 4) return p->kernel;
    if (p->requested == 4) return 4;
    if (p->requested == 3) return 3;
    return nrhs >= 2 ? 3 : 4;
}

// GSB_FUSED_END=1: the last CTA of the sweep's last colour phase ends the sweep itself (GsbEndArgs) instead of a
// kernel of its own.  Opt-in.  Measured on B200 (profiles/README.md, round 2): N = 2 strips +2.7 % (1086 against 1057
// Gnnz/s), N = 8 strips -14 % (2886 against 3344) -- with eight ranks the peer round trip of the stop-rule exchange
// then sits at the tail of a kernel whose CTAs have all retired, whereas the separate kernel is launched
// dependently (PDL) and overlaps the drain.  x is bit-identical either way.
static int fused_end_env() {
    static int env = -2;
    if (env == -2) {
        const char *e = getenv("GSB_FUSED_END");
        env = e ? atoi(e) : -1;
    }
    return env;
}
bool gsb_fused_end_enabled() { return fused_end_env() == 1; }
bool gsb_fused_end_enabled_strips() { return fused_end_env() == 1; }
bool gsb_plan_can_fuse_end(const GsbPlan *p, int nrhs) {
    const int eff = gsb_plan_effective_kernel(p, nrhs);
    const char *e = getenv("GSB_RING_STAGES"); // fused-end variants are built for the default stage count only
    if (e && atoi(e) != 0 && atoi(e) != GS_RING_STAGES_DEFAULT) return false;
    // the window kernel (4) is excluded: with the fused end its solution bits differed on B200 (round 1, (1500, k = 1));
    // until that is understood the separate end-of-sweep kernel stays the only path for it
    return eff == 3;
}

// stop-rule partial slots colour phase c writes (and gs_end_sweep folds) for `nrhs` right-hand sides
int gsb_plan_partial_slots(const GsbPlan *p, int c, int nrhs) {
    const int eff = gsb_plan_effective_kernel(p, nrhs);
    const int nb = p->blocks[c];
    if (eff == 3 || eff == 4) return nb < GS_RING_SLOTS_MAX ? nb : GS_RING_SLOTS_MAX;
    return nb;
}

int GsbPlan::total_blocks() const {
    int s = 0;
    for (int c = 0; c < n_colors; ++c) s += blocks[c];
    return s;
}

int gsb_plan_build(GsbPlan *p, const int *rp, const int *ci, const int *color_start, int n_colors,
                   int kernel_request, cudaStream_t st) {
    p->n_colors = n_colors;
    p->requested = kernel_request;
    for (int c = 0; c <= n_colors; ++c) p->color_start[c] = color_start[c];
    p->kernel = 1;
    p->tile_rows = GS_THREADS;
    p->smem_bytes = 0;
    p->fused_ok = false;
    if (kernel_request != 1) {
        GSB_TRY(p->tiny.alloc(32));
        struct { int *p; } mx = {p->tiny.p};
        for (int tile_rows = GS_THREADS; tile_rows >= 32; tile_rows >>= 1) {
            int total = 0;
            for (int c = 0; c < n_colors; ++c) {
                int rows = color_start[c + 1] - color_start[c];
                p->blocks[c] = (rows + tile_rows - 1) / tile_rows;
                p->tile_off[c] = total;
                total += p->blocks[c] + 1;
            }
            GSB_TRY(p->tile_k.alloc(total));
            GSB_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(int), st));
            for (int c = 0; c < n_colors; ++c) {
                if (!p->blocks[c]) continue;
                plan_tile_k<<<(p->blocks[c] + 1 + 255) / 256, 256, 0, st>>>(rp, color_start[c], color_start[c + 1],
                                                                          tile_rows, p->blocks[c],
                                                                          p->tile_k.p + p->tile_off[c], mx.p);
                GSB_KERNEL_CHECK();
            }
            int h = 0;
            GSB_CUDA(cudaMemcpyAsync(&h, mx.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            GSB_CUDA(cudaStreamSynchronize(st));
            if (h <= GS_TILE_CAP_MAX) {
                p->kernel = 2;
                p->tile_rows = tile_rows;
                p->cap = ((h + 8) + 3) & ~3;
                p->smem_bytes = 16 + p->cap * 12;
                // the ring kernel needs full 256-row tiles and STAGES stages within the 227 KB limit
                if (tile_rows == GS_THREADS && kernel_request != 2 &&
                    64 + 2 * ring_layout(p->cap, GSB_MAX_RHS, true, 0).stage_bytes <= 200 * 1024)
                    p->kernel = 3;
                break;
            }
        }
        // kernel 4: gather windows, if (nearly) every tile's gathers fit a few contiguous spans of x
        p->wcap = 0;
        if (p->kernel == 3) { // windows are probed even when kernel 3 was requested: one plan per matrix
            int total = 0;
            for (int c = 0; c < n_colors; ++c) {
                p->win_off[c] = total;
                total += p->blocks[c];
            }
            GSB_TRY(p->tiny.alloc(32));
            struct { int *p; } stats = {p->tiny.p + 4};
            GSB_CUDA(cudaMemsetAsync(stats.p, 0, 2 * sizeof(int), st));
            GSB_TRY(p->tile_win.alloc((int64_t)total * GS_WIN_DESC));
            for (int c = 0; c < n_colors; ++c) {
                if (!p->blocks[c]) continue;
                plan_tile_windows<<<p->blocks[c], GS_THREADS, 0, st>>>(rp, ci, color_start[c], color_start[c + 1],
                                                                      p->tile_win.p + (size_t)p->win_off[c] * GS_WIN_DESC,
                                                                      stats.p);
                GSB_KERNEL_CHECK();
            }
            int hs[2] = {0, 0};
            GSB_CUDA(cudaMemcpyAsync(hs, stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
            GSB_CUDA(cudaStreamSynchronize(st));
            const bool mostly = hs[1] * 8 <= total; // at most 1/8 of the tiles fall back to global gathers
            if (hs[0] > 0 && mostly &&
                64 + 2 * ring_layout(p->cap, GSB_MAX_RHS, true, hs[0]).stage_bytes <= 220 * 1024) {
                p->kernel = 4;
                p->wcap = hs[0];
                int nnz_h = 0;
                GSB_CUDA(cudaMemcpyAsync(&nnz_h, rp + color_start[n_colors], sizeof(int), cudaMemcpyDeviceToHost, st));
                GSB_CUDA(cudaStreamSynchronize(st));
                GSB_TRY(p->ci_slot.alloc((int64_t)nnz_h + 8));
                for (int c = 0; c < n_colors; ++c) {
                    if (!p->blocks[c]) continue;
                    plan_tile_slots<<<p->blocks[c], GS_THREADS, 0, st>>>(
                        rp, ci, color_start[c], color_start[c + 1], p->tile_win.p + (size_t)p->win_off[c] * GS_WIN_DESC,
                        p->ci_slot.p);
                    GSB_KERNEL_CHECK();
                }
                GSB_CUDA(cudaStreamSynchronize(st));
            } else {
                p->tile_win.release();
            }
        }
        // kernel 5: both colours of a two-colour system in one launch per sweep (gsb_fused.cu)
        if (p->kernel == 3 || p->kernel == 4) GSB_TRY(gsb_plan_build_fused(p, rp, ci, st));
        if (!p->fused_ok && kernel_request == 5) {
            gsb_set_error("fused sweep kernel unavailable: needs a two-colour system with banded coupling");
            return GSB_ERR_ARG;
        }
        if (p->kernel != 4 && kernel_request == 4) {
            gsb_set_error("window kernel unavailable: the gathers of this matrix do not form contiguous windows");
            return GSB_ERR_ARG;
        }
        if (p->kernel != 3 && p->kernel != 4 && kernel_request == 3) {
            gsb_set_error("ring kernel unavailable for this matrix (rows too long); use kernel 0/1/2");
            return GSB_ERR_ARG;
        }
        if (p->kernel != 2 && kernel_request == 2) {
            gsb_set_error("staged kernel needs <= %d entries per 32 rows; use kernel 0/1", GS_TILE_CAP_MAX);
            return GSB_ERR_ARG;
        }
    }
    if (p->kernel == 1) {
        p->tile_k.release();
        for (int c = 0; c < n_colors; ++c)
            p->blocks[c] = (color_start[c + 1] - color_start[c] + GS_THREADS - 1) / GS_THREADS;
    }
    p->valid = true;
    return GSB_OK;
}


// Resident CTAs per SM of a persistent kernel at `smem` bytes of dynamic shared memory, with the opt-in to the
// planner's upper bound made once per (function, device).  One small table for all ring / fused variants, guarded:
// the worker threads of the single-process multi-device solver launch concurrently.
int gsb_kernel_occupancy(const void *kern, int smem, int *occ, int threads) {
    struct Cfg { const void *fn; int smem, occ, dev; };
    static Cfg cfgs[256];
    static int ncfg = 0;
    static std::mutex mu;
    const int dev_now = gsb_current_device(); // function attributes and occupancy are per device
    std::lock_guard<std::mutex> lk(mu);
    for (int q = 0; q < ncfg; ++q)
        if (cfgs[q].fn == kern && cfgs[q].smem == smem && cfgs[q].dev == dev_now) {
            *occ = cfgs[q].occ;
            return GSB_OK;
        }
    // the attribute is per function, not per launch: setting it to this matrix' size would break a later launch for
    // a matrix with larger tiles, so every variant opts in to the upper bound once
    bool seen = false;
    for (int q = 0; q < ncfg; ++q) seen = seen || (cfgs[q].fn == kern && cfgs[q].dev == dev_now);
    if (!seen) GSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int o = 0;
    GSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem));
    if (ncfg == 256) ncfg = 0; // (entries overwritten after a wrap-around just get their attribute set again)
    Cfg *cf = &cfgs[ncfg++];
    cf->fn = kern;
    cf->smem = smem;
    cf->dev = dev_now;
    cf->occ = o < 1 ? 1 : o;
    *occ = cf->occ;
    return GSB_OK;
}

template <int NRHS>
static int plan_launch_t(const GsbPlan *p, int c, const int *rp, const int *ci, const double *va, const double *dg,
                         const double *b, double *x, int64_t ld, bool check, const GsCtl *ctl, double *partials,
                         cudaStream_t st, const GsbHaloArgs *halo_in, const GsbEndArgs *end_in) {
    GsbHaloArgs halo;
    memset(&halo, 0, sizeof(halo));
    if (halo_in) halo = *halo_in;
    GsbEndArgs end;
    memset(&end, 0, sizeof(end));
    if (end_in) end = *end_in;
    if (end.enabled && !gsb_plan_can_fuse_end(p, NRHS)) {
        gsb_set_error("fused end of sweep needs the ring kernels (3/4)");
        return GSB_ERR_ARG;
    }
    if (halo.enabled && p->kernel != 3 && p->kernel != 4) {
        gsb_set_error("fused halo exchange needs the ring kernels (3/4)");
        return GSB_ERR_ARG;
    }
    const int nb = p->blocks[c];
    if (nb <= 0) return GSB_OK;
    const int row0 = p->color_start[c], row1 = p->color_start[c + 1];
    const int eff = gsb_plan_effective_kernel(p, NRHS);
    if (eff == 3 || eff == 4) {
        // tuning knobs (defaults measured on B200, see profiles/README.md); overridable for experiments
        static int env_stages = -1, env_ctas = -1;
        if (env_stages < 0) {
            const char *e = getenv("GSB_RING_STAGES");
            env_stages = e ? atoi(e) : 0;
            e = getenv("GSB_RING_CTAS");
            env_ctas = e ? atoi(e) : 0;
        }
        const bool win = eff == 4;
        const int stage_bytes = ring_layout(p->cap, NRHS, check, win ? p->wcap : 0).stage_bytes;
        const int stages = GS_RING_STAGES_DEFAULT;
        (void)env_stages;
        const int smem = 64 + stages * stage_bytes;
        const int *tk = p->tile_k.p + p->tile_off[c];
        // window descriptors: kernel 4 stages the windows; kernel 3 can use them as L2 prefetch hints (GSB_X_PREFETCH=1|2)
        static int env_xpf = -1;
        if (env_xpf < 0) {
            const char *e = getenv("GSB_X_PREFETCH");
            env_xpf = e ? atoi(e) : 0; // measured: the kernel already runs at ~93 % of the copy bandwidth; the hints cost 4 %
        }
        const bool have_win = p->kernel == 4 && p->tile_win.p;
        const int *tw = (win || (have_win && env_xpf)) ? p->tile_win.p + (size_t)p->win_off[c] * GS_WIN_DESC : nullptr;
        const int wcap = win ? p->wcap : env_xpf; // kernel 3: the prefetch mode travels in the (unused) window capacity
        typedef void (*ring_fn)(const int *, const int *, const double *, const double *, const double *, double *,
                                int64_t, int, int, int, const int *, const int *, int, int, const GsCtl *, double *,
                                const GsbHaloArgs, const GsbEndArgs);
#define GSB_RING_PICK(ST, WN, HL, FE) \
    (check ? (ring_fn)gs_phase_ring<NRHS, true, ST, WN, HL, FE> : (ring_fn)gs_phase_ring<NRHS, false, ST, WN, HL, FE>)
        // (3- and 4-stage rings were measured in round 1 and never won: only the 2-stage variants are instantiated;
        // the fused end is never combined with the window kernel -- gsb_plan_can_fuse_end)
#define GSB_RING_PICK_ST(WN, HL, FE) GSB_RING_PICK(2, WN, HL, FE)
        // the fused-end variants exist for the default stage count only (gsb_plan_can_fuse_end checks it)
#define GSB_RING_PICK_FE(WN, HL) (end.enabled ? GSB_RING_PICK(2, WN, HL, true) : GSB_RING_PICK_ST(WN, HL, false))
        ring_fn kern = nullptr;
        if (halo.enabled)
            kern = win ? GSB_RING_PICK_ST(true, true, false) : GSB_RING_PICK_FE(false, true);
        else
            kern = win ? GSB_RING_PICK_ST(true, false, false) : GSB_RING_PICK_FE(false, false);
#undef GSB_RING_PICK_FE
#undef GSB_RING_PICK_ST
#undef GSB_RING_PICK
        int per_sm = 1;
        GSB_TRY(gsb_kernel_occupancy((const void *)kern, smem, &per_sm));
        if (env_ctas && env_ctas < per_sm) per_sm = env_ctas;
        // halo variant: n_halo_tiles one-tile CTAs first, then the persistent ring over the interior tiles
        // (the halo CTAs start first and occupy resident slots: the ring gets the remaining slots, so that every
        // ring CTA is resident from the start -- a ring CTA that had to wait for a halo CTA to retire would finish its
        // statically assigned tiles that much later and stretch the phase)
        const int ring_tiles = halo.enabled ? halo.n_interior : nb;
        int grid = gsb_sm_count() * per_sm;
        if (halo.enabled) {
            const int floor_grid = gsb_sm_count(); // many halo tiles: still at least one ring CTA per SM
            grid = grid - halo.n_halo_tiles > floor_grid ? grid - halo.n_halo_tiles : floor_grid;
        }
        if (grid > ring_tiles) grid = ring_tiles;
        if (halo.enabled) grid += halo.n_halo_tiles;
        if (grid > GS_RING_SLOTS_MAX) {
            if (halo.enabled) {
                gsb_set_error("fused halo exchange: %d halo tiles exceed the partial-slot budget", halo.n_halo_tiles);
                return GSB_ERR_ARG;
            }
            grid = GS_RING_SLOTS_MAX;
        }
        if (halo.enabled && (halo.interior_base < 0 || halo.n_halo_tiles + halo.n_interior != nb)) {
            gsb_set_error("fused halo exchange needs the interior tiles to be one contiguous range");
            return GSB_ERR_ARG;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(GS_THREADS);
        cfg.dynamicSmemBytes = (size_t)smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        const int pdl = gsb_pdl_mode((int64_t)row1 - row0);
        cfg.numAttrs = pdl ? 1 : 0;
        // the early prologue stages this colour's x_old while the predecessor may still be running; that is only safe
        // when the predecessor is ANOTHER colour's phase (with a single non-empty colour the predecessor, through the
        // end-of-sweep kernel's early release of its dependents, is the phase that writes exactly those values)
        int nonempty = 0;
        for (int q = 0; q < p->n_colors; ++q) nonempty += p->blocks[q] > 0 ? 1 : 0;
        halo.pdl_early = (pdl == 1 && nonempty >= 2) ? 1 : 0;
        GSB_CUDA(cudaLaunchKernelEx(&cfg, kern, rp, (const int *)(win ? p->ci_slot.p : ci), va, dg, b, x, ld, row0, row1,
                                    nb, tk, tw, p->cap, wcap, ctl, partials, (const GsbHaloArgs)halo,
                                    (const GsbEndArgs)end));
    } else if (eff == 2) {
        // L2 residency hints (see gs_phase_staged): GSB_STAGED_L2HINT=0 never, 1 always, default: when x of all right-
        // hand sides is at most GS_STAGED_HINT_X_BYTES and the CSR is several times that
        static const int hint_env = [] {
            const char *e = getenv("GSB_STAGED_L2HINT");
            return e ? atoi(e) : -1;
        }();
        const double x_bytes = 8.0 * (double)NRHS * (double)(p->color_start[p->n_colors] - p->color_start[0]);
        const double csr_bytes = 12.0 * (double)p->nnz_hint;
        const bool hint = hint_env >= 0 ? hint_env != 0
                                        : (GS_STAGED_HINT_AUTO && x_bytes <= GS_STAGED_HINT_X_BYTES && csr_bytes >= 4.0 * x_bytes);
        auto kt = hint ? gs_phase_staged<NRHS, true, true> : gs_phase_staged<NRHS, true, false>;
        auto kf = hint ? gs_phase_staged<NRHS, false, true> : gs_phase_staged<NRHS, false, false>;
        if (p->smem_bytes > 48 * 1024) {
            static bool set_t[8][2] = {{false}}, set_f[8][2] = {{false}};
            bool &flag = check ? set_t[NRHS][hint ? 1 : 0] : set_f[NRHS][hint ? 1 : 0];
            if (!flag) {
                GSB_CUDA(cudaFuncSetAttribute(check ? (const void *)kt : (const void *)kf,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 16 + GS_TILE_CAP_MAX * 12 + 256));
                flag = true;
            }
        }
        const int *tk = p->tile_k.p + p->tile_off[c];
        if (check)
            kt<<<nb, GS_THREADS, p->smem_bytes, st>>>(rp, ci, va, dg, b, x, ld, row0, row1, p->tile_rows, tk, p->cap, ctl,
                                                       partials);
        else
            kf<<<nb, GS_THREADS, p->smem_bytes, st>>>(rp, ci, va, dg, b, x, ld, row0, row1, p->tile_rows, tk, p->cap, ctl,
                                                       partials);
    } else {
        if (check)
            gs_phase_direct<NRHS, true><<<nb, GS_THREADS, 0, st>>>(rp, ci, va, dg, b, x, ld, row0, row1, ctl, partials);
        else
            gs_phase_direct<NRHS, false><<<nb, GS_THREADS, 0, st>>>(rp, ci, va, dg, b, x, ld, row0, row1, ctl, partials);
    }
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

int gsb_plan_launch(const GsbPlan *p, int c, const int *rp, const int *ci, const double *va, const double *dg,
                    const double *b, double *x, int64_t ld, int nrhs, bool check, const GsCtl *ctl, double *partials,
                    cudaStream_t st, const GsbHaloArgs *halo, const GsbEndArgs *end) {
    switch (nrhs) {
        case 1: return plan_launch_t<1>(p, c, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, halo, end);
        case 2: return plan_launch_t<2>(p, c, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, halo, end);
        case 3: return plan_launch_t<3>(p, c, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, halo, end);
        case 4: return plan_launch_t<4>(p, c, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, halo, end);
    }
    gsb_set_error("nrhs must be 1..%d", GSB_MAX_RHS);
    return GSB_ERR_ARG;
}
