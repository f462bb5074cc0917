// gsb_fused.cu -- kernel 5: one launch per Gauss-Seidel sweep for two-colour (red-black) systems.
//
// The per-colour phase kernels (gsb_phase.cu) move every byte at copy speed but run colour 0 over the whole system
// before colour 1 starts, so nothing survives in L2 between the phases (2.3 GB per sweep at 4096^2 against 126 MB):
// the values colour 1 gathers are read back from HBM after colour 0 wrote them, and the stop rule's x_old of a colour
// is a second HBM read of what the other colour gathered half a sweep earlier.  DRAM traffic = algorithmic + 8k B/row.
//
// Here both colours run in ONE persistent kernel as a software wavefront.  The work items of a sweep are the tiles
// of both colours in one static sequence
//       R_0 .. R_{L-1},  B_0, R_L, B_1, R_{L+1}, ...            (R = colour 0, B = colour 1, L = lead)
// walked round-robin by the resident CTAs (item p -> CTA p mod grid), each CTA keeping the ring of bulk-copy
// stages of kernel 3.  A colour-1 tile B_j may run once the colour-0 tiles in dep[j] = [lo, hi] have finished:
// those are the tiles whose new values it gathers (read after write) and the tiles that gather ITS rows' old values
// (write after read); the plan computes the range from the column indices, and L >= max_j(hi_j - j) + 1 puts every
// dependency earlier in the sequence; together with "a CTA that has to wait first retires and publishes its own
// current tile" (control warp, below) no cycle of waits can form (all CTAs are resident: grid = SMs x occupancy).
// Colour-0 tiles never wait: what they read is the previous sweep's output (kernel boundary).  A finished colour-0
// tile publishes flags[t] = sweep number (release); the control warp polls a colour-1 tile's range (relaxed loads,
// one acquire fence) before it releases the tile to the compute warps.  The reuse distance (~2L tiles, a few MB to
// tens of MB) sits inside L2: colour 1's gathers and x_old hit L2, every x is read from HBM once and written once
// per sweep -- the algorithmic count -- and the stop rule costs no traffic at all.
//
// Arithmetic is the row body of the phase kernels (gs_row_sigma, unfused, storage order): x after every sweep is
// bit-identical to kernels 1-4 and 6 (tests/test_gs_gpu.py::test_all_kernels_agree_bitwise).
#include "gsb_ring.cuh"

#include <limits.h>
#include <stdlib.h>

struct GsbFusedArgs { // (scalar members, selected with ?: -- indexing a by-value array would put it on the stack)
    int row0_0, row0_1, row1_0, row1_1, nt0, nt1;
    const int *tile_k0, *tile_k1; // CSR offset at every tile boundary of the colour
    const int2 *dep;      // per colour-1 tile: first / last colour-0 tile it waits for
    const int2 *span;     // per tile (colour 0 first): smallest / largest column it gathers ({INT_MAX, -1}: none)
    int *flags;           // per colour-0 tile
    int lead, nalt;       // colour 0 runs `lead` tiles ahead; `nalt` = alternating (B, R) pairs after the lead
};

// item p of the sweep's sequence -> (colour, tile)
__device__ __forceinline__ void fused_item(const GsbFusedArgs &fa, int p, int &c, int &t) {
    if (p < fa.lead) {
        c = 0;
        t = p;
        return;
    }
    const int q = p - fa.lead;
    if (q < 2 * fa.nalt) {
        c = (q & 1) ? 0 : 1;
        t = (q & 1) ? fa.lead + (q >> 1) : (q >> 1);
        return;
    }
    const int rest = q - 2 * fa.nalt;
    if (fa.lead + fa.nalt < fa.nt0) { // colour 1 exhausted first: the remaining colour-0 tiles
        c = 0;
        t = fa.lead + fa.nalt + rest;
    } else {
        c = 1;
        t = fa.nalt + rest;
    }
}

// One record per work item of the sweep, in sequence order, built once per (matrix, lead) by plan_fused_items: what
// the control warp needs to stage, release and retire a tile without computing or chasing anything at run time.
struct __align__(16) FusedItem {
    int k0, k1, r_begin, rows; // CSR span of the tile, its first row and row count
    int c, t, dep_lo, dep_hi;  // colour, tile, colour-0 tiles to wait for (dep_lo > dep_hi: none)
    // A colour-0 item waits for nothing; its dep fields carry an L2 prefetch hint instead: the x columns
    // [dep_lo, dep_lo + (-dep_hi - 1)) -- the top of its gather span, i.e. the values of the other colour that no
    // earlier tile of the sweep has touched (on a grid: the neighbours in the next image row).
};
#define GS_FUSED_GATHER_HINT_COLS 320 // a tile's rows + the raggedness of a 5-point row boundary

__global__ void __launch_bounds__(256) plan_fused_items(const GsbFusedArgs fa, FusedItem *__restrict__ items) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= fa.nt0 + fa.nt1) return;
    FusedItem it;
    fused_item(fa, p, it.c, it.t);
    const int *tk = it.c ? fa.tile_k1 : fa.tile_k0;
    it.k0 = tk[it.t];
    it.k1 = tk[it.t + 1];
    it.r_begin = (it.c ? fa.row0_1 : fa.row0_0) + it.t * GS_THREADS;
    it.rows = min(GS_THREADS, (it.c ? fa.row1_1 : fa.row1_0) - it.r_begin);
    it.dep_lo = 0;
    it.dep_hi = -1;
    if (it.c == 1) {
        const int2 d = fa.dep[it.t];
        it.dep_lo = d.x;
        it.dep_hi = d.y;
    } else if (fa.span) {
        const int2 sp = fa.span[it.t];
        if (sp.y >= sp.x) { // 16-byte aligned start, even count; the planes are padded to an even length
            const int lo = max(sp.x, sp.y + 1 - GS_FUSED_GATHER_HINT_COLS) & ~1;
            const int cnt = ((sp.y + 2) & ~1) - lo;
            it.dep_lo = lo;
            it.dep_hi = -cnt - 1;
        }
    }
    items[p] = it;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t *bar, uint32_t bytes) { // tx-count up, no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// Control warp, all 32 lanes: every colour-0 tile in [lo, hi] has published `epoch`.  The lanes poll one flag each
// (one coalesced L2 round trip for the usual ~20-tile range) with relaxed loads and fence once they have all seen it;
// the warp barrier then orders lane 0's arrive on the stage's `full` mbarrier after every lane's observation, so the
// compute warps' gathers are ordered after the colour-0 tiles' stores (release / acquire at GPU scope, cumulative
// through the warp and mbarrier synchronisation).  Bounded: a flag that never comes (it cannot, short of a device
// fault) raises ctl->error instead of hanging the GPU, and once one tile has given up nobody waits any more.
__device__ __forceinline__ void fused_acquire_fence(bool split) {
    // split: acquire only (SASS: CCTL.IVALL, the L1 invalidation) -- fence.acq_rel adds a MEMBAR.ALL.GPU in front of it,
    // i.e. the warp also waits for its outstanding accesses, which an observer of flags does not need
    if (split)
        asm volatile("fence.acquire.gpu;" ::: "memory");
    else
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ void fused_wait_tiles(const int *flags, int lo, int hi, int epoch, GsCtl *ctl, bool split) {
    const int lane = threadIdx.x & 31;
    for (int base = lo; base <= hi; base += 32) {
        const int i = base + lane;
        int spins = 0;
        for (;;) {
            int v = epoch;
            if (i <= hi) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
            if (__all_sync(0xffffffffu, v >= epoch)) break;
            if (++spins > (1 << 22) || ((spins & 255) == 0 && *(volatile int *)&ctl->error)) { // ~1 s
                if (lane == 0) ctl->error = 2;
                return;
            }
            __nanosleep(spins < 16 ? 20 : 200);
        }
    }
    fused_acquire_fence(split);
    __syncwarp();
}

// Warp-specialised: GS_THREADS compute threads (one row of the tile each) + one control warp.  The control warp is
// the only one that talks to the rest of the GPU: lane 0 issues the bulk copies of a tile's inputs (up to STAGES
// tiles ahead); the warp polls the dependency flags of a colour-1 tile and only then lane 0 arrives on the stage's
// `full` mbarrier (the copies complete its transaction count); once all compute warps have arrived on the stage's
// `empty` mbarrier -- the tile's rows are stored -- lane 0 refills the stage and publishes a colour-0 tile's flag.
// The compute warps never meet a CTA-wide barrier inside the loop; a warp moves on to the next tile as soon as that
// tile's stage is full.
//   v1 (thread 0 of the compute warps did the control work between two __syncthreads per tile): 18 of 27 issue-stall
//      cycles on those barriers, 54 % of the DRAM throughput (profiles/r02_sweep_fused_v1_rhs3.txt).
//   v2 (one control THREAD, polling its ~20 flags one after another and computing every descriptor on the fly): the
//      compute warps starved on `full`, 43 % (profiles/r02_sweep_fused_v2_rhs3.txt).
#define GS_FUSED_THREADS (GS_THREADS + 32)
#define GS_FUSED_WARPS (GS_THREADS / 32)

template <int NRHS, bool CHECK, int STAGES>
__global__ void __launch_bounds__(GS_FUSED_THREADS, (NRHS >= 4 ? 3 : 4))
    gs_sweep_fused(const int *__restrict__ rp, const int *__restrict__ ci, const double *__restrict__ va,
                   const double *__restrict__ dg, const double *__restrict__ b, double *x, int64_t n, int cap, GsCtl *ctl,
                   double *__restrict__ partials, const FusedItem *__restrict__ items, int total, int *flags, int debug,
                   int pubk, int hints) {
    static_assert(STAGES == 2, "the control warp's item pipeline (r0..r3) is written for two stages");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const RingLayout L = ring_layout(cap, NRHS, CHECK, 0);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *empty = full + STAGES;
    unsigned char *stage0 = smem_raw + 64;
    const int tid = threadIdx.x, bid = blockIdx.x, gsz = gridDim.x;
    __shared__ double red_ws[NRHS][GS_FUSED_WARPS];

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GS_FUSED_WARPS);
        }
    }
    __syncthreads();
    // the previous sweep (and its end-of-sweep kernel) is complete and visible from here on
    pdl_wait();
    pdl_launch_dependents();
    if (*(volatile const int *)&ctl->done) return;
    const int epoch = *(volatile const int *)&ctl->sweeps + 1;
    const int my_items = bid < total ? (total - bid + gsz - 1) / gsz : 0;

    if (tid >= GS_THREADS) {
        // ------------------------------------ control warp -------------------------------------
        const int lane = tid - GS_THREADS;
        const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
        auto load_item = [&](int j) -> FusedItem { // all lanes read the same 32 bytes (one broadcast transaction)
            const int4 *src = reinterpret_cast<const int4 *>(items + (bid + j * gsz));
            const int4 a = __ldg(src), c4 = __ldg(src + 1);
            FusedItem it;
            it.k0 = a.x; it.k1 = a.y; it.r_begin = a.z; it.rows = a.w;
            it.c = c4.x; it.t = c4.y; it.dep_lo = c4.z; it.dep_hi = c4.w;
            return it;
        };
        auto issue = [&](const FusedItem &d, int s) { // lane 0: every per-tile input is one contiguous span
            unsigned char *st = stage0 + (size_t)s * L.stage_bytes;
            const int r_begin = d.r_begin, rows = d.rows, k0 = d.k0, k1 = d.k1;
            const int kv0 = k0 & ~1, kc0 = k0 & ~3;
            const uint32_t bytes_v = (uint32_t)(((k1 + 1) & ~1) - kv0) * 8u;
            const uint32_t bytes_c = (uint32_t)(((k1 + 3) & ~3) - kc0) * 4u;
            const int ra = r_begin & ~3;
            const uint32_t bytes_r = (uint32_t)(((r_begin + rows + 1 + 3) & ~3) - ra) * 4u;
            const int ea = r_begin & ~1; // the leading dimension is even: the same alignment for every plane
            const uint32_t bytes_p = (uint32_t)(((r_begin + rows + 1) & ~1) - ea) * 8u;
            const uint32_t tx = bytes_v + bytes_c + bytes_r + bytes_p + (uint32_t)NRHS * bytes_p * (CHECK ? 2u : 1u);
            int *hdr = reinterpret_cast<int *>(st + L.hdr_off);
            hdr[0] = k0;
            hdr[1] = r_begin;
            hdr[2] = rows;
            mbar_expect_tx_only(&full[s], tx);
            // the matrix, the diagonal and b are streamed once per sweep: evict-first, so that they do not push the
            // x lines out of L2 on whose reuse this kernel lives (hints & 1; measured, profiles/README.md)
            if (hints & 1) {
                if (bytes_v) bulk_g2s_hint(st + L.va_off, va + kv0, bytes_v, &full[s], pol_stream);
                if (bytes_c) bulk_g2s_hint(st + L.ci_off, ci + kc0, bytes_c, &full[s], pol_stream);
                bulk_g2s_hint(st + L.rp_off, rp + ra, bytes_r, &full[s], pol_stream);
                bulk_g2s_hint(st + L.dg_off, dg + ea, bytes_p, &full[s], pol_stream);
            } else {
                if (bytes_v) bulk_g2s(st + L.va_off, va + kv0, bytes_v, &full[s]);
                if (bytes_c) bulk_g2s(st + L.ci_off, ci + kc0, bytes_c, &full[s]);
                bulk_g2s(st + L.rp_off, rp + ra, bytes_r, &full[s]);
                bulk_g2s(st + L.dg_off, dg + ea, bytes_p, &full[s]);
            }
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                if (hints & 1)
                    bulk_g2s_hint(st + L.b_off + r * L.plane * 8, b + r * n + ea, bytes_p, &full[s], pol_stream);
                else
                    bulk_g2s(st + L.b_off + r * L.plane * 8, b + r * n + ea, bytes_p, &full[s]);
                // x_old: this tile's own rows, which nobody but this tile writes during the sweep
                if (CHECK) {
                    if (hints & 2)
                        bulk_g2s_hint(st + L.xo_off + r * L.plane * 8, x + r * n + ea, bytes_p, &full[s], pol_keep);
                    else
                        bulk_g2s(st + L.xo_off + r * L.plane * 8, x + r * n + ea, bytes_p, &full[s]);
                }
            }
        };
        // hints & 16 -- lane 0: pull the inputs of the item AFTER the next refill into L2 (no completion tracking).  The
        // ring has two stages, so a refill has one tile time to land; issued a tile time before the refill, the
        // prefetch turns the refill's HBM round trip (~3 us under load) into an L2 hit.
        auto prefetch_l2 = [&](const FusedItem &d) {
            const int r_begin = d.r_begin, rows = d.rows, k0 = d.k0, k1 = d.k1;
            const int kv0 = k0 & ~1, kc0 = k0 & ~3;
            const uint32_t bytes_v = (uint32_t)(((k1 + 1) & ~1) - kv0) * 8u;
            const uint32_t bytes_c = (uint32_t)(((k1 + 3) & ~3) - kc0) * 4u;
            const int ea = r_begin & ~1;
            const uint32_t bytes_p = (uint32_t)(((r_begin + rows + 1) & ~1) - ea) * 8u;
            if (hints & 16) {
                if (bytes_v) bulk_prefetch_l2(va + kv0, bytes_v);
                if (bytes_c) bulk_prefetch_l2(ci + kc0, bytes_c);
            }
            if ((hints & 64) && d.c == 0 && d.dep_hi < -1) { // the gathers' compulsory misses (see FusedItem)
                const uint32_t bytes_g = (uint32_t)(-d.dep_hi - 1) * 8u;
#pragma unroll
                for (int r = 0; r < NRHS; ++r) bulk_prefetch_l2(x + r * n + d.dep_lo, bytes_g);
            }
            if (hints & 32) {
                const int ra = r_begin & ~3;
                bulk_prefetch_l2(rp + ra, (uint32_t)(((r_begin + rows + 1 + 3) & ~3) - ra) * 4u);
                bulk_prefetch_l2(dg + ea, bytes_p);
#pragma unroll
                for (int r = 0; r < NRHS; ++r) {
                    bulk_prefetch_l2(b + r * n + ea, bytes_p);
                    if (CHECK) bulk_prefetch_l2(x + r * n + ea, bytes_p);
                }
            }
        };
        // ---- dependency polls, issued one iteration before they are needed ------------------------------------
        // start: every lane loads one flag of the item's range (relaxed; the load completes in the background -- its
        // value is first looked at an iteration later).  finish: all lanes saw `epoch` -> one acquire fence; else the
        // blocking path.  Ranges wider than the warp always take the blocking path.
        auto start_poll = [&](const FusedItem &d) -> int {
            int v = epoch;
            const int i = d.dep_lo + lane;
            if (!(debug & 1) && i <= d.dep_hi)
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
            return v;
        };
        // ---- colour-0 tiles this CTA has finished but not yet published -------------------------------------
        // One release fence covers up to `pubk` tiles (the fence has to wait for the SM's outstanding stores: ~0.6 us
        // each time).  Never held across a blocking wait: a CTA that waits has published everything it computed, so
        // the in-order argument for deadlock freedom is unchanged.
        int pend_n = 0, pend_t0 = 0, pend_t1 = 0, pend_t2 = 0, pend_t3 = 0;
        auto flush = [&]() {
            if (pend_n && lane == 0 && !(debug & 2)) {
                // release: cumulative over the compute warps' stores observed through the `empty` mbarriers.
                // hints & 8: fence.release (MEMBAR.ALL.GPU only) instead of fence.acq_rel (MEMBAR + CCTL.IVALL): the
                // publisher has nothing to acquire, and the L1 invalidation costs the SM's other CTAs their gather hits
                if (hints & 8)
                    asm volatile("fence.release.gpu;" ::: "memory");
                else
                    asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flags + pend_t0), "r"(epoch) : "memory");
                if (pend_n > 1) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flags + pend_t1), "r"(epoch) : "memory");
                if (pend_n > 2) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flags + pend_t2), "r"(epoch) : "memory");
                if (pend_n > 3) asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(flags + pend_t3), "r"(epoch) : "memory");
            }
            pend_n = 0;
        };
        // blocking completion of a poll (the slow path); fences and re-converges the warp itself
        auto wait_blocking = [&](const FusedItem &d) {
            flush();
            fused_wait_tiles(flags, d.dep_lo, d.dep_hi, epoch, ctl, (hints & 8) != 0);
        };
        // did the poll started an iteration ago see every flag?  (uniform over the warp)
        auto poll_ready = [&](const FusedItem &d, int v) -> bool {
            if (d.dep_hi < d.dep_lo || (debug & 1)) return true; // a colour-0 item: nothing to wait for
            if (d.dep_hi - d.dep_lo >= 32) return false;          // wider than the warp: always the blocking path
            return __all_sync(0xffffffffu, v >= epoch);
        };
        auto acquire = [&](const FusedItem &d) {
            if (d.dep_hi >= d.dep_lo && !(debug & 1)) {
                fused_acquire_fence((hints & 8) != 0);
                __syncwarp();
            }
        };
        // items in flight: r0 = item j (being computed), r1 = item j+1 (staged, released this iteration), r2 = item
        // j+2 (issued this iteration), r3 = item j+3 (record being fetched); pv1 / pv2 = the polls of r1 / r2
        FusedItem r0, r1, r2, r3;
        r0 = r1 = r2 = r3 = FusedItem{0, 0, 0, 0, 0, 0, 0, -1};
        if (my_items > 0) r0 = load_item(0);
        if (my_items > 1) r1 = load_item(1);
        if (my_items > 2) r2 = load_item(2);
        if (lane == 0) {
            if (my_items > 0) issue(r0, 0);
            if (my_items > 1) issue(r1, 1);
        }
        int pv1 = epoch, pv2 = epoch;
        if (my_items > 0) { // (nothing of this CTA is unpublished yet: blocking here is harmless)
            if (poll_ready(r0, start_poll(r0)))
                acquire(r0);
            else
                wait_blocking(r0);
            if (lane == 0) mbar_arrive(&full[0]);
        }
        if (my_items > 1) pv1 = start_poll(r1);
        for (int j = 0; j < my_items; ++j) {
            const int s = j % STAGES;
            // retire item j: every compute warp has stored its rows -> refill the stage (first: the copies are what the
            // pipeline waits for), then make a colour-0 tile's values public
            auto retire = [&]() {
                mbar_wait(&empty[s], (uint32_t)(j / STAGES) & 1u);
                if (lane == 0 && j + 2 < my_items) issue(r2, s);
                if (r0.c == 0) {
                    if (pend_n == 0) pend_t0 = r0.t;
                    else if (pend_n == 1) pend_t1 = r0.t;
                    else if (pend_n == 2) pend_t2 = r0.t;
                    else pend_t3 = r0.t;
                    if (++pend_n >= pubk) flush();
                }
            };
            bool retired = false;
            if (j + 1 < my_items) { // release item j+1 to the compute warps once its dependencies are met
                if (poll_ready(r1, pv1)) {
                    acquire(r1);
                } else {
                    // NEVER block while this CTA holds a tile that is computed (or being computed) but not yet
                    // published: item j will finish -- it was released -- so retire and publish it first.  A CTA
                    // that waits has then published every tile it is responsible for up to j, and the earliest
                    // unpublished colour-0 tile of the whole sweep always belongs to a CTA that is not waiting:
                    // no cycle of CTAs waiting on each other's current tile can form, whatever the lead.
                    // (Without this a lead of about half the grid size deadlocked: tile j+1 of CTA X needed the
                    // current tile of CTA Y, whose tile j+1 needed the current tile of X -- measured, round 2.)
                    retire();
                    retired = true;
                    wait_blocking(r1);
                }
                if (lane == 0) mbar_arrive(&full[(j + 1) % STAGES]);
            }
            if (!retired) {
                // (r2 = item j+2 was fetched an iteration ago; it is refilled into stage s right after the wait below)
                if ((hints & (16 | 32 | 64)) && lane == 0 && j + 2 < my_items) prefetch_l2(r2);
                retire();
            }
            if (j + 2 < my_items) pv2 = start_poll(r2);
            if (j + 3 < my_items) r3 = load_item(j + 3);
            r0 = r1;
            r1 = r2;
            r2 = r3;
            pv1 = pv2;
        }
        flush();
    } else {
        // ------------------------------------ compute warps ------------------------------------
        double acc[NRHS];
#pragma unroll
        for (int r = 0; r < NRHS; ++r) acc[r] = 0.0;
        const uint64_t pol_x = l2_policy_evict_last();
        for (int j = 0; j < my_items; ++j) {
            const int s = j % STAGES;
            unsigned char *st = stage0 + (size_t)s * L.stage_bytes;
            mbar_wait(&full[s], (uint32_t)(j / STAGES) & 1u);
            const int *hdr = reinterpret_cast<const int *>(st + L.hdr_off);
            const int k0 = hdr[0], r_begin = hdr[1], rows = hdr[2];
            const double *va_s = reinterpret_cast<const double *>(st + L.va_off) - (k0 & ~1);
            const int *ci_s = reinterpret_cast<const int *>(st + L.ci_off) - (k0 & ~3);
            const int *rp_s = reinterpret_cast<const int *>(st + L.rp_off) + (r_begin & 3);
            const int i = r_begin + tid;
            if (tid < rows) {
                const int rs = rp_s[tid], len = rp_s[tid + 1] - rs;
                const int po = (r_begin & 1) + tid;
                const double d = reinterpret_cast<const double *>(st + L.dg_off)[po];
                double sig[NRHS];
                gs_row_sigma<NRHS>(ci_s + rs, va_s + rs, len, [&](int col, int r) { return x[r * n + col]; }, sig);
                if (d != 0.0) { // zero or absent diagonal: row skipped, x_i unchanged (v2 :360-363)
#pragma unroll
                    for (int r = 0; r < NRHS; ++r) {
                        const double bb = reinterpret_cast<const double *>(st + L.b_off)[r * L.plane + po];
                        const double xn = __ddiv_rn(__dsub_rn(bb, sig[r]), d);
                        if (CHECK) acc[r] += fabs(xn - reinterpret_cast<const double *>(st + L.xo_off)[r * L.plane + po]);
                        if (hints & 4) // keep the new value in L2: the other colour gathers it a few hundred tiles later
                            asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(x + r * n + i), "d"(xn), "l"(pol_x)
                                         : "memory");
                        else
                            x[r * n + i] = xn;
                    }
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]); // this warp is done with stage s and has stored its rows
        }
        if (CHECK) {
#pragma unroll
            for (int r = 0; r < NRHS; ++r) {
                double t = acc[r];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
                if ((tid & 31) == 0) red_ws[r][tid >> 5] = t;
            }
        }
    }
    if (CHECK) { // fixed-order fold of the CTA's stop-rule partial -> slot blockIdx.x
        __syncthreads();
        if (tid < NRHS) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < GS_FUSED_WARPS; ++w) sum += red_ws[tid][w];
            partials[(size_t)blockIdx.x * NRHS + tid] = sum;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// plan: dependency ranges
// ---------------------------------------------------------------------------------------------
// smallest / largest column of the tile's off-diagonal entries ({INT_MAX, -1}: none)
__global__ void __launch_bounds__(GS_THREADS) plan_tile_colspan(const int *__restrict__ rp, const int *__restrict__ ci,
                                                                int row0, int row1, int2 *__restrict__ span) {
    __shared__ int lo_s, hi_s;
    if (threadIdx.x == 0) {
        lo_s = INT_MAX;
        hi_s = -1;
    }
    __syncthreads();
    const int i = row0 + blockIdx.x * GS_THREADS + threadIdx.x;
    if (i < row1) {
        int lo = INT_MAX, hi = -1;
        for (int k = rp[i]; k < rp[i + 1]; ++k) {
            const int c = ci[k];
            lo = min(lo, c);
            hi = max(hi, c);
        }
        if (hi >= 0) {
            atomicMin(&lo_s, lo);
            atomicMax(&hi_s, hi);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) span[blockIdx.x] = make_int2(lo_s, hi_s);
}

// dep[j] over the colour-1 tiles.  side 0: the colour-0 tiles a colour-1 tile reads; side 1: the colour-1 tiles a
// colour-0 tile reads (whose old values must be consumed before they are overwritten).  stats[0] |= 1 when a
// column lies outside the other colour's rows (not a two-colour system this kernel can run).
__global__ void __launch_bounds__(256) plan_fused_deps(const int2 *__restrict__ span, int ntiles, int side, int o_row0,
                                                       int o_row1, int o_ntiles, int *__restrict__ dep,
                                                       int *__restrict__ stats) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= ntiles) return;
    const int2 sp = span[t];
    if (sp.y < 0) return; // no off-diagonal entries
    if (sp.x < o_row0 || sp.y >= o_row1) {
        atomicOr(&stats[0], 1);
        return;
    }
    const int lo = (sp.x - o_row0) / GS_THREADS, hi = min((sp.y - o_row0) / GS_THREADS, o_ntiles - 1);
    if (side == 0) {
        atomicMin(&dep[2 * t], lo);
        atomicMax(&dep[2 * t + 1], hi);
    } else {
        if (hi - lo > 4096) { // not a banded system: the fused walk would serialise
            atomicOr(&stats[0], 2);
            return;
        }
        for (int j = lo; j <= hi; ++j) {
            atomicMin(&dep[2 * j], t);
            atomicMax(&dep[2 * j + 1], t);
        }
    }
}

// stats[1] = max_j (hi_j - j), stats[2] = max_j (hi_j - lo_j + 1)
__global__ void __launch_bounds__(256) plan_fused_stats(int *__restrict__ dep, int ntiles, int *__restrict__ stats) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= ntiles) return;
    const int lo = dep[2 * j], hi = dep[2 * j + 1];
    if (hi < 0) { // the tile neither reads nor is read: an empty range the wait loop skips
        dep[2 * j] = 0;
        return;
    }
    atomicMax(&stats[1], hi - j);
    atomicMax(&stats[2], hi - lo + 1);
}

__global__ void __launch_bounds__(256) plan_fused_init(int *__restrict__ dep, int ntiles) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= ntiles) return;
    dep[2 * j] = INT_MAX;
    dep[2 * j + 1] = -1;
}

#define GS_FUSED_DEP_WIDTH_MAX 256 // colour-0 tiles one colour-1 tile may wait for (8 polls per lane)

// Called by gsb_plan_build once the ring kernel (3/4) is available: decides p->fused_ok.
int gsb_plan_build_fused(GsbPlan *p, const int *rp, const int *ci, cudaStream_t st) {
    p->fused_ok = false;
    p->fused_items_lead = -1; // the item table belongs to the previous matrix
    if (!p->fused_allowed || p->n_colors != 2 || p->tile_rows != GS_THREADS) return GSB_OK;
    const int nt0 = p->blocks[0], nt1 = p->blocks[1];
    if (nt0 <= 0 || nt1 <= 0) return GSB_OK;
    DevBuf<int2> &span = p->fused_span; // kept: the item table takes the colour-0 tiles' gather hints from it
    GSB_TRY(span.alloc((int64_t)nt0 + nt1));
    GSB_TRY(p->tiny.alloc(32));
    struct { int *p; } stats = {p->tiny.p + 8};
    GSB_TRY(p->fused_dep.alloc((int64_t)2 * nt1));
    GSB_TRY(p->fused_flags.alloc(nt0));
    GSB_CUDA(cudaMemsetAsync(stats.p, 0, 4 * sizeof(int), st));
    GSB_CUDA(cudaMemsetAsync(p->fused_flags.p, 0, sizeof(int) * (size_t)nt0, st));
    const int *cs = p->color_start;
    plan_tile_colspan<<<nt0, GS_THREADS, 0, st>>>(rp, ci, cs[0], cs[1], span.p);
    plan_tile_colspan<<<nt1, GS_THREADS, 0, st>>>(rp, ci, cs[1], cs[2], span.p + nt0);
    plan_fused_init<<<(nt1 + 255) / 256, 256, 0, st>>>(p->fused_dep.p, nt1);
    plan_fused_deps<<<(nt1 + 255) / 256, 256, 0, st>>>(span.p + nt0, nt1, 0, cs[0], cs[1], nt0, p->fused_dep.p, stats.p);
    plan_fused_deps<<<(nt0 + 255) / 256, 256, 0, st>>>(span.p, nt0, 1, cs[1], cs[2], nt1, p->fused_dep.p, stats.p);
    plan_fused_stats<<<(nt1 + 255) / 256, 256, 0, st>>>(p->fused_dep.p, nt1, stats.p);
    GSB_KERNEL_CHECK();
    int h[4] = {0, 0, 0, 0};
    GSB_CUDA(cudaMemcpyAsync(h, stats.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h[0] != 0 || h[2] > GS_FUSED_DEP_WIDTH_MAX) {
        p->fused_dep.release();
        p->fused_flags.release();
        p->fused_span.release();
        return GSB_OK;
    }
    p->fused_lead_min = h[1] + 1 > 1 ? h[1] + 1 : 1;
    p->fused_ok = true;
    return GSB_OK;
}

#define GS_FUSED_L2HINT_DEFAULT 59 // measured (r02_call9): 0 -> 0.836, 1 -> 0.848, 3 -> 0.855, 7 -> 0.856 of peak; calls 17 / 18,
                                   // same box, alternating: 3 -> 0.834 / 0.833 / 0.842, 59 -> 0.857 / 0.838 / 0.848 (bits 8 + 16 + 32:
                                   // +0.5 .. 2.7 %); bit 64 (gather windows) loses 2 % (67: 0.824 / 0.837, 123: 0.822 / 0.821)
#define GS_FUSED_LEAD_DEFAULT 128 // beyond the dependency distance: long enough for the dependencies of a tile to be
                                  // finished when it comes up, short enough for the reuse to stay in L2 (measured)
#define GS_FUSED_PUBK_DEFAULT 1 // measured: batching delays availability by a whole item of the owning CTA (-13 %)
// tuning knobs of kernel 5 (defaults measured on B200, profiles/README.md); read once
struct FusedEnv {
    int ctas, lead, debug, pubk, hints;
};
static const FusedEnv &fused_env() {
    static FusedEnv e = {-1, 0, 0, 0, 0};
    if (e.ctas < 0) {
        const char *v = getenv("GSB_FUSED_DEBUG"); // measurement aid, WRONG RESULTS: 1 = no dependency waits, 2 = no flags
        e.debug = v ? atoi(v) : 0;
        v = getenv("GSB_FUSED_LEAD"); // tiles colour 0 runs ahead beyond the dependency distance
        e.lead = v ? atoi(v) : 0;
        v = getenv("GSB_FUSED_PUBK"); // colour-0 tiles published per release fence (1..4)
        e.pubk = v ? atoi(v) : 0;
        // 1: matrix / b copies evict-first; 2: x_old copies evict-last; 4: x stores evict-last; 8: release / acquire
        // fences instead of acq_rel; 16: L2 prefetch of the next refill's matrix spans; 32: ... and of its vectors;
        // 64: L2 prefetch of the x columns a colour-0 tile is the first to gather
        v = getenv("GSB_FUSED_L2HINT");
        e.hints = v ? atoi(v) : GS_FUSED_L2HINT_DEFAULT;
        v = getenv("GSB_RING_CTAS");
        e.ctas = v ? atoi(v) : 0;
    }
    return e;
}

// Once per solve, OUTSIDE any graph capture: reset the tile flags (ctl->sweeps restarts at 0) and (re)build the item
// table when the lead changed.
int gsb_plan_fused_reset(const GsbPlan *p, cudaStream_t st) {
    if (!p->fused_ok) return GSB_OK;
    GsbPlan *pm = const_cast<GsbPlan *>(p); // the table is a cache that belongs to (matrix, lead)
    const int nt0 = p->blocks[0], nt1 = p->blocks[1];
    GSB_CUDA(cudaMemsetAsync(p->fused_flags.p, 0, sizeof(int) * (size_t)nt0, st));
    const FusedEnv &env = fused_env();
    int lead = p->fused_lead_min + (p->fused_lead_extra > 0 ? p->fused_lead_extra : env.lead > 0 ? env.lead : GS_FUSED_LEAD_DEFAULT);
    if (lead > nt0) lead = nt0;
    if (pm->fused_items.p && pm->fused_items_lead == lead) return GSB_OK;
    GsbFusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.row0_0 = p->color_start[0];
    fa.row1_0 = fa.row0_1 = p->color_start[1];
    fa.row1_1 = p->color_start[2];
    fa.nt0 = nt0;
    fa.nt1 = nt1;
    fa.tile_k0 = p->tile_k.p + p->tile_off[0];
    fa.tile_k1 = p->tile_k.p + p->tile_off[1];
    fa.dep = reinterpret_cast<const int2 *>(p->fused_dep.p);
    fa.span = p->fused_span.p;
    fa.flags = p->fused_flags.p;
    fa.lead = lead;
    fa.nalt = nt1 < nt0 - lead ? nt1 : nt0 - lead;
    GSB_TRY(pm->fused_items.alloc((int64_t)(nt0 + nt1) * (int64_t)(sizeof(FusedItem) / sizeof(int4))));
    plan_fused_items<<<(nt0 + nt1 + 255) / 256, 256, 0, st>>>(fa, reinterpret_cast<FusedItem *>(pm->fused_items.p));
    GSB_KERNEL_CHECK();
    pm->fused_items_lead = lead;
    return GSB_OK;
}

template <int NRHS>
static int launch_fused_t(const GsbPlan *p, const int *rp, const int *ci, const double *va, const double *dg,
                          const double *b, double *x, int64_t ld, bool check, GsCtl *ctl, double *partials,
                          cudaStream_t st, int *slots) {
    typedef void (*fused_fn)(const int *, const int *, const double *, const double *, const double *, double *, int64_t,
                             int, GsCtl *, double *, const FusedItem *, int, int *, int, int, int);
    fused_fn kern = check ? (fused_fn)gs_sweep_fused<NRHS, true, GS_RING_STAGES_DEFAULT>
                          : (fused_fn)gs_sweep_fused<NRHS, false, GS_RING_STAGES_DEFAULT>;
    const int smem = 64 + GS_RING_STAGES_DEFAULT * ring_layout(p->cap, NRHS, check, 0).stage_bytes;
    int per_sm = 1;
    GSB_TRY(gsb_kernel_occupancy((const void *)kern, smem, &per_sm, GS_FUSED_THREADS));
    const FusedEnv &env = fused_env();
    if (env.ctas && env.ctas < per_sm) per_sm = env.ctas;
    const int nt0 = p->blocks[0], nt1 = p->blocks[1];
    int grid = gsb_sm_count() * per_sm; // every CTA must be resident: the walk waits on other CTAs' tiles
    if (grid > nt0 + nt1) grid = nt0 + nt1;
    if (grid > GSB_RING_SLOTS_MAX) grid = GSB_RING_SLOTS_MAX;
    if (!p->fused_items.p || p->fused_items_lead < 0) {
        gsb_set_error("internal: kernel 5 launched without gsb_plan_fused_reset");
        return GSB_ERR_STATE;
    }
    int pubk = env.pubk > 0 ? env.pubk : GS_FUSED_PUBK_DEFAULT;
    if (pubk > 4) pubk = 4;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(GS_FUSED_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gsb_pdl_mode((int64_t)p->color_start[2] - p->color_start[0]) ? 1 : 0;
    GSB_CUDA(cudaLaunchKernelEx(&cfg, kern, rp, ci, va, dg, b, x, ld, p->cap, ctl, partials,
                                (const FusedItem *)p->fused_items.p, nt0 + nt1, p->fused_flags.p, env.debug, pubk, env.hints));
    if (slots) *slots = grid;
    return GSB_OK;
}

int gsb_plan_launch_fused(const GsbPlan *p, const int *rp, const int *ci, const double *va, const double *dg,
                          const double *b, double *x, int64_t ld, int nrhs, bool check, GsCtl *ctl, double *partials,
                          cudaStream_t st, int *slots) {
    if (!p->fused_ok) {
        gsb_set_error("fused sweep kernel unavailable for this matrix (needs two colours and banded coupling)");
        return GSB_ERR_ARG;
    }
    switch (nrhs) {
        case 1: return launch_fused_t<1>(p, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, slots);
        case 2: return launch_fused_t<2>(p, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, slots);
        case 3: return launch_fused_t<3>(p, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, slots);
        case 4: return launch_fused_t<4>(p, rp, ci, va, dg, b, x, ld, check, ctl, partials, st, slots);
    }
    gsb_set_error("nrhs must be 1..%d", GSB_MAX_RHS);
    return GSB_ERR_ARG;
}
