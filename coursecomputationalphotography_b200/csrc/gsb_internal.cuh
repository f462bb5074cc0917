// gsb_internal.cuh -- shared declarations of libgsb200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/gsb200.h"

#define GSB_SM_COUNT_FALLBACK 148

void gsb_set_error(const char *fmt, ...);
int gsb_current_device();
cudaStream_t gsb_cur_stream();
cudaStream_t gsb_copy_stream(); // per device, for uploads that overlap work on the main stream
int gsb_sm_count();
int gsb_ensure_device(); // GSB_OK or GSB_ERR_NO_DEVICE

#define GSB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            gsb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
            return (e_ == cudaErrorMemoryAllocation) ? GSB_ERR_ALLOC : GSB_ERR_CUDA;                \
        }                                                                                           \
    } while (0)

#define GSB_TRY(call)                 \
    do {                              \
        int s_ = (call);              \
        if (s_ != GSB_OK) return s_;  \
    } while (0)

#define GSB_KERNEL_CHECK()                                                                          \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        if (e_ != cudaSuccess) {                                                                    \
            gsb_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_));\
            return GSB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// Device buffer with RAII; freed on the owning device.
void gsb_count_alloc(int frees); // gsb_prims.cu: process-wide counters behind gsb_alloc_counters()
template <typename T>
struct DevBuf {
    T *p = nullptr;
    int64_t n = 0;   // elements requested by the last alloc()
    int64_t cap = 0; // elements actually allocated (>= n)
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            cudaFree(p);
            gsb_count_alloc(1);
        }
        p = nullptr;
        n = 0;
        cap = 0;
    }
    // Contents are undefined after alloc().  An existing allocation is kept when it is large enough and not more
    // than twice what is asked for: re-importing / re-analysing a matrix of the same shape into the same handle
    // then costs no cudaFree + cudaMalloc of GB-sized buffers (each of which synchronises the device).
    int alloc(int64_t count) {
        if (count <= 0) count = 1;
        if (p && cap >= count && cap <= 2 * count + 4096) {
            n = count;
            return GSB_OK;
        }
        release();
        cudaError_t e = cudaMalloc((void **)&p, sizeof(T) * (size_t)count);
        if (e != cudaSuccess) {
            p = nullptr;
            gsb_set_error("cudaMalloc(%lld bytes) -> %s", (long long)(sizeof(T) * (size_t)count),
                          cudaGetErrorString(e));
            cudaGetLastError();
            return GSB_ERR_ALLOC;
        }
        gsb_count_alloc(0);
        n = count;
        cap = count;
        return GSB_OK;
    }
    void swap(DevBuf &o) {
        T *tp = p; p = o.p; o.p = tp;
        int64_t tn = n; n = o.n; o.n = tn;
        int64_t tc = cap; cap = o.cap; o.cap = tc;
    }
};

// ---- device-wide primitives (gsb_prims.cu) ---------------------------------------------
// exclusive prefix sum of int32 -> int32 (out may alias in); *total_dev (optional, device) gets the sum
int gsb_exclusive_scan_i32(const int *in, int *out, int64_t n, int *total_dev, cudaStream_t st, int *scratch = nullptr,
                           int64_t scratch_ints = 0); // with enough scratch (gsb_scan_scratch_ints): no allocation, no sync
int64_t gsb_scan_scratch_ints(int64_t n);
// deterministic sum reduction helpers: result written to out_dev[0]
int gsb_reduce_max_i32(const int *in, int64_t n, int *out_dev, cudaStream_t st);
int gsb_l1_dist_dev(const double *a, const double *b, int64_t n, double *out_dev, cudaStream_t st);
int gsb_dot_dev(const double *a, const double *b, int64_t n, double *out_dev, cudaStream_t st);
// scratch for the deterministic reductions (per device, grown on demand)
double *gsb_reduce_scratch(int64_t n_doubles);

static inline int gsb_blocks_for(int64_t n, int threads, int max_blocks = 1 << 30) {
    int64_t b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

#define GSB_DIST_MAX_WORLD_DECL 16
struct gsb_dist_group;
// process-wide device list of the host entry points (gsb_set_devices / GSB_DEVICES); returns the count
int gsb_devices(int *out, int cap);

// ---- the matrix handle -------------------------------------------------------------------
struct gsb_matrix {
    int vtype = GSB_F64;
    int device = 0;
    // reference layout ("slack CSR"), device resident, natural row order.
    // values are kept in the caller's element type for bit-exact download, plus an FP64 view
    // for the solvers (same buffer when vtype == GSB_F64).
    DevBuf<unsigned char> values_raw; // store * elem_size
    DevBuf<double> values_f64;        // only when vtype == GSB_I32
    DevBuf<int> cols, row_begin, row_nnz, row_left;
    int64_t store = 0;
    int n_rows = 0, n_cols = 0;
    int64_t nnz = 0; // sum(row_nnz)
    bool has_layout = false;

    const double *vals() const {
        return vtype == GSB_F64 ? (const double *)values_raw.p : values_f64.p;
    }

    // solver format: colour-major compact CSR (gsb_matrix.cu)
    bool analyzed = false;
    int ordering_used = 0;
    int n_colors = 0;
    int grid_width = 0;
    int color_start[66];   // row ranges per colour in permuted order
    DevBuf<int> perm;      // perm[new] = old
    DevBuf<int> iperm;     // iperm[old] = new
    DevBuf<int> colors;    // colour[old]
    DevBuf<int> rp;        // n_rows+1
    DevBuf<int> ci;        // off-diagonal entries only (permuted column ids, ascending per row)
    DevBuf<double> va;     // their values
    DevBuf<double> dg;     // dg[new row] = diagonal value, 0 when the row stores none (row skipped, v2 :360-363)
    int64_t nnz_off = 0;   // entries in ci/va
    // staged-kernel tiling (gsb_solve.cu)
    DevBuf<int4> tiles;    // per tile: {row0, nrows, nnz0_aligned, nnz_count_aligned}
    int tile_rows = 0;
    int tiles_per_color[66];
    int tile_start[66];
    int max_row_nnz = 0;

    // solver workspaces (lazily sized for nrhs)
    DevBuf<double> xw, bw; // permuted x and b, nrhs * n_rows
    int ws_nrhs = 0;
    DevBuf<double> stage_b, stage_x; // natural-order device copies of host b / x for the host-pointer entry points
    DevBuf<int> tiny;                // 64 ints of device scratch in fixed slots (GSB_TINY_*) for the few-word results of
                                     // import / analysis kernels: a steady-state import + analysis + solve then makes no
                                     // cudaMalloc / cudaFree at all (each is a trip into the driver that, on a box shared
                                     // with other processes, was seen to take 100-900 ms every few calls)
    DevBuf<int> scan_scratch;        // block sums of the analysis' prefix scans (gsb_scan_scratch_ints(n_rows + 1))
    DevBuf<int> scratch_rows;        // n_rows + 1 ints of scratch for import / analysis (kept: a re-import of the same
                                     // shape then makes no GB-scale cudaMalloc / cudaFree, each a device-wide sync)
    void *ev_t0 = nullptr, *ev_t1 = nullptr; // cudaEvent_t pair timing the sweep loop (kept: a small solve is a few hundred us)
    void *b_ready_event = nullptr;   // cudaEvent_t: stage_b's upload on the copy stream (host entry point, first solve)
    bool b_upload_pending = false;   // the solver core has to wait for b_ready_event before it reads stage_b
    DevBuf<double> partials;   // per-block partial sums of the stop rule
    DevBuf<unsigned char> ctl; // GsCtl
    DevBuf<unsigned> small_bar; // kernel 6: grid-barrier counter + generation
    void *ctl_host = nullptr;  // pinned mirror
    // cached CUDA graph of one batch
    void *graph_exec = nullptr;
    int graph_key[6] = {0, 0, 0, 0, 0, 0};
    struct GsbPlan *plan = nullptr; // colour-phase launch plan (gsb_phase.cu)
    // conjugate-gradient workspaces (gsb_cg.cu): vectors, per-CTA partials, the device-resident scalar block
    DevBuf<double> cg_ws, cg_partials;
    DevBuf<unsigned char> cg_state;
    void *cg_state_host = nullptr;
    // multi-device solve through the host entry points (gsb_set_devices): the group and whether it holds this matrix
    struct gsb_dist_group *group = nullptr;
    bool group_built = false;
    int group_key[GSB_DIST_MAX_WORLD_DECL + 1] = {0}; // device list the group was made for ([0] = count)

    void drop_analysis();
    ~gsb_matrix();
};

int gsb_matrix_finish_layout(gsb_matrix *m); // compute nnz, f64 view after the five arrays are set
// slots of gsb_matrix::tiny (ints; 8-byte values on even slots)
#define GSB_TINY_FIRST 0
#define GSB_TINY_TOT64 2
#define GSB_TINY_COLOR_TOT 4
#define GSB_TINY_INFO 8
#define GSB_TINY_INTS 64
static inline int gsb_tiny_alloc(gsb_matrix *m) { return m->tiny.alloc(GSB_TINY_INTS); }

// gsb_assembly.cu: sorted COO (device pointers) -> slack CSR in m
template <typename T>
int gsb_assemble_sorted_device(gsb_matrix *m, const int *d_rows, const int *d_cols_in, const T *d_vals_in, int64_t n,
                               int n_rows, int n_cols_override);

// ---- solver control block + launchers shared by the single-GPU and the strip solver -------
#define GSB_MAX_RHS 4
struct GsCtl {
    int done;
    int sweeps;
    int max_iter;
    int check_every;
    double epsilon;
    double eps_last[GSB_MAX_RHS];
    int error; // 1: the peer stop-rule exchange timed out (a rank is missing); the solve stops
    int ticket; // fused end of sweep: CTAs of the sweep's last colour phase that have retired
    double eps_local[GSB_MAX_RHS]; // strip solver, ncclAllReduce path: this rank's share of eps_last
};

// Ring kernels: a colour phase writes one stop-rule partial per CTA into at most GSB_RING_SLOTS_MAX slots (the
// end-of-sweep kernel folds all of them, so the bound is kept close to the largest grid a B200 runs: 148 SMs x 4
// resident CTAs = 592, plus the halo CTAs of the strip solver, at most GSB_HALO_TILES_MAX per colour).
#define GSB_RING_SLOTS_MAX 1024
#define GSB_HALO_TILES_MAX 256

// how the colour phases of one colour-major CSR are launched (gsb_phase.cu)
struct GsbPlan {
    bool valid = false;
    int requested = 0;  // kernel the caller asked for (0 auto)
    int kernel = 1;     // resolved: 1 = row-per-thread direct, 2 = bulk-copy staged tiles
    int tile_rows = 256;
    int cap = 0;        // shared-memory capacity of one tile, in CSR entries
    int smem_bytes = 0;
    int n_colors = 0;
    int blocks[66];     // CTAs (== partial-sum slots) per colour phase
    int tile_off[66];   // offset of the colour's segment in tile_k
    int color_start[66];
    DevBuf<int> tile_k; // per colour: CSR offset at every tile boundary (blocks[c] + 1 entries)
    // kernel 4: per tile up to 4 windows of x columns (64-column granules) that cover every gather of the tile
    DevBuf<int> tile_win; // 12 ints per tile: {nwin, lo[4], len[4], pad[3]}; same indexing as partial slots
    DevBuf<int> ci_slot;  // kernel 4's index array: shared-memory slots for window tiles (-1 = diagonal)
    int win_off[66];      // tile index of the colour's first tile
    int wcap = 0;         // doubles per right-hand side reserved per stage for the windows
    // kernel 5 (two colours, one launch per sweep): per tile of colour 1 the range of colour-0 tiles it has to see
    // finished (the ones it reads, and the ones that read its rows' old values); per tile of colour 0 a flag
    int64_t nnz_hint = 0;       // stored entries of the matrix (set by the single-GPU solver: kernel 2's L2-hint policy)
    DevBuf<int> tiny;           // 32 ints of scratch for the plan kernels' few-word results (slots 0, 4, 8), kept
    bool fused_allowed = false; // set by the single-GPU solver before gsb_plan_build (strip plans never fuse)
    bool fused_ok = false;
    int fused_lead_min = 0;     // colour 0 has to run at least this many tiles ahead of colour 1
    int fused_lead_extra = 0;   // caller's gsb_gs_options.fused_lead (0 = default)
    DevBuf<int> fused_dep;      // 2 ints per colour-1 tile: first, last colour-0 tile
    DevBuf<int> fused_flags;    // per colour-0 tile: number of the last sweep that updated it (reset per solve)
    DevBuf<int2> fused_span;    // per tile (colour 0 first): smallest / largest column among its off-diagonal entries
    DevBuf<int4> fused_items;   // per work item of the sweep, sequence order: two int4 (FusedItem, gsb_fused.cu)
    int fused_items_lead = -1;  // the lead the table was built for
    int total_blocks() const;
};
int gsb_plan_build(GsbPlan *p, const int *rp, const int *ci, const int *color_start, int n_colors,
                   int kernel_request, cudaStream_t st);
int gsb_plan_effective_kernel(const GsbPlan *p, int nrhs);
// kernel 5: the whole sweep (both colours) in one launch; *slots = stop-rule partial slots written (== grid size).
// gsb_plan_fused_reset must be enqueued once before the first sweep of every solve (ctl->sweeps restarts at 0).
int gsb_plan_fused_reset(const GsbPlan *p, cudaStream_t st);
int gsb_plan_launch_fused(const GsbPlan *p, const int *rp, const int *ci, const double *va, const double *dg,
                          const double *b, double *x, int64_t ld, int nrhs, bool check, GsCtl *ctl, double *partials,
                          cudaStream_t st, int *slots);
int gsb_plan_partial_slots(const GsbPlan *p, int c, int nrhs); // stop-rule partial slots of colour phase c
// programmatic dependent launch of the ring kernels / gs_end_sweep (GSB_PDL=0 disables; suppressed during graph capture)
bool gsb_pdl_enabled();
int gsb_kernel_occupancy(const void *kern, int smem, int *occ, int threads = 256); // resident CTAs per SM (cached, thread-safe)
int gsb_pdl_mode(int64_t phase_rows); // decision for the launch about to be made (remembered for gs_end_sweep)
int gsb_plan_build_fused(GsbPlan *p, const int *rp, const int *ci, cudaStream_t st); // gsb_fused.cu
void gsb_pdl_suppress(int on);
static inline int64_t gsb_padded_ld(int64_t n) { return (n + 1) & ~(int64_t)1; }
// Fused halo exchange of the strip solver (gsb_dist.cu): the phase kernel itself writes the boundary values a
// neighbour GPU reads straight into that neighbour's ghost slots (peer-mapped memory over NVLink) and
// raises a flag there; tiles that read ghosts are processed first and wait on the flag the neighbour
// raised in its previous phase.  Passed by value to the ring kernels; enabled == 0 on a single GPU.
struct GsbHaloArgs {
    int enabled;
    int n_halo_tiles;          // tiles (first in `order`) that read ghosts and/or own rows a neighbour reads
    const int *order;          // (host-side bookkeeping; the kernels number the tiles arithmetically)
    const unsigned char *info; // per tile: bit0 reads ghosts, bit1 has rows to push (likewise)
    const int *push_map[2];    // per neighbour: row (permuted local index) -> slot in the neighbour's ghost range, or -1
    double *peer_x[2];         // neighbour's x workspace (peer mapping)
    long long peer_ld[2];
    int peer_gs[2];            // start of the neighbour's ghost range that holds this colour's values from this rank
    int *peer_flag[2];         // neighbour's flag to raise once every halo tile of this phase is done
    const int *wait_flag[2];   // own flags the neighbours raise: the other colour's values have arrived
    int wait_epoch, signal_epoch;
    int *counter;              // halo tiles finished in this phase
    int has_peer[2];
    int pdl_early;             // set by the launcher: stage the first tiles before griddepcontrol.wait (kernel 3)
    int interior_base;         // the tiles that are not halo tiles: [interior_base, interior_base + n_interior)
    int n_interior;
};

// (layout of the stop-rule exchange: see gsb_launch_end_sweep_peer below)
#define GSB_DIST_MAX_WORLD 16
struct GsbEpsExchange {
    int world, rank;
    int epoch;  // exchanges issued so far + 1 (monotonic over the lifetime of the handle); parity = epoch & 1
    double *box[GSB_DIST_MAX_WORLD];
};

// Fused end of sweep (opt-in: GSB_FUSED_END=1; ring kernels only).  The launch of the sweep's last colour phase
// carries these arguments; every CTA takes a ticket when it retires and the last one does what gs_end_sweep /
// gs_end_sweep_peer would do in a kernel of their own: fold the partial slots of all colour phases, (strip
// solver) exchange the sums with the peers, bump the sweep counter, decide.  Saves one launch and one
// drain / fill of the GPU per sweep.  Not yet measured (written after round 1's GPU budget was spent).
struct GsbEndArgs {
    int enabled;
    int checked;            // the stop rule is evaluated this sweep
    int n_partials;         // slots of all colour phases of the sweep, this launch's included
    int exchange;           // 1: strip solver, sums are exchanged through `ex` (GsbEpsExchange)
    GsCtl *ctl;
    const double *partials; // slot 0 of the sweep's first colour phase
    GsbEpsExchange ex;
};
bool gsb_fused_end_enabled();                          // GSB_FUSED_END=1
bool gsb_fused_end_enabled_strips();                   // strip solver: on unless GSB_FUSED_END=0
bool gsb_plan_can_fuse_end(const GsbPlan *p, int nrhs); // the effective kernel is a ring kernel

// one colour phase; x and b have leading dimension ld; partials: blocks[c] * nrhs doubles
// rp/ci/va: off-diagonal CSR in colour-major order; dg: the diagonal (0 = row skipped)
int gsb_plan_launch(const GsbPlan *p, int c, const int *rp, const int *ci, const double *va, const double *dg,
                    const double *b, double *x, int64_t ld, int nrhs, bool check, const GsCtl *ctl, double *partials,
                    cudaStream_t st, const GsbHaloArgs *halo = nullptr, const GsbEndArgs *end = nullptr);
// kernel 6 (gsb_small.cu): small systems, up to `max_sweeps` sweeps incl. the stop rule in one persistent launch
bool gsb_small_auto(int64_t n_rows, int n_colors, int check_every);
int gsb_launch_small_persistent(const int *rp, const int *ci, const double *va, const double *dg, const double *b, double *x,
                                int64_t ld, int nrhs, const int *color_start, int n_colors, GsCtl *ctl, double *partials,
                                unsigned *bar, int max_sweeps, cudaStream_t st, int *slots);
// end of sweep.  mode 0: fold partials, bump the counter, decide (single GPU)
//                mode 1: fold partials into ctl->eps_last only (strip solver, before the all-reduce)
//                mode 2: bump the counter and decide from ctl->eps_last (after the all-reduce)
int gsb_launch_end_sweep(GsCtl *ctl, const double *partials, int n_partials, int nrhs, int checked, int mode,
                         cudaStream_t st);
// Strip solver: end of a checked sweep with the all-reduce of the stop rule fused in.  Every rank owns a small
// "box" in device memory that all ranks have peer-mapped: [2 parities][world][GSB_MAX_RHS] doubles followed by
// [2 parities][world] int flags.  The kernel folds the local partials, stores the k sums into slot `rank` of every
// rank's box (NVLink peer stores) and raises that slot's flag (release, system scope) to `epoch`; it then waits for
// all `world` flags of its own box, adds the slots in rank order (identical on every rank -> identical decision),
// bumps the sweep counter and decides.  One launch replaces fold + ncclAllReduce + decide.
int gsb_launch_end_sweep_peer(GsCtl *ctl, const double *partials, int n_partials, int nrhs, const GsbEpsExchange *ex,
                              cudaStream_t st);

// device-pointer solver cores shared with the gradient-domain-fusion driver (gsb_gdf.cu)
int gsb_gs_solve_device_x0(gsb_matrix *m, const double *b_dev, const double *x0_dev, int nrhs, double epsilon,
                           int max_iteration, const gsb_gs_options *opts, double *x_dev, gsb_gs_stats *stats);
int gsb_cg_solve_device(gsb_matrix *m, const double *b_dev, const double *x0_dev, double epsilon, int max_iteration,
                        double *x_dev, int *iters);
int gsb_cg_solve_device_multi(gsb_matrix *m, const double *b_dev, const double *x0_dev, int nrhs, double epsilon,
                              int max_iteration, double *x_dev, int *iters);

// gsb_poisson.cu: rows [p0,p1) of the reference's Poisson matrix (global columns)
int gsb_poisson_launch_row_len(int W, int H, int64_t p0, int64_t p1, int *len, cudaStream_t st);
int gsb_poisson_launch_fill(int W, int H, int64_t p0, int64_t p1, const int *rp, int *ci, double *va,
                            cudaStream_t st);

#ifdef __CUDACC__
// fixed-order block reduction of NRHS values per thread -> out[0..NRHS) (deterministic partials)
template <int NRHS, int THREADS>
__device__ __forceinline__ void gsb_block_reduce_store(double (&v)[NRHS], double *__restrict__ out) {
    __shared__ double ws[NRHS][THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NRHS; ++r) {
        double t = v[r];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
        if (lane == 0) ws[r][wid] = t;
    }
    __syncthreads();
    if (threadIdx.x < NRHS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) s += ws[threadIdx.x][w];
        out[threadIdx.x] = s;
    }
}
#endif
