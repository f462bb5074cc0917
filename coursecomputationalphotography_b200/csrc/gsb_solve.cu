// gsb_solve.cu -- Gauss-Seidel sweeps, SpMV and the residual on the device.
//
// gaussSeidel (v2 :350-380) becomes: x = 1; repeat { for each colour: one phase kernel over the
// colour's contiguous row range of the colour-major CSR; gs_end_sweep folds the per-block partial
// L1 sums of the sweep's update in a fixed order, bumps the sweep counter and raises `done` }.
// The stop decision lives on the device (GsCtl); the host enqueues batches of sweeps and reads
// one small struct per batch.  After `done` is raised the remaining kernels of the batch return
// at their first instruction, so the returned iterate is exactly the one the stop rule accepted.
//
// Arithmetic: products and sums are rounded separately (__dmul_rn/__dadd_rn, never an FMA) in
// storage order, as the reference's x64 build does, so a sweep is bit-identical to the
// reference's sweep over the permuted matrix.
#include "gsb_internal.cuh"

#include <cooperative_groups.h>
#include <math.h>
#include <new>

#define MAX_RHS GSB_MAX_RHS

// ---------------------------------------------------------------------------------------------
// vector permutation helpers
// ---------------------------------------------------------------------------------------------
// natural order (stride n) -> colour-major workspace (stride ld)
__global__ void __launch_bounds__(256) gather_perm(const double *__restrict__ src, const int *__restrict__ perm,
                                                   int64_t n, int64_t ld, int nrhs, double *__restrict__ dst) {
    int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    int o = perm[p];
    for (int r = 0; r < nrhs; ++r) dst[r * ld + p] = src[r * n + o];
}

__global__ void __launch_bounds__(256) scatter_perm(const double *__restrict__ src, const int *__restrict__ perm,
                                                    int64_t n, int64_t ld, int nrhs, double *__restrict__ dst) {
    int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= n) return;
    int o = perm[p];
    for (int r = 0; r < nrhs; ++r) dst[r * n + o] = src[r * ld + p];
}

__global__ void __launch_bounds__(256) fill_f64(double *__restrict__ p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
extern "C" void gsb_gs_default_options(gsb_gs_options *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->ordering = GSB_ORDER_AUTO;
    o->check_every = 1;
    o->batch_sweeps = 0;
    o->use_graph = -1;
    o->kernel = 0;
    o->compute_residual = 0;
}

static int enqueue_sweep(gsb_matrix *m, int nrhs, bool check, cudaStream_t st, int64_t *launches) {
    GsCtl *ctl = (GsCtl *)m->ctl.p;
    const int64_t n = gsb_padded_ld(m->n_rows);
    if (gsb_plan_effective_kernel(m->plan, nrhs) == 5) { // both colours in one launch (gsb_fused.cu)
        int slots = 0;
        GSB_TRY(gsb_plan_launch_fused(m->plan, m->rp.p, m->ci.p, m->va.p, m->dg.p, m->bw.p, m->xw.p, n, nrhs, check, ctl,
                                      m->partials.p, st, &slots));
        GSB_TRY(gsb_launch_end_sweep(ctl, m->partials.p, slots, nrhs, check ? 1 : 0, 0, st));
        *launches += 2;
        return GSB_OK;
    }
    // opt-in (GSB_FUSED_END=1): the last non-empty colour phase ends the sweep itself (GsbEndArgs)
    int carrier = -1, total = 0;
    for (int c = 0; c < m->n_colors; ++c)
        if (m->plan->blocks[c] > 0) {
            carrier = c;
            total += gsb_plan_partial_slots(m->plan, c, nrhs);
        }
    const bool fuse = carrier >= 0 && gsb_fused_end_enabled() && gsb_plan_can_fuse_end(m->plan, nrhs);
    int poff = 0;
    for (int c = 0; c < m->n_colors; ++c) {
        const int nb = m->plan->blocks[c];
        if (nb == 0) continue;
        GsbEndArgs ea;
        memset(&ea, 0, sizeof(ea));
        if (fuse && c == carrier) {
            ea.enabled = 1;
            ea.checked = check ? 1 : 0;
            ea.n_partials = total;
            ea.ctl = ctl;
            ea.partials = m->partials.p;
        }
        GSB_TRY(gsb_plan_launch(m->plan, c, m->rp.p, m->ci.p, m->va.p, m->dg.p, m->bw.p, m->xw.p, n, nrhs, check, ctl,
                                m->partials.p + (size_t)poff * nrhs, st, nullptr, ea.enabled ? &ea : nullptr));
        poff += gsb_plan_partial_slots(m->plan, c, nrhs);
        ++*launches;
    }
    if (!fuse) {
        GSB_TRY(gsb_launch_end_sweep(ctl, m->partials.p, poff, nrhs, check ? 1 : 0, 0, st));
        ++*launches;
    }
    return GSB_OK;
}

static void drop_graph(gsb_matrix *m) {
    if (m->graph_exec) {
        cudaGraphExecDestroy((cudaGraphExec_t)m->graph_exec);
        m->graph_exec = nullptr;
    }
}

static int ensure_workspace(gsb_matrix *m, int nrhs, int kernel_request, cudaStream_t st) {
    const int64_t n = m->n_rows;
    if (!m->plan) m->plan = new (std::nothrow) GsbPlan();
    if (!m->plan) return GSB_ERR_ALLOC;
    if (!m->plan->valid || m->plan->requested != kernel_request) {
        m->plan->fused_allowed = true;
        m->plan->nnz_hint = m->nnz;
        GSB_TRY(gsb_plan_build(m->plan, m->rp.p, m->ci.p, m->color_start, m->n_colors, kernel_request, st));
        GSB_TRY(m->partials.alloc((int64_t)(m->plan->total_blocks() + 1 + 64) * MAX_RHS)); // +64: second-level fold
        drop_graph(m);
    }
    if (m->ws_nrhs < nrhs) {
        const int64_t ld = gsb_padded_ld(n); // even stride: every plane starts 16-byte aligned
        GSB_TRY(m->xw.alloc(ld * nrhs + 128)); // +128: aligned bulk copies / 64-column windows may over-read
        GSB_TRY(m->bw.alloc(ld * nrhs + 128));
        m->ws_nrhs = nrhs;
        drop_graph(m);
    }
    if (!m->ctl.p) GSB_TRY(m->ctl.alloc(sizeof(GsCtl)));
    if (!m->ctl_host) GSB_CUDA(cudaHostAlloc(&m->ctl_host, sizeof(GsCtl), cudaHostAllocDefault));
    return GSB_OK;
}

// b_dev/x0_dev/x_dev: natural order device vectors (x0_dev may be null -> 1.0)
static int gs_solve_device(gsb_matrix *m, const double *b_dev, const double *x0_dev, int nrhs, double epsilon,
                           int max_iteration, const gsb_gs_options *opts_in, double *x_dev, gsb_gs_stats *stats) {
    gsb_gs_options opts;
    gsb_gs_default_options(&opts);
    if (opts_in) opts = *opts_in;
    if (opts.check_every < 1) opts.check_every = 1;
    cudaStream_t st = gsb_cur_stream();
    double setup_ms = 0.0;
    if (!m->ev_t0) { // one pair of timing events per handle (analysis, then the sweep loop)
        cudaEvent_t a = nullptr, b2 = nullptr;
        GSB_CUDA(cudaEventCreate(&a));
        m->ev_t0 = a;
        GSB_CUDA(cudaEventCreate(&b2));
        m->ev_t1 = b2;
    }
    if (!m->analyzed) {
        cudaEvent_t e0 = (cudaEvent_t)m->ev_t0, e1 = (cudaEvent_t)m->ev_t1;
        GSB_CUDA(cudaEventRecord(e0, st));
        GSB_TRY(gsb_matrix_analyze(m, opts.ordering, nullptr));
        GSB_CUDA(cudaEventRecord(e1, st));
        GSB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        setup_ms = ms;
    }
    const int64_t n = m->n_rows;
    GSB_TRY(ensure_workspace(m, nrhs, opts.kernel, st));
    if (m->plan->fused_lead_extra != opts.fused_lead) {
        m->plan->fused_lead_extra = opts.fused_lead;
        drop_graph(m); // the lead is a launch argument baked into a captured batch
    }
    const int nbv = (int)((n + 255) / 256);
    const int64_t ld = gsb_padded_ld(n);
    if (m->b_upload_pending) { // the host entry point uploaded b on the copy stream while the analysis ran here
        GSB_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)m->b_ready_event, 0));
        m->b_upload_pending = false;
    }
    gather_perm<<<nbv, 256, 0, st>>>(b_dev, m->perm.p, n, ld, nrhs, m->bw.p);
    GSB_KERNEL_CHECK();
    if (x0_dev)
        gather_perm<<<nbv, 256, 0, st>>>(x0_dev, m->perm.p, n, ld, nrhs, m->xw.p);
    else
        fill_f64<<<gsb_blocks_for(ld * nrhs, 256 * 4, gsb_sm_count() * 16), 256, 0, st>>>(m->xw.p, ld * nrhs,
                                                                                      1.0); // v2 :352
    GSB_KERNEL_CHECK();
    int64_t launches = 2;

    GsCtl h;
    memset(&h, 0, sizeof(h));
    h.max_iter = max_iteration;
    h.check_every = opts.check_every;
    h.epsilon = epsilon;
    for (int r = 0; r < MAX_RHS; ++r) h.eps_last[r] = 10.0;          // v2 :354
    h.done = !(10.0 > epsilon && 0 < max_iteration) ? 1 : 0;          // v2 :356
    GsCtl *hp = (GsCtl *)m->ctl_host;
    *hp = h;
    GSB_CUDA(cudaMemcpyAsync(m->ctl.p, hp, sizeof(GsCtl), cudaMemcpyHostToDevice, st));
    GSB_TRY(gsb_plan_fused_reset(m->plan, st)); // kernel 5's tile flags count sweeps from 0 again

    int batch = opts.batch_sweeps;
    if (batch <= 0) {
        // aim at ~2 ms of sweeps per host round trip, assuming ~3 TB/s effective
        double bytes = 12.0 * (double)m->nnz + (4.0 + 24.0 * nrhs) * (double)n;
        double est_ms = bytes / 3.0e9 + 0.004 * (m->n_colors + 1);
        batch = (int)(2.0 / est_ms);
        if (batch < 4) batch = 4;
        if (batch > 256) batch = 256;
    }
    bool use_graph = opts.use_graph == 1 || (opts.use_graph == -1 && n * (int64_t)m->n_colors < (int64_t)1 << 20);
    if (opts.check_every != 1) use_graph = false; // keep the captured batch simple: one sweep shape
    // small systems: the whole solve in one persistent launch (kernel 6) instead of ~(colours + 1) graph nodes per sweep
    const bool use_small = (opts.kernel == 6 || (opts.kernel == 0 && opts.use_graph == -1)) &&
                           gsb_small_auto(n, m->n_colors, opts.check_every);
    if (opts.kernel == 6 && !use_small) {
        gsb_set_error("kernel 6 (persistent small-system kernel) needs check_every = 1 and at most 64 colours");
        return GSB_ERR_ARG;
    }
    if (use_small) {
        use_graph = false;
        if (!m->small_bar.p) GSB_TRY(m->small_bar.alloc(2));
        GSB_CUDA(cudaMemsetAsync(m->small_bar.p, 0, 2 * sizeof(unsigned), st)); // (a solve that gave up may have left it mid-count)
    }

    cudaEvent_t ev0 = (cudaEvent_t)m->ev_t0, ev1 = (cudaEvent_t)m->ev_t1;
    GSB_CUDA(cudaEventRecord(ev0, st));
    int status = GSB_OK;
    int issued = 0;
    while (!h.done && status == GSB_OK) {
        int todo = max_iteration - issued;
        if (todo > batch) todo = batch;
        if (todo <= 0) todo = 1; // cannot happen (done would be set) -- guards an endless loop
        if (use_small) {
            int slots = 0;
            status = gsb_launch_small_persistent(m->rp.p, m->ci.p, m->va.p, m->dg.p, m->bw.p, m->xw.p, ld, nrhs, m->color_start,
                                                 m->n_colors, (GsCtl *)m->ctl.p, m->partials.p, m->small_bar.p,
                                                 max_iteration - issued, st, &slots);
            ++launches;
            issued = max_iteration; // the launch runs until the stop rule fires or max_iteration is reached
        } else if (use_graph) {
            int key[6] = {nrhs, batch, 1, gsb_plan_effective_kernel(m->plan, nrhs), m->n_colors, 1};
            if (!m->graph_exec || memcmp(key, m->graph_key, sizeof(key)) != 0) {
                if (m->graph_exec) {
                    cudaGraphExecDestroy((cudaGraphExec_t)m->graph_exec);
                    m->graph_exec = nullptr;
                }
                cudaGraph_t g = nullptr;
                int64_t dummy = 0;
                GSB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                gsb_pdl_suppress(1); // plain kernel nodes inside the graph
                for (int s = 0; s < batch && status == GSB_OK; ++s) status = enqueue_sweep(m, nrhs, true, st, &dummy);
                gsb_pdl_suppress(0);
                cudaError_t ce = cudaStreamEndCapture(st, &g);
                if (status != GSB_OK) break;
                if (ce != cudaSuccess) {
                    gsb_set_error("graph capture failed: %s", cudaGetErrorString(ce));
                    status = GSB_ERR_CUDA;
                    break;
                }
                cudaGraphExec_t ge = nullptr;
                ce = cudaGraphInstantiate(&ge, g, 0);
                cudaGraphDestroy(g);
                if (ce != cudaSuccess) {
                    gsb_set_error("graph instantiate failed: %s", cudaGetErrorString(ce));
                    status = GSB_ERR_CUDA;
                    break;
                }
                m->graph_exec = ge;
                memcpy(m->graph_key, key, sizeof(key));
            }
            // the graph always holds `batch` sweeps; max_iter is enforced on the device
            cudaError_t ce = cudaGraphLaunch((cudaGraphExec_t)m->graph_exec, st);
            if (ce != cudaSuccess) {
                gsb_set_error("graph launch failed: %s", cudaGetErrorString(ce));
                status = GSB_ERR_CUDA;
                break;
            }
            launches += (int64_t)batch * (m->n_colors + 1);
            issued += batch;
        } else {
            for (int s = 0; s < todo && status == GSB_OK; ++s) {
                bool check = ((issued + s + 1) % opts.check_every) == 0 || (issued + s + 1) == max_iteration;
                status = enqueue_sweep(m, nrhs, check, st, &launches);
            }
            issued += todo;
        }
        if (status != GSB_OK) break;
        cudaError_t ce = cudaMemcpyAsync(hp, m->ctl.p, sizeof(GsCtl), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {
            gsb_set_error("sweep batch failed: %s", cudaGetErrorString(ce));
            status = GSB_ERR_CUDA;
            break;
        }
        h = *hp;
        if (h.error) {
            gsb_set_error("gauss_seidel: the sweep kernel gave up waiting (code %d)", h.error);
            status = GSB_ERR_CUDA;
        }
    }
    cudaEventRecord(ev1, st);
    cudaEventSynchronize(ev1);
    float solve_ms = 0.f;
    cudaEventElapsedTime(&solve_ms, ev0, ev1);
    if (status != GSB_OK) return status;

    scatter_perm<<<nbv, 256, 0, st>>>(m->xw.p, m->perm.p, n, ld, nrhs, x_dev);
    GSB_KERNEL_CHECK();
    ++launches;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->sweeps = h.sweeps;
        stats->n_colors = m->n_colors;
        stats->ordering_used = m->ordering_used;
        stats->kernel_used = use_small ? 6 : gsb_plan_effective_kernel(m->plan, nrhs);
        stats->kernel_launches = launches;
        for (int r = 0; r < MAX_RHS; ++r) stats->last_eps[r] = r < nrhs ? h.eps_last[r] : 0.0;
        stats->solve_ms = solve_ms;
        stats->setup_ms = setup_ms;
        if (opts.compute_residual) {
            for (int r = 0; r < nrhs; ++r)
                GSB_TRY(gsb_residual_l2_dev(m, b_dev + r * n, x_dev + r * n, &stats->residual_l2[r]));
        }
    }
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

static int gs_check_args(gsb_matrix *m, const double *b, int nrhs, double *x) {
    if (!m || !b || !x || nrhs < 1 || nrhs > MAX_RHS) {
        gsb_set_error("gauss_seidel: bad argument (nrhs must be 1..%d)", MAX_RHS);
        return GSB_ERR_ARG;
    }
    if (!m->has_layout) {
        gsb_set_error("gauss_seidel: matrix holds no layout yet");
        return GSB_ERR_STATE;
    }
    if (m->n_rows != m->n_cols) {
        gsb_set_error("gauss_seidel: matrix must be square (have %d x %d)", m->n_rows, m->n_cols);
        return GSB_ERR_SHAPE;
    }
    return gsb_set_device(m->device);
}

extern "C" int gsb_gauss_seidel_dev(gsb_matrix *m, const double *b_dev, int nrhs, double epsilon, int max_iteration,
                                    const gsb_gs_options *opts, double *x_dev, gsb_gs_stats *stats) {
    GSB_TRY(gs_check_args(m, b_dev, nrhs, x_dev));
    return gs_solve_device(m, b_dev, nullptr, nrhs, epsilon, max_iteration, opts, x_dev, stats);
}

// device pointers with an initial guess (x0_dev may alias x_dev, or be null for the reference's 1.0): used by the
// gradient-domain-fusion driver (gsb_gdf.cu)
int gsb_gs_solve_device_x0(gsb_matrix *m, const double *b_dev, const double *x0_dev, int nrhs, double epsilon,
                           int max_iteration, const gsb_gs_options *opts, double *x_dev, gsb_gs_stats *stats) {
    GSB_TRY(gs_check_args(m, b_dev, nrhs, x_dev));
    return gs_solve_device(m, b_dev, x0_dev, nrhs, epsilon, max_iteration, opts, x_dev, stats);
}

// Multi-device path of the host entry point: when the process has a device list of two or more
// (gsb_set_devices / GSB_DEVICES) and the matrix takes a two-colouring, the solve runs on row strips, one per
// device, from this one blocking call (gsb_dist_init_local: worker thread per device, peer-memory halo).  *took = 0
// when the matrix does not shard that way (more than two colours, fewer rows than GS_MGPU_MIN_ROWS per device): the
// caller then runs the single-device path.
#define GS_MGPU_MIN_ROWS 32768
static int gs_host_multi(gsb_matrix *m, const double *b, int nrhs, double epsilon, int max_iteration,
                         const gsb_gs_options *opts_in, double *x_out, gsb_gs_stats *stats, int *took) {
    *took = 0;
    int devs[GSB_DIST_MAX_WORLD_DECL];
    const int nd = gsb_devices(devs, GSB_DIST_MAX_WORLD_DECL);
    if (nd < 2 || (int64_t)m->n_rows < (int64_t)GS_MGPU_MIN_ROWS * nd) return GSB_OK;
    gsb_gs_options opts;
    gsb_gs_default_options(&opts);
    if (opts_in) opts = *opts_in;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double setup_ms = 0.0;
    if (!m->group_built) {
        GSB_CUDA(cudaEventCreate(&e0));
        GSB_CUDA(cudaEventCreate(&e1));
        GSB_CUDA(cudaEventRecord(e0, gsb_cur_stream()));
    }
    if (!m->analyzed) GSB_TRY(gsb_matrix_analyze(m, opts.ordering, nullptr));
    if (m->n_colors > 2) { // replicas only: stays on one device
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        return GSB_OK;
    }
    bool same = m->group && m->group_key[0] == nd;
    for (int i = 0; same && i < nd; ++i) same = m->group_key[1 + i] == devs[i];
    if (!same) {
        if (m->group) gsb_dist_group_finalize(m->group);
        m->group = nullptr;
        m->group_built = false;
        GSB_TRY(gsb_dist_init_local(&m->group, devs, nd));
        m->group_key[0] = nd;
        for (int i = 0; i < nd; ++i) m->group_key[1 + i] = devs[i];
    }
    if (!m->group_built) {
        GSB_TRY(gsb_dist_group_matrix(m->group, m));
        m->group_built = true;
        GSB_CUDA(cudaEventRecord(e1, gsb_cur_stream()));
        GSB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        setup_ms = ms;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    GSB_TRY(gsb_dist_group_gauss_seidel(m->group, b, nrhs, epsilon, max_iteration, &opts, x_out, stats));
    if (stats) {
        stats->setup_ms = setup_ms;
        stats->n_colors = m->n_colors;
        stats->ordering_used = m->ordering_used;
        if (opts.compute_residual)
            for (int r = 0; r < nrhs; ++r)
                GSB_TRY(gsb_dist_group_residual_l2(m->group, b + (size_t)r * m->n_rows, x_out + (size_t)r * m->n_rows,
                                                   &stats->residual_l2[r]));
    }
    *took = 1;
    return GSB_OK;
}

static int gs_host(gsb_matrix *m, const double *b, const double *x0, int nrhs, double epsilon, int max_iteration,
                   const gsb_gs_options *opts, double *x_out, gsb_gs_stats *stats) {
    GSB_TRY(gs_check_args(m, b, nrhs, x_out));
    if (!x0) { // (the initial-guess extension stays single-device)
        int took = 0;
        GSB_TRY(gs_host_multi(m, b, nrhs, epsilon, max_iteration, opts, x_out, stats, &took));
        if (took) return GSB_OK;
    }
    cudaStream_t st = gsb_cur_stream();
    const int64_t n = m->n_rows;
    // device staging of the caller's host vectors: kept with the handle (repeated solves reuse it)
    DevBuf<double> &db = m->stage_b, &dx = m->stage_x;
    GSB_TRY(db.alloc(n * nrhs));
    GSB_TRY(dx.alloc(n * nrhs));
    size_t bytes = sizeof(double) * (size_t)(n * nrhs);
    static const int b_overlap = [] { // GSB_B_OVERLAP=0: upload b on the main stream (measurement aid)
        const char *e = getenv("GSB_B_OVERLAP");
        return e ? atoi(e) : 1;
    }();
    if (!m->analyzed && b_overlap) {
        // a freshly imported matrix: the ordering analysis and the launch plan (several device round trips) come
        // first in the solver core -- b travels on the copy stream meanwhile and the core waits for it only where it
        // first reads it (the previous solve's use of stage_b is over: every entry point ends with a stream sync)
        if (!m->b_ready_event) {
            cudaEvent_t ev = nullptr;
            GSB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            m->b_ready_event = ev;
        }
        cudaStream_t cs = gsb_copy_stream();
        GSB_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, cs));
        GSB_CUDA(cudaEventRecord((cudaEvent_t)m->b_ready_event, cs));
        m->b_upload_pending = true;
    } else {
        GSB_CUDA(cudaMemcpyAsync(db.p, b, bytes, cudaMemcpyHostToDevice, st));
    }
    const double *x0_dev = nullptr;
    if (x0) {
        GSB_CUDA(cudaMemcpyAsync(dx.p, x0, bytes, cudaMemcpyHostToDevice, st));
        x0_dev = dx.p;
    }
    // x0 lives in dx and the result is scattered into dx: gs_solve_device permutes x0 into its
    // workspace before it writes dx, so the aliasing is safe.
    const int solved = gs_solve_device(m, db.p, x0_dev, nrhs, epsilon, max_iteration, opts, dx.p, stats);
    if (m->b_upload_pending) { // the core failed before it consumed the upload: do not leave a copy in flight
        cudaStreamSynchronize(gsb_copy_stream());
        m->b_upload_pending = false;
    }
    GSB_TRY(solved);
    GSB_CUDA(cudaMemcpyAsync(x_out, dx.p, bytes, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_gauss_seidel(gsb_matrix *m, const double *b, int nrhs, double epsilon, int max_iteration,
                                const gsb_gs_options *opts, double *x_out, gsb_gs_stats *stats) {
    return gs_host(m, b, nullptr, nrhs, epsilon, max_iteration, opts, x_out, stats);
}

extern "C" int gsb_gauss_seidel_x0(gsb_matrix *m, const double *b, const double *x0, int nrhs, double epsilon,
                                   int max_iteration, const gsb_gs_options *opts, double *x_out,
                                   gsb_gs_stats *stats) {
    if (!x0) return GSB_ERR_ARG;
    return gs_host(m, b, x0, nrhs, epsilon, max_iteration, opts, x_out, stats);
}

// ---------------------------------------------------------------------------------------------
// A7: SpMV over the reference layout, storage order, unfused (bit-exact with applyToVector)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) spmv_row_thread(const double *__restrict__ vals, const int *__restrict__ cols,
                                                       const int *__restrict__ row_begin,
                                                       const int *__restrict__ row_nnz, int n_rows,
                                                       const double *__restrict__ in, double *__restrict__ out) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_rows) return;
    int k = row_begin[i];
    const int e = k + row_nnz[i];
    double s = 0.0;
    for (; k < e; ++k) s = __dadd_rn(s, __dmul_rn(vals[k], in[cols[k]]));
    out[i] = s;
}

// residual partials: (b_i - (A x)_i)^2.  Rows with > 8 entries on average go row-per-warp
// (lanes stride the row, shuffle reduction); the norm is a diagnostic, not a parity quantity.
__global__ void __launch_bounds__(256) resid_row_thread(const double *__restrict__ vals, const int *__restrict__ cols,
                                                        const int *__restrict__ row_begin,
                                                        const int *__restrict__ row_nnz, int n_rows,
                                                        const double *__restrict__ b, const double *__restrict__ x,
                                                        double *__restrict__ partial) {
    double acc = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_rows; i += gridDim.x * 256) {
        int k = row_begin[i];
        const int e = k + row_nnz[i];
        double s = 0.0;
        for (; k < e; ++k) s += vals[k] * x[cols[k]];
        double r = b[i] - s;
        acc += r * r;
    }
    double a1[1] = {acc};
    gsb_block_reduce_store<1, 256>(a1, partial + blockIdx.x);
}

__global__ void __launch_bounds__(256) resid_row_warp(const double *__restrict__ vals, const int *__restrict__ cols,
                                                      const int *__restrict__ row_begin,
                                                      const int *__restrict__ row_nnz, int n_rows,
                                                      const double *__restrict__ b, const double *__restrict__ x,
                                                      double *__restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * 256) >> 5;
    double acc = 0.0;
    for (int i = warp; i < n_rows; i += nwarps) {
        const int k0 = row_begin[i], e = k0 + row_nnz[i];
        double s = 0.0;
        for (int k = k0 + lane; k < e; k += 32) s += vals[k] * x[cols[k]];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
        if (lane == 0) {
            double r = b[i] - s;
            acc += r * r;
        }
    }
    double a1[1] = {acc};
    gsb_block_reduce_store<1, 256>(a1, partial + blockIdx.x);
}

__global__ void __launch_bounds__(256) finish_sqrt(const double *__restrict__ partial, int np, double *out) {
    __shared__ double ws[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += 256) s += partial[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        out[0] = sqrt(t);
    }
}

static int spmv_check(gsb_matrix *m, const void *a, const void *b) {
    if (!m || !a || !b) return GSB_ERR_ARG;
    if (!m->has_layout) {
        gsb_set_error("spmv: matrix holds no layout yet");
        return GSB_ERR_STATE;
    }
    return gsb_set_device(m->device);
}

extern "C" int gsb_spmv_dev(gsb_matrix *m, const double *in_dev, double *out_dev) {
    GSB_TRY(spmv_check(m, in_dev, out_dev));
    cudaStream_t st = gsb_cur_stream();
    spmv_row_thread<<<(m->n_rows + 255) / 256, 256, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p,
                                                            m->n_rows, in_dev, out_dev);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

extern "C" int gsb_spmv(gsb_matrix *m, const double *in, double *out) {
    GSB_TRY(spmv_check(m, in, out));
    cudaStream_t st = gsb_cur_stream();
    DevBuf<double> di, dout;
    GSB_TRY(di.alloc(m->n_cols));
    GSB_TRY(dout.alloc(m->n_rows));
    GSB_CUDA(cudaMemcpyAsync(di.p, in, sizeof(double) * (size_t)m->n_cols, cudaMemcpyHostToDevice, st));
    GSB_TRY(gsb_spmv_dev(m, di.p, dout.p));
    GSB_CUDA(cudaMemcpyAsync(out, dout.p, sizeof(double) * (size_t)m->n_rows, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_residual_l2_dev(gsb_matrix *m, const double *b_dev, const double *x_dev, double *out) {
    GSB_TRY(spmv_check(m, b_dev, x_dev));
    if (!out) return GSB_ERR_ARG;
    cudaStream_t st = gsb_cur_stream();
    const int n = m->n_rows;
    const bool by_warp = m->nnz > (int64_t)8 * n;
    int nb = gsb_blocks_for(by_warp ? (int64_t)n * 32 : n, 256, gsb_sm_count() * 8);
    double *scratch = gsb_reduce_scratch(nb + 1);
    if (!scratch) return GSB_ERR_ALLOC;
    if (by_warp)
        resid_row_warp<<<nb, 256, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p, n, b_dev, x_dev, scratch);
    else
        resid_row_thread<<<nb, 256, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, m->row_nnz.p, n, b_dev, x_dev,
                                            scratch);
    GSB_KERNEL_CHECK();
    finish_sqrt<<<1, 256, 0, st>>>(scratch, nb, scratch + nb);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaMemcpyAsync(out, scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_residual_l2(gsb_matrix *m, const double *b, const double *x, double *out) {
    GSB_TRY(spmv_check(m, b, x));
    cudaStream_t st = gsb_cur_stream();
    DevBuf<double> db, dx;
    GSB_TRY(db.alloc(m->n_rows));
    GSB_TRY(dx.alloc(m->n_cols));
    GSB_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * (size_t)m->n_rows, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(dx.p, x, sizeof(double) * (size_t)m->n_cols, cudaMemcpyHostToDevice, st));
    return gsb_residual_l2_dev(m, db.p, dx.p, out);
}
