// gsb_prims.cu -- runtime plumbing (device, stream, errors) and the device-wide primitives the
// assembly and solver kernels share: warp-shuffle block scan -> multi-level exclusive scan,
// deterministic two-pass reductions, and the reference's free vector helpers (A8).
#include "gsb_internal.cuh"

#include <atomic>

#include <stdlib.h>

#include <mutex>

// ---------------------------------------------------------------------------------------------
// errors / device / stream
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void gsb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *gsb_last_error(void) { return g_err; }
extern "C" int gsb_version(void) { return 100; }

// the current device is per host thread (as CUDA's own): the single-process multi-device solver runs one worker
// thread per device; streams, SM counts and reduction scratch are per device and shared
static thread_local int g_device = 0;
static cudaStream_t g_streams[64] = {nullptr};
static int g_sm_count[64] = {0};
static double *g_scratch[64] = {nullptr};
static int64_t g_scratch_n[64] = {0};
static std::mutex g_mu;

extern "C" int gsb_device_count(int *count) {
    if (!count) return GSB_ERR_ARG;
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    *count = c;
    return GSB_OK;
}

int gsb_ensure_device() {
    int c = 0;
    gsb_device_count(&c);
    if (c <= 0) {
        gsb_set_error("no CUDA device visible: libgsb200 has no CPU fallback");
        return GSB_ERR_NO_DEVICE;
    }
    if (g_device >= c) {
        gsb_set_error("device %d selected but only %d visible", g_device, c);
        return GSB_ERR_NO_DEVICE;
    }
    GSB_CUDA(cudaSetDevice(g_device));
    return GSB_OK;
}

extern "C" int gsb_set_device(int device) {
    int c = 0;
    gsb_device_count(&c);
    if (c <= 0) {
        gsb_set_error("no CUDA device visible: libgsb200 has no CPU fallback");
        return GSB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= c || device >= 64) {
        gsb_set_error("gsb_set_device(%d): %d devices visible", device, c);
        return GSB_ERR_ARG;
    }
    g_device = device;
    GSB_CUDA(cudaSetDevice(device));
    return GSB_OK;
}

int gsb_current_device() { return g_device; }

// ---- device list of the host entry points (multi-device solve behind SparseMatrix::gaussSeidel) ----------------
static int g_devlist[GSB_DIST_MAX_WORLD_DECL];
static int g_devlist_n = -1; // -1: not set yet -> GSB_DEVICES is consulted once

extern "C" int gsb_set_devices(const int *devices, int n) {
    if (n < 0 || n > GSB_DIST_MAX_WORLD_DECL || (n > 0 && !devices)) {
        gsb_set_error("gsb_set_devices: 0..%d devices", GSB_DIST_MAX_WORLD_DECL);
        return GSB_ERR_ARG;
    }
    int c = 0;
    gsb_device_count(&c);
    for (int i = 0; i < n; ++i) {
        if (devices[i] < 0 || devices[i] >= c) {
            gsb_set_error("gsb_set_devices: device %d, %d visible", devices[i], c);
            return c <= 0 ? GSB_ERR_NO_DEVICE : GSB_ERR_ARG;
        }
        for (int j = 0; j < i; ++j)
            if (devices[j] == devices[i]) {
                gsb_set_error("gsb_set_devices: device %d listed twice", devices[i]);
                return GSB_ERR_ARG;
            }
    }
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < n; ++i) g_devlist[i] = devices[i];
    g_devlist_n = n;
    return GSB_OK;
}

int gsb_devices(int *out, int cap) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_devlist_n < 0) { // GSB_DEVICES=0,1,2,3 (unset: single device, the one gsb_set_device selected)
        g_devlist_n = 0;
        const char *e = getenv("GSB_DEVICES");
        int c = 0;
        cudaError_t ce = cudaGetDeviceCount(&c);
        if (ce != cudaSuccess) {
            cudaGetLastError();
            c = 0;
        }
        while (e && *e && g_devlist_n < GSB_DIST_MAX_WORLD_DECL) {
            char *end = nullptr;
            long v = strtol(e, &end, 10);
            if (end == e) break;
            if (v >= 0 && v < c) g_devlist[g_devlist_n++] = (int)v;
            e = *end == ',' ? end + 1 : end;
        }
    }
    for (int i = 0; i < g_devlist_n && i < cap; ++i) out[i] = g_devlist[i];
    return g_devlist_n;
}

extern "C" int gsb_get_devices(int *devices, int cap) { return gsb_devices(devices, devices ? cap : 0); }

cudaStream_t gsb_cur_stream() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_streams[g_device]) {
        cudaSetDevice(g_device);
        cudaStreamCreateWithFlags(&g_streams[g_device], cudaStreamNonBlocking);
    }
    return g_streams[g_device];
}

// second stream per device for host -> device copies that can overlap work on the main stream (the upload of b
// while the ordering analysis of a freshly imported matrix runs)
static cudaStream_t g_copy_streams[64] = {nullptr};
cudaStream_t gsb_copy_stream() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_copy_streams[g_device]) {
        cudaSetDevice(g_device);
        cudaStreamCreateWithFlags(&g_copy_streams[g_device], cudaStreamNonBlocking);
    }
    return g_copy_streams[g_device];
}

extern "C" void *gsb_stream(void) {
    if (gsb_ensure_device() != GSB_OK) return nullptr;
    return (void *)gsb_cur_stream();
}

int gsb_sm_count() {
    if (!g_sm_count[g_device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, g_device) != cudaSuccess || v <= 0)
            v = GSB_SM_COUNT_FALLBACK;
        g_sm_count[g_device] = v;
    }
    return g_sm_count[g_device];
}

double *gsb_reduce_scratch(int64_t n_doubles) {
    std::lock_guard<std::mutex> lk(g_mu);
    int d = g_device;
    if (g_scratch_n[d] < n_doubles) {
        if (g_scratch[d]) cudaFree(g_scratch[d]);
        g_scratch[d] = nullptr;
        int64_t want = n_doubles < 65536 ? 65536 : n_doubles;
        if (cudaMalloc((void **)&g_scratch[d], sizeof(double) * (size_t)want) != cudaSuccess) {
            cudaGetLastError();
            g_scratch_n[d] = 0;
            return nullptr;
        }
        g_scratch_n[d] = want;
    }
    return g_scratch[d];
}

extern "C" int gsb_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    GSB_CUDA(cudaHostAlloc(ptr, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault));
    return GSB_OK;
}

static std::atomic<long long> g_dev_allocs{0}, g_dev_frees{0};
void gsb_count_alloc(int frees) { (frees ? g_dev_frees : g_dev_allocs).fetch_add(1, std::memory_order_relaxed); }
extern "C" int gsb_alloc_counters(int64_t *device_allocs, int64_t *device_frees) {
    if (device_allocs) *device_allocs = (int64_t)g_dev_allocs.load();
    if (device_frees) *device_frees = (int64_t)g_dev_frees.load();
    return GSB_OK;
}

extern "C" int gsb_host_free(void *ptr) {
    if (!ptr) return GSB_OK;
    GSB_CUDA(cudaFreeHost(ptr));
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan (int32): warp-shuffle block scan, three-phase multi-level
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int warp_incl_scan(int v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, *block_total = sum
__device__ __forceinline__ int block_excl_scan(int v, int *block_total) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    __shared__ int total_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int wi = warp_incl_scan(w);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) total_s = wi;
    }
    __syncthreads();
    int r = incl - v + warp_sums[wid];
    *block_total = total_s;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(const int *__restrict__ in, int64_t n,
                                                                int *__restrict__ sums) {
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    int tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// each thread owns SCAN_ITEMS consecutive items (blocked arrangement) so the scan is a true prefix
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const int *in, int *out, int64_t n,
                                                           const int *__restrict__ block_offsets,
                                                           int *total_out) {
    __shared__ int tile[SCAN_TILE + SCAN_TILE / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    // coalesced load into (padded) shared memory
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int j = k * SCAN_THREADS + threadIdx.x;
        int64_t i = base + j;
        tile[j + (j >> 5)] = (i < n) ? in[i] : 0;
    }
    __syncthreads();
    int vals[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int j = threadIdx.x * SCAN_ITEMS + k;
        vals[k] = tile[j + (j >> 5)];
        s += vals[k];
    }
    int tot;
    int excl = block_excl_scan(s, &tot);
    int off = block_offsets ? block_offsets[blockIdx.x] : 0;
    int run = excl + off;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int j = threadIdx.x * SCAN_ITEMS + k;
        tile[j + (j >> 5)] = run;
        run += vals[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int j = k * SCAN_THREADS + threadIdx.x;
        int64_t i = base + j;
        if (i < n) out[i] = tile[j + (j >> 5)];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total_out = off + tot;
}

// ints of caller-provided scratch that make gsb_exclusive_scan_i32 allocation-free for n elements
int64_t gsb_scan_scratch_ints(int64_t n) {
    int64_t tot = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        tot += n;
    }
    return tot + 1;
}

int gsb_exclusive_scan_i32(const int *in, int *out, int64_t n, int *total_dev, cudaStream_t st, int *scratch,
                           int64_t scratch_ints) {
    if (n <= 0) {
        if (total_dev) GSB_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int), st));
        return GSB_OK;
    }
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        scan_apply<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, total_dev);
        GSB_KERNEL_CHECK();
        return GSB_OK;
    }
    if (scratch && scratch_ints >= gsb_scan_scratch_ints(n)) { // the block sums of every level live in the caller's buffer
        scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, scratch);
        GSB_KERNEL_CHECK();
        GSB_TRY(gsb_exclusive_scan_i32(scratch, scratch, nb, nullptr, st, scratch + nb, scratch_ints - nb));
        scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, scratch, total_dev);
        GSB_KERNEL_CHECK();
        return GSB_OK; // (stream-ordered: nothing to free, no host synchronisation)
    }
    DevBuf<int> sums;
    GSB_TRY(sums.alloc(nb));
    scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, n, sums.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(gsb_exclusive_scan_i32(sums.p, sums.p, nb, nullptr, st));
    scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums.p, total_dev);
    GSB_KERNEL_CHECK();
    GSB_CUDA(cudaStreamSynchronize(st)); // sums is freed on return
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// deterministic reductions: fixed grid, per-block partials, one finishing block (fixed order)
// ---------------------------------------------------------------------------------------------
#define RED_THREADS 256

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double ws[RED_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) ws[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < RED_THREADS / 32 ? ws[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r; // valid in thread 0
}

template <int OP> // 0: |a-b|   1: a*b
__global__ void __launch_bounds__(RED_THREADS) reduce2_partial(const double *__restrict__ a,
                                                               const double *__restrict__ b, int64_t n,
                                                               double *__restrict__ partial) {
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * RED_THREADS) {
        double x = a[i], y = b[i];
        s += (OP == 0) ? fabs(x - y) : x * y;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void __launch_bounds__(RED_THREADS) reduce_finish(const double *__restrict__ partial, int np,
                                                             double *__restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < np; i += RED_THREADS) s += partial[i];
    s = block_sum(s);
    if (threadIdx.x == 0) out[0] = s;
}

template <int OP>
static int reduce2(const double *a, const double *b, int64_t n, double *out_dev, cudaStream_t st) {
    int nb = gsb_blocks_for(n, RED_THREADS * 8, gsb_sm_count() * 8);
    double *scratch = gsb_reduce_scratch(nb);
    if (!scratch) {
        gsb_set_error("reduction scratch allocation failed");
        return GSB_ERR_ALLOC;
    }
    reduce2_partial<OP><<<nb, RED_THREADS, 0, st>>>(a, b, n, scratch);
    GSB_KERNEL_CHECK();
    reduce_finish<<<1, RED_THREADS, 0, st>>>(scratch, nb, out_dev);
    GSB_KERNEL_CHECK();
    return GSB_OK;
}

int gsb_l1_dist_dev(const double *a, const double *b, int64_t n, double *out_dev, cudaStream_t st) {
    return reduce2<0>(a, b, n, out_dev, st);
}
int gsb_dot_dev(const double *a, const double *b, int64_t n, double *out_dev, cudaStream_t st) {
    return reduce2<1>(a, b, n, out_dev, st);
}

__global__ void __launch_bounds__(RED_THREADS) max_i32_partial(const int *__restrict__ in, int64_t n,
                                                               int *__restrict__ out) {
    int m = INT32_MIN;
    for (int64_t i = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * RED_THREADS)
        m = max(m, in[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void set_i32(int *p, int v) { *p = v; }

int gsb_reduce_max_i32(const int *in, int64_t n, int *out_dev, cudaStream_t st) {
    set_i32<<<1, 1, 0, st>>>(out_dev, INT32_MIN);
    GSB_KERNEL_CHECK();
    if (n > 0) {
        int nb = gsb_blocks_for(n, RED_THREADS * 8, gsb_sm_count() * 8);
        max_i32_partial<<<nb, RED_THREADS, 0, st>>>(in, n, out_dev);
        GSB_KERNEL_CHECK();
    }
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// A8: the reference's free vector helpers with host pointers (v2 :45-105)
// ---------------------------------------------------------------------------------------------
template <int OP> // 0: a + s*b (unfused, as the reference's lambda rounds the product first)  1: a*b
__global__ void __launch_bounds__(256) vec_binary(const double *a, const double *b,
                                                  double s, int64_t n, double *out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        double x = a[i], y = b[i];
        out[i] = (OP == 0) ? __dadd_rn(x, __dmul_rn(s, y)) : __dmul_rn(x, y);
    }
}

static int host_reduce2(int op, const double *a, const double *b, int64_t n, double *out) {
    if (!a || !b || !out || n < 0) return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    cudaStream_t st = gsb_cur_stream();
    DevBuf<double> da, db, dr;
    GSB_TRY(da.alloc(n));
    GSB_TRY(db.alloc(n));
    GSB_TRY(dr.alloc(1));
    GSB_CUDA(cudaMemcpyAsync(da.p, a, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    GSB_TRY(op == 0 ? gsb_l1_dist_dev(da.p, db.p, n, dr.p, st) : gsb_dot_dev(da.p, db.p, n, dr.p, st));
    GSB_CUDA(cudaMemcpyAsync(out, dr.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_l1_dist(const double *a, const double *b, int64_t n, double *out) {
    return host_reduce2(0, a, b, n, out);
}
extern "C" int gsb_dot(const double *a, const double *b, int64_t n, double *out) {
    return host_reduce2(1, a, b, n, out);
}

static int host_binary(int op, const double *a, const double *b, double s, int64_t n, double *out) {
    if (!a || !b || !out || n < 0) return GSB_ERR_ARG;
    GSB_TRY(gsb_ensure_device());
    cudaStream_t st = gsb_cur_stream();
    DevBuf<double> da, db;
    GSB_TRY(da.alloc(n));
    GSB_TRY(db.alloc(n));
    GSB_CUDA(cudaMemcpyAsync(da.p, a, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemcpyAsync(db.p, b, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, st));
    if (n > 0) {
        int nb = gsb_blocks_for(n, 256 * 4, gsb_sm_count() * 16);
        if (op == 0)
            vec_binary<0><<<nb, 256, 0, st>>>(da.p, db.p, s, n, da.p);
        else
            vec_binary<1><<<nb, 256, 0, st>>>(da.p, db.p, s, n, da.p);
        GSB_KERNEL_CHECK();
    }
    GSB_CUDA(cudaMemcpyAsync(out, da.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_axpy(const double *a, const double *b, double scale_b, int64_t n, double *out) {
    return host_binary(0, a, b, scale_b, n, out);
}
extern "C" int gsb_vecmul(const double *a, const double *b, int64_t n, double *out) {
    return host_binary(1, a, b, 0.0, n, out);
}
