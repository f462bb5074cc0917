// gsb_dist.cu -- multi-GPU row strips.  One rank per GPU; each rank owns a contiguous
// block of matrix rows (image rows [y0,y1) of the grid), coloured by GLOBAL pixel parity so strips
// agree, stored colour-major with the ghost unknowns (the neighbour strips' boundary rows) appended.
// After every colour phase the values of that colour which a neighbour reads are packed and moved
// with grouped ncclSend/ncclRecv over NVLink; every checked sweep one ncclAllReduce of the per-RHS
// L1 update norms keeps the stop decision identical on all ranks.
//
// Entries inside a row are ordered by (colour, global column) -- the order the single-GPU solver
// uses -- so an N-strip solve is bit-identical to the 1-GPU solve of the same system.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy torch already loaded, else the system
// one), so libgsb200.so itself has no link dependency on it.
#include "gsb_internal.cuh"

#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>

#include <condition_variable>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

// ---------------------------------------------------------------------------------------------
// NCCL binding
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return GSB_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        gsb_set_error("NCCL not found (dlopen libnccl.so.2): %s", dlerror());
        return GSB_ERR_NCCL;
    }
#define LOAD(field, sym)                                              \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                        \
    if (!g_nccl.field) {                                              \
        gsb_set_error("NCCL symbol %s missing", sym);                 \
        return GSB_ERR_NCCL;                                          \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(Send, "ncclSend")
    LOAD(Recv, "ncclRecv")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(AllGather, "ncclAllGather")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    g_nccl.handle = h;
    return GSB_OK;
}

#define GSB_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            gsb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return GSB_ERR_NCCL;                                                                    \
        }                                                                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Single-process mode (gsb_dist_init_local): the ranks are worker threads of one process, one per device.  What the
// one-process-per-GPU mode does with NCCL during SETUP (neighbours exchanging id lists and pointers, small
// all-gathers / all-reduces of host values, barriers) is done here through this shared block: a barrier, a few
// published pointers per rank, and peer copies (cudaMemcpyPeerAsync) issued by the receiving rank.  The data path of
// a solve needs none of it: halo values and the stop-rule sums travel as peer stores inside the kernels.
// ---------------------------------------------------------------------------------------------
struct LocalGroup {
    int world = 1;
    int dev[GSB_DIST_MAX_WORLD] = {0};
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    long long generation = 0;
    int aborted = 0; // a rank failed: every barrier returns an error instead of waiting for it
    const void *ptr[GSB_DIST_MAX_WORLD][2] = {{nullptr}};
    unsigned char blob[GSB_DIST_MAX_WORLD][256];
    int barrier() {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return GSB_ERR_NCCL;
        const long long gen = generation;
        if (++arrived == world) {
            arrived = 0;
            ++generation;
            cv.notify_all();
            return GSB_OK;
        }
        cv.wait(lk, [&] { return generation != gen || aborted; });
        return generation != gen ? GSB_OK : GSB_ERR_NCCL;
    }
    void abort() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = 1;
        cv.notify_all();
    }
    void reset() {
        std::lock_guard<std::mutex> lk(mu);
        aborted = 0;
        arrived = 0;
    }
};

// ---------------------------------------------------------------------------------------------
// the per-rank handle
// ---------------------------------------------------------------------------------------------
struct gsb_dist {
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    LocalGroup *lg = nullptr;           // single-process mode: shared with the other ranks (owned by the group)
    DevBuf<unsigned char> colors8;      // optional: colour (0/1) of every GLOBAL row; absent -> pixel parity with W
    DevBuf<double> stage_b, stage_x;    // single-process mode: device copies of the caller's host b / x slices
    // local system
    bool built = false;
    int64_t row0 = 0, n_global = 0;
    int n_local = 0, n_ghost = 0, W = 1;
    int64_t ld = 0; // n_local + n_ghost: leading dimension of xw / bw
    int64_t nnz_local = 0;
    int color_start[3] = {0, 0, 0};
    DevBuf<int> perm, iperm, rp, ci; // rp/ci/va: off-diagonal entries only
    DevBuf<double> va, dg;          // dg: diagonal per (permuted) local row
    // halo: peer 0 = rank-1 (rows above), peer 1 = rank+1 (rows below); second index = colour
    int need_cnt[2][2] = {{0, 0}, {0, 0}};
    int ghost_start[2][2] = {{0, 0}, {0, 0}};
    int send_cnt[2][2] = {{0, 0}, {0, 0}};
    DevBuf<int> send_idx[2][2];
    DevBuf<double> sendbuf[2][2];
    // natural-order copy for the residual
    DevBuf<int> nat_rp, nat_cg;
    DevBuf<double> nat_va;
    // workspaces
    DevBuf<double> xw, bw, partials;
    DevBuf<unsigned char> ctl;
    void *ctl_host = nullptr;
    int ws_nrhs = 0;
    GsbPlan plan;
    // fused halo exchange over peer-mapped memory (NVLink): see GsbHaloArgs
    bool peer_ready = false;   // peer pointers valid for the current xw allocation
    bool peer_failed = false;  // IPC unavailable: stay on the NCCL send/recv path
    bool halo_meta = false;    // tile order / info / push maps built for the current plan
    bool halo_agreed_valid = false, halo_agreed = false; // all ranks can fuse the exchange for the current plan
    DevBuf<int> flags;         // [peer][colour] raised by the neighbours, then [4..5] halo-tile counters per colour
    double *peer_x[2] = {nullptr, nullptr};
    int *peer_flags[2] = {nullptr, nullptr};
    long long peer_ld[2] = {0, 0};
    int peer_gs[2][2] = {{0, 0}, {0, 0}};
    DevBuf<int> tile_order[2], push_map[2];
    DevBuf<unsigned char> tile_info[2];
    int n_halo_tiles[2] = {0, 0};
    int interior_base[2] = {-1, -1}; // first interior tile when the interior tiles are one contiguous range
    int n_interior[2] = {0, 0};
    long long epoch = 0;       // flags only ever grow: epoch of the last sweep issued so far
    int used_peer = 0;
    // fused stop-rule all-reduce (GsbEpsExchange): this rank's box and every rank's box peer-mapped
    DevBuf<double> eps_box;
    double *peer_box[GSB_DIST_MAX_WORLD] = {nullptr};
    bool box_ready = false, box_failed = false;
    long long xcount = 0;      // stop-rule exchanges issued so far (identical on every rank)
    int used_fused_eps = 0;
    int peer_rank(int p) const { return p == 0 ? rank - 1 : rank + 1; }
    bool has_peer(int p) const { return p == 0 ? rank > 0 : rank < world - 1; }
};


// ---------------------------------------------------------------------------------------------
// setup-time communication: NCCL (one process per GPU) or the LocalGroup (single process)
// ---------------------------------------------------------------------------------------------
#define GSB_LOCAL(call)                                                                              \
    do {                                                                                             \
        if ((call) != GSB_OK) {                                                                      \
            gsb_set_error("dist (single process): another rank failed, rank %d gives up", d->rank); \
            return GSB_ERR_NCCL;                                                                     \
        }                                                                                            \
    } while (0)

// Device buffers to / from the two neighbours (p = 0: rank - 1, p = 1: rank + 1); zero-length sides are skipped.
// NCCL: stream-ordered.  Local: on return the data has arrived and the send buffers may be reused.
static int comm_exchange_neighbours(gsb_dist *d, const void *const send[2], const size_t sbytes[2], void *const recv[2],
                                    const size_t rbytes[2], cudaStream_t st) {
    if (d->lg) {
        LocalGroup *g = d->lg;
        GSB_CUDA(cudaStreamSynchronize(st)); // what is sent has been produced
        for (int p = 0; p < 2; ++p) g->ptr[d->rank][p] = send[p];
        GSB_LOCAL(g->barrier());
        for (int p = 0; p < 2; ++p)
            if (d->has_peer(p) && rbytes[p]) {
                const int q = d->peer_rank(p);
                GSB_CUDA(cudaMemcpyPeerAsync(recv[p], d->device, g->ptr[q][1 - p], g->dev[q], rbytes[p], st));
            }
        GSB_CUDA(cudaStreamSynchronize(st));
        GSB_LOCAL(g->barrier());
        return GSB_OK;
    }
    GSB_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < 2; ++p)
        if (d->has_peer(p)) {
            if (sbytes[p]) GSB_NCCL(g_nccl.Send(send[p], sbytes[p], ncclInt8, d->peer_rank(p), d->comm, st));
            if (rbytes[p]) GSB_NCCL(g_nccl.Recv(recv[p], rbytes[p], ncclInt8, d->peer_rank(p), d->comm, st));
        }
    GSB_NCCL(g_nccl.GroupEnd());
    return GSB_OK;
}

// all-gather of a small host struct (bytes <= 256): all[q * bytes ...] = rank q's `mine`
static int comm_allgather_host(gsb_dist *d, const void *mine, void *all, size_t bytes, cudaStream_t st) {
    if (d->lg) {
        LocalGroup *g = d->lg;
        if (bytes > sizeof(g->blob[0])) return GSB_ERR_ARG;
        memcpy(g->blob[d->rank], mine, bytes);
        GSB_LOCAL(g->barrier());
        for (int q = 0; q < d->world; ++q) memcpy((unsigned char *)all + (size_t)q * bytes, g->blob[q], bytes);
        GSB_LOCAL(g->barrier());
        return GSB_OK;
    }
    DevBuf<unsigned char> dm, da;
    GSB_TRY(dm.alloc((int64_t)bytes));
    GSB_TRY(da.alloc((int64_t)bytes * d->world));
    GSB_CUDA(cudaMemcpyAsync(dm.p, mine, bytes, cudaMemcpyHostToDevice, st));
    GSB_NCCL(g_nccl.AllGather(dm.p, da.p, bytes, ncclInt8, d->comm, st));
    GSB_CUDA(cudaMemcpyAsync(all, da.p, bytes * (size_t)d->world, cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

// *v = min over the ranks of *v (the "can everybody do X" agreements)
static int comm_allreduce_min(gsb_dist *d, int *v, cudaStream_t st) {
    if (d->world == 1) return GSB_OK;
    if (d->lg) {
        int all[GSB_DIST_MAX_WORLD];
        GSB_TRY(comm_allgather_host(d, v, all, sizeof(int), st));
        for (int q = 0; q < d->world; ++q) *v = all[q] < *v ? all[q] : *v;
        return GSB_OK;
    }
    DevBuf<int> agree;
    GSB_TRY(agree.alloc(1));
    GSB_CUDA(cudaMemcpyAsync(agree.p, v, sizeof(int), cudaMemcpyHostToDevice, st));
    GSB_NCCL(g_nccl.AllReduce(agree.p, agree.p, 1, ncclInt32, ncclMin, d->comm, st));
    GSB_CUDA(cudaMemcpyAsync(v, agree.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

// every rank's stream has reached this point before any rank's stream goes on
static int comm_stream_barrier(gsb_dist *d, int *scratch_dev, cudaStream_t st) {
    if (d->world == 1) return GSB_OK;
    if (d->lg) {
        GSB_CUDA(cudaStreamSynchronize(st));
        GSB_LOCAL(d->lg->barrier());
        return GSB_OK;
    }
    GSB_NCCL(g_nccl.AllReduce(scratch_dev, scratch_dev, 1, ncclInt32, ncclMax, d->comm, st));
    return GSB_OK;
}

// in place sum of one device double over the ranks (rank order -> the same bits everywhere)
static int comm_allreduce_sum_dev(gsb_dist *d, double *v_dev, cudaStream_t st) {
    if (d->world == 1) return GSB_OK;
    if (d->lg) {
        double mine = 0.0, all[GSB_DIST_MAX_WORLD];
        GSB_CUDA(cudaMemcpyAsync(&mine, v_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        GSB_TRY(comm_allgather_host(d, &mine, all, sizeof(double), st));
        double s = 0.0;
        for (int q = 0; q < d->world; ++q) s += all[q];
        GSB_CUDA(cudaMemcpyAsync(v_dev, &s, sizeof(double), cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        return GSB_OK;
    }
    GSB_NCCL(g_nccl.AllReduce(v_dev, v_dev, 1, ncclFloat64, ncclSum, d->comm, st));
    return GSB_OK;
}

extern "C" int gsb_dist_unique_id(unsigned char id[GSB_UNIQUE_ID_BYTES]) {
    if (!id) return GSB_ERR_ARG;
    GSB_TRY(nccl_load());
    ncclUniqueId u;
    GSB_NCCL(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(u) == GSB_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &u, sizeof(u));
    return GSB_OK;
}

extern "C" int gsb_dist_init(gsb_dist **out, const unsigned char id[GSB_UNIQUE_ID_BYTES], int rank, int world,
                             int device) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) {
        gsb_set_error("dist_init: bad argument");
        return GSB_ERR_ARG;
    }
    GSB_TRY(nccl_load());
    GSB_TRY(gsb_set_device(device));
    gsb_dist *d = new (std::nothrow) gsb_dist();
    if (!d) return GSB_ERR_ALLOC;
    d->rank = rank;
    d->world = world;
    d->device = device;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = g_nccl.CommInitRank(&d->comm, world, u, rank);
    if (r != ncclSuccess) {
        gsb_set_error("ncclCommInitRank(rank %d of %d) -> %s", rank, world, g_nccl.GetErrorString(r));
        delete d;
        return GSB_ERR_NCCL;
    }
    *out = d;
    return GSB_OK;
}

extern "C" int gsb_dist_finalize(gsb_dist *d) {
    if (!d) return GSB_OK;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(gsb_cur_stream());
    if (!d->lg) { // (single-process mode holds the neighbours' raw pointers: nothing to unmap)
        for (int p = 0; p < 2; ++p) {
            if (d->peer_x[p]) cudaIpcCloseMemHandle(d->peer_x[p]);
            if (d->peer_flags[p]) cudaIpcCloseMemHandle(d->peer_flags[p]);
        }
        for (int q = 0; q < d->world && q < GSB_DIST_MAX_WORLD; ++q)
            if (q != d->rank && d->peer_box[q]) cudaIpcCloseMemHandle(d->peer_box[q]);
    }
    if (d->comm) g_nccl.CommDestroy(d->comm);
    if (d->ctl_host) cudaFreeHost(d->ctl_host);
    delete d;
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// local build kernels
// ---------------------------------------------------------------------------------------------
// colour of GLOBAL row g: the caller's two-colouring when one was given (c8), else pixel parity on a W-wide grid
struct ColorOf {
    const unsigned char *c8;
    int W;
    __device__ __forceinline__ int operator()(int64_t g) const {
        return c8 ? (int)c8[g] : (int)(((g % W) + (g / W)) & 1);
    }
};

__global__ void __launch_bounds__(256) d_minmax_col(const int *__restrict__ cg, int64_t nnz, int *__restrict__ mm) {
    int lo = INT32_MAX, hi = -1;
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * 256) {
        int c = cg[k];
        lo = min(lo, c);
        hi = max(hi, c);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, __shfl_down_sync(0xffffffffu, lo, d));
        hi = max(hi, __shfl_down_sync(0xffffffffu, hi, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}

// flags of the ghost columns in the two windows [row0-halo_lo,row0) and [row1,row1+halo_hi), split by colour:
// flag[(w*2+colour)][j]
__global__ void __launch_bounds__(256) d_mark_ghosts(const int *__restrict__ cg, int64_t nnz, int64_t row0,
                                                     int64_t row1, int halo_lo, int halo_hi, ColorOf col,
                                                     int *__restrict__ f_lo0, int *__restrict__ f_lo1,
                                                     int *__restrict__ f_hi0, int *__restrict__ f_hi1) {
    for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * 256) {
        int64_t c = cg[k];
        if (c < row0) {
            int j = (int)(c - (row0 - halo_lo));
            (col(c) ? f_lo1 : f_lo0)[j] = 1;
        } else if (c >= row1) {
            int j = (int)(c - row1);
            (col(c) ? f_hi1 : f_hi0)[j] = 1;
        }
    }
}

__global__ void __launch_bounds__(256) d_owned_flag(int64_t row0, int n_local, ColorOf col, int *__restrict__ f) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_local) f[i] = col(row0 + i) == 0 ? 1 : 0;
    if (i == n_local) f[i] = 0;
}

__global__ void __launch_bounds__(256) d_owned_place(int64_t row0, int n_local, ColorOf col, const int *__restrict__ scan0,
                                                     int n0, int *__restrict__ perm, int *__restrict__ iperm) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_local) return;
    int s = scan0[i];
    int p = col(row0 + i) == 0 ? s : n0 + (i - s);
    perm[p] = i;
    iperm[i] = p;
}

// window flag scan -> ghost map (local id of window slot j, -1 if not a ghost) and the ordered id list
__global__ void __launch_bounds__(256) d_ghost_place(const int *__restrict__ scan, int len, int64_t gid0, int base,
                                                     int *__restrict__ map, int *__restrict__ ids) {
    int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= len) return;
    int s = scan[j];
    if (scan[j + 1] != s) {
        map[j] = base + s;
        ids[s] = (int)(gid0 + j);
    }
}

// off-diagonal row lengths in permuted order
__global__ void __launch_bounds__(256) d_perm_len(const int *__restrict__ perm, const int *__restrict__ rp_nat,
                                                  const int *__restrict__ cg_nat, int64_t row0, int n_local,
                                                  int *__restrict__ len) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p < n_local) {
        int o = perm[p];
        int off = 0;
        for (int k = rp_nat[o]; k < rp_nat[o + 1]; ++k) off += cg_nat[k] != row0 + o;
        len[p] = off;
    }
    if (p == n_local) len[p] = 0;
}

// copy rows in permuted order, sorting each row by (colour, global column); then map to local ids
__global__ void __launch_bounds__(128) d_fill_rows(const int *__restrict__ perm, const int *__restrict__ iperm,
                                                   const int *__restrict__ rp_nat, const int *__restrict__ cg_nat,
                                                   const double *__restrict__ va_nat, int n_local, int64_t row0,
                                                   int64_t row1, int halo_lo, ColorOf col, const int *__restrict__ map_lo,
                                                   const int *__restrict__ map_hi, const int *__restrict__ rp,
                                                   int *__restrict__ ci, double *__restrict__ va,
                                                   double *__restrict__ dg, int *__restrict__ bad) {
    int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_local) return;
    const int o = perm[p];
    const int src = rp_nat[o], len_all = rp_nat[o + 1] - src, dst = rp[p];
    const int len = rp[p + 1] - dst;
    const int my_par = col(row0 + o);
    double d = 0.0;
    int w = 0;
    for (int k = 0; k < len_all; ++k) { // insertion sort on key (colour, global id); ci temporarily holds global ids
        int g = cg_nat[src + k];
        double v = va_nat[src + k];
        if (g == row0 + o) { // the diagonal goes to its own array
            d = v;
            continue;
        }
        int gp = col(g);
        if (gp == my_par) atomicOr(bad, 1); // parity colouring improper for this matrix
        int q = dst + w;
        while (q > dst) {
            int h = ci[q - 1];
            int hp = col(h);
            if (hp < gp || (hp == gp && h <= g)) break;
            ci[q] = h;
            va[q] = va[q - 1];
            --q;
        }
        ci[q] = g;
        va[q] = v;
        ++w;
    }
    dg[p] = d;
    for (int k = 0; k < len; ++k) {
        int64_t g = ci[dst + k];
        int l;
        if (g < row0)
            l = map_lo[(int)(g - (row0 - halo_lo))];
        else if (g >= row1)
            l = map_hi[(int)(g - row1)];
        else
            l = iperm[(int)(g - row0)];
        ci[dst + k] = l;
    }
}

__global__ void __launch_bounds__(256) d_ids_to_local(const int *__restrict__ ids, int cnt, int64_t row0, int n_local,
                                                      const int *__restrict__ iperm, int *__restrict__ out,
                                                      int *__restrict__ bad) {
    int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= cnt) return;
    int64_t l = (int64_t)ids[j] - row0;
    if (l < 0 || l >= n_local) {
        atomicOr(bad, 2);
        out[j] = 0;
    } else {
        out[j] = iperm[l];
    }
}

__global__ void __launch_bounds__(256) d_pack(const double *__restrict__ x, int64_t ld, const int *__restrict__ idx,
                                              int cnt, int nrhs, double *__restrict__ buf) {
    int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= cnt) return;
    int s = idx[j];
    for (int r = 0; r < nrhs; ++r) buf[(size_t)r * cnt + j] = x[r * ld + s];
}

__global__ void __launch_bounds__(256) d_gather(const double *__restrict__ src, int64_t src_ld,
                                                const int *__restrict__ perm, int n_local, int nrhs, int64_t dst_ld,
                                                double *__restrict__ dst) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= n_local) return;
    int o = perm[p];
    for (int r = 0; r < nrhs; ++r) dst[r * dst_ld + p] = src[r * src_ld + o];
}

__global__ void __launch_bounds__(256) d_scatter(const double *__restrict__ src, int64_t src_ld,
                                                 const int *__restrict__ perm, int n_local, int nrhs, int64_t dst_ld,
                                                 double *__restrict__ dst) {
    int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= n_local) return;
    int o = perm[p];
    for (int r = 0; r < nrhs; ++r) dst[r * dst_ld + o] = src[r * src_ld + p];
}

__global__ void __launch_bounds__(256) d_fill(double *__restrict__ p, int64_t n, double v) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// build from device-resident natural-order local rows (global columns)
// ---------------------------------------------------------------------------------------------
static int scan_count(DevBuf<int> &f, int len, cudaStream_t st, int *count) {
    // f has len+1 entries with f[len] = 0; after the scan f[len] is the count
    GSB_TRY(gsb_exclusive_scan_i32(f.p, f.p, (int64_t)len + 1, nullptr, st));
    GSB_CUDA(cudaMemcpyAsync(count, f.p + len, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

static int dist_build(gsb_dist *d, int64_t row0, int n_local, int64_t n_global, int W) {
    cudaStream_t st = gsb_cur_stream();
    const int64_t row1 = row0 + n_local;
    d->row0 = row0;
    d->n_local = n_local;
    d->n_global = n_global;
    d->W = W < 1 ? 1 : W;
    W = d->W;
    const ColorOf col = {d->colors8.n >= n_global ? d->colors8.p : nullptr, W};
    int nnz = 0;
    GSB_CUDA(cudaMemcpyAsync(&nnz, d->nat_rp.p + n_local, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    d->nnz_local = nnz;

    // 1. column range -> halo widths
    DevBuf<int> mm;
    GSB_TRY(mm.alloc(4));
    int init[4] = {INT32_MAX, -1, 0, 0};
    GSB_CUDA(cudaMemcpyAsync(mm.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        d_minmax_col<<<gsb_blocks_for(nnz, 256 * 8, gsb_sm_count() * 8), 256, 0, st>>>(d->nat_cg.p, nnz, mm.p);
        GSB_KERNEL_CHECK();
    }
    int h_mm[2];
    GSB_CUDA(cudaMemcpyAsync(h_mm, mm.p, sizeof(h_mm), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    int halo_lo = 0, halo_hi = 0;
    if (nnz > 0) {
        if (h_mm[0] < 0 || h_mm[1] >= n_global) {
            gsb_set_error("dist: column index out of range [0,%lld)", (long long)n_global);
            return GSB_ERR_SHAPE;
        }
        if (h_mm[0] < row0) halo_lo = (int)(row0 - h_mm[0]);
        if (h_mm[1] >= row1) halo_hi = (int)(h_mm[1] - (row1 - 1));
    }
    if ((halo_lo && d->rank == 0) || (halo_hi && d->rank == d->world - 1)) {
        gsb_set_error("dist: rank %d of %d references rows outside the partition", d->rank, d->world);
        return GSB_ERR_SHAPE;
    }

    // 2./3. ghost flags per window and colour; owned ordering
    DevBuf<int> f[4], owned;
    const int flen[4] = {halo_lo, halo_lo, halo_hi, halo_hi};
    for (int q = 0; q < 4; ++q) {
        GSB_TRY(f[q].alloc((int64_t)flen[q] + 1));
        GSB_CUDA(cudaMemsetAsync(f[q].p, 0, sizeof(int) * (size_t)(flen[q] + 1), st));
    }
    if (nnz > 0 && (halo_lo || halo_hi)) {
        d_mark_ghosts<<<gsb_blocks_for(nnz, 256 * 8, gsb_sm_count() * 8), 256, 0, st>>>(
            d->nat_cg.p, nnz, row0, row1, halo_lo, halo_hi, col, f[0].p, f[1].p, f[2].p, f[3].p);
        GSB_KERNEL_CHECK();
    }
    GSB_TRY(owned.alloc((int64_t)n_local + 1));
    d_owned_flag<<<(n_local + 1 + 255) / 256, 256, 0, st>>>(row0, n_local, col, owned.p);
    GSB_KERNEL_CHECK();
    int n0 = 0;
    GSB_TRY(scan_count(owned, n_local, st, &n0));
    GSB_TRY(d->perm.alloc(n_local));
    GSB_TRY(d->iperm.alloc(n_local));
    d_owned_place<<<(n_local + 255) / 256, 256, 0, st>>>(row0, n_local, col, owned.p, n0, d->perm.p, d->iperm.p);
    GSB_KERNEL_CHECK();
    d->color_start[0] = 0;
    d->color_start[1] = n0;
    d->color_start[2] = n_local;

    // 4. ghost numbering: [lo colour0 | lo colour1 | hi colour0 | hi colour1] after the owned rows
    DevBuf<int> map_lo, map_hi, need_ids[2][2];
    GSB_TRY(map_lo.alloc(halo_lo));
    GSB_TRY(map_hi.alloc(halo_hi));
    GSB_CUDA(cudaMemsetAsync(map_lo.p, 0xff, sizeof(int) * (size_t)(halo_lo > 0 ? halo_lo : 1), st));
    GSB_CUDA(cudaMemsetAsync(map_hi.p, 0xff, sizeof(int) * (size_t)(halo_hi > 0 ? halo_hi : 1), st));
    int base = n_local;
    for (int w = 0; w < 2; ++w)
        for (int c = 0; c < 2; ++c) {
            int q = w * 2 + c, len = flen[q], cnt = 0;
            if (len > 0) GSB_TRY(scan_count(f[q], len, st, &cnt));
            d->need_cnt[w][c] = cnt;
            // every ghost segment starts on its own 128-byte line: a line of x then holds either owned values or
            // the values one neighbour pushes in one colour phase, never both (the gathers may cache lines in L1)
            base = (base + 15) & ~15;
            d->ghost_start[w][c] = base;
            GSB_TRY(need_ids[w][c].alloc(cnt));
            if (cnt > 0) {
                d_ghost_place<<<(len + 255) / 256, 256, 0, st>>>(f[q].p, len, w == 0 ? row0 - halo_lo : row1, base,
                                                                 w == 0 ? map_lo.p : map_hi.p, need_ids[w][c].p);
                GSB_KERNEL_CHECK();
            }
            base += cnt;
        }
    d->n_ghost = base - n_local;
    d->ld = gsb_padded_ld(base);

    // 5. colour-major local CSR
    GSB_TRY(d->rp.alloc((int64_t)n_local + 1 + 8));
    d_perm_len<<<(n_local + 1 + 255) / 256, 256, 0, st>>>(d->perm.p, d->nat_rp.p, d->nat_cg.p, row0, n_local, d->rp.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(gsb_exclusive_scan_i32(d->rp.p, d->rp.p, (int64_t)n_local + 1, nullptr, st));
    GSB_TRY(d->ci.alloc((int64_t)nnz + 8));
    GSB_TRY(d->va.alloc((int64_t)nnz + 8));
    GSB_TRY(d->dg.alloc((int64_t)n_local + 8));
    GSB_CUDA(cudaMemsetAsync(mm.p + 2, 0, sizeof(int), st));
    d_fill_rows<<<(n_local + 127) / 128, 128, 0, st>>>(d->perm.p, d->iperm.p, d->nat_rp.p, d->nat_cg.p, d->nat_va.p,
                                                      n_local, row0, row1, halo_lo, col, map_lo.p, map_hi.p, d->rp.p,
                                                      d->ci.p, d->va.p, d->dg.p, mm.p + 2);
    GSB_KERNEL_CHECK();
    int h_bad = 0;
    GSB_CUDA(cudaMemcpyAsync(&h_bad, mm.p + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        gsb_set_error(col.c8 ? "dist: the given two-colouring is not proper for this matrix (grid width %d unused)"
                             : "dist: parity colouring with grid width %d is not proper for this matrix", W);
        return GSB_ERR_COLORING;
    }

    // 6. tell the neighbours which of their unknowns this rank reads (counts, then id lists)
    DevBuf<int> cnt_out, cnt_in;
    GSB_TRY(cnt_out.alloc(4));
    GSB_TRY(cnt_in.alloc(4));
    int h_out[4] = {d->need_cnt[0][0], d->need_cnt[0][1], d->need_cnt[1][0], d->need_cnt[1][1]};
    GSB_CUDA(cudaMemcpyAsync(cnt_out.p, h_out, sizeof(h_out), cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaMemsetAsync(cnt_in.p, 0, 4 * sizeof(int), st));
    {
        const void *sp[2] = {cnt_out.p, cnt_out.p + 2};
        void *rp2[2] = {cnt_in.p, cnt_in.p + 2};
        const size_t nb[2] = {2 * sizeof(int), 2 * sizeof(int)};
        GSB_TRY(comm_exchange_neighbours(d, sp, nb, rp2, nb, st));
    }
    int h_in[4];
    GSB_CUDA(cudaMemcpyAsync(h_in, cnt_in.p, sizeof(h_in), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    DevBuf<int> req_ids[2][2];
    for (int p = 0; p < 2; ++p)
        for (int c = 0; c < 2; ++c) {
            d->send_cnt[p][c] = d->has_peer(p) ? h_in[2 * p + c] : 0;
            GSB_TRY(req_ids[p][c].alloc(d->send_cnt[p][c]));
            GSB_TRY(d->send_idx[p][c].alloc(d->send_cnt[p][c]));
            GSB_TRY(d->sendbuf[p][c].alloc((int64_t)d->send_cnt[p][c] * GSB_MAX_RHS));
        }
    for (int c = 0; c < 2; ++c) { // id lists, one colour at a time: what this rank reads <-> what it has to provide
        const void *sp[2] = {need_ids[0][c].p, need_ids[1][c].p};
        void *rp2[2] = {req_ids[0][c].p, req_ids[1][c].p};
        const size_t sb[2] = {sizeof(int) * (size_t)d->need_cnt[0][c], sizeof(int) * (size_t)d->need_cnt[1][c]};
        const size_t rb[2] = {sizeof(int) * (size_t)d->send_cnt[0][c], sizeof(int) * (size_t)d->send_cnt[1][c]};
        GSB_TRY(comm_exchange_neighbours(d, sp, sb, rp2, rb, st));
    }
    GSB_CUDA(cudaMemsetAsync(mm.p + 3, 0, sizeof(int), st));
    for (int p = 0; p < 2; ++p)
        for (int c = 0; c < 2; ++c)
            if (d->send_cnt[p][c]) {
                d_ids_to_local<<<(d->send_cnt[p][c] + 255) / 256, 256, 0, st>>>(req_ids[p][c].p, d->send_cnt[p][c], row0,
                                                                             n_local, d->iperm.p, d->send_idx[p][c].p,
                                                                             mm.p + 3);
                GSB_KERNEL_CHECK();
            }
    GSB_CUDA(cudaMemcpyAsync(&h_bad, mm.p + 3, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        gsb_set_error("dist: a neighbour requested rows this rank does not own (strips thinner than the halo?)");
        return GSB_ERR_SHAPE;
    }
    d->ws_nrhs = 0;
    d->plan.valid = false; // the launch plan belongs to the previous matrix
    d->halo_meta = false;
    d->halo_agreed_valid = false;
    d->built = true;
    return GSB_OK;
}

extern "C" int gsb_dist_poisson_strip(gsb_dist *d, int W, int H, int y0, int y1) {
    if (!d || W < 1 || H < 1 || y0 < 0 || y1 > H || y0 >= y1) {
        gsb_set_error("dist_poisson_strip: bad argument");
        return GSB_ERR_ARG;
    }
    const int64_t n_global = (int64_t)W * H;
    if (n_global > INT32_MAX - 1) {
        gsb_set_error("dist_poisson_strip: %d x %d exceeds the int32 column range", W, H);
        return GSB_ERR_OVERFLOW;
    }
    GSB_TRY(gsb_set_device(d->device));
    cudaStream_t st = gsb_cur_stream();
    const int64_t p0 = (int64_t)y0 * W, p1 = (int64_t)y1 * W;
    const int n_local = (int)(p1 - p0);
    GSB_TRY(d->nat_rp.alloc((int64_t)n_local + 1));
    GSB_TRY(gsb_poisson_launch_row_len(W, H, p0, p1, d->nat_rp.p, st));
    GSB_TRY(gsb_exclusive_scan_i32(d->nat_rp.p, d->nat_rp.p, (int64_t)n_local + 1, nullptr, st));
    int nnz = 0;
    GSB_CUDA(cudaMemcpyAsync(&nnz, d->nat_rp.p + n_local, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    GSB_TRY(d->nat_cg.alloc(nnz));
    GSB_TRY(d->nat_va.alloc(nnz));
    GSB_TRY(gsb_poisson_launch_fill(W, H, p0, p1, d->nat_rp.p, d->nat_cg.p, d->nat_va.p, st));
    return dist_build(d, p0, n_local, n_global, W);
}

extern "C" int gsb_dist_set_colors(gsb_dist *d, const unsigned char *colors, int64_t n_global) {
    if (!d || n_global < 0 || (n_global > 0 && !colors)) return GSB_ERR_ARG;
    GSB_TRY(gsb_set_device(d->device));
    cudaStream_t st = gsb_cur_stream();
    if (n_global == 0) {
        d->colors8.release();
        return GSB_OK;
    }
    GSB_TRY(d->colors8.alloc(n_global));
    GSB_CUDA(cudaMemcpyAsync(d->colors8.p, colors, (size_t)n_global, cudaMemcpyHostToDevice, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

extern "C" int gsb_dist_matrix_rows(gsb_dist *d, const double *values, const int *row_off, const int *col_idx,
                                    int64_t row0, int n_local, int64_t n_global, int grid_width) {
    if (!d || !row_off || n_local <= 0 || row0 < 0 || row0 + n_local > n_global || grid_width < 1) {
        gsb_set_error("dist_matrix_rows: bad argument");
        return GSB_ERR_ARG;
    }
    if (n_global > INT32_MAX - 1) return GSB_ERR_OVERFLOW;
    const int nnz = row_off[n_local] - row_off[0];
    if (nnz < 0 || (nnz > 0 && (!values || !col_idx))) return GSB_ERR_ARG;
    GSB_TRY(gsb_set_device(d->device));
    cudaStream_t st = gsb_cur_stream();
    GSB_TRY(d->nat_rp.alloc((int64_t)n_local + 1));
    GSB_TRY(d->nat_cg.alloc(nnz));
    GSB_TRY(d->nat_va.alloc(nnz));
    if (row_off[0] != 0) {
        gsb_set_error("dist_matrix_rows: row_off must start at 0 (offsets are local to the strip)");
        return GSB_ERR_ARG;
    }
    GSB_CUDA(cudaMemcpyAsync(d->nat_rp.p, row_off, sizeof(int) * (size_t)(n_local + 1), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        GSB_CUDA(cudaMemcpyAsync(d->nat_cg.p, col_idx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaMemcpyAsync(d->nat_va.p, values, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    }
    return dist_build(d, row0, n_local, n_global, grid_width);
}

// ---------------------------------------------------------------------------------------------
// fused halo exchange: peer mappings and per-tile metadata
// ---------------------------------------------------------------------------------------------
struct PeerHello { // what a rank tells a neighbour so that it can write this rank's ghost slots
    cudaIpcMemHandle_t hx, hf;
    void *raw_x, *raw_f; // single-process mode: the pointers themselves
    long long ld;
    int gs[2];
    int ok;
    int pad;
};

// allocates this rank's stop-rule box and maps every other rank's (all-gather of the IPC handles); collective.
// Any failure anywhere keeps every rank on the ncclAllReduce path.
static int dist_box_setup(gsb_dist *d, cudaStream_t st) {
    if (d->box_ready || d->box_failed) return GSB_OK;
    if (d->world > GSB_DIST_MAX_WORLD) {
        d->box_failed = true;
        return GSB_OK;
    }
    // one dedicated allocation, padded to the 2 MiB granule the driver sub-allocates from: the IPC handle then
    // exports this box and nothing else
    const int64_t box_doubles = (int64_t)2 * d->world * GSB_MAX_RHS + d->world + 8; // flags: 2 * world ints
    const int64_t alloc_doubles = ((box_doubles * 8 + (2 << 20) - 1) / (2 << 20)) * ((2 << 20) / 8);
    int ok = 1;
    if (d->eps_box.alloc(alloc_doubles) != GSB_OK) ok = 0;
    struct Hello { cudaIpcMemHandle_t h; void *raw; int ok; int pad; };
    Hello hm;
    memset(&hm, 0, sizeof(hm));
    if (ok) {
        GSB_CUDA(cudaMemsetAsync(d->eps_box.p, 0, sizeof(double) * (size_t)box_doubles, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        hm.raw = d->eps_box.p;
        if (!d->lg && cudaIpcGetMemHandle(&hm.h, d->eps_box.p) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
    }
    hm.ok = ok;
    std::vector<Hello> all((size_t)d->world);
    GSB_TRY(comm_allgather_host(d, &hm, all.data(), sizeof(Hello), st));
    for (int q = 0; q < d->world; ++q) ok = ok && all[(size_t)q].ok;
    for (int q = 0; q < d->world && ok; ++q) {
        if (q == d->rank) {
            d->peer_box[q] = d->eps_box.p;
            continue;
        }
        if (d->lg) { // same process: the other rank's pointer is valid here once peer access is on
            d->peer_box[q] = (double *)all[(size_t)q].raw;
            continue;
        }
        void *pb = nullptr;
        if (cudaIpcOpenMemHandle(&pb, all[(size_t)q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        d->peer_box[q] = (double *)pb;
    }
    GSB_TRY(comm_allreduce_min(d, &ok, st));
    if (ok)
        d->box_ready = true;
    else {
        d->box_failed = true;
        fprintf(stderr, "libgsb200: rank %d: stop-rule exchange falls back to ncclAllReduce (peer mapping failed)\n", d->rank);
    }
    return GSB_OK;
}

// (re)maps the neighbours' x workspaces and flag words; collective over neighbours (same order everywhere)
static int dist_peer_setup(gsb_dist *d, cudaStream_t st) {
    if (d->world == 1 || d->peer_failed) return GSB_OK;
    if (!d->flags.p) {
        // a dedicated 2 MiB allocation: what the IPC handle exports is these flag words only
        GSB_TRY(d->flags.alloc((2 << 20) / (int64_t)sizeof(int)));
        GSB_CUDA(cudaMemsetAsync(d->flags.p, 0, 8 * sizeof(int), st));
    }
    if (!d->lg)
        for (int p = 0; p < 2; ++p) {
            if (d->peer_x[p]) cudaIpcCloseMemHandle(d->peer_x[p]);
            d->peer_x[p] = nullptr;
        }
    PeerHello mine[2], theirs[2];
    memset(mine, 0, sizeof(mine));
    memset(theirs, 0, sizeof(theirs));
    int ok = 1;
    cudaIpcMemHandle_t hx, hf;
    memset(&hx, 0, sizeof(hx));
    memset(&hf, 0, sizeof(hf));
    if (d->lg) {
        // same process: direct peer access between the devices replaces the IPC mappings -- to EVERY rank's device,
        // not only the neighbours': the stop-rule exchange stores into all ranks' boxes
        for (int q = 0; q < d->world && ok; ++q) {
            if (q == d->rank) continue;
            const int pd = d->lg->dev[q];
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d->device, pd) != cudaSuccess || !can) ok = 0;
            if (ok) {
                cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
                cudaGetLastError();
            }
        }
    } else if (cudaIpcGetMemHandle(&hx, d->xw.p) != cudaSuccess || cudaIpcGetMemHandle(&hf, d->flags.p) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    for (int p = 0; p < 2; ++p) {
        mine[p].hx = hx;
        mine[p].hf = hf;
        mine[p].raw_x = d->xw.p;
        mine[p].raw_f = d->flags.p;
        mine[p].ld = d->ld;
        mine[p].gs[0] = d->ghost_start[p][0];
        mine[p].gs[1] = d->ghost_start[p][1];
        mine[p].ok = ok;
    }
    DevBuf<unsigned char> dm, dt;
    GSB_TRY(dm.alloc(2 * sizeof(PeerHello)));
    GSB_TRY(dt.alloc(2 * sizeof(PeerHello)));
    GSB_CUDA(cudaMemcpyAsync(dm.p, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    {
        const void *sp[2] = {dm.p, dm.p + sizeof(PeerHello)};
        void *rp2[2] = {dt.p, dt.p + sizeof(PeerHello)};
        const size_t nb[2] = {sizeof(PeerHello), sizeof(PeerHello)};
        GSB_TRY(comm_exchange_neighbours(d, sp, nb, rp2, nb, st));
    }
    GSB_CUDA(cudaMemcpyAsync(theirs, dt.p, sizeof(theirs), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < 2 && ok; ++p) {
        if (!d->has_peer(p)) continue;
        if (!theirs[p].ok) {
            ok = 0;
            break;
        }
        if (d->lg) {
            d->peer_x[p] = (double *)theirs[p].raw_x;
            d->peer_flags[p] = (int *)theirs[p].raw_f;
        } else {
            void *px = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&px, theirs[p].hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
                break;
            }
            d->peer_x[p] = (double *)px;
            if (!d->peer_flags[p]) {
                if (cudaIpcOpenMemHandle(&pf, theirs[p].hf, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    cudaGetLastError();
                    ok = 0;
                    break;
                }
                d->peer_flags[p] = (int *)pf;
            }
        }
        d->peer_ld[p] = theirs[p].ld;
        // this rank is the neighbour's other-side peer: it writes the ghost range the neighbour keeps for it
        d->peer_gs[p][0] = theirs[p].gs[0];
        d->peer_gs[p][1] = theirs[p].gs[1];
    }
    // all ranks must agree on the transport: one failure anywhere keeps everybody on NCCL send/recv
    GSB_TRY(comm_allreduce_min(d, &ok, st));
    if (!ok) {
        d->peer_failed = true;
        d->peer_ready = false;
        fprintf(stderr, "libgsb200: rank %d: halo exchange falls back to ncclSend/ncclRecv (peer mapping failed)\n", d->rank);
        return GSB_OK;
    }
    d->peer_ready = true;
    return GSB_OK;
}

__global__ void __launch_bounds__(256) d_tile_needs_ghost(const int *__restrict__ rp, const int *__restrict__ ci,
                                                          int row0, int row1, int n_local,
                                                          unsigned char *__restrict__ info) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int i = row0 + blockIdx.x * 256 + threadIdx.x;
    if (i < row1) {
        bool g = false;
        for (int k = rp[i]; k < rp[i + 1]; ++k) g = g || ci[k] >= n_local;
        if (g) any = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && any) info[blockIdx.x] |= 1;
}

__global__ void __launch_bounds__(256) d_mark_push(const int *__restrict__ send_idx, int cnt, int row0,
                                                   int *__restrict__ push_map, unsigned char *__restrict__ info) {
    int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= cnt) return;
    int row = send_idx[j];
    push_map[row] = j;
    info[(row - row0) / 256] = info[(row - row0) / 256] | 2; // benign race: every writer sets bit 1 of a byte that
                                                             // d_tile_needs_ghost finished writing earlier
}

// tile order (halo tiles first), info and push maps for the current plan; host builds the small order array
static int dist_halo_meta(gsb_dist *d, cudaStream_t st) {
    if (d->halo_meta) return GSB_OK;
    if (d->plan.kernel != 3 && d->plan.kernel != 4) return GSB_OK;
    for (int p = 0; p < 2; ++p) {
        GSB_TRY(d->push_map[p].alloc(d->n_local));
        GSB_CUDA(cudaMemsetAsync(d->push_map[p].p, 0xff, sizeof(int) * (size_t)d->n_local, st));
    }
    for (int c = 0; c < 2; ++c) {
        const int nt = d->plan.blocks[c];
        GSB_TRY(d->tile_info[c].alloc(nt));
        GSB_TRY(d->tile_order[c].alloc(nt));
        d->n_halo_tiles[c] = 0;
        if (nt == 0) continue;
        GSB_CUDA(cudaMemsetAsync(d->tile_info[c].p, 0, (size_t)nt, st));
        d_tile_needs_ghost<<<nt, 256, 0, st>>>(d->rp.p, d->ci.p, d->color_start[c], d->color_start[c + 1], d->n_local,
                                              d->tile_info[c].p);
        GSB_KERNEL_CHECK();
        GSB_CUDA(cudaStreamSynchronize(st));
        for (int p = 0; p < 2; ++p)
            if (d->has_peer(p) && d->send_cnt[p][c]) {
                d_mark_push<<<(d->send_cnt[p][c] + 255) / 256, 256, 0, st>>>(d->send_idx[p][c].p, d->send_cnt[p][c],
                                                                           d->color_start[c], d->push_map[p].p,
                                                                           d->tile_info[c].p);
                GSB_KERNEL_CHECK();
                GSB_CUDA(cudaStreamSynchronize(st)); // the two neighbours' marks touch the same bytes: serialise
            }
        std::vector<unsigned char> hi((size_t)nt);
        GSB_CUDA(cudaMemcpyAsync(hi.data(), d->tile_info[c].p, (size_t)nt, cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        std::vector<int> order;
        order.reserve((size_t)nt);
        for (int t = 0; t < nt; ++t)
            if (hi[t]) order.push_back(t | ((int)hi[t] << 24)); // tile | info << 24: one load in the kernel
        d->n_halo_tiles[c] = (int)order.size();
        // interior tiles: one contiguous range [a, b) when the halo tiles are a prefix and a suffix of the colour's
        // tiles (row strips) -> the kernel numbers them arithmetically
        int a = 0, b = nt;
        while (a < nt && hi[(size_t)a]) ++a;
        while (b > a && hi[(size_t)b - 1]) --b;
        bool contiguous = true;
        for (int t = a; t < b; ++t)
            if (hi[(size_t)t]) contiguous = false;
        d->interior_base[c] = contiguous ? a : -1;
        d->n_interior[c] = contiguous ? b - a : 0;
        for (int t = 0; t < nt; ++t)
            if (!hi[t]) order.push_back(t);
        GSB_CUDA(cudaMemcpyAsync(d->tile_order[c].p, order.data(), sizeof(int) * (size_t)nt, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaStreamSynchronize(st));
    }
    d->halo_meta = true;
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// solve
// ---------------------------------------------------------------------------------------------
static int dist_exchange(gsb_dist *d, int c, int nrhs, cudaStream_t st, int64_t *launches) {
    bool any = false;
    for (int p = 0; p < 2; ++p) {
        if (!d->has_peer(p)) continue;
        if (d->send_cnt[p][c]) {
            d_pack<<<(d->send_cnt[p][c] + 255) / 256, 256, 0, st>>>(d->xw.p, d->ld, d->send_idx[p][c].p,
                                                                   d->send_cnt[p][c], nrhs, d->sendbuf[p][c].p);
            GSB_KERNEL_CHECK();
            ++*launches;
        }
        any = any || d->send_cnt[p][c] || d->need_cnt[p][c];
    }
    if (d->lg) { // single-process mode (the residual only; solves always use the fused peer stores)
        for (int r = 0; r < nrhs; ++r) {
            const void *sp[2];
            void *rp2[2];
            size_t sb[2], rb[2];
            for (int p = 0; p < 2; ++p) {
                sp[p] = d->sendbuf[p][c].p + (size_t)r * d->send_cnt[p][c];
                rp2[p] = d->xw.p + r * d->ld + d->ghost_start[p][c];
                sb[p] = sizeof(double) * (size_t)d->send_cnt[p][c];
                rb[p] = sizeof(double) * (size_t)d->need_cnt[p][c];
            }
            GSB_TRY(comm_exchange_neighbours(d, sp, sb, rp2, rb, st));
        }
        return GSB_OK;
    }
    if (!any) return GSB_OK;
    GSB_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < 2; ++p) {
        if (!d->has_peer(p)) continue;
        const int peer = d->peer_rank(p);
        for (int r = 0; r < nrhs; ++r) {
            if (d->send_cnt[p][c])
                GSB_NCCL(g_nccl.Send(d->sendbuf[p][c].p + (size_t)r * d->send_cnt[p][c], d->send_cnt[p][c], ncclFloat64,
                                     peer, d->comm, st));
            if (d->need_cnt[p][c])
                GSB_NCCL(g_nccl.Recv(d->xw.p + r * d->ld + d->ghost_start[p][c], d->need_cnt[p][c], ncclFloat64, peer,
                                     d->comm, st));
        }
    }
    GSB_NCCL(g_nccl.GroupEnd());
    return GSB_OK;
}

// arguments of the fused halo exchange for colour phase c of the sweep with index `sidx` (sweeps of this call issued
// before it); flag epochs only ever grow: epoch_base = sweeps issued by earlier calls
static void dist_fill_halo_args(const gsb_dist *d, int c, long long sidx, long long epoch_base, GsbHaloArgs *ha) {
    memset(ha, 0, sizeof(*ha));
    ha->enabled = 1;
    ha->n_halo_tiles = d->n_halo_tiles[c];
    ha->interior_base = d->interior_base[c];
    ha->n_interior = d->n_interior[c];
    ha->order = d->tile_order[c].p;
    ha->info = d->tile_info[c].p; // (host-side bookkeeping; the kernels number the tiles arithmetically)
    for (int p = 0; p < 2; ++p) {
        ha->has_peer[p] = d->has_peer(p) ? 1 : 0;
        ha->push_map[p] = d->push_map[p].p;
        ha->peer_x[p] = d->peer_x[p];
        ha->peer_ld[p] = d->peer_ld[p];
        ha->peer_gs[p] = d->peer_gs[p][c];
        // this rank is peer (1-p) in the neighbour's numbering
        ha->peer_flag[p] = d->peer_flags[p] ? d->peer_flags[p] + ((1 - p) * 2 + c) : nullptr;
        ha->wait_flag[p] = d->flags.p + (p * 2 + (1 - c));
    }
    // phase 0 reads what the neighbours' phase 1 of the previous sweep pushed (nothing before the first sweep: the
    // ghosts were filled with the start value); phase 1 reads what their phase 0 of this sweep pushed
    ha->wait_epoch = c == 0 ? (sidx == 0 ? 0 : (int)(epoch_base + sidx)) : (int)(epoch_base + sidx + 1);
    ha->signal_epoch = (int)(epoch_base + sidx + 1);
    ha->counter = d->flags.p + 4 + c;
}

// the next stop-rule exchange of this handle (GsbEpsExchange): monotonic count, every rank's box
static void dist_fill_exchange(gsb_dist *d, GsbEpsExchange *ex) {
    memset(ex, 0, sizeof(*ex));
    ex->world = d->world;
    ex->rank = d->rank;
    ex->epoch = (int)(++d->xcount);
    for (int q = 0; q < d->world; ++q) ex->box[q] = d->peer_box[q];
}

extern "C" int gsb_dist_gauss_seidel_dev(gsb_dist *d, const double *b_dev, int nrhs, double epsilon,
                                         int max_iteration, const gsb_gs_options *opts_in, double *x_dev,
                                         gsb_gs_stats *stats) {
    if (!d || !b_dev || !x_dev || nrhs < 1 || nrhs > GSB_MAX_RHS) return GSB_ERR_ARG;
    if (!d->built) {
        gsb_set_error("dist_gauss_seidel: no matrix on this rank yet");
        return GSB_ERR_STATE;
    }
    GSB_TRY(gsb_set_device(d->device));
    cudaStream_t st = gsb_cur_stream();
    gsb_gs_options opts;
    gsb_gs_default_options(&opts);
    if (opts_in) opts = *opts_in;
    if (opts.check_every < 1) opts.check_every = 1;
    const int n_local = d->n_local;
    const int64_t ld = d->ld;
    if (!d->plan.valid || d->plan.requested != opts.kernel) {
        GSB_TRY(gsb_plan_build(&d->plan, d->rp.p, d->ci.p, d->color_start, 2, opts.kernel, st));
        GSB_TRY(d->partials.alloc((int64_t)(d->plan.total_blocks() + 1 + 64) * GSB_MAX_RHS));
        d->halo_meta = false;
        d->halo_agreed_valid = false;
    }
    if (d->ws_nrhs < nrhs) {
        GSB_TRY(d->xw.alloc(ld * nrhs + 128));
        GSB_TRY(d->bw.alloc(ld * nrhs + 128));
        d->ws_nrhs = nrhs;
        d->peer_ready = false; // x moved: the neighbours' mappings of it are stale
    }
    // transport of the halo: fused peer stores inside the ring kernels (default), or packed ncclSend/ncclRecv
    bool use_peer = false;
    {
        bool want = d->world > 1 && !d->peer_failed;
        const char *e = getenv("GSB_DIST_TRANSPORT");
        if (e && strcmp(e, "nccl") == 0 && !d->lg) want = false;
        use_peer = false;
        if (want) {
            if (!d->peer_ready) GSB_TRY(dist_peer_setup(d, st)); // collective; agrees on peer_failed itself
            if (d->peer_ready && !d->peer_failed) {
                if (!d->halo_agreed_valid) {
                    int cap = (d->plan.kernel == 3 || d->plan.kernel == 4) ? 1 : 0;
                    if (cap) {
                        GSB_TRY(dist_halo_meta(d, st));
                        if (!d->halo_meta || d->n_halo_tiles[0] == 0 || d->n_halo_tiles[1] == 0) cap = 0;
                        // the kernels number the interior tiles arithmetically and give every halo tile its own CTA
                        for (int c = 0; c < 2; ++c)
                            if (d->interior_base[c] < 0 || d->n_halo_tiles[c] > GSB_HALO_TILES_MAX) cap = 0;
                    }
                    GSB_TRY(comm_allreduce_min(d, &cap, st));
                    d->halo_agreed = cap != 0;
                    d->halo_agreed_valid = true;
                }
                use_peer = d->halo_agreed;
            }
        }
    }
    // measurement aid: one rank, but the halo variant of the ring kernels (no tile is a halo tile, no flag is ever
    // touched) -- isolates what the variant itself costs from what the exchange costs
    bool force_halo = false;
    if (d->world == 1 && (d->plan.kernel == 3 || d->plan.kernel == 4)) {
        const char *e = getenv("GSB_DIST_FORCE_HALO");
        if (e && atoi(e) == 1) {
            GSB_TRY(dist_halo_meta(d, st));
            if (!d->flags.p) {
                GSB_TRY(d->flags.alloc(8));
                GSB_CUDA(cudaMemsetAsync(d->flags.p, 0, 8 * sizeof(int), st));
            }
            force_halo = d->halo_meta && d->interior_base[0] >= 0 && d->interior_base[1] >= 0;
        }
    }
    if (d->lg && d->world > 1 && !use_peer) {
        gsb_set_error("dist (single process): rank %d cannot reach its neighbours' memory (peer access) or the strips "
                      "are too thin for the fused halo exchange; there is no NCCL transport in this mode", d->rank);
        return GSB_ERR_NCCL;
    }
    d->used_peer = use_peer ? 1 : 0;
    // stop-rule all-reduce: fused into the end-of-sweep kernel over peer memory, or fold + ncclAllReduce + decide
    bool fused_eps = use_peer;
    {
        const char *e = getenv("GSB_DIST_EPS");
        if (e && strcmp(e, "nccl") == 0 && !d->lg) fused_eps = false;
    }
    if (fused_eps) {
        GSB_TRY(dist_box_setup(d, st));
        fused_eps = d->box_ready;
    }
    if (d->lg && d->world > 1 && !fused_eps) {
        gsb_set_error("dist (single process): the stop-rule boxes could not be set up on rank %d", d->rank);
        return GSB_ERR_NCCL;
    }
    d->used_fused_eps = fused_eps ? 1 : 0;
    if (!d->ctl.p) GSB_TRY(d->ctl.alloc(sizeof(GsCtl)));
    if (!d->ctl_host) GSB_CUDA(cudaHostAlloc(&d->ctl_host, sizeof(GsCtl), cudaHostAllocDefault));
    GsCtl *ctl = (GsCtl *)d->ctl.p;
    GsCtl *hp = (GsCtl *)d->ctl_host;

    int64_t launches = 0;
    d_gather<<<(n_local + 255) / 256, 256, 0, st>>>(b_dev, n_local, d->perm.p, n_local, nrhs, ld, d->bw.p);
    GSB_KERNEL_CHECK();
    d_fill<<<gsb_blocks_for(ld * nrhs, 256 * 4, gsb_sm_count() * 16), 256, 0, st>>>(d->xw.p, ld * nrhs, 1.0);
    GSB_KERNEL_CHECK();
    launches += 2;
    GsCtl h;
    memset(&h, 0, sizeof(h));
    h.max_iter = max_iteration;
    h.check_every = opts.check_every;
    h.epsilon = epsilon;
    for (int r = 0; r < GSB_MAX_RHS; ++r) h.eps_last[r] = 10.0;
    h.done = !(10.0 > epsilon && 0 < max_iteration) ? 1 : 0;
    *hp = h;
    GSB_CUDA(cudaMemcpyAsync(ctl, hp, sizeof(GsCtl), cudaMemcpyHostToDevice, st));

    int batch = opts.batch_sweeps;
    if (batch <= 0) batch = 32; // identical on every rank (the flag epochs below count issued sweeps)
    const long long epoch_base = d->epoch;
    if (use_peer) {
        // every rank has refilled its ghosts (stream order) before any neighbour starts pushing into them
        GSB_TRY(comm_stream_barrier(d, (int *)ctl, st));
    }
    cudaEvent_t ev0, ev1;
    GSB_CUDA(cudaEventCreate(&ev0));
    GSB_CUDA(cudaEventCreate(&ev1));
    GSB_CUDA(cudaEventRecord(ev0, st));
    int issued = 0, status = GSB_OK;
    // GSB_TRACE_PHASES=1: CUDA events around every launch of the first sweeps, averages printed to stderr
    // (a measurement aid; the events break programmatic dependent launch, so use it with GSB_PDL=0)
    std::vector<cudaEvent_t> tev;
    const int trace_sweeps = 48;
    {
        const char *e = getenv("GSB_TRACE_PHASES");
        if (e && atoi(e) == 1) tev.reserve(4 * trace_sweeps + 4);
    }
    const bool tracing = tev.capacity() > 0;
    auto trace_mark = [&]() {
        if (!tracing || (int)tev.size() >= 4 * trace_sweeps) return;
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) == cudaSuccess) {
            cudaEventRecord(ev, st);
            tev.push_back(ev);
        }
    };
    while (!h.done && status == GSB_OK) {
        int todo = max_iteration - issued;
        if (todo > batch) todo = batch;
        if (todo <= 0) todo = 1;
        for (int s = 0; s < todo && status == GSB_OK; ++s) {
            const int sweep_no = issued + s + 1;
            trace_mark(); // 4 marks per sweep: start, after phase 0, after phase 1, after the end-of-sweep step
            const bool check = (sweep_no % opts.check_every) == 0 || sweep_no == max_iteration;
            int poff = 0;
            // opt-in (GSB_FUSED_END=1; measured slower at N = 8): the second colour phase ends the sweep itself -- fold, peer exchange, decision
            // (GsbEndArgs); needs the fused stop-rule exchange on checked sweeps and both colours non-empty
            const bool fuse_end = use_peer && fused_eps && gsb_fused_end_enabled_strips() &&
                                  gsb_plan_can_fuse_end(&d->plan, nrhs) && d->color_start[1] > d->color_start[0] &&
                                  d->color_start[2] > d->color_start[1];
            for (int c = 0; c < 2 && status == GSB_OK; ++c) {
                const int r0 = d->color_start[c], r1 = d->color_start[c + 1];
                if (r1 > r0) {
                    GsbHaloArgs ha;
                    memset(&ha, 0, sizeof(ha));
                    if (use_peer || force_halo) dist_fill_halo_args(d, c, (long long)issued + s, epoch_base, &ha);
                    GsbEndArgs ea;
                    memset(&ea, 0, sizeof(ea));
                    if (fuse_end && c == 1) {
                        ea.enabled = 1;
                        ea.checked = check ? 1 : 0;
                        ea.n_partials = gsb_plan_partial_slots(&d->plan, 0, nrhs) + gsb_plan_partial_slots(&d->plan, 1, nrhs);
                        ea.ctl = ctl;
                        ea.partials = d->partials.p;
                        if (check) {
                            ea.exchange = 1;
                            dist_fill_exchange(d, &ea.ex);
                        }
                    }
                    status = gsb_plan_launch(&d->plan, c, d->rp.p, d->ci.p, d->va.p, d->dg.p, d->bw.p, d->xw.p, ld, nrhs, check,
                                             ctl, d->partials.p + (size_t)poff * nrhs, st,
                                             (use_peer || force_halo) ? &ha : nullptr, ea.enabled ? &ea : nullptr);
                    poff += gsb_plan_partial_slots(&d->plan, c, nrhs);
                    ++launches;
                }
                if (status == GSB_OK && !use_peer) status = dist_exchange(d, c, nrhs, st, &launches);
                trace_mark();
            }
            if (status != GSB_OK) break;
            if (fuse_end) {
                // the sweep was ended by the second colour phase
            } else if (check && fused_eps) {
                GsbEpsExchange ex;
                dist_fill_exchange(d, &ex);
                status = gsb_launch_end_sweep_peer(ctl, d->partials.p, poff, nrhs, &ex, st);
                ++launches;
            } else if (check && d->world == 1) {
                status = gsb_launch_end_sweep(ctl, d->partials.p, poff, nrhs, 1, 0, st);
                ++launches;
            } else if (check) {
                status = gsb_launch_end_sweep(ctl, d->partials.p, poff, nrhs, 1, 1, st);
                if (status == GSB_OK) {
                    ncclResult_t r = g_nccl.AllReduce(ctl->eps_local, ctl->eps_last, nrhs, ncclFloat64, ncclSum, d->comm, st);
                    if (r != ncclSuccess) {
                        gsb_set_error("ncclAllReduce -> %s", g_nccl.GetErrorString(r));
                        status = GSB_ERR_NCCL;
                    }
                }
                if (status == GSB_OK) status = gsb_launch_end_sweep(ctl, d->partials.p, 0, nrhs, 1, 2, st);
                launches += 2;
            } else {
                status = gsb_launch_end_sweep(ctl, d->partials.p, 0, nrhs, 0, 2, st);
                ++launches;
            }
            trace_mark();
        }
        issued += todo;
        d->epoch = epoch_base + issued;
        if (status != GSB_OK) break;
        cudaError_t ce = cudaMemcpyAsync(hp, ctl, sizeof(GsCtl), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {
            gsb_set_error("dist sweep batch failed: %s", cudaGetErrorString(ce));
            status = GSB_ERR_CUDA;
            break;
        }
        h = *hp;
        if (h.error) {
            gsb_set_error("dist: the peer stop-rule exchange timed out on rank %d (a rank is missing)", d->rank);
            status = GSB_ERR_NCCL;
        }
    }
    cudaEventRecord(ev1, st);
    cudaEventSynchronize(ev1);
    float solve_ms = 0.f;
    cudaEventElapsedTime(&solve_ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (tracing && tev.size() >= 8) {
        double acc[3] = {0, 0, 0};
        const int ns = (int)tev.size() / 4;
        for (int q = 1; q < ns; ++q) // skip the first sweep (cold)
            for (int j = 0; j < 3; ++j) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, tev[(size_t)q * 4 + j], tev[(size_t)q * 4 + j + 1]);
                acc[j] += ms;
            }
        float span = 0.f;
        cudaEventElapsedTime(&span, tev[4], tev[(size_t)(ns - 1) * 4 + 3]);
        fprintf(stderr, "gsb trace: rank %d of %d, %d sweeps: phase0 %.1f us  phase1 %.1f us  end-of-sweep %.1f us  "
                        "sweep (span) %.1f us\n", d->rank, d->world, ns - 1, 1e3 * acc[0] / (ns - 1),
                1e3 * acc[1] / (ns - 1), 1e3 * acc[2] / (ns - 1), 1e3 * span / (ns - 1));
    }
    for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
    if (status != GSB_OK) return status;
    d_scatter<<<(n_local + 255) / 256, 256, 0, st>>>(d->xw.p, ld, d->perm.p, n_local, nrhs, n_local, x_dev);
    GSB_KERNEL_CHECK();
    ++launches;
    GSB_CUDA(cudaStreamSynchronize(st));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->sweeps = h.sweeps;
        stats->n_colors = 2;
        stats->ordering_used = GSB_ORDER_REDBLACK;
        // +10: fused peer-memory halo exchange; +20 more: the stop-rule all-reduce fused into the end-of-sweep kernel
        stats->kernel_used = gsb_plan_effective_kernel(&d->plan, nrhs) + (d->used_peer ? 10 : 0) + (d->used_fused_eps ? 20 : 0);
        stats->kernel_launches = launches;
        for (int r = 0; r < nrhs; ++r) stats->last_eps[r] = h.eps_last[r];
        stats->solve_ms = solve_ms;
    }
    return GSB_OK;
}

// ||b - A x||_2 over all strips: x's ghost values are fetched from the neighbours first
__global__ void __launch_bounds__(256) d_resid(const int *__restrict__ rp, const int *__restrict__ ci,
                                               const double *__restrict__ va, const double *__restrict__ dg,
                                               const double *__restrict__ b, const double *__restrict__ x, int n_local,
                                               double *__restrict__ acc) {
    double s2 = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n_local; i += gridDim.x * 256) {
        double s = dg[i] * x[i];
        for (int k = rp[i]; k < rp[i + 1]; ++k) s += va[k] * x[ci[k]];
        double r = b[i] - s;
        s2 += r * r;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s2 += __shfl_down_sync(0xffffffffu, s2, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, s2);
}

extern "C" int gsb_dist_residual_l2_dev(gsb_dist *d, const double *b_dev, const double *x_dev, double *out) {
    if (!d || !b_dev || !x_dev || !out) return GSB_ERR_ARG;
    if (!d->built) return GSB_ERR_STATE;
    GSB_TRY(gsb_set_device(d->device));
    cudaStream_t st = gsb_cur_stream();
    const int n_local = d->n_local;
    const int64_t ld = d->ld;
    if (d->ws_nrhs < 1) {
        GSB_TRY(d->xw.alloc(ld + 128));
        GSB_TRY(d->bw.alloc(ld + 128));
        d->ws_nrhs = 1;
    }
    int64_t launches = 0;
    d_gather<<<(n_local + 255) / 256, 256, 0, st>>>(x_dev, n_local, d->perm.p, n_local, 1, ld, d->xw.p);
    GSB_KERNEL_CHECK();
    d_gather<<<(n_local + 255) / 256, 256, 0, st>>>(b_dev, n_local, d->perm.p, n_local, 1, ld, d->bw.p);
    GSB_KERNEL_CHECK();
    for (int c = 0; c < 2; ++c) GSB_TRY(dist_exchange(d, c, 1, st, &launches));
    DevBuf<double> acc;
    GSB_TRY(acc.alloc(1));
    GSB_CUDA(cudaMemsetAsync(acc.p, 0, sizeof(double), st));
    d_resid<<<gsb_blocks_for(n_local, 256, gsb_sm_count() * 8), 256, 0, st>>>(d->rp.p, d->ci.p, d->va.p, d->dg.p, d->bw.p,
                                                                             d->xw.p, n_local, acc.p);
    GSB_KERNEL_CHECK();
    GSB_TRY(comm_allreduce_sum_dev(d, acc.p, st));
    double h = 0.0;
    GSB_CUDA(cudaMemcpyAsync(&h, acc.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    *out = sqrt(h);
    return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// Single-process multi-device solve (gsb_dist_init_local): N devices of this box, one row strip and one worker
// thread each, behind one blocking call -- the reference's call model (one host thread, synchronous solve,
// project/src/PhotoMontage/main.cpp:581).  No NCCL, no torch, no IPC: setup talks through the LocalGroup, the data
// path is peer stores inside the kernels.  Worker threads (rather than the calling thread enqueuing on N streams)
// because a 4096^2 strip of an 8-GPU solve finishes a sweep in ~70 us, i.e. 24 launches per 70 us over 8 devices --
// more than one thread can enqueue; a thread per device also reuses the per-rank code of the multi-process mode
// unchanged, so both modes produce the same bits.
// ---------------------------------------------------------------------------------------------
struct gsb_dist_group {
    int n = 0;
    gsb_dist *rank[GSB_DIST_MAX_WORLD] = {nullptr};
    LocalGroup lg;
    int64_t n_global = 0;
    int64_t row0[GSB_DIST_MAX_WORLD + 1] = {0}; // strip r owns global rows [row0[r], row0[r+1])
    bool built = false;
};

// runs f(rank) on one thread per device; first failure wins (its message becomes the caller's last error)
template <typename F>
static int group_run(gsb_dist_group *g, F f) {
    std::vector<int> rc((size_t)g->n, GSB_OK);
    std::vector<std::string> msg((size_t)g->n);
    g->lg.reset();
    auto body = [&](int r) {
        int s = gsb_set_device(g->rank[r]->device);
        if (s == GSB_OK) s = f(r);
        rc[(size_t)r] = s;
        if (s != GSB_OK) {
            msg[(size_t)r] = gsb_last_error();
            g->lg.abort(); // the other ranks' barriers return instead of waiting for this one
        }
    };
    if (g->n == 1) {
        body(0);
    } else {
        std::vector<std::thread> th;
        for (int r = 0; r < g->n; ++r) th.emplace_back(body, r);
        for (auto &t : th) t.join();
    }
    // report the failure that caused the abort, not the "another rank failed" echoes of the others
    int first = -1;
    for (int r = 0; r < g->n; ++r)
        if (rc[(size_t)r] != GSB_OK && (first < 0 || (msg[(size_t)first].find("another rank failed") != std::string::npos &&
                                                     msg[(size_t)r].find("another rank failed") == std::string::npos)))
            first = r;
    if (first >= 0) {
        gsb_set_error("rank %d (device %d): %s", first, g->rank[first]->device, msg[(size_t)first].c_str());
        return rc[(size_t)first];
    }
    return GSB_OK;
}

extern "C" int gsb_dist_init_local(gsb_dist_group **out, const int *devices, int n) {
    if (!out || !devices || n < 1 || n > GSB_DIST_MAX_WORLD) {
        gsb_set_error("dist_init_local: need 1..%d devices", GSB_DIST_MAX_WORLD);
        return GSB_ERR_ARG;
    }
    int count = 0;
    gsb_device_count(&count);
    if (count <= 0) {
        gsb_set_error("no CUDA device visible: libgsb200 has no CPU fallback");
        return GSB_ERR_NO_DEVICE;
    }
    for (int r = 0; r < n; ++r) {
        if (devices[r] < 0 || devices[r] >= count) {
            gsb_set_error("dist_init_local: device %d of %d visible", devices[r], count);
            return GSB_ERR_ARG;
        }
        for (int q = 0; q < r; ++q)
            if (devices[q] == devices[r]) {
                gsb_set_error("dist_init_local: device %d listed twice", devices[r]);
                return GSB_ERR_ARG;
            }
    }
    gsb_dist_group *g = new (std::nothrow) gsb_dist_group();
    if (!g) return GSB_ERR_ALLOC;
    g->n = n;
    g->lg.world = n;
    for (int r = 0; r < n; ++r) {
        gsb_dist *d = new (std::nothrow) gsb_dist();
        if (!d) {
            for (int q = 0; q < r; ++q) delete g->rank[q];
            delete g;
            return GSB_ERR_ALLOC;
        }
        d->rank = r;
        d->world = n;
        d->device = devices[r];
        d->lg = &g->lg;
        g->lg.dev[r] = devices[r];
        g->rank[r] = d;
    }
    *out = g;
    return GSB_OK;
}

extern "C" int gsb_dist_group_finalize(gsb_dist_group *g) {
    if (!g) return GSB_OK;
    const int keep = gsb_current_device();
    for (int r = 0; r < g->n; ++r) gsb_dist_finalize(g->rank[r]);
    gsb_set_device(keep);
    delete g;
    return GSB_OK;
}

extern "C" int gsb_dist_group_size(const gsb_dist_group *g) { return g ? g->n : 0; }

// config C4: the reference-faithful full-grid Poisson system, every strip generated on its own device
extern "C" int gsb_dist_group_poisson(gsb_dist_group *g, int W, int H) {
    if (!g || W < 1 || H < g->n) {
        gsb_set_error("dist_group_poisson: bad argument (need at least one image row per device)");
        return GSB_ERR_ARG;
    }
    const int keep = gsb_current_device();
    g->built = false;
    g->n_global = (int64_t)W * H;
    int y = 0;
    std::vector<int> y0((size_t)g->n + 1);
    for (int r = 0; r < g->n; ++r) { // as even as possible (workloads.strip_bounds)
        y0[(size_t)r] = y;
        y += H / g->n + (r < H % g->n ? 1 : 0);
        g->row0[r] = (int64_t)y0[(size_t)r] * W;
    }
    y0[(size_t)g->n] = H;
    g->row0[g->n] = g->n_global;
    for (int r = 0; r < g->n; ++r) g->rank[r]->colors8.release(); // pixel parity
    const int s = group_run(g, [&](int r) { return gsb_dist_poisson_strip(g->rank[r], W, H, y0[(size_t)r], y0[(size_t)r + 1]); });
    gsb_set_device(keep);
    g->built = s == GSB_OK;
    return s;
}

// rows [r0, r0 + n_local) of a slack CSR -> compact CSR (local offsets)
__global__ void __launch_bounds__(256) d_strip_row_len(const int *__restrict__ row_nnz, int64_t r0, int n_local,
                                                       int *__restrict__ len) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n_local) len[i] = row_nnz[r0 + i];
    if (i == n_local) len[i] = 0;
}
__global__ void __launch_bounds__(256) d_strip_copy_rows(const double *__restrict__ vals, const int *__restrict__ cols,
                                                         const int *__restrict__ row_begin, int64_t r0, int n_local,
                                                         const int *__restrict__ rp, int *__restrict__ cg,
                                                         double *__restrict__ va) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_local) return;
    const int src = row_begin[r0 + i], dst = rp[i], len = rp[i + 1] - dst;
    for (int k = 0; k < len; ++k) {
        cg[dst + k] = cols[src + k];
        va[dst + k] = vals[src + k];
    }
}
__global__ void __launch_bounds__(256) d_colors_to_u8(const int *__restrict__ c, int64_t n, unsigned char *__restrict__ o) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) o[i] = (unsigned char)c[i];
}

// Shard an assembled matrix (any entry point: initializeFromVector / Triplets / EigenRowMajor) by rows over the
// group's devices.  Needs the two-colouring of the analysis (red-black probe or the caller's colours); matrices that
// take more colours do not shard this way (replicas only).
extern "C" int gsb_dist_group_matrix(gsb_dist_group *g, gsb_matrix *m) {
    if (!g || !m) return GSB_ERR_ARG;
    if (!m->has_layout || m->n_rows != m->n_cols) {
        gsb_set_error("dist_group_matrix: needs an assembled square matrix");
        return GSB_ERR_STATE;
    }
    const int keep = gsb_current_device();
    GSB_TRY(gsb_set_device(m->device));
    if (!m->analyzed) GSB_TRY(gsb_matrix_analyze(m, GSB_ORDER_AUTO, nullptr));
    if (m->n_colors > 2) {
        gsb_set_error("dist_group_matrix: the ordering needs %d colours; row strips need a two-colouring (replicas only)",
                      m->n_colors);
        gsb_set_device(keep);
        return GSB_ERR_COLORING;
    }
    const int64_t n = m->n_rows;
    if (n < g->n) {
        gsb_set_error("dist_group_matrix: %lld rows over %d devices", (long long)n, g->n);
        gsb_set_device(keep);
        return GSB_ERR_ARG;
    }
    g->built = false;
    g->n_global = n;
    // partition: whole image rows when the matrix is a W-wide grid, else equal row counts
    const int W = m->grid_width > 1 && n % m->grid_width == 0 && n / m->grid_width >= g->n ? m->grid_width : 1;
    const int64_t units = n / W;
    int64_t u = 0;
    for (int r = 0; r < g->n; ++r) {
        g->row0[r] = u * W;
        u += units / g->n + (r < units % g->n ? 1 : 0);
    }
    g->row0[g->n] = n;
    cudaStream_t st = gsb_cur_stream();
    // the colour of every global row as bytes, once on the matrix' device, then copied to every rank's device
    DevBuf<unsigned char> c8;
    GSB_TRY(c8.alloc(n));
    d_colors_to_u8<<<gsb_blocks_for(n, 256), 256, 0, st>>>(m->colors.p, n, c8.p);
    GSB_KERNEL_CHECK();
    DevBuf<int> rp;
    DevBuf<int> cg;
    DevBuf<double> va;
    int status = GSB_OK;
    for (int r = 0; r < g->n && status == GSB_OK; ++r) {
        gsb_dist *d = g->rank[r];
        const int64_t r0 = g->row0[r];
        const int n_local = (int)(g->row0[r + 1] - r0);
        auto step = [&]() -> int {
            GSB_TRY(gsb_set_device(m->device));
            GSB_TRY(rp.alloc((int64_t)n_local + 1));
            d_strip_row_len<<<(n_local + 1 + 255) / 256, 256, 0, st>>>(m->row_nnz.p, r0, n_local, rp.p);
            GSB_KERNEL_CHECK();
            GSB_TRY(gsb_exclusive_scan_i32(rp.p, rp.p, (int64_t)n_local + 1, nullptr, st));
            int nnz = 0;
            GSB_CUDA(cudaMemcpyAsync(&nnz, rp.p + n_local, sizeof(int), cudaMemcpyDeviceToHost, st));
            GSB_CUDA(cudaStreamSynchronize(st));
            GSB_TRY(cg.alloc(nnz));
            GSB_TRY(va.alloc(nnz));
            d_strip_copy_rows<<<(n_local + 255) / 256, 256, 0, st>>>(m->vals(), m->cols.p, m->row_begin.p, r0, n_local,
                                                                    rp.p, cg.p, va.p);
            GSB_KERNEL_CHECK();
            GSB_CUDA(cudaStreamSynchronize(st));
            GSB_TRY(gsb_set_device(d->device));
            GSB_TRY(d->nat_rp.alloc((int64_t)n_local + 1));
            GSB_TRY(d->nat_cg.alloc(nnz));
            GSB_TRY(d->nat_va.alloc(nnz));
            GSB_TRY(d->colors8.alloc(n));
            GSB_CUDA(cudaMemcpyPeer(d->nat_rp.p, d->device, rp.p, m->device, sizeof(int) * (size_t)(n_local + 1)));
            if (nnz > 0) {
                GSB_CUDA(cudaMemcpyPeer(d->nat_cg.p, d->device, cg.p, m->device, sizeof(int) * (size_t)nnz));
                GSB_CUDA(cudaMemcpyPeer(d->nat_va.p, d->device, va.p, m->device, sizeof(double) * (size_t)nnz));
            }
            GSB_CUDA(cudaMemcpyPeer(d->colors8.p, d->device, c8.p, m->device, (size_t)n));
            return GSB_OK;
        };
        status = step();
    }
    gsb_set_device(m->device);
    rp.release();
    cg.release();
    va.release();
    c8.release();
    if (status == GSB_OK)
        status = group_run(g, [&](int r) {
            return dist_build(g->rank[r], g->row0[r], (int)(g->row0[r + 1] - g->row0[r]), n, W);
        });
    gsb_set_device(keep);
    g->built = status == GSB_OK;
    return status;
}

// b / x: HOST vectors, nrhs x n_global doubles (one vector after another), as gsb_gauss_seidel takes them
extern "C" int gsb_dist_group_gauss_seidel(gsb_dist_group *g, const double *b, int nrhs, double epsilon,
                                           int max_iteration, const gsb_gs_options *opts, double *x_out,
                                           gsb_gs_stats *stats) {
    if (!g || !b || !x_out || nrhs < 1 || nrhs > GSB_MAX_RHS) return GSB_ERR_ARG;
    if (!g->built) {
        gsb_set_error("dist_group_gauss_seidel: no matrix on the group yet");
        return GSB_ERR_STATE;
    }
    const int keep = gsb_current_device();
    std::vector<gsb_gs_stats> st_r((size_t)g->n);
    const int64_t n = g->n_global;
    const int s = group_run(g, [&](int r) -> int {
        gsb_dist *d = g->rank[r];
        cudaStream_t st = gsb_cur_stream();
        const int64_t r0 = g->row0[r];
        const int n_local = d->n_local;
        GSB_TRY(d->stage_b.alloc((int64_t)n_local * nrhs));
        GSB_TRY(d->stage_x.alloc((int64_t)n_local * nrhs));
        for (int c = 0; c < nrhs; ++c)
            GSB_CUDA(cudaMemcpyAsync(d->stage_b.p + (size_t)c * n_local, b + (size_t)c * n + r0, sizeof(double) * (size_t)n_local,
                                     cudaMemcpyHostToDevice, st));
        GSB_TRY(gsb_dist_gauss_seidel_dev(d, d->stage_b.p, nrhs, epsilon, max_iteration, opts, d->stage_x.p, &st_r[(size_t)r]));
        for (int c = 0; c < nrhs; ++c)
            GSB_CUDA(cudaMemcpyAsync(x_out + (size_t)c * n + r0, d->stage_x.p + (size_t)c * n_local,
                                     sizeof(double) * (size_t)n_local, cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        return GSB_OK;
    });
    gsb_set_device(keep);
    if (s != GSB_OK) return s;
    if (stats) {
        *stats = st_r[0];
        for (int r = 1; r < g->n; ++r) { // slowest rank's device time, all ranks' launches
            if (st_r[(size_t)r].solve_ms > stats->solve_ms) stats->solve_ms = st_r[(size_t)r].solve_ms;
            stats->kernel_launches += st_r[(size_t)r].kernel_launches;
        }
    }
    return GSB_OK;
}

// ||b - A x||_2 over all strips (host vectors of n_global doubles)
extern "C" int gsb_dist_group_residual_l2(gsb_dist_group *g, const double *b, const double *x, double *out) {
    if (!g || !b || !x || !out) return GSB_ERR_ARG;
    if (!g->built) return GSB_ERR_STATE;
    const int keep = gsb_current_device();
    std::vector<double> res((size_t)g->n, 0.0);
    const int s = group_run(g, [&](int r) -> int {
        gsb_dist *d = g->rank[r];
        cudaStream_t st = gsb_cur_stream();
        const int n_local = d->n_local;
        GSB_TRY(d->stage_b.alloc(n_local));
        GSB_TRY(d->stage_x.alloc(n_local));
        GSB_CUDA(cudaMemcpyAsync(d->stage_b.p, b + g->row0[r], sizeof(double) * (size_t)n_local, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaMemcpyAsync(d->stage_x.p, x + g->row0[r], sizeof(double) * (size_t)n_local, cudaMemcpyHostToDevice, st));
        return gsb_dist_residual_l2_dev(d, d->stage_b.p, d->stage_x.p, &res[(size_t)r]);
    });
    gsb_set_device(keep);
    if (s == GSB_OK) *out = res[0];
    return s;
}
