// gsb_sort.cu -- unsorted COO (triplets) -> CSR.  Replaces initializeFromTriplets (v2 :249-263),
// which upstream is a loop of insert() calls on a matrix whose row_num_nze_ was never sized
// (it segfaults and has no caller).  The semantics kept are those of the
// insert() loop on a well-formed empty matrix: the last triplet of a coordinate wins, a zero
// value leaves the coordinate empty, columns end up ascending.
//
// Device pipeline: pack (row, col) into one key -> stable LSD radix sort (8-bit digits, only the
// passes the key width needs, warp-level match_any ranking) carrying the original position ->
// keep the last entry of every key run if it is nonzero -> scan-compact -> the sorted-COO build.
#include "gsb_internal.cuh"

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_BINS 256

__global__ void __launch_bounds__(256) pack_keys(const int *__restrict__ rows, const int *__restrict__ cols, int64_t n,
                                                 int n_rows, int n_cols, unsigned long long *__restrict__ keys,
                                                 int *__restrict__ idx, int *__restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int r = rows[i], c = cols[i];
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
        atomicOr(bad, 1);
        r = 0;
        c = 0;
    }
    keys[i] = (unsigned long long)r * (unsigned long long)n_cols + (unsigned long long)c;
    idx[i] = (int)i;
}

// Tile order: warp w owns the contiguous chunk [w*32*RS_ITEMS, (w+1)*32*RS_ITEMS) of the tile and
// walks it in RS_ITEMS rounds of 32 consecutive elements, so (warp, round, lane) order == input order.
__device__ __forceinline__ int64_t rs_elem(int64_t tile_base, int warp, int round, int lane) {
    return tile_base + (int64_t)warp * 32 * RS_ITEMS + round * 32 + lane;
}

// pass 1: per-block digit histogram -> table[digit * nblocks + block]
__global__ void __launch_bounds__(RS_THREADS) rs_histogram(const unsigned long long *__restrict__ keys, int64_t n,
                                                           int shift, int nblocks, int *__restrict__ table) {
    __shared__ int cnt[RS_BINS];
    for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) cnt[d] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = rs_elem(base, warp, k, lane);
        if (i < n) atomicAdd(&cnt[(int)((keys[i] >> shift) & 0xff)], 1);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) table[(size_t)d * nblocks + blockIdx.x] = cnt[d];
}

// pass 2 (after an exclusive scan of the table): stable scatter
__global__ void __launch_bounds__(RS_THREADS) rs_scatter(const unsigned long long *__restrict__ keys_in,
                                                         const int *__restrict__ idx_in, int64_t n, int shift,
                                                         int nblocks, const int *__restrict__ table,
                                                         unsigned long long *__restrict__ keys_out,
                                                         int *__restrict__ idx_out) {
    __shared__ int wcnt[RS_WARPS][RS_BINS]; // per-warp digit counts, then running offsets
    __shared__ int goff[RS_BINS];
    for (int t = threadIdx.x; t < RS_WARPS * RS_BINS; t += RS_THREADS) (&wcnt[0][0])[t] = 0;
    for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) goff[d] = table[(size_t)d * nblocks + blockIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    unsigned long long key[RS_ITEMS];
    int id[RS_ITEMS];
    bool valid[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int64_t i = rs_elem(base, warp, k, lane);
        valid[k] = i < n;
        key[k] = valid[k] ? keys_in[i] : ~0ull;
        id[k] = valid[k] ? idx_in[i] : 0;
    }
    // count per warp
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int d = valid[k] ? (int)((key[k] >> shift) & 0xff) : -1;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid[k] && (peers & lt_mask) == 0) wcnt[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps for every digit (column-wise), in place
    for (int d = threadIdx.x; d < RS_BINS; d += RS_THREADS) {
        int run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            int c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // rank and scatter
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        int d = valid[k] ? (int)((key[k] >> shift) & 0xff) : -1;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int rank = 0;
        if (valid[k]) rank = wcnt[warp][d] + __popc(peers & lt_mask);
        __syncwarp();
        if (valid[k] && (peers & lt_mask) == 0) wcnt[warp][d] += __popc(peers);
        __syncwarp();
        if (valid[k]) {
            int64_t dst = (int64_t)goff[d] + rank;
            keys_out[dst] = key[k];
            idx_out[dst] = id[k];
        }
    }
}

// keep[i] = 1 iff entry i is the last of its key run and its value is nonzero; keep[n] = 0
template <typename T>
__global__ void __launch_bounds__(256) mark_last_nonzero(const unsigned long long *__restrict__ keys,
                                                         const int *__restrict__ idx, const T *__restrict__ vals,
                                                         int64_t n, int *__restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i > n) return;
    if (i == n) {
        keep[n] = 0;
        return;
    }
    bool last = (i == n - 1) || keys[i + 1] != keys[i];
    keep[i] = (last && vals[idx[i]] != T(0)) ? 1 : 0;
}

template <typename T>
__global__ void __launch_bounds__(256) emit_kept(const unsigned long long *__restrict__ keys,
                                                 const int *__restrict__ idx, const T *__restrict__ vals, int64_t n,
                                                 int n_cols, const int *__restrict__ pos, int *__restrict__ rows_out,
                                                 int *__restrict__ cols_out, T *__restrict__ vals_out) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int p = pos[i];
    if (pos[i + 1] == p) return;
    unsigned long long k = keys[i];
    rows_out[p] = (int)(k / (unsigned long long)n_cols);
    cols_out[p] = (int)(k % (unsigned long long)n_cols);
    vals_out[p] = vals[idx[i]];
}

static int radix_sort_pairs(DevBuf<unsigned long long> &keys, DevBuf<int> &idx, int64_t n, int key_bits,
                            cudaStream_t st) {
    if (n <= 1) return GSB_OK;
    DevBuf<unsigned long long> keys2;
    DevBuf<int> idx2, table;
    GSB_TRY(keys2.alloc(n));
    GSB_TRY(idx2.alloc(n));
    const int nblocks = (int)((n + RS_TILE - 1) / RS_TILE);
    const int64_t tsize = (int64_t)RS_BINS * nblocks;
    GSB_TRY(table.alloc(tsize));
    for (int shift = 0; shift < key_bits; shift += 8) {
        rs_histogram<<<nblocks, RS_THREADS, 0, st>>>(keys.p, n, shift, nblocks, table.p);
        GSB_KERNEL_CHECK();
        GSB_TRY(gsb_exclusive_scan_i32(table.p, table.p, tsize, nullptr, st));
        rs_scatter<<<nblocks, RS_THREADS, 0, st>>>(keys.p, idx.p, n, shift, nblocks, table.p, keys2.p, idx2.p);
        GSB_KERNEL_CHECK();
        keys.swap(keys2);
        idx.swap(idx2);
    }
    GSB_CUDA(cudaStreamSynchronize(st));
    return GSB_OK;
}

template <typename T>
static int assemble_coo_t(gsb_matrix *m, const int *rows, const int *cols, const T *vals, int64_t n, int n_rows,
                          int n_cols) {
    cudaStream_t st = gsb_cur_stream();
    DevBuf<int> d_rows, d_cols, idx, keep, bad;
    DevBuf<T> d_vals;
    DevBuf<unsigned long long> keys;
    GSB_TRY(d_rows.alloc(n));
    GSB_TRY(d_cols.alloc(n));
    GSB_TRY(d_vals.alloc(n));
    GSB_TRY(idx.alloc(n));
    GSB_TRY(keys.alloc(n));
    GSB_TRY(keep.alloc(n + 1));
    GSB_TRY(bad.alloc(1));
    GSB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    if (n > 0) {
        GSB_CUDA(cudaMemcpyAsync(d_rows.p, rows, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaMemcpyAsync(d_cols.p, cols, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
        GSB_CUDA(cudaMemcpyAsync(d_vals.p, vals, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, st));
        pack_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_rows.p, d_cols.p, n, n_rows, n_cols, keys.p, idx.p,
                                                              bad.p);
        GSB_KERNEL_CHECK();
    }
    int h_bad = 0;
    GSB_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSB_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        gsb_set_error("assemble_coo: a triplet lies outside the %d x %d matrix", n_rows, n_cols);
        return GSB_ERR_SHAPE;
    }
    int key_bits = 1;
    {
        unsigned long long span = (unsigned long long)n_rows * (unsigned long long)n_cols;
        while (key_bits < 64 && (span >> key_bits) != 0) ++key_bits;
    }
    GSB_TRY(radix_sort_pairs(keys, idx, n, key_bits, st));
    int64_t kept = 0;
    if (n > 0) {
        mark_last_nonzero<T><<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(keys.p, idx.p, d_vals.p, n, keep.p);
        GSB_KERNEL_CHECK();
        GSB_TRY(gsb_exclusive_scan_i32(keep.p, keep.p, n + 1, nullptr, st));
        int h = 0;
        GSB_CUDA(cudaMemcpyAsync(&h, keep.p + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        GSB_CUDA(cudaStreamSynchronize(st));
        kept = h;
    }
    DevBuf<int> s_rows, s_cols;
    DevBuf<T> s_vals;
    GSB_TRY(s_rows.alloc(kept));
    GSB_TRY(s_cols.alloc(kept));
    GSB_TRY(s_vals.alloc(kept));
    if (kept > 0) {
        emit_kept<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys.p, idx.p, d_vals.p, n, n_cols, keep.p, s_rows.p,
                                                                 s_cols.p, s_vals.p);
        GSB_KERNEL_CHECK();
    }
    return gsb_assemble_sorted_device<T>(m, s_rows.p, s_cols.p, s_vals.p, kept, n_rows, n_cols);
}

extern "C" int gsb_matrix_assemble_coo(gsb_matrix *m, const int *rows, const int *cols, const void *vals, int64_t n,
                                       int n_rows, int n_cols) {
    if (!m || n < 0 || n_rows <= 0 || n_cols <= 0 || (n > 0 && (!rows || !cols || !vals))) {
        gsb_set_error("assemble_coo: bad argument");
        return GSB_ERR_ARG;
    }
    if (n > (int64_t)INT32_MAX - 1) return GSB_ERR_OVERFLOW;
    GSB_TRY(gsb_set_device(m->device));
    return m->vtype == GSB_I32 ? assemble_coo_t<int>(m, rows, cols, (const int *)vals, n, n_rows, n_cols)
                               : assemble_coo_t<double>(m, rows, cols, (const double *)vals, n, n_rows, n_cols);
}
