// gsb_ring.cuh -- pieces shared by the colour-phase kernels (gsb_phase.cu) and the fused two-colour sweep
// (gsb_fused.cu): the row body, the mbarrier / bulk-copy (TMA) wrappers and the shared-memory stage layout.
#pragma once
#include "gsb_internal.cuh"

#define GS_THREADS 256
#define GS_TILE_CAP_MAX 6144 // CSR entries per tile that still leave >= 3 CTAs per SM (72 KB each)
#define GS_UNROLL 4          // rows with up to this many off-diagonal entries take the gather-prefetch path (5-point rows)

// sigma_r = sum_j v_j * x_r[c_j] over the row's off-diagonal entries, storage order, product and sum rounded
// separately.  XV(c, r) yields x_r[c] (global memory or a shared-memory window).  For short rows all gathers are
// issued before the first is consumed; padded positions load index 0 (always valid) and are not accumulated.
template <int NRHS, typename XV>
__device__ __forceinline__ void gs_row_sigma(const int *__restrict__ crow, const double *__restrict__ vrow, int len,
                                             XV xv, double (&sig)[NRHS]) {
#pragma unroll
    for (int r = 0; r < NRHS; ++r) sig[r] = 0.0;
    if (len <= GS_UNROLL) {
        int cc[GS_UNROLL];
        double vv[GS_UNROLL], xg[GS_UNROLL][NRHS];
#pragma unroll
        for (int j = 0; j < GS_UNROLL; ++j) {
            cc[j] = j < len ? crow[j] : 0;
            vv[j] = j < len ? vrow[j] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < GS_UNROLL; ++j)
#pragma unroll
            for (int r = 0; r < NRHS; ++r) xg[j][r] = xv(cc[j], r);
#pragma unroll
        for (int j = 0; j < GS_UNROLL; ++j)
            if (j < len) {
#pragma unroll
                for (int r = 0; r < NRHS; ++r) sig[r] = __dadd_rn(sig[r], __dmul_rn(vv[j], xg[j][r]));
            }
    } else {
        for (int j = 0; j < len; ++j) {
            const int c = crow[j];
            const double v = vrow[j];
#pragma unroll
            for (int r = 0; r < NRHS; ++r) sig[r] = __dadd_rn(sig[r], __dmul_rn(v, xv(c, r)));
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// the same with an L2 eviction policy for the lines the copy brings in (createpolicy, e.g. evict_first for data
// that is streamed once, so that it does not push reusable lines out of L2)
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; pdl_wait() blocks until the
// predecessor has completed and its writes are visible.  Both are no-ops for a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// hint: bring [p, p + bytes) into L2 (bytes a multiple of 16, p 16-byte aligned); no completion tracking
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

#define GS_RING_STAGES_DEFAULT 2
#define GS_RING_STAGES_MAX 4
#define GS_WIN_MAX 4           // windows per tile
#define GS_WIN_GRANULE 64      // columns per granule (512 bytes: keeps every window 16-byte aligned)
#define GS_WIN_CAP_MAX 2048    // doubles per right-hand side per stage
#define GS_WIN_DESC 12         // ints per tile descriptor
#define GS_RING_SLOTS_MAX GSB_RING_SLOTS_MAX // upper bound of the grid = stop-rule partial slots per colour phase

struct RingLayout {
    int va_off, dg_off, b_off, xo_off, xw_off, ci_off, rp_off, hdr_off, stage_bytes, plane;
};

__host__ __device__ inline RingLayout ring_layout(int cap, int nrhs, bool check, int wcap) {
    RingLayout L;
    L.plane = GS_THREADS + 2;
    L.va_off = 0;
    L.dg_off = L.va_off + cap * 8;
    L.b_off = L.dg_off + L.plane * 8;
    L.xo_off = L.b_off + nrhs * L.plane * 8;
    L.xw_off = L.xo_off + (check ? nrhs * L.plane * 8 : 0);
    L.ci_off = L.xw_off + nrhs * wcap * 8;
    L.rp_off = L.ci_off + cap * 4;
    L.hdr_off = L.rp_off + (GS_THREADS + 8) * 4;
    L.stage_bytes = L.hdr_off + 64;
    return L;
}

