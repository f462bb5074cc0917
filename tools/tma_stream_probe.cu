// tma_stream_probe.cu -- how fast can persistent CTAs stream HBM through shared memory with 1-D bulk copies?
// Same structure as gs_phase_ring (one elected producer thread per CTA, STAGES stages, mbarrier completion,
// __syncthreads hand-back), no Gauss-Seidel math: consumers read every staged byte once.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_probe tools/tma_stream_probe.cu
// usage: tma_stream_probe                      (prints a table: tile KB x copies x stages x CTAs/SM -> GB/s)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}

template <int STAGES>
__global__ void __launch_bounds__(256) probe(const double *__restrict__ src, int64_t ntiles, int tile_bytes, int ncopies,
                                             double *__restrict__ sink, int work) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    unsigned char *st0 = smem + 64;
    const int tid = threadIdx.x;
    const int cb = tile_bytes / ncopies;
    auto issue = [&](int64_t t, int s) {
        mbar_expect_tx(&full[s], (uint32_t)(cb * ncopies));
        for (int c = 0; c < ncopies; ++c)
            bulk_g2s(st0 + (size_t)s * tile_bytes + (size_t)c * cb,
                     reinterpret_cast<const unsigned char *>(src) + t * tile_bytes + (size_t)c * cb, cb, &full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < STAGES; ++s) {
            int64_t t = blockIdx.x + (int64_t)s * gridDim.x;
            if (t < ntiles) issue(t, s);
        }
    double acc = 0.0;
    int k = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
        const int s = k % STAGES;
        mbar_wait(&full[s], (k / STAGES) & 1);
        const double *d = reinterpret_cast<const double *>(st0 + (size_t)s * tile_bytes);
        for (int i = tid; i < tile_bytes / 8; i += 256) acc += d[i];
        for (int w = 0; w < work; ++w) acc = acc * 1.0000001 + 1e-9; // dependent FP64 chain: emulated compute
        __syncthreads();
        int64_t tn = t + (int64_t)STAGES * gridDim.x;
        if (tid == 0 && tn < ntiles) issue(tn, s);
    }
    if (acc == 123.456) sink[0] = acc;
}

int main() {
    const size_t bytes = (size_t)2 << 30;
    double *src, *sink;
    cudaMalloc(&src, bytes);
    cudaMalloc(&sink, 8);
    cudaMemset(src, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("tileKB copies stages ctas work  GB/s\n");
    const int tiles_kb[] = {16, 29, 49};
    const int copies[] = {1, 4, 18};
    for (int tk : tiles_kb)
        for (int nc : copies)
            for (int stages = 2; stages <= 4; ++stages)
                for (int ctas = 1; ctas <= 6; ++ctas)
                    for (int work : {0, 200}) {
                        int tile_bytes = tk * 1024 / (16 * nc) * (16 * nc);
                        size_t smem = 64 + (size_t)stages * tile_bytes;
                        if (smem * ctas > 225 * 1024) continue;
                        if (work && !(nc == 4 && stages == 2)) continue;
                        int64_t ntiles = bytes / tile_bytes;
                        int grid = 148 * ctas;
                        float best = 1e9f;
                        for (int rep = 0; rep < 3; ++rep) {
                            cudaEventRecord(e0);
                            if (stages == 2) {
                                cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                                probe<2><<<grid, 256, smem>>>(src, ntiles, tile_bytes, nc, sink, work);
                            } else if (stages == 3) {
                                cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                                probe<3><<<grid, 256, smem>>>(src, ntiles, tile_bytes, nc, sink, work);
                            } else {
                                cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                                probe<4><<<grid, 256, smem>>>(src, ntiles, tile_bytes, nc, sink, work);
                            }
                            cudaEventRecord(e1);
                            cudaEventSynchronize(e1);
                            float ms;
                            cudaEventElapsedTime(&ms, e0, e1);
                            if (ms < best) best = ms;
                        }
                        cudaError_t err = cudaGetLastError();
                        if (err != cudaSuccess) {
                            printf("error %s\n", cudaGetErrorString(err));
                            return 1;
                        }
                        printf("%6d %6d %6d %4d %4d  %7.0f\n", tk, nc, stages, ctas, work,
                               (double)ntiles * tile_bytes / (best * 1e-3) / 1e9);
                    }
    return 0;
}
