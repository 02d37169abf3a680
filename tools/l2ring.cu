// l2ring.cu -- can a small, continuously recycled record ring stay resident in the 126 MB L2 while a
// 16 B/point input stream flows through it?  (design question behind the fused bin+reduce kernel)
//
// Every CTA streams its slice of a large input (16 B per element, read once) and, per element, writes
// one 4 B record into its private slice of a ring buffer; `lag` elements later it reads the record
// back (ld.global.cg: L2 only).  The ring wraps, so its lines are overwritten in place while dirty.
// Run under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`: if the ring stays in L2 the
// DRAM traffic is the input alone (no record write-back, no record read).
//
//   l2ring <ring_MB> <lag_fraction_of_ring 0..1> <policy> [input_MB=1600] [ctas_per_sm=4]
//   policy: 0 plain loads / stores          1 input evict_first (createpolicy)
//           2 input evict_first + ring evict_last (stores and loads)       3 input ld.global.cs (__ldcs)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_in(const uint4 *p, int policy, uint64_t pol_first) {
    uint4 v;
    if (policy == 1 || policy == 2) {
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol_first));
    } else if (policy == 3) {
        v = __ldcs(p);
    } else {
        v = *p;
    }
    return v;
}
__device__ __forceinline__ void st_ring(uint32_t *p, uint32_t v, int policy, uint64_t pol_last) {
    if (policy == 2) asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol_last) : "memory");
    else *p = v;
}
__device__ __forceinline__ uint32_t ld_ring(const uint32_t *p, int policy, uint64_t pol_last) {
    uint32_t v;
    if (policy == 2) asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol_last));
    else asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256) stream_ring(const uint4 *__restrict__ in, size_t n, uint32_t *ring, size_t ring_per_cta,
                                                   size_t lag, int policy, unsigned long long *sink) {
    uint64_t pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    const size_t per = (n + gridDim.x - 1) / gridDim.x;
    const size_t b0 = per * blockIdx.x, b1 = b0 + per < n ? b0 + per : n;
    uint32_t *my = ring + ring_per_cta * blockIdx.x;
    unsigned long long acc = 0;
    constexpr int U = 4;
    for (size_t k = threadIdx.x; b0 + k + (U - 1) * 256 < b1; k += U * 256) {
        uint4 v[U];
        uint32_t rd[U] = {0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ld_in(in + b0 + k + u * 256, policy, pol_first);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t kk = k + u * 256;
            st_ring(my + (kk & (ring_per_cta - 1)), v[u].x ^ v[u].y ^ v[u].z ^ v[u].w, policy, pol_last);
            if (kk >= lag) rd[u] = ld_ring(my + ((kk - lag) & (ring_per_cta - 1)), policy, pol_last);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += rd[u];
    }
    if (acc == 0x1234567ull) *sink = acc;
}

// L2-resident read bandwidth: every CTA re-reads a slice of a small buffer `reps` times (ld.global.cg)
__global__ void __launch_bounds__(256) l2_read(const uint4 *__restrict__ buf, size_t n, int reps, unsigned long long *sink) {
    unsigned long long acc = 0;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc += v.x ^ v.w;
        }
    if (acc == 0x1234567ull) *sink = acc;
}

// reference points: the input stream alone, and input + a NON-recycled record pool (what bin_points does today)
__global__ void __launch_bounds__(256) stream_only(const uint4 *__restrict__ in, size_t n, uint32_t *pool, int write_pool,
                                                   unsigned long long *sink) {
    const size_t per = (n + gridDim.x - 1) / gridDim.x;
    const size_t b0 = per * blockIdx.x, b1 = b0 + per < n ? b0 + per : n;
    unsigned long long acc = 0;
    constexpr int U = 4;
    for (size_t k = threadIdx.x; b0 + k + (U - 1) * 256 < b1; k += U * 256) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = in[b0 + k + u * 256];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            if (write_pool) pool[b0 + k + u * 256] = r;
            else acc += r;
        }
    }
    if (acc == 0x1234567ull) *sink = acc;
}

int main(int argc, char **argv) {
    const double ring_mb = argc > 1 ? atof(argv[1]) : 32.0;
    const double lag_frac = argc > 2 ? atof(argv[2]) : 0.5;
    const int policy = argc > 3 ? atoi(argv[3]) : 0;
    const double in_mb = argc > 4 ? atof(argv[4]) : 1600.0;
    const int cps = argc > 5 ? atoi(argv[5]) : 4;
    const size_t n = (size_t)(in_mb * 1e6 / 16);
    const int grid = 148 * cps;
    size_t ring_per_cta = 2048;                       // words per CTA, a power of two
    while (ring_per_cta * 2 * grid * 4 <= (size_t)(ring_mb * 1e6)) ring_per_cta *= 2;
    const size_t lag = (size_t)(ring_per_cta * lag_frac) / 1024 * 1024;
    uint4 *in;
    uint32_t *ring, *pool;
    unsigned long long *sink;
    CK(cudaMalloc(&in, n * 16));
    CK(cudaMalloc(&ring, ring_per_cta * grid * 4));
    CK(cudaMalloc(&pool, n * 4));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(in, 1, n * 16));
    CK(cudaMemset(ring, 0, ring_per_cta * grid * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        stream_only<<<grid, 256>>>(in, n, pool, 0, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 2) printf("read-only stream              : %.3f ms  %.0f GB/s\n", ms, n * 16 / ms / 1e6);
    }
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        stream_only<<<grid, 256>>>(in, n, pool, 1, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 2) printf("stream + 4 B/elt pool write   : %.3f ms  %.0f GB/s (16 B alg)\n", ms, n * 16 / ms / 1e6);
    }
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        stream_ring<<<grid, 256>>>(in, n, ring, ring_per_cta, lag, policy, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 2)
            printf("ring %.1f MB (%zu KB/CTA) lag %.2f policy %d: %.3f ms  %.0f GB/s (16 B alg)\n", ring_per_cta * grid * 4 / 1e6,
                   ring_per_cta * 4 / 1024, lag_frac, policy, ms, n * 16 / ms / 1e6);
    }
    for (int mb = 8; mb <= 64; mb *= 2) {
        const size_t m = (size_t)mb * 1000000 / 16;
        l2_read<<<grid, 256>>>(in, m, 2, sink);
        CK(cudaEventRecord(e0));
        l2_read<<<grid, 256>>>(in, m, 20, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("L2-resident re-read of %d MB x 20: %.3f ms  %.0f GB/s\n", mb, ms, 20.0 * m * 16 / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
