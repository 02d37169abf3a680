// ubench.cu -- design-parameter microbenchmarks for the BEV rasteriser (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Prints: shared-memory atomic rates (add / max / 64-bit add, random vs hot addresses) and the
// bandwidth of the 16 B-read + 4 B-write streaming pattern that bounds bin_points.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// MODE 0: atomicAdd u32, 1: atomicMax u32, 2: add + add + max (3 planes), 3: atomicAdd u64, 4: plain ++ (non-atomic RMW)
template <int MODE>
__global__ void __launch_bounds__(512, 1) smem_atomics(int iters, int cells, int hot_mask, uint32_t *sink) {
    extern __shared__ uint32_t s[];
    for (int i = threadIdx.x; i < cells * 3; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t seed = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    for (int it = 0; it < iters; ++it) {
        const uint32_t r = lcg(seed);
        const uint32_t cell = (r & hot_mask) % cells;
        const uint32_t v = r >> 16 & 255;
        if (MODE == 0) atomicAdd(&s[cell], v);
        if (MODE == 1) atomicMax(&s[cell], v);
        if (MODE == 2) { atomicAdd(&s[cell], 1u); atomicAdd(&s[cells + cell], v); atomicMax(&s[2 * cells + cell], v); }
        if (MODE == 3) atomicAdd(reinterpret_cast<unsigned long long *>(s) + (cell >> 1) , (unsigned long long)v | (1ull << 40));
        if (MODE == 4) s[cell] += v;
    }
    __syncthreads();
    uint32_t acc = 0;
    for (int i = threadIdx.x; i < cells * 3; i += blockDim.x) acc += s[i];
    if (acc == 0xdeadbeef) sink[0] = acc;
}

__global__ void __launch_bounds__(256) stream_16r_4w(const float4 *__restrict__ in, uint32_t *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = __ldcs(in + i);
        out[i] = __float_as_uint(p.x) ^ __float_as_uint(p.y) ^ __float_as_uint(p.z) ^ __float_as_uint(p.w);
    }
}
__global__ void __launch_bounds__(256) stream_16r_4w_x4(const float4 *__restrict__ in, uint4 *__restrict__ out, long long n4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = __ldcs(in + 4 * i + j);
        uint4 o;
        o.x = __float_as_uint(p[0].x) ^ __float_as_uint(p[0].w);
        o.y = __float_as_uint(p[1].x) ^ __float_as_uint(p[1].w);
        o.z = __float_as_uint(p[2].x) ^ __float_as_uint(p[2].w);
        o.w = __float_as_uint(p[3].x) ^ __float_as_uint(p[3].w);
        out[i] = o;
    }
}
__global__ void __launch_bounds__(256) copy16(const float4 *__restrict__ in, float4 *__restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __ldcs(in + i);
}
__global__ void __launch_bounds__(256) read16(const float4 *__restrict__ in, uint32_t *sink, long long n) {
    uint32_t a = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float4 p = __ldcs(in + i);
        a ^= __float_as_uint(p.x) ^ __float_as_uint(p.w);
    }
    if (a == 0xdeadbeef) sink[0] = a;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s, %d SMs, smem/block optin %zu\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerBlockOptin);
    uint32_t *sink; CK(cudaMalloc(&sink, 64));
    const int cells = 16384, iters = 4096, threads = 512;
    const size_t smem = (size_t)cells * 3 * 4;
    const int grid = pr.multiProcessorCount;
#define RUN(MODE, name, hot) { \
        CK(cudaFuncSetAttribute(smem_atomics<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        float ms = time_ms([&] { smem_atomics<MODE><<<grid, threads, smem>>>(iters, cells, hot, sink); }); \
        CK(cudaGetLastError()); \
        double ops = (double)grid * threads * iters; \
        printf("%-34s %8.3f ms  %7.2f Gthread-iter/s  %6.2f iter/clk/SM@1.9GHz\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / grid / 1.9); }
    RUN(0, "smem atomicAdd random 16K cells", 0xFFFFFF);
    RUN(1, "smem atomicMax random 16K cells", 0xFFFFFF);
    RUN(2, "smem add+add+max random (3 planes)", 0xFFFFFF);
    RUN(3, "smem atomicAdd u64 random", 0xFFFFFF);
    RUN(4, "smem plain RMW random (non-atomic)", 0xFFFFFF);
    RUN(0, "smem atomicAdd hot (16 cells)", 0xF);
    RUN(2, "smem add+add+max hot (16 cells)", 0xF);
    RUN(0, "smem atomicAdd hot (256 cells)", 0xFF);

    const long long n = 100000000LL;
    float4 *in; uint32_t *out; float4 *out16;
    CK(cudaMalloc(&in, n * 16)); CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&out16, n * 16));
    CK(cudaMemset(in, 1, n * 16));
    for (int mult : {4, 8, 16, 32}) {
        const int g = pr.multiProcessorCount * mult;
        float ms = time_ms([&] { stream_16r_4w<<<g, 256>>>(in, out, n); });
        printf("stream 16B read + 4B write  grid=%5d  %7.3f ms  %7.1f GB/s\n", g, ms, n * 20.0 / ms / 1e6);
    }
    {
        const int g = pr.multiProcessorCount * 8;
        float ms = time_ms([&] { stream_16r_4w_x4<<<g, 256>>>(in, (uint4 *)out, n / 4); });
        printf("stream 16B read + 4B write (4/thread, 16B stores)  %7.3f ms  %7.1f GB/s\n", ms, n * 20.0 / ms / 1e6);
        ms = time_ms([&] { copy16<<<g, 256>>>(in, out16, n); });
        printf("copy 16B->16B                %7.3f ms  %7.1f GB/s\n", ms, n * 32.0 / ms / 1e6);
        ms = time_ms([&] { read16<<<g, 256>>>(in, sink, n); });
        printf("read-only 16B                %7.3f ms  %7.1f GB/s\n", ms, n * 16.0 / ms / 1e6);
        ms = time_ms([&] { cudaMemsetAsync(out16, 0, n * 16); });
        printf("memset 1.6 GB                %7.3f ms  %7.1f GB/s\n", ms, n * 16.0 / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
