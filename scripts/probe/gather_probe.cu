// Micro-probe: what does a uniform-random 4-byte gather from a 4 GiB source cost on B200, and does
// the L2 fetch granularity limit / load flavour change it?  (development aid, not product code)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
__global__ void fill_idx(uint64_t* idx, uint64_t n, uint64_t mod) { uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i < n) idx[i] = mix(i + 12345) % mod; }

template <int MODE, int V>
__global__ void __launch_bounds__(256) gather(const uint64_t* __restrict__ idx, const float* __restrict__ src, float* __restrict__ out, uint64_t n) {
    uint64_t g = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * V;
    if (g >= n) return;
    uint64_t k[V];
#pragma unroll
    for (int i = 0; i < V; i += 2) {
        uint4 t;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(idx + g + i));
        k[i] = t.x | ((uint64_t)t.y << 32); k[i + 1] = t.z | ((uint64_t)t.w << 32);
    }
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float* p = src + k[i];
        if (MODE == 0) v[i] = __ldg(p);
        else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
        else if (MODE == 2) asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
        else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
        else if (MODE == 4) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
        else if (MODE == 5) asm volatile("ld.global.lu.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
    }
    if (V == 4) *(float4*)(out + g) = make_float4(v[0], v[1], v[2], v[3]);
    else if (V == 2) *(float2*)(out + g) = make_float2(v[0], v[1]);
    else { for (int i = 0; i < V; i += 4) *(float4*)(out + g + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]); }
}

template <int MODE, int V> void run(const char* name, const uint64_t* idx, const float* src, float* out, uint64_t n) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    unsigned grid = (unsigned)((n / V + 255) / 256);
    gather<MODE, V><<<grid, 256>>>(idx, src, out, n);
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) gather<MODE, V><<<grid, 256>>>(idx, src, out, n);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("  %-28s V=%d  %.3f ms  %.1f Ggather/s  alg %.0f GB/s\n", name, V, ms, n / ms / 1e6, 16.0 * n / ms / 1e6);
}

int main(int argc, char** argv) {
    size_t lim = 0;
    if (argc > 1) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1])); printf("set limit %s -> %s\n", argv[1], cudaGetErrorString(e)); }
    CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity)); printf("cudaLimitMaxL2FetchGranularity = %zu\n", lim);
    const uint64_t n = 1ull << 28, m = 1ull << 30;
    uint64_t* idx; float *src, *out;
    CK(cudaMalloc(&idx, n * 8)); CK(cudaMalloc(&src, m * 4)); CK(cudaMalloc(&out, n * 4));
    CK(cudaMemset(src, 0, m * 4));
    for (int span = 30; span >= 22; span -= 4) {
        fill_idx<<<(unsigned)(n / 256), 256>>>(idx, n, 1ull << span);
        printf("source span 2^%d elements (%.0f MiB)\n", span, (4.0 * (1ull << span)) / (1 << 20));
        run<0, 4>("ldg", idx, src, out, n);
        if (span == 30) {
            run<1, 4>("nc.L1::no_allocate", idx, src, out, n);
            run<2, 4>("ld.cv", idx, src, out, n);
            run<3, 4>("nc.no_allocate.L2::64B", idx, src, out, n);
            run<4, 4>("ld.cs", idx, src, out, n);
            run<5, 4>("ld.lu", idx, src, out, n);
            run<0, 2>("ldg", idx, src, out, n);
            run<0, 8>("ldg", idx, src, out, n);
        }
    }
    return 0;
}
