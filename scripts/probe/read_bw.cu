// Probe: what does a pure READ stream reach on this GPU?  (The fold kernels are read-only streams; the measured
// "copy" peak in MEASURED_PEAKS.json is a 1:1 read/write figure and understates it.)
// Each thread loads 32 B (one 256-bit load) per trip, grid-stride or one-shot, and XORs into a register.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int UNROLL>
__global__ void __launch_bounds__(256) read_stream(const uint4* __restrict__ src, uint64_t n32, uint32_t* sink) {
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n32; i += UNROLL * stride) {
        uint32_t r[UNROLL][8];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7])
                         : "l"((const char*)src + (i + u * stride) * 32));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc ^= r[u][k];
    }
    for (; i < n32; i += stride) {
        const uint4 a = src[2 * i], b = src[2 * i + 1];
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int UNROLL> int run(const uint4* s, uint64_t bytes, int ctas_per_sm, uint32_t* sink) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const uint64_t n32 = bytes / 32;
    const int grid = ctas_per_sm > 0 ? 148 * ctas_per_sm : (int)((n32 / UNROLL + 255) / 256);
    read_stream<UNROLL><<<grid, 256>>>(s, n32, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) read_stream<UNROLL><<<grid, 256>>>(s, n32, sink);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    printf("  unroll %d  grid %8d: %.4f ms  %.0f GB/s\n", UNROLL, grid, ms, bytes / ms / 1e6);
    return 0;
}

int main() {
    const uint64_t bytes = 4ull << 30;
    uint4* s; uint32_t* sink;
    CK(cudaMalloc(&s, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(s, 1, bytes));
    for (int c : {0, 8, 16, 32}) { run<1>(s, bytes, c, sink); run<2>(s, bytes, c, sink); run<4>(s, bytes, c, sink); }
    return 0;
}
