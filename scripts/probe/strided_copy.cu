// Probe: how does DRAM throughput depend on the contiguous run length of a tiled access pattern?
// Copies a (N x N) f32 matrix tile by tile; a tile is R rows x W bytes, tiles visited b-fastest as the
// transpose kernel does.  mode 0: read tiled / write tiled (same place); mode 1: read tiled, write the
// transposed tile position with the SAME run length (pure pattern cost, no smem transpose).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int W>  // bytes per run; tile = 16 KB => R = 16384 / W rows; 256 threads move 16 B each per pass
__global__ void __launch_bounds__(256) tiled_copy(const char* __restrict__ src, char* __restrict__ dst, uint64_t n, int mode) {
    constexpr int R = 16384 / W, LPR = W / 16, RPP = 256 / LPR, PASSES = R / RPP;
    const uint64_t row_bytes = n * 4, tiles_x = row_bytes / W, tiles_y = n / R;
    for (uint64_t t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
        const uint64_t ty = t % tiles_y, tx = t / tiles_y;
        uint4 v[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const uint64_t r = ty * R + p * RPP + threadIdx.x / LPR, c = tx * W + (threadIdx.x % LPR) * 16;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[p].x), "=r"(v[p].y), "=r"(v[p].z), "=r"(v[p].w) : "l"(src + r * row_bytes + c));
        }
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            uint64_t r = ty * R + p * RPP + threadIdx.x / LPR, c = tx * W + (threadIdx.x % LPR) * 16;
            if (mode == 1) { const uint64_t ty2 = tx % tiles_y, tx2 = (tx / tiles_y) * tiles_y + ty; r = ty2 * R + p * RPP + threadIdx.x / LPR; c = (tx2 % tiles_x) * W + (threadIdx.x % LPR) * 16; }
            asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + r * row_bytes + c), "r"(v[p].x), "r"(v[p].y), "r"(v[p].z), "r"(v[p].w) : "memory");
        }
    }
}

template <int W> void run(const char* s, char* d, uint64_t n, int mode) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    tiled_copy<W><<<148 * 8, 256>>>(s, d, n, mode);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) tiled_copy<W><<<148 * 8, 256>>>(s, d, n, mode);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    printf("  run %4d B  mode %d: %.4f ms  %.0f GB/s\n", W, mode, ms, 8.0 * n * n / ms / 1e6);
}

int main() {
    const uint64_t n = 16384;
    char *s, *d;
    CK(cudaMalloc(&s, n * n * 4)); CK(cudaMalloc(&d, n * n * 4));
    CK(cudaMemset(s, 1, n * n * 4));
    for (int mode = 0; mode < 2; ++mode) { run<128>(s, d, n, mode); run<256>(s, d, n, mode); run<512>(s, d, n, mode); run<1024>(s, d, n, mode); run<2048>(s, d, n, mode); }
    return 0;
}
