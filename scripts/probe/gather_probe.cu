// Micro-probe (round 2): what does a uniform-random 4-byte gather cost on B200 as a function of the source size, the load
// flavour and cudaLimitMaxL2FetchGranularity — and what would the two halves of a PARTITIONED gather cost (indices
// bucketed by source slab so that each slab is L2-resident; results scattered back to their positions)?
// Development aid, not product code.  Run plain for the times, and under
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
// for the DRAM bytes per launch (scripts/probe/gather_probe_table.py turns both logs into profiles/r2_gather_probe.md).
// usage: gather_probe [l2_fetch_granularity] [reps]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x; }
__global__ void fill_idx(uint64_t* idx, uint64_t n, uint64_t mod) { uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; if (i < n) idx[i] = mix(i + 12345) % mod; }
// indices already bucketed by source slab: element i reads slab (i / per_bucket), a random offset inside it
__global__ void fill_bucketed(uint64_t* idx, uint32_t* pos, uint64_t n, uint64_t slab, uint64_t per_bucket) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) { idx[i] = (i / per_bucket) * slab + mix(i + 777) % slab; pos[i] = (uint32_t)(mix(i + 999) % n); }
}

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
        else if (MODE == 6) asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
        else if (MODE == 7) asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
        else if (MODE == 8) asm volatile("ld.global.nc.L1::evict_last.L2::64B.f32 %0, [%1];" : "=f"(v[i]) : "l"(p));
    }
    if (V == 4) *(float4*)(out + g) = make_float4(v[0], v[1], v[2], v[3]);
    else if (V == 2) *(float2*)(out + g) = make_float2(v[0], v[1]);
    else { for (int i = 0; i < V; i += 4) *(float4*)(out + g + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]); }
}

// second half of a partitioned gather: bucket-ordered reads (L2-resident slab), results scattered to their positions
template <int V>
__global__ void __launch_bounds__(256) gather_scatter(const uint64_t* __restrict__ idx, const uint32_t* __restrict__ pos, const float* __restrict__ src,
                                                      float* __restrict__ out, uint64_t n) {
    uint64_t g = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * V;
    if (g >= n) return;
    float v[V]; uint32_t p[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { v[i] = __ldg(src + idx[g + i]); p[i] = pos[g + i]; }
#pragma unroll
    for (int i = 0; i < V; ++i) out[p[i]] = v[i];
}
// the scatter alone: random 4-byte stores over the output
template <int V>
__global__ void __launch_bounds__(256) scatter_only(const uint32_t* __restrict__ pos, float* __restrict__ out, uint64_t n) {
    uint64_t g = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * V;
    if (g >= n) return;
#pragma unroll
    for (int i = 0; i < V; ++i) out[pos[g + i]] = (float)i;
}

static int g_reps = 3;
template <int MODE, int V> void run(const char* name, const uint64_t* idx, const float* src, float* out, uint64_t n, int span) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    unsigned grid = (unsigned)((n / V + 255) / 256);
    gather<MODE, V><<<grid, 256>>>(idx, src, out, n);
    cudaEventRecord(a);
    for (int r = 0; r < g_reps; ++r) gather<MODE, V><<<grid, 256>>>(idx, src, out, n);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= g_reps;
    printf("ROW gather span=2^%d flavour=%s V=%d ms=%.3f ggather_s=%.1f alg_gbs=%.0f\n", span, name, V, ms, n / ms / 1e6, 16.0 * n / ms / 1e6);
}

int main(int argc, char** argv) {
    size_t lim = 0;
    if (argc > 1 && atoi(argv[1]) > 0) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, atoi(argv[1])); printf("set limit %s -> %s\n", argv[1], cudaGetErrorString(e)); }
    if (argc > 2) g_reps = atoi(argv[2]);
    CK(cudaDeviceGetLimit(&lim, cudaLimitMaxL2FetchGranularity)); printf("CONFIG l2_fetch_granularity=%zu reps=%d\n", lim, g_reps);
    const uint64_t n = 1ull << 28, m = 1ull << 30;
    uint64_t* idx; float *src, *out; uint32_t* pos;
    CK(cudaMalloc(&idx, n * 8)); CK(cudaMalloc(&src, m * 4)); CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&pos, n * 4));
    CK(cudaMemset(src, 0, m * 4));
    for (int span = 30; span >= 22; span -= 2) {
        fill_idx<<<(unsigned)(n / 256), 256>>>(idx, n, 1ull << span);
        run<0, 4>("ldg", idx, src, out, n, span);
        if (span == 30 || span == 26 || span == 24) {
            run<1, 4>("nc.L1::no_allocate", idx, src, out, n, span);
            run<2, 4>("ld.cv", idx, src, out, n, span);
            run<3, 4>("nc.no_allocate.L2::64B", idx, src, out, n, span);
            run<4, 4>("ld.cs", idx, src, out, n, span);
            run<5, 4>("ld.lu", idx, src, out, n, span);
            run<6, 4>("nc.L2::64B", idx, src, out, n, span);
            run<7, 4>("ld.L2::64B", idx, src, out, n, span);
            run<8, 4>("nc.L1::evict_last.L2::64B", idx, src, out, n, span);
            run<0, 2>("ldg", idx, src, out, n, span);
            run<0, 8>("ldg", idx, src, out, n, span);
        }
    }
    // the halves of a partitioned gather
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int slab_log2 = 22; slab_log2 <= 24; ++slab_log2) {  // 16, 32, 64 MiB slabs
        const uint64_t slab = 1ull << slab_log2, buckets = m / slab, per_bucket = n / buckets;
        fill_bucketed<<<(unsigned)(n / 256), 256>>>(idx, pos, n, slab, per_bucket);
        unsigned grid = (unsigned)((n / 4 + 255) / 256);
        gather<0, 4><<<grid, 256>>>(idx, src, out, n);
        cudaEventRecord(a);
        for (int r = 0; r < g_reps; ++r) gather<0, 4><<<grid, 256>>>(idx, src, out, n);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= g_reps;
        printf("ROW bucketed_gather_ordered_out slab=2^%d ms=%.3f ggather_s=%.1f\n", slab_log2, ms, n / ms / 1e6);
        gather_scatter<4><<<grid, 256>>>(idx, pos, src, out, n);
        cudaEventRecord(a);
        for (int r = 0; r < g_reps; ++r) gather_scatter<4><<<grid, 256>>>(idx, pos, src, out, n);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        cudaEventElapsedTime(&ms, a, b); ms /= g_reps;
        printf("ROW bucketed_gather_scatter_out slab=2^%d ms=%.3f ggather_s=%.1f\n", slab_log2, ms, n / ms / 1e6);
    }
    {
        unsigned grid = (unsigned)((n / 4 + 255) / 256);
        scatter_only<4><<<grid, 256>>>(pos, out, n);
        cudaEventRecord(a);
        for (int r = 0; r < g_reps; ++r) scatter_only<4><<<grid, 256>>>(pos, out, n);
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= g_reps;
        printf("ROW scatter_only_random_4B_over_1GiB ms=%.3f gscatter_s=%.1f\n", ms, n / ms / 1e6);
    }
    return 0;
}
