// Probe (round 2): tensor-map TMA transpose variants, timed back to back on rotating buffers > L2.
//   load : cp.async.bulk.tensor (SWIZZLE_128B boxes of 128-byte rows) -> shared, mbarrier completion
//   turn : every thread moves CH x 16-byte blocks shared -> registers -> shared (conflict-free diagonal lane map)
//   store: cp.async.bulk.tensor shared -> global (bulk_group)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tr_tma_probe tr_tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

struct TrP {
    uint32_t tiles_a, tiles_b, n_tiles;
    int32_t b_fastest, skip_wait, stages, out_stages, hint;
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load2(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar, uint64_t pol, bool hint) {
    if (hint)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store2(const CUtensorMap* m, int c0, int c1, uint32_t src, uint64_t pol, bool hint) {
    if (hint)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(m), "r"(c0), "r"(c1), "r"(src), "l"(pol) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(m), "r"(c0), "r"(c1), "r"(src) : "memory");
}

template <int ES, int GA, int GB, int NT>
__global__ void __launch_bounds__(NT) k_tr_tma(const __grid_constant__ CUtensorMap src_map, const __grid_constant__ CUtensorMap dst_map, const __grid_constant__ TrP P) {
    constexpr int E = 128 / ES, CH = 16 / ES, SUB = E * 128, TILE = GA * GB * SUB, NBLK = GA * GB * 64;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (s32(smem_raw) + 1023u) & ~1023u;
    const int S = P.stages, OS = P.out_stages;
    const uint32_t in0 = base, out0 = base + S * TILE, bars = out0 + OS * TILE;
    const int tid = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t n_my = (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    uint64_t pol = 0;
    auto coords = [&](uint32_t it, int& a0, int& b0) {
        const uint32_t t = blockIdx.x + it * gridDim.x;
        uint32_t ta, tb;
        if (P.b_fastest) { tb = t % P.tiles_b; ta = t / P.tiles_b; } else { ta = t % P.tiles_a; tb = t / P.tiles_a; }
        a0 = ta * GA; b0 = tb * GB;
    };
    auto issue_load = [&](uint32_t it) {
        const int s = it % S;
        int a0, b0; coords(it, a0, b0);
        const uint32_t bar = bars + 8 * s;
        mbar_expect(bar, TILE);
#pragma unroll
        for (int ga = 0; ga < GA; ++ga) tma_load2(in0 + s * TILE + ga * (GB * SUB), &src_map, (a0 + ga) * 32, b0 * E, bar, pol, P.hint & 1);
    };
    if (tid == 0) {
        if (P.hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (!P.skip_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
        for (uint32_t it = 0; it < (uint32_t)S && it < n_my; ++it) issue_load(it);
    }
    __syncthreads();
    for (uint32_t it = 0; it < n_my; ++it) {
        const int s = it % S, o = it % OS;
        if (OS == 1 && it > 0) {
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
        }
        mbar_wait(bars + 8 * s, (it / S) & 1);
        const uint32_t ib = in0 + s * TILE, ob = out0 + o * TILE;
#pragma unroll
        for (int w0 = 0; w0 < NBLK; w0 += NT) {
            const int w = w0 + tid;
            if (NBLK % NT != 0 && w >= NBLK) break;
            const int k = w & 7, q = (w >> 3) & 3, u = w >> 5, st = u & 1, sub = u >> 1;
            const int ga = sub % GA, gb = sub / GA;
            const int c = k, bb = (k + q + 4 * st) & 7;
            uint32_t v[CH][4];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int r = gb * E + bb * CH + i;
                const uint32_t addr = ib + ga * (GB * SUB) + r * 128 + ((c ^ (r & 7)) << 4);
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]) : "r"(addr));
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int r = ga * E + c * CH + j;
                const uint32_t addr = ob + gb * (GA * SUB) + r * 128 + ((bb ^ (r & 7)) << 4);
                if constexpr (ES == 4)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[0][j]), "r"(v[1][j]), "r"(v[2][j]), "r"(v[3][j]) : "memory");
                else
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[0][2 * j]), "r"(v[0][2 * j + 1]), "r"(v[1][2 * j]), "r"(v[1][2 * j + 1]) : "memory");
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (OS > 1 && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // store(it-1) has left out[(it+1)%OS]
        __syncthreads();
        if (tid == 0) {
            int a0, b0; coords(it, a0, b0);
#pragma unroll
            for (int gb = 0; gb < GB; ++gb) tma_store2(&dst_map, (b0 + gb) * 32, a0 * E, ob + gb * (GA * SUB), pol, P.hint & 2);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (it + S < n_my) issue_load(it + S);
        }
    }
    if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (P.skip_wait) asm volatile("griddepcontrol.wait;" ::: "memory");  // completion of this grid still implies completion of its predecessors
    }
}

// ---- host ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encode_fn() {
    static EncodeFn f = [] {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        return (EncodeFn)p;
    }();
    return f;
}
static CUtensorMap make_map(void* base, uint64_t inner_words, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_rows, int promo) {
    CUtensorMap m;
    cuuint64_t dims[2] = {inner_words, rows}, strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {32, box_rows}, es[2] = {1, 1};
    CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

__global__ void fill_iota(uint32_t* p, uint64_t n) { for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = (uint32_t)i; }
// out[x * lb + y] must be src[y * la + x]   (la = len_a = source row length, lb = len_b = source rows)
template <int W>
__global__ void check_tr(const uint32_t* out, uint64_t la, uint64_t lb, unsigned long long* bad) {
    const uint64_t n = la * lb;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t x = i / lb, y = i % lb;
        for (int w = 0; w < W; ++w)
            if (out[i * W + w] != (uint32_t)((y * la + x) * W + w)) atomicAdd(bad, 1ull);
    }
}

struct Variant { const char* name; int es, ga, gb, nt; };

template <int ES, int GA, int GB, int NT>
static void launch(const CUtensorMap& sm, const CUtensorMap& dm, const TrP& P, int grid, size_t smem, bool pdl, cudaStream_t st) {
    static bool once = [] { CK(cudaFuncSetAttribute(k_tr_tma<ES, GA, GB, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); return true; }();
    (void)once;
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute attr;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization; attr.val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_tr_tma<ES, GA, GB, NT>, sm, dm, P));
}

using LaunchFn = void (*)(const CUtensorMap&, const CUtensorMap&, const TrP&, int, size_t, bool, cudaStream_t);
struct Shape { int es, ga, gb, nt; LaunchFn fn; };
#define SHAPE(ES, GA, GB, NT) {ES, GA, GB, NT, launch<ES, GA, GB, NT>}
static const Shape kShapes[] = {SHAPE(4, 2, 2, 256), SHAPE(4, 2, 2, 128), SHAPE(4, 1, 1, 64), SHAPE(4, 1, 2, 128), SHAPE(4, 2, 1, 128), SHAPE(4, 4, 2, 256), SHAPE(4, 2, 4, 256),
                                SHAPE(4, 4, 4, 512), SHAPE(4, 4, 4, 256), SHAPE(8, 2, 2, 256), SHAPE(8, 4, 4, 256), SHAPE(4, 1, 4, 128), SHAPE(4, 4, 1, 128)};

int main(int argc, char** argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);
    cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    unsigned long long* bad; CK(cudaMalloc(&bad, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    struct Size { uint64_t la, lb; int pairs; };
    const Size sizes[] = {{4096, 4096, 16}, {16384, 16384, 2}, {4096 + 64, 2048 + 32, 8}};
    for (const Size& Z : sizes) {
        for (int es : {4, 8}) {
            if (es == 8 && Z.la == 16384) continue;
            const uint64_t bytes = Z.la * Z.lb * es;
            std::vector<void*> src(Z.pairs), dst(Z.pairs);
            for (int i = 0; i < Z.pairs; ++i) {
                CK(cudaMalloc(&src[i], bytes)); CK(cudaMalloc(&dst[i], bytes));
                fill_iota<<<sms * 8, 256, 0, st>>>((uint32_t*)src[i], bytes / 4);
                CK(cudaMemsetAsync(dst[i], 0xff, bytes, st));
            }
            CK(cudaStreamSynchronize(st));
            printf("== %llu x %llu, es %d, %d rotating pairs (%.0f MiB each way)\n", (unsigned long long)Z.la, (unsigned long long)Z.lb, es, Z.pairs, bytes / 1048576.0);
            {   // context: device-to-device copy of the same bytes
                for (int i = 0; i < 3; ++i) CK(cudaMemcpyAsync(dst[i % Z.pairs], src[i % Z.pairs], bytes, cudaMemcpyDeviceToDevice, st));
                CK(cudaEventRecord(e0, st));
                for (int i = 0; i < reps; ++i) CK(cudaMemcpyAsync(dst[i % Z.pairs], src[i % Z.pairs], bytes, cudaMemcpyDeviceToDevice, st));
                CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("   cudaMemcpyAsync D2D: %.2f us  %.0f GB/s\n", ms * 1e3 / reps, 2.0 * bytes * reps / ms / 1e6);
                for (int i = 0; i < Z.pairs; ++i) CK(cudaMemsetAsync(dst[i], 0xff, bytes, st));
            }
            for (const Shape& sh : kShapes) {
                if (sh.es != es) continue;
                const int E = 128 / es, sub = E * 128, tile = sh.ga * sh.gb * sub;
                const uint32_t tiles_a = (uint32_t)((Z.la + sh.ga * E - 1) / (sh.ga * E)), tiles_b = (uint32_t)((Z.lb + sh.gb * E - 1) / (sh.gb * E));
                struct Mode { const char* name; int ctas_per_sm; int stages; int out_stages; int skip; int hint; int promo; int bfast; };
                std::vector<Mode> modes;
                const int max_res = (227 * 1024 - 2048) / (2 * tile + 1024 + 64);
                modes.push_back({"1tile/cta", 0, 1, 1, 0, 0, 2, 1});
                modes.push_back({"1tile/cta skipwait", 0, 1, 1, 1, 0, 2, 1});
                modes.push_back({"1tile/cta skipwait hint3", 0, 1, 1, 1, 3, 2, 1});
                modes.push_back({"1tile/cta skipwait promo256", 0, 1, 1, 1, 0, 3, 1});
                modes.push_back({"1tile/cta skipwait afast", 0, 1, 1, 1, 0, 2, 0});
                for (int cps : {1, 2, 4}) {
                    const int budget = (227 * 1024 - 1024) / cps - 2048;
                    int stg = (budget - 2 * tile) / tile;
                    if (stg < 2) continue;
                    if (stg > 12) stg = 12;
                    modes.push_back({"persist", cps, stg, 2, 0, 0, 2, 1});
                    modes.push_back({"persist skipwait", cps, stg, 2, 1, 0, 2, 1});
                }
                for (int mult : {16, 32}) {  // non-persistent, several tiles per CTA through a short ring
                    const int stg = 2;
                    if (max_res < 1) continue;
                    modes.push_back({mult == 16 ? "grid16/sm ring2" : "grid32/sm ring2", -mult, stg, 2, 1, 0, 2, 1});
                }
                for (const Mode& M : modes) {
                    TrP P = {};
                    P.tiles_a = tiles_a; P.tiles_b = tiles_b; P.n_tiles = tiles_a * tiles_b; P.b_fastest = M.bfast; P.skip_wait = M.skip; P.stages = M.stages; P.out_stages = M.out_stages; P.hint = M.hint;
                    const size_t smem = (size_t)(M.stages + M.out_stages) * tile + 1024 + 8 * M.stages + 64;
                    if (smem > 227 * 1024) continue;
                    int grid = M.ctas_per_sm == 0 ? (int)P.n_tiles : M.ctas_per_sm > 0 ? sms * M.ctas_per_sm : sms * (-M.ctas_per_sm);
                    if ((uint32_t)grid > P.n_tiles) grid = (int)P.n_tiles;
                    std::vector<CUtensorMap> smaps(Z.pairs), dmaps(Z.pairs);
                    const int W = es / 4;
                    for (int i = 0; i < Z.pairs; ++i) {
                        smaps[i] = make_map(src[i], Z.la * W, Z.lb, Z.la * es, sh.gb * E, M.promo);
                        dmaps[i] = make_map(dst[i], Z.lb * W, Z.la, Z.lb * es, sh.ga * E, M.promo);
                    }
                    // correctness (pair 0)
                    CK(cudaMemsetAsync(bad, 0, 8, st));
                    sh.fn(smaps[0], dmaps[0], P, grid, smem, true, st);
                    if (es == 4) check_tr<1><<<sms * 8, 256, 0, st>>>((const uint32_t*)dst[0], Z.la, Z.lb, bad);
                    else check_tr<2><<<sms * 8, 256, 0, st>>>((const uint32_t*)dst[0], Z.la, Z.lb, bad);
                    unsigned long long hbad = 0;
                    CK(cudaMemcpyAsync(&hbad, bad, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
                    CK(cudaMemsetAsync(dst[0], 0xff, bytes, st));
                    for (int i = 0; i < 5; ++i) sh.fn(smaps[i % Z.pairs], dmaps[i % Z.pairs], P, grid, smem, true, st);
                    CK(cudaEventRecord(e0, st));
                    for (int i = 0; i < reps; ++i) sh.fn(smaps[i % Z.pairs], dmaps[i % Z.pairs], P, grid, smem, true, st);
                    CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("   es%d tile %3dx%-3d nt%3d %-28s grid %6d stg %2d smem %6zu: %8.2f us %6.0f GB/s %s\n", es, sh.ga * E, sh.gb * E, sh.nt, M.name, grid, M.stages, smem,
                           ms * 1e3 / reps, 2.0 * bytes * reps / ms / 1e6, hbad ? "MISMATCH" : "ok");
                    fflush(stdout);
                }
            }
            for (int i = 0; i < Z.pairs; ++i) { CK(cudaFree(src[i])); CK(cudaFree(dst[i])); }
        }
    }
    return 0;
}
