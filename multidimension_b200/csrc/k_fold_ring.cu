// k_fold_ring.cu — fold over the SHARDED (outermost) axis, bit-identical to the reference's sequential add chain,
// as ONE fused compute + exchange kernel per GPU over NVLink peer memory (no NCCL call on the data path).
//
// The reference sums  s = init; for i in 0..I { s = s (op) x[i][c] }  strictly in index order (src/view.rs:250-252,
// 617-622).  With the I axis cut over N GPUs (rank r owns rows [r I/N, (r+1) I/N)) that chain runs THROUGH the ranks:
// rank 0 folds its rows from `init`, hands the running values to rank 1, which continues with its rows, ... rank N-1
// ends with the result — the same sequence of roundings as one GPU folding all the rows, so the result is bit-exact
// (an all-reduce reassociates the sum: 1e-6 relative).  The chain is pipelined over column slices:
//
//   CTA (persistent, one per SM) loops over slices of 256 columns; per slice
//     producer thread : TMA (cp.async.bulk.tensor.2d, UTMALDG) streams the rank's rows of that slice through a ring of
//                       32 KB shared-memory stages (full / empty mbarriers), far ahead of the adds;
//     consumers       : 64 threads, 4 columns each: wait for the previous rank's running values of this slice (a flag in
//                       THIS GPU's memory that the previous rank sets after writing them into this GPU's inbox), then
//                       fold the staged rows in order, then store the running values into the NEXT rank's inbox
//                       (plain stores to its peer-mapped HBM), __threadfence_system, and release its flag.
//   The last rank writes the finished slice into every rank's result area and releases their final flags; every CTA
//   copies its slices from the local result area to `out` once their final flags are up, so stream order still means
//   "the kernel has completed => out is complete on this GPU".
//
// Rank r starts slice s as soon as rank r-1 has finished it, so with S slices per CTA the whole fold takes about
// T_local (1 + (N-1)/S) + (N-1) hops instead of N T_local; measured on 8 GPUs: see profiles/r2_scale_ops.md.
// Spins are bounded (~2 s): a rank that never arrives turns into MDIM_ERR_NCCL on the others, not a hung GPU.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "exec.cuh"
#include "kernels.cuh"

namespace mdim {

namespace {

constexpr int kRingThreads = 128;       // warp 0: producer (one lane); warps 1-2: consumers; warp 3: publisher (fences and flags, off the adders' path)
constexpr int kRingCons = 64;
constexpr int kRingStageBytes = 32 * 1024;
constexpr int kRingStages = 6;
constexpr int kRingMaxLocalSlices = 64;  // slices one CTA may own: each has its OWN publish barrier (used once per launch, so the
                                         // adders can run any number of slices ahead of the publisher without overrunning a phase)
constexpr unsigned long long kSpinLimit = 40ull * 1000 * 1000;  // polls of a local flag (~50 ns each): ~2 s

__device__ __forceinline__ uint32_t rsm(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n RW_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra RD_%=;\n bra RW_%=;\n RD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
// -> false on timeout
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch) {
    for (unsigned long long spins = 0; spins < kSpinLimit; ++spins) {
        if ((int32_t)(ld_acquire_sys(flag) - epoch) >= 0) return true;
        if (spins > 16) __nanosleep(40);
    }
    return false;
}

// One staged tile folded into the running values, strictly in row order: this IS the reference's add chain.  The element type
// and operator are compile-time here (bin_op's switches fold away): dispatched once per tile, not once per element.
template <class S, int DT, int OP>
__device__ __forceinline__ void fold_tile(S (&acc)[4], uint32_t p, uint32_t rows) {
    constexpr int ES = (int)sizeof(S), SC = kRingCons * 4;
#pragma unroll 8
    for (uint32_t r = 0; r < rows; ++r) {
        S x[4];
        if constexpr (ES == 4) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(p + r * SC * ES));
        else {
            asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x[0]), "=l"(x[1]) : "r"(p + r * SC * ES));
            asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(x[2]), "=l"(x[3]) : "r"(p + r * SC * ES + 16));
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { bool arith = false; acc[e] = bin_op<S>(DT, OP, DT, acc[e], x[e], arith); }
    }
}
template <class S, int DT>
__device__ __forceinline__ void fold_tile_op(int op, S (&acc)[4], uint32_t p, uint32_t rows) {
    switch (op) {
        case MDIM_ADD: fold_tile<S, DT, MDIM_ADD>(acc, p, rows); break;
        case MDIM_SUB: fold_tile<S, DT, MDIM_SUB>(acc, p, rows); break;
        case MDIM_MUL: fold_tile<S, DT, MDIM_MUL>(acc, p, rows); break;
        case MDIM_AND: fold_tile<S, DT, MDIM_AND>(acc, p, rows); break;
        case MDIM_OR: fold_tile<S, DT, MDIM_OR>(acc, p, rows); break;
        default: fold_tile<S, DT, MDIM_XOR>(acc, p, rows); break;
    }
}
template <class S>
__device__ __forceinline__ void fold_tile_dispatch(int dtype, int op, S (&acc)[4], uint32_t p, uint32_t rows) {
    if constexpr (sizeof(S) == 4) {
        if (dtype == MDIM_F32) fold_tile_op<S, MDIM_F32>(op, acc, p, rows);
        else if (dtype == MDIM_I32) fold_tile_op<S, MDIM_I32>(op, acc, p, rows);
        else fold_tile_op<S, MDIM_U32>(op, acc, p, rows);
    } else {
        if (dtype == MDIM_F64) fold_tile_op<S, MDIM_F64>(op, acc, p, rows);
        else if (dtype == MDIM_I64) fold_tile_op<S, MDIM_I64>(op, acc, p, rows);
        else fold_tile_op<S, MDIM_U64>(op, acc, p, rows);
    }
}

template <class S>  // S = uint32_t (4-byte elements) or uint64_t (8-byte)
__global__ void __launch_bounds__(kRingThreads) k_fold_ring(const __grid_constant__ CUtensorMap rows_map, const __grid_constant__ FoldRingArgs A) {
    constexpr int ES = (int)sizeof(S), EPT = 4;            // elements per consumer thread
    constexpr int SC = kRingCons * EPT;                     // columns per slice (256)
    constexpr int ROWS = kRingStageBytes / (SC * ES);       // rows per stage (32 or 16)
    extern __shared__ uint8_t ring_raw[];
    const uint32_t base = (rsm(ring_raw) + 127u) & ~127u;
    const uint32_t bars = base + kRingStages * kRingStageBytes;  // full[s] at bars + 16 s, empty[s] at bars + 16 s + 8
    const uint32_t pub_bars = bars + 16 * kRingStages;           // publish[i] for the i-th slice of this CTA: 64 consumer arrivals
    __shared__ int timed_out;
    const int tid = threadIdx.x, warp = tid >> 5;
    pdl_entry(false);  // reads the caller's rows: always waits for its predecessor in the stream
    if (tid == 0) {
        timed_out = 0;
        for (int s = 0; s < kRingStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 16 * s) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 16 * s + 8), "r"(kRingCons) : "memory");
        }
        for (int i = 0; i < kRingMaxLocalSlices; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pub_bars + 8 * i), "r"(kRingCons) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t n_slices = (A.n_cols + SC - 1) / SC;
    const uint32_t n_row_tiles = (uint32_t)((A.n_rows + ROWS - 1) / ROWS);
    uint32_t fill = 0, drain = 0;  // stage counters (producer / consumers): stage = k % kRingStages, parity = (k / kRingStages) & 1
    if (warp == 0) {
        if (tid == 0) {  // ---- producer: every row tile of every slice of this CTA, in order ------------------------------
            for (uint64_t s = blockIdx.x; s < n_slices; s += gridDim.x) {
                for (uint32_t t = 0; t < n_row_tiles; ++t, ++fill) {
                    const uint32_t st = fill % kRingStages;
                    if (fill >= kRingStages) {  // the consumers have left this stage (or given up: a peer never arrived)
                        const uint32_t parity = ((fill / kRingStages) - 1) & 1;
                        uint32_t done = 0;
                        while (!done && !*(volatile int*)&timed_out)
                            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bars + 16 * st + 8), "r"(parity) : "memory");
                        if (!done) goto producer_done;
                    }
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 16 * st), "r"((uint32_t)kRingStageBytes) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(base + st * kRingStageBytes),
                                 "l"(&rows_map), "r"((int)(s * SC)), "r"((int)(t * ROWS)), "r"(bars + 16 * st)
                                 : "memory");
                }
            }
        producer_done:
            if (*(volatile int*)&timed_out) __nanosleep(200000);  // let the loads still in flight land before the CTA gives its shared memory back
        }
    } else if (warp == 3) {  // ---- publisher: once the consumers have stored a slice's values, make them visible and raise the flag(s)
        const bool last = A.rank == A.world - 1;
        uint32_t i = 0;
        for (uint64_t s = blockIdx.x; s < n_slices; s += gridDim.x, ++i) {
            if (tid != 96) continue;
            mbar_wait_parity(pub_bars + 8 * i, 0);  // all 64 consumers have stored slice i's values (or given up)
            if (!*(volatile int*)&timed_out) {
                // the consumers' stores happen before their arrival, the arrival before this fence: ONE system-scope fence by ONE
                // thread publishes the whole slice (fence + relaxed store = release); the adders never wait for it
                __threadfence_system();
                if (last) { for (int d = 0; d < A.world; ++d) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(A.flag_final[d] + s), "r"(A.epoch) : "memory"); }
                else asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(A.next_flag_in + s), "r"(A.epoch) : "memory");
            }
        }
    } else {  // ---- consumers: 64 threads, EPT consecutive columns each -----------------------------------------------------
        const int ct = tid - 32;
        const bool first = A.rank == 0, last = A.rank == A.world - 1;
        uint32_t li = 0;  // index of the slice among this CTA's
        for (uint64_t s = blockIdx.x; s < n_slices; s += gridDim.x, ++li) {
            const uint64_t col = s * SC + (uint64_t)ct * EPT;
            S acc[EPT];
            if (first) {
#pragma unroll
                for (int e = 0; e < EPT; ++e) acc[e] = (S)A.init;
            } else {
                if (ct == 0 && !timed_out && !wait_flag(A.flag_in + s, A.epoch)) timed_out = 1;
                asm volatile("bar.sync 1, %0;" ::"n"(kRingCons) : "memory");  // consumers only
                if (timed_out) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pub_bars + 8 * li) : "memory"); continue; }  // the publisher must not wait forever
#pragma unroll
                for (int e = 0; e < EPT; ++e) {  // written by the previous rank over NVLink: bypass L1
                    if (col + e < A.n_cols) {
                        if constexpr (ES == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(acc[e]) : "l"((const S*)A.inbox + col + e));
                        else asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(acc[e]) : "l"((const S*)A.inbox + col + e));
                    } else acc[e] = 0;
                }
            }
            for (uint32_t t = 0; t < n_row_tiles; ++t, ++drain) {
                const uint32_t st = drain % kRingStages;
                mbar_wait_parity(bars + 16 * st, (drain / kRingStages) & 1);
                const uint32_t rows = min((uint64_t)ROWS, A.n_rows - (uint64_t)t * ROWS);
                const uint32_t p = base + st * kRingStageBytes + (uint32_t)ct * EPT * ES;
                fold_tile_dispatch<S>(A.dtype, A.op, acc, p, rows);
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bars + 16 * st + 8) : "memory");
            }
            // hand the running values on (or, on the last rank, publish the result to everyone)
            const int n_dst = last ? A.world : 1;
            for (int d = 0; d < n_dst; ++d) {
                S* dst = (S*)(last ? A.result[d] : A.next_inbox) + col;
#pragma unroll
                for (int e = 0; e < EPT; ++e) if (col + e < A.n_cols) dst[e] = acc[e];
            }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pub_bars + 8 * li) : "memory");  // hand the slice to the publisher and move on
        }
    }
    __syncthreads();
    // ---- every slice of this CTA: wait for the finished values (thread i polls the i-th slice's flag), then ONE parallel
    //      pass copies them from the local result area to `out` ------------------------------------------------------------
    const uint32_t my_slices = blockIdx.x < n_slices ? (uint32_t)((n_slices - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    for (uint32_t i = tid; i < my_slices && !timed_out; i += kRingThreads)
        if (!wait_flag(A.flag_final[A.rank] + (blockIdx.x + (uint64_t)i * gridDim.x), A.epoch)) timed_out = 1;
    __syncthreads();
    if (!timed_out) {
        const uint32_t total = my_slices * SC;
        for (uint32_t k = tid; k < total; k += kRingThreads) {
            const uint64_t c = (blockIdx.x + (uint64_t)(k / SC) * gridDim.x) * SC + (k % SC);
            if (c < A.n_cols) {
                S v;
                if constexpr (ES == 4) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"((const S*)A.result[A.rank] + c));
                else asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"((const S*)A.result[A.rank] + c));
                ((S*)A.out)[c] = v;
            }
        }
    }
    if (timed_out && tid == 0) atomicExch(A.error, 1u);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

// -> 0 ok; cudaError / -1 when the rows cannot be described by a tensor map (misaligned pointer or row pitch)
int launch_fold_ring(const FoldRingArgs& A, const void* rows, int sm_count, cudaStream_t stream) {
    static const EncodeTiledFn2 encode = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn2)p;
    }();
    const int es = A.esize;
    if (!encode || ((uintptr_t)rows & 15) || (A.n_cols * (uint64_t)es) % 16 != 0 || A.n_rows == 0 || A.n_cols == 0) return -1;
    const uint32_t words = (uint32_t)(es / 4);
    CUtensorMap map;
    cuuint64_t dims[2] = {A.n_cols * words, A.n_rows}, strides[1] = {A.n_cols * (uint64_t)es};
    cuuint32_t box[2] = {(cuuint32_t)(kRingCons * 4 * words), (cuuint32_t)(kRingStageBytes / (kRingCons * 4 * es))}, estr[2] = {1, 1};
    if (box[0] > 256) {  // 8-byte elements: 512 words per row exceed the 256-element box limit -> describe pairs of words as one 8-byte element
        dims[0] = A.n_cols; box[0] = kRingCons * 4;
        if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(rows), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -1;
    } else if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(rows), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -1;
    const size_t smem = (size_t)kRingStages * kRingStageBytes + 16 * kRingStages + 8 * kRingMaxLocalSlices + 256;
    const uint64_t n_slices = (A.n_cols + kRingCons * 4 - 1) / (kRingCons * 4);
    const int grid = (int)std::min<uint64_t>(n_slices, (uint64_t)sm_count);  // one CTA per SM: every CTA is resident, so waiting on peers cannot starve anyone
    if ((n_slices + grid - 1) / grid > (uint64_t)kRingMaxLocalSlices) return -1;
    cudaError_t e;
    if (es == 4) {
        static const bool ok = cudaFuncSetAttribute(k_fold_ring<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
        if (!ok) return -1;
        e = launch_pdl(k_fold_ring<uint32_t>, dim3(grid), dim3(kRingThreads), smem, stream, map, A);
    } else {
        static const bool ok = cudaFuncSetAttribute(k_fold_ring<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
        if (!ok) return -1;
        e = launch_pdl(k_fold_ring<uint64_t>, dim3(grid), dim3(kRingThreads), smem, stream, map, A);
    }
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace mdim
