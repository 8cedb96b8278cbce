// k_fold.cu — K4: last-axis fold in the reference's SEQUENTIAL order, optionally fused with a
// broadcast epilogue, one HBM pass (sm_100a; TMA bulk copies + mbarrier pipeline).
//
// The reference has no reduce API; a fold over an axis is spelled
//     a.rows::<I,J>().map(|row| { let mut s = init; row.each(|x| s = s (op) x); s })
// (src/view.rs:617-622 rows, :1341 Rows::at, :250-252 each → src/int.rs:23-25), i.e. a strictly
// left-to-right accumulation in index order.  For f32 that order IS the parity contract
// (SURVEY.md fact 3 / §7 hard part 6), so this kernel never uses a shuffle tree: it parallelises
// ACROSS rows and keeps each row's add chain serial, which is bit-exact with the reference.
//
// Layout / schedule.  Rows are contiguous (n_rows x row_len 4-byte elements).  Each WARP owns a
// ring of STAGES shared-memory tiles of 32 rows (pitch = row_len + pad words, pitch/4 odd).  Lane 0
// issues one `cp.async.bulk` (TMA, 1-D) per row into the tile and arms the tile's mbarrier with
// the byte count; while that lands, the warp works on the previous tile:
//   sum      : lane l walks row l with 16-byte shared loads (conflict-free because pitch/4 is odd)
//              and a dependent chain of row_len adds — 32 independent chains per warp;
//   epilogue : (fused form, BASELINE config 4) the whole warp re-reads the tile row by row,
//              applies  x (eop) g(fold[row])  with the row's fold broadcast by shuffle, and writes
//              coalesced 16-byte streaming stores; or (fold-only form) lane l stores fold[l].
// The input is read from HBM exactly once for fold + broadcast-subtract: algorithmic bytes
// = 2 x n_rows x row_len x 4.  No block-level barrier anywhere: warps only use
// their own mbarriers and __syncwarp.
#include "kernels.cuh"

namespace mdim {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA: global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float apply_f32(int op, float a, float b) {
    switch (op) {
        case MDIM_ADD: return __fadd_rn(a, b);
        case MDIM_SUB: return __fsub_rn(a, b);
        case MDIM_MUL: return __fmul_rn(a, b);
        case MDIM_DIV: return __fdiv_rn(a, b);
        case MDIM_REM: return fmodf(a, b);
    }
    return a;
}

// generic element op through the shared value semantics (ints wrap; errors cannot be reported from
// here, so the planner only routes float ops and non-trapping integer ops to this kernel)
__device__ __forceinline__ uint32_t apply_any(int dtype, int op, uint32_t a, uint32_t b) {
    if (dtype == MDIM_F32) return __float_as_uint(apply_f32(op, __uint_as_float(a), __uint_as_float(b)));
    bool arith = false;
    return bin_op<uint32_t>(dtype, op, dtype, a, b, arith);
}

constexpr int kRowsPerTile = 32;  // one row per lane

// FAST = f32 with fold op ADD and epilogue op SUB (or none): the shape of BASELINE config 4.
template <bool FAST, bool TMA>
__global__ void __launch_bounds__(kFoldThreads)
k_fold_rows(const __grid_constant__ FoldRowsPlan R, void* __restrict__ out_v, int n_warps, int stages, int pitch) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= n_warps) return;
    const uint32_t tile_words = (uint32_t)kRowsPerTile * (uint32_t)pitch;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // [n_warps][stages], 8 B each (<= 128 B)
    uint32_t* tiles = reinterpret_cast<uint32_t*>(smem_raw + 128) + (size_t)warp * stages * tile_words;
    uint64_t* my_bars = bars + warp * stages;

    const uint32_t row_len = R.row_len, q4 = row_len >> 2, row_bytes = row_len * 4u;
    const char* __restrict__ src = (const char*)R.src + R.src_offset * 4;
    const uint64_t n_tiles = (R.n_rows + kRowsPerTile - 1) / kRowsPerTile;
    const uint64_t total_warps = (uint64_t)gridDim.x * n_warps;
    const uint64_t gw = (uint64_t)blockIdx.x * n_warps + warp;

    auto issue = [&](uint64_t tile, int s) {  // lane 0 only
        const uint64_t row0 = tile * kRowsPerTile;
        const uint32_t nrows = (uint32_t)min((uint64_t)kRowsPerTile, R.n_rows - row0);
        mbar_expect_tx(&my_bars[s], nrows * row_bytes);
        uint32_t* dst = tiles + (size_t)s * tile_words;
        for (uint32_t r = 0; r < nrows; ++r) tma_load_1d(dst + r * pitch, src + (row0 + r) * (uint64_t)row_bytes, row_bytes, &my_bars[s]);
    };

    if constexpr (TMA) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) mbar_init(&my_bars[s], 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // make the inits visible to the TMA unit
        }
        __syncwarp();
        if (lane == 0)
            for (int s = 0; s < stages; ++s) {
                const uint64_t t = gw + (uint64_t)s * total_warps;
                if (t < n_tiles) issue(t, s);
            }
    }

    uint64_t it = 0;
    for (uint64_t tile = gw; tile < n_tiles; tile += total_warps, ++it) {
        const int s = (int)(it % (uint64_t)stages);
        const uint32_t parity = (uint32_t)((it / (uint64_t)stages) & 1u);
        const uint64_t row0 = tile * kRowsPerTile;
        const uint32_t nrows = (uint32_t)min((uint64_t)kRowsPerTile, R.n_rows - row0);
        uint32_t* tile_s = tiles + (size_t)s * tile_words;

        if constexpr (TMA) {
            mbar_wait(&my_bars[s], parity);
        } else {
            // synchronous fallback: the warp copies its tile with 128-bit loads/stores
            for (uint32_t r = 0; r < nrows; ++r)
                for (uint32_t c = lane; c < q4; c += 32) {
                    uint4 v;
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src + (row0 + r) * (uint64_t)row_bytes + c * 16u));
                    *reinterpret_cast<uint4*>(tile_s + r * pitch + c * 4) = v;
                }
            __syncwarp();
        }

        // ---- sum: lane l folds row l, strictly left to right --------------------------------------
        uint32_t fold = (uint32_t)R.init;
        if ((uint32_t)lane < nrows) {
            const uint4* rowp = reinterpret_cast<const uint4*>(tile_s + lane * pitch);
            if constexpr (FAST) {
                float acc = __uint_as_float(fold);
#pragma unroll 4
                for (uint32_t k = 0; k < q4; ++k) {
                    const uint4 v = rowp[k];
                    acc = __fadd_rn(acc, __uint_as_float(v.x));
                    acc = __fadd_rn(acc, __uint_as_float(v.y));
                    acc = __fadd_rn(acc, __uint_as_float(v.z));
                    acc = __fadd_rn(acc, __uint_as_float(v.w));
                }
                fold = __float_as_uint(acc);
            } else {
                for (uint32_t k = 0; k < q4; ++k) {
                    const uint4 v = rowp[k];
                    fold = apply_any(R.dtype, R.op, fold, v.x);
                    fold = apply_any(R.dtype, R.op, fold, v.y);
                    fold = apply_any(R.dtype, R.op, fold, v.z);
                    fold = apply_any(R.dtype, R.op, fold, v.w);
                }
            }
        }
        uint32_t g = fold;
        if (R.has_post) g = apply_any(R.dtype, R.post_op, g, (uint32_t)R.post_imm);

        if (R.epilogue == 0) {
            // fold only: one value per row, coalesced 128 B per warp
            if ((uint32_t)lane < nrows) reinterpret_cast<uint32_t*>(out_v)[row0 + lane] = fold;
        } else {
            // out[row][k] = src[row][k] (eop) g[row]; Zip over (I, J) x (I, ()) — the () axis of the
            // fold is expanded by Broadcast (src/broadcast.rs:54-60), i.e. g is constant along k
            char* __restrict__ out = (char*)out_v;
            for (uint32_t r = 0; r < nrows; ++r) {
                const uint32_t gr = __shfl_sync(0xffffffffu, g, (int)r);
                const uint4* rowp = reinterpret_cast<const uint4*>(tile_s + r * pitch);
                char* orow = out + (row0 + r) * (uint64_t)row_bytes;
                for (uint32_t c = lane; c < q4; c += 32) {
                    uint4 v = rowp[c];
                    if constexpr (FAST) {
                        const float m = __uint_as_float(gr);
                        v.x = __float_as_uint(__fsub_rn(__uint_as_float(v.x), m));
                        v.y = __float_as_uint(__fsub_rn(__uint_as_float(v.y), m));
                        v.z = __float_as_uint(__fsub_rn(__uint_as_float(v.z), m));
                        v.w = __float_as_uint(__fsub_rn(__uint_as_float(v.w), m));
                    } else {
                        v.x = apply_any(R.dtype, R.eop, v.x, gr);
                        v.y = apply_any(R.dtype, R.eop, v.y, gr);
                        v.z = apply_any(R.dtype, R.eop, v.z, gr);
                        v.w = apply_any(R.dtype, R.eop, v.w, gr);
                    }
                    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(orow + c * 16u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                }
            }
        }
        __syncwarp();  // every lane is done reading this stage before it is refilled
        if constexpr (TMA) {
            const uint64_t next = tile + (uint64_t)stages * total_warps;
            if (lane == 0 && next < n_tiles) issue(next, s);
        }
    }
}

}  // namespace

void launch_fold_rows(const FoldRowsPlan& R, void* out, int sm_count, cudaStream_t stream) {
    // pitch: row_len padded so that pitch/4 is odd (16-byte shared loads of 8 consecutive rows hit
    // 8 distinct bank groups) and rows stay 16-byte aligned for the bulk copies
    int pitch = (int)R.row_len;
    if (((pitch / 4) & 1) == 0) pitch += 4;
    const size_t tile_bytes = (size_t)kRowsPerTile * pitch * 4;
    const size_t budget = 200 * 1024;
    int stages = 2, n_warps = (int)(budget / (tile_bytes * 2));
    if (n_warps < 1) { stages = 1; n_warps = (int)(budget / tile_bytes); }
    if (n_warps < 1) n_warps = 1;
    if (n_warps > kFoldThreads / 32) n_warps = kFoldThreads / 32;
    const size_t smem = 128 + (size_t)n_warps * stages * tile_bytes;
    const uint64_t n_tiles = (R.n_rows + kRowsPerTile - 1) / kRowsPerTile;
    int grid = (int)std::min<uint64_t>((n_tiles + n_warps - 1) / n_warps, (uint64_t)sm_count);
    if (grid < 1) grid = 1;
    const bool fast = R.dtype == MDIM_F32 && R.op == MDIM_ADD && (R.epilogue == 0 || R.eop == MDIM_SUB);
    static const bool use_tma = [] { const char* e = getenv("MDIM_FOLD_TMA"); return !(e && e[0] == '0'); }();
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, kFoldThreads, smem, stream>>>(R, out, n_warps, stages, pitch);
    };
    if (fast) { if (use_tma) go(k_fold_rows<true, true>); else go(k_fold_rows<true, false>); }
    else { if (use_tma) go(k_fold_rows<false, true>); else go(k_fold_rows<false, false>); }
}

}  // namespace mdim
