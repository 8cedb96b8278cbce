/*
 * ref_shaped.c — CPU ORACLE / BASELINE (test infrastructure, NOT product code).
 *
 * The five BASELINE.json configs written the way rustc monomorphises the reference crate's
 * `collect()` for them: one specialised loop nest per expression type, no interpretation.
 * It is the fair single-threaded CPU baseline beside the GPU numbers (mdim_oracle.c interprets a
 * descriptor per element and is ~10x slower than compiled Rust would be), and it is cross-checked
 * against mdim_oracle.c in tests/test_ref_shaped.py.
 *
 * What is kept from the reference, per element:
 *   - Index::each nested loops, last axis fastest            src/index.rs:122-124, src/int.rs:23-25
 *   - one bounds assert per usize index component            src/int.rs:16-19
 *   - the slice bounds check of items[..]                    src/array.rs:86
 *   - Vec::push with its capacity check into a with_capacity buffer   src/array.rs:99-113
 *   - f32 multiply and add rounded separately (no FMA), sequential-order folds   src/view.rs:250-252
 * Compile with -O2 -fno-tree-vectorize -ffp-contract=off: README.md:11-12 "not SIMD optimized".
 * Parity status: pinned by semantics only (the reference has no test at these sizes), and by
 * agreement with oracle/mdim_oracle.c + oracle/reference_model.py, which ARE pinned by the doctests.
 */
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>

/* The product x * y must be rounded to f32 before the add (Rust never contracts x * y + 1.0 into an FMA).  The default
 * build forces that with a volatile temporary; the -O3 -march=x86-64-v3 build (libmdim_refshaped_o3.so, the most the
 * reference could get from `cargo build --release` with the vectoriser on) relies on -ffp-contract=off instead. */
#ifdef REF_NO_VOLATILE
#define SEPARATELY_ROUNDED float
#else
#define SEPARATELY_ROUNDED volatile float
#endif

typedef struct { float* ptr; uint64_t len, cap; } vec_f32;

static void grow(vec_f32* v) { /* never reached: capacity is exact, as with Vec::with_capacity */
    v->cap = v->cap ? v->cap * 2 : 4;
    v->ptr = (float*)realloc(v->ptr, v->cap * sizeof(float));
    if (!v->ptr) abort();
}
static inline void push(vec_f32* v, float x) {
    if (__builtin_expect(v->len == v->cap, 0)) grow(v);
    v->ptr[v->len++] = x;
}
static void oob(uint64_t i, uint64_t n) {
    fprintf(stderr, "Index %llu is out of bounds for size %llu\n", (unsigned long long)i, (unsigned long long)n);
    abort();
}
#define CHECK(i, n) do { if (__builtin_expect(!((i) < (n)), 0)) oob((i), (n)); } while (0)

/* C2: a.zip(b).map(|(x, y)| x * y + 1.0).collect() over Array<usize, f32> */
int ref_c2_zip_map(const float* a, const float* b, uint64_t n, float* out) {
    vec_f32 v = { out, 0, n };
    for (uint64_t i = 0; i < n; ++i) {
        CHECK(i, n); CHECK(i, n);          /* usize::to_usize assert + slice check, operand a */
        float x = a[i];
        CHECK(i, n); CHECK(i, n);          /* operand b */
        float y = b[i];
        SEPARATELY_ROUNDED m = x * y;          /* separately rounded */
        push(&v, m + 1.0f);
    }
    return v.len == n ? 0 : 2;             /* Array::new_inner assert, src/array.rs:12 */
}

/* C1: a.transpose::<(), usize, usize, ()>().collect(), a: Array<(usize, usize), f32> of size (Y, X) */
int ref_c1_transpose(const float* a, uint64_t Y, uint64_t X, float* out) {
    vec_f32 v = { out, 0, X * Y };
    for (uint64_t x = 0; x < X; ++x)
        for (uint64_t y = 0; y < Y; ++y) {
            CHECK(y, Y); CHECK(x, X);      /* (y, x).to_usize((Y, X)) */
            uint64_t k = y * X + x;
            CHECK(k, X * Y);
            push(&v, a[k]);
        }
    return v.len == X * Y ? 0 : 2;
}

/* C3: idx.compose(src).collect(), idx: Array<usize, usize>, src: Array<usize, f32> */
int ref_c3_compose(const uint64_t* idx, uint64_t n, const float* src, uint64_t m, float* out) {
    vec_f32 v = { out, 0, n };
    for (uint64_t i = 0; i < n; ++i) {
        CHECK(i, n); CHECK(i, n);
        uint64_t k = idx[i];
        CHECK(k, m); CHECK(k, m);          /* the gather's bounds assert, src/int.rs:17 */
        push(&v, src[k]);
    }
    return v.len == n ? 0 : 2;
}

/* C4a: a.rows::<(usize,usize),usize>().map(|r| { let mut s = 0f32; r.each(|x| s += x); s }).collect() */
int ref_c4_fold(const float* a, uint64_t I, uint64_t J, uint64_t K, float* sums) {
    vec_f32 v = { sums, 0, I * J };
    for (uint64_t i = 0; i < I; ++i)
        for (uint64_t j = 0; j < J; ++j) {
            float s = 0.0f;
            for (uint64_t k = 0; k < K; ++k) {
                CHECK(i, I); CHECK(j, J); CHECK(k, K);
                uint64_t p = (i * J + j) * K + k;
                CHECK(p, I * J * K);
                s += a[p];                 /* sequential, index order */
            }
            push(&v, s);
        }
    return v.len == I * J ? 0 : 2;
}

/* C4b: (a - mean.iso::<(usize,usize,())>()).collect() */
int ref_c4_sub(const float* a, const float* mean, uint64_t I, uint64_t J, uint64_t K, float* out) {
    vec_f32 v = { out, 0, I * J * K };
    for (uint64_t i = 0; i < I; ++i)
        for (uint64_t j = 0; j < J; ++j)
            for (uint64_t k = 0; k < K; ++k) {
                CHECK(i, I); CHECK(j, J); CHECK(k, K);
                uint64_t p = (i * J + j) * K + k;
                CHECK(p, I * J * K);
                CHECK(i, I); CHECK(j, J);
                uint64_t q = i * J + j;
                CHECK(q, I * J);
                push(&v, a[p] - mean[q]);
            }
    return v.len == I * J * K ? 0 : 2;
}

/* C5: t = a.transpose(); d = t.diagonal(0.0); z = d.iso().zip(w.iso()); z.map(|(x, y)| x * y + 1.0)
 * out[q,p,q',p',r] = ((q,p) == (q',p') ? a[p,q] : 0) * w[r] + 1,   a: (P, Q), w: R */
int ref_c5_chain(const float* a, uint64_t P, uint64_t Q, const float* w, uint64_t R, float* out) {
    vec_f32 v = { out, 0, Q * P * Q * P * R };
    for (uint64_t q = 0; q < Q; ++q)
        for (uint64_t p = 0; p < P; ++p)
            for (uint64_t q2 = 0; q2 < Q; ++q2)
                for (uint64_t p2 = 0; p2 < P; ++p2)
                    for (uint64_t r = 0; r < R; ++r) {
                        float x = 0.0f;
                        if (q == q2 && p == p2) {  /* Diagonal::at evaluates its inner view only here */
                            CHECK(p, P); CHECK(q, Q);
                            uint64_t k = p * Q + q;
                            CHECK(k, P * Q);
                            x = a[k];
                        }
                        CHECK(r, R); CHECK(r, R);
                        SEPARATELY_ROUNDED m = x * w[r];
                        push(&v, m + 1.0f);
                    }
    return v.len == Q * P * Q * P * R ? 0 : 2;
}
