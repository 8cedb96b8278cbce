/*
 * mdim_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of apt1002/multidimension's `View::collect()` over the position-space
 * descriptor of include/mdim.h, with HOST pointers.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this; the product library
 * (multidimension_b200/csrc) never links or calls it and has no CPU fallback.
 *
 * Parity status: PINNED for integer/index/gather/transpose/broadcast work by the reference's own
 * doctest vectors (tests/golden/doctests.json, transcribed from src/view.rs, src/array.rs with
 * file:line), which this file and oracle/reference_model.py both reproduce.  Floating-point
 * results and reductions are "pinned by semantics only": the reference has no float test and no
 * reduce API, so the contract is the evaluation order cited below.
 *
 * Structure follows the reference deliberately:
 *   collect  = new_view(size, |buf| each(|t| buf.push(t)))          src/view.rs:146-150
 *   each     = I::each(size, |i| f(self.at(i)))                     src/view.rs:250-252
 *   I::each  = nested loops, LAST axis fastest                      src/index.rs:122-124,151-153; src/int.rs:23-25
 *   at       = recursive descent through the node structs           src/view.rs:846-1408
 *   push     = Vec::push into a with_capacity buffer + final length assert   src/array.rs:11-14,99-113
 */
#include "../include/mdim.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    const mdim_expr* e;
    int child[MDIM_MAX_NODES][MDIM_MAX_RANK]; /* child node indices, left to right */
    int n_child[MDIM_MAX_NODES];
    uint64_t coord[MDIM_MAX_RANK]; /* current iteration coordinates: out axes then reduction axes */
    uint64_t position;             /* linear output position of the element being computed */
    mdim_error_info* err;
    int failed;
} oracle_t;

static size_t dtype_size(int dt) {
    switch (dt) {
        case MDIM_U8: return 1;
        case MDIM_I32: case MDIM_U32: case MDIM_F32: return 4;
        case MDIM_I64: case MDIM_U64: case MDIM_F64: return 8;
    }
    return 0;
}

static int is_int(int dt) { return dt == MDIM_U8 || dt == MDIM_I32 || dt == MDIM_U32 || dt == MDIM_I64 || dt == MDIM_U64; }

static int arity(const mdim_node* n) {
    switch (n->kind) {
        case MDIM_NODE_LEAF: case MDIM_NODE_IOTA: case MDIM_NODE_CONST: return 0;
        case MDIM_NODE_UNARY: case MDIM_NODE_DIAG: return 1;
        case MDIM_NODE_FOLD: return n->n_comp == 2 ? 2 : 1; /* (init view, body) or body alone */
        case MDIM_NODE_BINARY: case MDIM_NODE_CONCAT: return 2;
        case MDIM_NODE_GATHER: return n->n_comp;
        case MDIM_NODE_TUPLE: return n->n_comp; /* root only: one output run per child (structure of arrays) */
    }
    return -1;
}

static void fail(oracle_t* o, int status, int node, uint64_t value, uint64_t bound, int comp) {
    if (o->failed) return;
    o->failed = 1;
    if (!o->err) return;
    mdim_error_info* x = o->err;
    memset(x, 0, sizeof *x);
    x->status = status; x->node = node; x->position = o->position;
    x->value = value; x->bound = bound; x->component = comp;
    if (status == MDIM_ERR_OOB) /* src/int.rs:17 */
        snprintf(x->message, sizeof x->message, "Index %llu is out of bounds for size %llu",
                 (unsigned long long)value, (unsigned long long)bound);
    else if (status == MDIM_ERR_ARITH)
        snprintf(x->message, sizeof x->message, "attempt to divide by zero or with overflow");
}

static mdim_scalar load(const void* base, int64_t idx, int dt) {
    mdim_scalar s; s.u64 = 0;
    switch (dt) {
        case MDIM_U8: s.u8 = ((const uint8_t*)base)[idx]; break;
        case MDIM_I32: s.i32 = ((const int32_t*)base)[idx]; break;
        case MDIM_U32: s.u32 = ((const uint32_t*)base)[idx]; break;
        case MDIM_F32: s.f32 = ((const float*)base)[idx]; break;
        default: s.u64 = ((const uint64_t*)base)[idx]; break;
    }
    return s;
}

static void store(void* base, uint64_t idx, int dt, mdim_scalar s) {
    switch (dt) {
        case MDIM_U8: ((uint8_t*)base)[idx] = s.u8; break;
        case MDIM_I32: case MDIM_U32: case MDIM_F32: ((uint32_t*)base)[idx] = s.u32; break;
        default: ((uint64_t*)base)[idx] = s.u64; break;
    }
}

/* Rust `as` (saturating float→int, NaN→0). */
static int64_t f_to_i64(double f, int64_t lo, int64_t hi) {
    if (f != f) return 0;
    if (f <= (double)lo) return lo;
    if (f >= (double)hi) return hi; /* hi itself may not be representable; >= catches 2^63 */
    return (int64_t)f;
}
static uint64_t f_to_u64(double f, uint64_t hi) {
    if (f != f || f <= 0.0) return 0;
    if (f >= (double)hi) return hi;
    return (uint64_t)f;
}

static mdim_scalar cast(mdim_scalar a, int from, int to) {
    mdim_scalar r; r.u64 = 0;
    /* widen to a canonical triple */
    int fromf = (from == MDIM_F32 || from == MDIM_F64);
    double d = 0; int64_t si = 0; uint64_t ui = 0; int is_signed = 0;
    switch (from) {
        case MDIM_U8: ui = a.u8; break;
        case MDIM_I32: si = a.i32; is_signed = 1; break;
        case MDIM_U32: ui = a.u32; break;
        case MDIM_I64: si = a.i64; is_signed = 1; break;
        case MDIM_U64: ui = a.u64; break;
        case MDIM_F32: d = a.f32; break;
        case MDIM_F64: d = a.f64; break;
    }
    if (fromf) {
        switch (to) {
            case MDIM_U8: r.u8 = (uint8_t)f_to_u64(d, 255); break;
            case MDIM_I32: r.i32 = (int32_t)f_to_i64(d, INT32_MIN, INT32_MAX); break;
            case MDIM_U32: r.u32 = (uint32_t)f_to_u64(d, UINT32_MAX); break;
            case MDIM_I64: r.i64 = f_to_i64(d, INT64_MIN, INT64_MAX); break;
            case MDIM_U64: r.u64 = f_to_u64(d, UINT64_MAX); break;
            case MDIM_F32: r.f32 = (from == MDIM_F32) ? a.f32 : (float)a.f64; break;
            case MDIM_F64: r.f64 = d; break;
        }
        return r;
    }
    uint64_t bits = is_signed ? (uint64_t)si : ui; /* two's complement truncation for int→int */
    switch (to) {
        case MDIM_U8: r.u8 = (uint8_t)bits; break;
        case MDIM_I32: r.i32 = (int32_t)(uint32_t)bits; break;
        case MDIM_U32: r.u32 = (uint32_t)bits; break;
        case MDIM_I64: r.i64 = (int64_t)bits; break;
        case MDIM_U64: r.u64 = bits; break;
        case MDIM_F32: r.f32 = is_signed ? (float)si : (float)ui; break;
        case MDIM_F64: r.f64 = is_signed ? (double)si : (double)ui; break;
    }
    return r;
}

/* B::call for the ops.rs vocabulary (src/ops.rs:31-129), Rust release-build integer semantics:
 * add/sub/mul wrap, shifts mask the amount, x/0 and MIN/-1 panic (→ MDIM_ERR_ARITH). */
static mdim_scalar binary(oracle_t* o, int node, int op, int dt, mdim_scalar a, mdim_scalar b, int bdt) {
    mdim_scalar r; r.u64 = 0;
    if (dt == MDIM_F32) {
        volatile float x = a.f32, y = b.f32; /* volatile: each op separately rounded, never fused */
        switch (op) {
            case MDIM_ADD: r.f32 = x + y; break;
            case MDIM_SUB: r.f32 = x - y; break;
            case MDIM_MUL: r.f32 = x * y; break;
            case MDIM_DIV: r.f32 = x / y; break;
            case MDIM_REM: r.f32 = fmodf(x, y); break;
            default: o->failed = 1; break;
        }
        return r;
    }
    if (dt == MDIM_F64) {
        volatile double x = a.f64, y = b.f64;
        switch (op) {
            case MDIM_ADD: r.f64 = x + y; break;
            case MDIM_SUB: r.f64 = x - y; break;
            case MDIM_MUL: r.f64 = x * y; break;
            case MDIM_DIV: r.f64 = x / y; break;
            case MDIM_REM: r.f64 = fmod(x, y); break;
            default: o->failed = 1; break;
        }
        return r;
    }
    /* integers: compute in 64 bits then truncate to the width */
    int bits = (int)dtype_size(dt) * 8;
    int sgn = (dt == MDIM_I32 || dt == MDIM_I64);
    uint64_t ua, ub; int64_t sa, sb;
    switch (dt) {
        case MDIM_U8: ua = a.u8; ub = b.u8; break;
        case MDIM_U32: ua = a.u32; ub = b.u32; break;
        case MDIM_I32: ua = (uint64_t)(int64_t)a.i32; ub = (uint64_t)(int64_t)b.i32; break;
        default: ua = a.u64; ub = b.u64; break;
    }
    sa = (int64_t)ua; sb = (int64_t)ub;
    uint64_t res = 0;
    if (op == MDIM_SHL || op == MDIM_SHR) {
        /* shift amount comes from the right operand's own dtype, masked to the left width */
        uint64_t amt;
        switch (bdt) {
            case MDIM_U8: amt = b.u8; break;
            case MDIM_I32: case MDIM_U32: amt = b.u32; break;
            default: amt = b.u64; break;
        }
        amt &= (uint64_t)(bits - 1);
        if (op == MDIM_SHL) res = ua << amt;
        else if (sgn) res = (uint64_t)(sa >> amt);
        else {
            uint64_t m = bits == 64 ? ~0ull : ((1ull << bits) - 1);
            res = (ua & m) >> amt;
        }
    } else switch (op) {
        case MDIM_ADD: res = ua + ub; break;
        case MDIM_SUB: res = ua - ub; break;
        case MDIM_MUL: res = ua * ub; break;
        case MDIM_AND: res = ua & ub; break;
        case MDIM_OR: res = ua | ub; break;
        case MDIM_XOR: res = ua ^ ub; break;
        case MDIM_DIV: case MDIM_REM: {
            uint64_t m = bits == 64 ? ~0ull : ((1ull << bits) - 1);
            if ((ub & m) == 0) { fail(o, MDIM_ERR_ARITH, node, ub & m, 0, 0); return r; }
            if (sgn) {
                int64_t mn = bits == 64 ? INT64_MIN : (int64_t)INT32_MIN;
                if (sa == mn && sb == -1) { fail(o, MDIM_ERR_ARITH, node, ub & m, 0, 0); return r; }
                res = (uint64_t)(op == MDIM_DIV ? sa / sb : sa % sb);
            } else {
                res = op == MDIM_DIV ? (ua & m) / (ub & m) : (ua & m) % (ub & m);
            }
            break;
        }
        default: o->failed = 1; break;
    }
    switch (dt) {
        case MDIM_U8: r.u8 = (uint8_t)res; break;
        case MDIM_I32: case MDIM_U32: r.u32 = (uint32_t)res; break;
        default: r.u64 = res; break;
    }
    return r;
}

static mdim_scalar unary(oracle_t* o, int op, int dt, int src_dt, mdim_scalar a) {
    mdim_scalar r; r.u64 = 0;
    if (op == MDIM_CAST) return cast(a, src_dt, dt);
    switch (dt) {
        case MDIM_F32:
            switch (op) {
                case MDIM_NEG: r.f32 = -a.f32; break;
                case MDIM_ABS: r.f32 = fabsf(a.f32); break;
                case MDIM_SQRT: r.f32 = sqrtf(a.f32); break;
                default: o->failed = 1;
            }
            break;
        case MDIM_F64:
            switch (op) {
                case MDIM_NEG: r.f64 = -a.f64; break;
                case MDIM_ABS: r.f64 = fabs(a.f64); break;
                case MDIM_SQRT: r.f64 = sqrt(a.f64); break;
                default: o->failed = 1;
            }
            break;
        case MDIM_U8:
            switch (op) {
                case MDIM_NEG: r.u8 = (uint8_t)(0u - a.u8); break;
                case MDIM_NOT: r.u8 = (uint8_t)~a.u8; break;
                case MDIM_ABS: r.u8 = a.u8; break;
                default: o->failed = 1;
            }
            break;
        case MDIM_I32: case MDIM_U32:
            switch (op) {
                case MDIM_NEG: r.u32 = 0u - a.u32; break;
                case MDIM_NOT: r.u32 = ~a.u32; break;
                case MDIM_ABS: r.u32 = (dt == MDIM_I32 && a.i32 < 0) ? 0u - a.u32 : a.u32; break;
                default: o->failed = 1;
            }
            break;
        default:
            switch (op) {
                case MDIM_NEG: r.u64 = 0ull - a.u64; break;
                case MDIM_NOT: r.u64 = ~a.u64; break;
                case MDIM_ABS: r.u64 = (dt == MDIM_I64 && a.i64 < 0) ? 0ull - a.u64 : a.u64; break;
                default: o->failed = 1;
            }
            break;
    }
    return r;
}

static int64_t linear(const oracle_t* o, const mdim_node* n) {
    int64_t idx = n->offset;
    int total = o->e->rank + o->e->red_rank;
    for (int a = 0; a < total; ++a) idx += (int64_t)o->coord[a] * n->stride[a];
    return idx;
}

/* node.at(i): one recursive call per node per element, like the reference's nested `at`s. */
static mdim_scalar at(oracle_t* o, int ni) {
    const mdim_node* n = &o->e->nodes[ni];
    mdim_scalar r; r.u64 = 0;
    if (o->failed) return r;
    switch (n->kind) {
        case MDIM_NODE_LEAF: /* src/array.rs:81,86: items[index.to_usize(size)].clone() */
            if (n->n_peers > 1) { /* the Array's items are split into equal blocks, block p at peer[p] */
                uint64_t idx = (uint64_t)linear(o, n), p = idx / n->peer_block;
                return load(n->peer[p], (int64_t)(idx % n->peer_block), n->dtype);
            }
            return load(n->data, linear(o, n), n->dtype);
        case MDIM_NODE_IOTA: /* src/index.rs:185: at(index) = index */
            r.u64 = (uint64_t)linear(o, n);
            return cast(r, MDIM_U64, n->dtype);
        case MDIM_NODE_CONST: /* src/view.rs:1407 */
            return n->imm;
        case MDIM_NODE_UNARY: /* src/view.rs:888: f(v.at(i)) */
            return unary(o, n->op, n->dtype, n->src_dtype, at(o, o->child[ni][0]));
        case MDIM_NODE_BINARY: { /* src/view.rs:1194-1197: B::call(v.at(vi), w.at(wi)), left first */
            mdim_scalar a = at(o, o->child[ni][0]);
            mdim_scalar b = at(o, o->child[ni][1]);
            if (o->failed) return r;
            return binary(o, ni, n->op, n->dtype, a, b, o->e->nodes[o->child[ni][1]].dtype);
        }
        case MDIM_NODE_DIAG: { /* src/view.rs:854-856: inner evaluated ONLY on the diagonal */
            for (int p = 0; p < n->n_comp; ++p) {
                uint64_t lhs = o->coord[n->axis_a[p]];
                uint64_t rhs = n->axis_b[p] >= 0 ? o->coord[n->axis_b[p]] + n->axis_c[p] /* wrapping: signed offset */ : n->axis_c[p];
                if (lhs != rhs) return n->imm;
            }
            return at(o, o->child[ni][0]);
        }
        case MDIM_NODE_CONCAT: /* src/view.rs:938-945: only the selected side is evaluated */
            if (o->coord[n->axis_a[0]] < n->axis_c[0]) return at(o, o->child[ni][0]);
            return at(o, o->child[ni][1]);
        case MDIM_NODE_GATHER: { /* src/view.rs:905,911: w.at(v.at(i)); bounds per component src/int.rs:16-19 */
            int64_t idx = linear(o, n);
            for (int c = 0; c < n->n_comp; ++c) {
                mdim_scalar k = at(o, o->child[ni][c]);
                if (o->failed) return r;
                if (!(k.u64 < n->bound[c])) { fail(o, MDIM_ERR_OOB, ni, k.u64, n->bound[c], c); return r; }
                idx += (int64_t)k.u64 * n->gstride[c];
            }
            if (n->n_peers > 1) {
                uint64_t p = (uint64_t)idx / n->peer_block;
                return load(n->peer[p], (int64_t)((uint64_t)idx % n->peer_block), n->dtype);
            }
            return load(n->data, idx, n->dtype);
        }
        case MDIM_NODE_FOLD: { /* rows().map(|row| { let mut s = init; row.each(|x| s = s ⊕ x); s })
                                   src/view.rs:617-622,1341,250-252: sequential, in index order */
            int rank = o->e->rank, rr = o->e->red_rank;
            int body = o->child[ni][n->n_comp == 2 ? 1 : 0];
            mdim_scalar acc = n->imm;
            if (n->n_comp == 2) { /* `let mut s = init.at(i)`: the initial value is itself a view over the output index */
                acc = at(o, o->child[ni][0]);
                if (o->failed) return r;
            }
            uint64_t count = 1;
            for (int a = 0; a < rr; ++a) count *= o->e->length[rank + a];
            for (int a = 0; a < rr; ++a) o->coord[rank + a] = 0;
            int cdt = o->e->nodes[body].dtype;
            for (uint64_t k = 0; k < count; ++k) {
                mdim_scalar x = at(o, body);
                if (o->failed) return r;
                acc = binary(o, ni, n->op, n->dtype, acc, x, cdt);
                if (o->failed) return r;
                for (int a = rr - 1; a >= 0; --a) { /* last reduction axis fastest */
                    if (++o->coord[rank + a] < o->e->length[rank + a]) break;
                    o->coord[rank + a] = 0;
                }
            }
            return acc;
        }
    }
    o->failed = 1;
    return r;
}

/* a `Vec::with_capacity` + `push` sink, src/array.rs:99-113 */
typedef struct { void* items; uint64_t len, cap; int dtype; } sink_t;

static void push(sink_t* s, mdim_scalar v) {
    if (s->len == s->cap) abort(); /* the reference would reallocate; capacity is exact here */
    store(s->items, s->len++, s->dtype, v);
}

/* I::each, nested loops, last axis fastest (src/index.rs:122-124) */
static void each(oracle_t* o, int axis, sink_t* s) {
    if (o->failed) return;
    if (axis == o->e->rank) {
        int root = o->e->n_nodes - 1;
        if (o->e->nodes[root].kind == MDIM_NODE_TUPLE) { /* a tuple-typed element: every scalar leaf goes to its own run */
            for (int c = 0; c < o->n_child[root] && !o->failed; ++c) {
                mdim_scalar v = at(o, o->child[root][c]);
                if (!o->failed) push(&s[c], v);
            }
            if (!o->failed) o->position++;
            return;
        }
        mdim_scalar v = at(o, root);
        if (!o->failed) { push(s, v); o->position++; }
        return;
    }
    for (uint64_t i = 0; i < o->e->length[axis] && !o->failed; ++i) {
        o->coord[axis] = i;
        each(o, axis + 1, s);
    }
}

static int validate(oracle_t* o) {
    const mdim_expr* e = o->e;
    if (e->abi_version != MDIM_ABI_VERSION) return MDIM_ERR_INVALID;
    if (e->rank < 0 || e->red_rank < 0 || e->rank + e->red_rank > MDIM_MAX_RANK) return MDIM_ERR_INVALID;
    if (e->n_nodes < 1 || e->n_nodes > MDIM_MAX_NODES || !e->nodes) return MDIM_ERR_INVALID;
    int stack[MDIM_MAX_NODES]; int sp = 0;
    for (int i = 0; i < e->n_nodes; ++i) {
        const mdim_node* n = &e->nodes[i];
        int k = arity(n);
        if (k < 0 || k > sp || k > MDIM_MAX_RANK) return MDIM_ERR_INVALID;
        if (n->dtype < 0 || n->dtype >= MDIM_DTYPE_COUNT) return MDIM_ERR_INVALID;
        o->n_child[i] = k;
        for (int c = 0; c < k; ++c) o->child[i][c] = stack[sp - k + c];
        sp -= k;
        stack[sp++] = i;
        if ((n->kind == MDIM_NODE_LEAF || n->kind == MDIM_NODE_GATHER) && n->n_peers <= 1 && !n->data)
            return MDIM_ERR_INVALID;
        if ((n->kind == MDIM_NODE_LEAF || n->kind == MDIM_NODE_GATHER) && n->n_peers > 1) {
            if (n->n_peers > MDIM_MAX_PEERS || n->peer_block == 0) return MDIM_ERR_INVALID;
            for (int q = 0; q < n->n_peers; ++q) if (!n->peer[q]) return MDIM_ERR_INVALID;
        }
        if (n->kind == MDIM_NODE_BINARY) {
            int l = e->nodes[o->child[i][0]].dtype, r = e->nodes[o->child[i][1]].dtype;
            if (l != n->dtype) return MDIM_ERR_INVALID;
            if (n->op == MDIM_SHL || n->op == MDIM_SHR) { if (!is_int(l) || !is_int(r)) return MDIM_ERR_INVALID; }
            else if (r != n->dtype) return MDIM_ERR_INVALID;
            if (!is_int(n->dtype) && n->op >= MDIM_AND) return MDIM_ERR_INVALID;
        }
        if (n->kind == MDIM_NODE_GATHER)
            for (int c = 0; c < k; ++c) if (e->nodes[o->child[i][c]].dtype != MDIM_U64) return MDIM_ERR_INVALID;
    }
    return sp == 1 ? MDIM_OK : MDIM_ERR_INVALID;
}

/* View::collect of a tuple-typed view (root = MDIM_NODE_TUPLE): one dense row-major run per scalar leaf (include/mdim.h). */
int mdim_oracle_collect_tuple(const mdim_expr* e, void* const* outs, int n_outs, mdim_error_info* err) {
    oracle_t* o = (oracle_t*)calloc(1, sizeof *o);
    if (!o) return MDIM_ERR_NOMEM;
    o->e = e; o->err = err;
    if (err) memset(err, 0, sizeof *err);
    int st = validate(o);
    const mdim_node* root = &e->nodes[e->n_nodes - 1];
    if (st == MDIM_OK && (root->kind != MDIM_NODE_TUPLE || root->n_comp != n_outs || n_outs > MDIM_MAX_OUTS)) st = MDIM_ERR_INVALID;
    if (st != MDIM_OK) { free(o); if (err) err->status = st; return st; }
    uint64_t len = 1;
    for (int a = 0; a < e->rank; ++a) len *= e->length[a];
    sink_t s[MDIM_MAX_OUTS];
    for (int c = 0; c < n_outs; ++c) { s[c].items = outs[c]; s[c].len = 0; s[c].cap = len; s[c].dtype = e->nodes[o->child[e->n_nodes - 1][c]].dtype; }
    each(o, 0, s);
    if (o->failed) { st = (err && err->status) ? err->status : MDIM_ERR_INVALID; free(o); return st; }
    free(o);
    return MDIM_OK;
}

/* View::collect (src/view.rs:146-150) into a dense row-major host buffer. */
int mdim_oracle_collect(const mdim_expr* e, void* out, mdim_error_info* err) {
    oracle_t* o = (oracle_t*)calloc(1, sizeof *o);
    if (!o) return MDIM_ERR_NOMEM;
    o->e = e; o->err = err;
    if (err) memset(err, 0, sizeof *err);
    int st = validate(o);
    if (st == MDIM_OK && e->nodes[e->n_nodes - 1].kind == MDIM_NODE_TUPLE) st = MDIM_ERR_INVALID; /* needs mdim_oracle_collect_tuple */
    if (st != MDIM_OK) { free(o); if (err) err->status = st; return st; }
    uint64_t len = 1;
    for (int a = 0; a < e->rank; ++a) len *= e->length[a]; /* I::length, src/index.rs:104-107 */
    sink_t s = { out, 0, len, e->nodes[e->n_nodes - 1].dtype };
    each(o, 0, &s);
    if (o->failed) {
        st = (err && err->status) ? err->status : MDIM_ERR_INVALID;
        free(o);
        return st;
    }
    if (s.len != len) { free(o); return MDIM_ERR_SIZE; } /* src/array.rs:12 */
    free(o);
    return MDIM_OK;
}

size_t mdim_oracle_dtype_size(int dt) { return dtype_size(dt); }
size_t mdim_oracle_sizeof_node(void) { return sizeof(mdim_node); }
size_t mdim_oracle_sizeof_expr(void) { return sizeof(mdim_expr); }
