"""Differential parity test that does NOT share the product's lowering with the checker.

A seeded generator builds the SAME random View chain twice:

  * with `oracle.reference_model` — the literal, index-tuple-level restatement of the reference's
    `Index` / `View` / `Array` (nested-tuple indices, `Isomorphic::to_iso`, `Broadcast::index`,
    per-component bounds asserts; /root/reference/src/view.rs:846-1408, src/index.rs:42-276), and
  * with `multidimension_b200` — the product's host mirror, whose `_lower()` turns the chain into the
    position-space descriptor that the kernels run,

and compares the product's collect() (`[emu]`: planner + evaluator compiled for the host; `[gpu]`: the
sm_100a kernels through the C ABI) bit for bit with the model's collect().  The model never sees a
stride, an offset or an `mdim_node`: a wrong stride / offset / predicate rule in view.py or
lowering.py changes the product's output only, and the test fails
(`test_harness_detects_a_broken_stride_rule` proves that on purpose).

Chains cover: transpose with compound / empty I, X, Y, J; iso; row / column; diagonal; zip and every
arithmetic operator with `()`-broadcasting; compose with 1-3 index components (built with zip and
broadcasting); map_axis; concat; from_usize / to_usize; insert_one / remove_one; enumerate; traced
map closures; sequential folds over trailing axes; shard_view; bool / Fixed / Reversed / Option leaves.
"""
import os
import random

import numpy as np
import pytest

import multidimension_b200 as P
from multidimension_b200 import index as PX
from multidimension_b200 import _ffi as F
from multidimension_b200 import sharding
from oracle import reference_model as M

from helpers import emu_collect, emu_collect_tuple, assert_same_bits, CheckerPanic
from multidimension_b200 import lowering as L

MAX_ELEMS = 6000      # per collected view (the model is pure Python)
MAX_AXES = 8          # MDIM_MAX_RANK
N_BATCHES, PER_BATCH = 20, 30  # 600 chains per backend

_gpu_ctx = []


@pytest.fixture(params=[pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "gpu" and not _gpu_ctx:
        _gpu_ctx.append(P.Context(0))
        P.set_default_context(_gpu_ctx[0])
    return request.param


def product_collect(view, backend):
    if backend == "emu":
        if isinstance(view.T, tuple) and len(L.flatten_value(view._lower()[1])) <= F.MAX_OUTS:
            return emu_collect_tuple(view)  # tuple-typed elements: ONE descriptor with an MDIM_NODE_TUPLE root, as on the GPU
        return emu_collect(view)
    return view.collect(location="device", ctx=_gpu_ctx[0]).as_ref()


# ---- type / value translation between the two mirrors ----------------------------------------------------
def mt(T):
    """product index type -> model index type"""
    if T is P.usize:
        return M.usize
    if T is P.Reversed:
        return M.Reversed
    if T is bool:
        return bool
    if isinstance(T, P.Fixed):
        return M.Fixed(T.n)
    if isinstance(T, P.Option):
        return M.Option(mt(T.inner))
    if isinstance(T, tuple):
        return tuple(mt(t) for t in T)
    raise TypeError(T)


def mv(T, v):
    """product index VALUE of type T -> model index value"""
    if isinstance(T, tuple):
        return tuple(mv(t, x) for t, x in zip(T, v))
    if T is P.Reversed:
        return M.Rev(v)
    if isinstance(T, P.Option):
        return None if v is None else M.Some(mv(T.inner, v.i))
    return v


def leaf_list(I, size):
    return list(zip(PX.type_leaves(I), PX.size_leaves(I, size)))


def leaf_len(t, s):
    (n,) = PX.leaf_lengths(t, s)
    return n


def n_axes(I, size):
    return len(leaf_list(I, size))


def random_tree(rng, items, allow_unit=True):
    """A random index type (and its size) whose flattening is exactly `items` = [(leaf type, leaf size)]:
    tuples of arity 1-3, nested at random, with an occasional `()` component."""
    n = len(items)
    if n == 0:
        return (), ()
    if n == 1 and rng.random() < 0.7:
        return items[0]
    arity = rng.choice([1, 2, 2, 3]) if n > 1 else (rng.choice([1, 2]) if allow_unit else 1)
    if not allow_unit:
        arity = min(arity, n)
    if allow_unit:
        cuts = sorted(rng.randint(0, n) if rng.random() < 0.15 else rng.randint(1, max(1, n - 1)) for _ in range(arity - 1))
    else:
        cuts = sorted(rng.sample(range(1, n), arity - 1))  # distinct interior cuts: no empty part
    parts, lo = [], 0
    for c in cuts + [n]:
        parts.append(items[lo:c])
        lo = c
    sub = [random_tree(rng, p, allow_unit) for p in parts]
    return tuple(t for t, _ in sub), tuple(s for _, s in sub)


def random_index(rng, T, size):
    """A random in-range index value of product type T (None if some axis is empty)."""
    if isinstance(T, tuple):
        out = []
        for t, s in zip(T, size):
            v = random_index(rng, t, s)
            if v is _EMPTY:
                return _EMPTY
            out.append(v)
        return tuple(out)
    if T is P.usize or T is P.Reversed:
        return rng.randrange(size) if size > 0 else _EMPTY
    if T is bool:
        return rng.random() < 0.5
    if isinstance(T, P.Fixed):
        return rng.randrange(T.n) if T.n > 0 else _EMPTY
    if isinstance(T, P.Option):
        if rng.random() < 0.3:
            return None
        v = random_index(rng, T.inner, size)
        return None if v is _EMPTY else P.Some(v)
    raise TypeError(T)


_EMPTY = object()


# ---- a view built in both mirrors ----------------------------------------------------------------------------
class Both:
    """The same view in the model (`m`) and in the product (`p`), with the product-side index type / size
    (checked against the model's after every step) and the element kind ('i64', 'f32', 'usize' or a tuple)."""

    def __init__(self, m, p, T, note):
        self.m, self.p, self.T, self.note = m, p, T, note
        self.I, self.size = p.I, p.size()
        assert mt(self.I) == m.I or M.isomorphic(mt(self.I), m.I), f"{note}: index types diverge: {self.I!r} vs {m.I!r}"
        assert _plain(m.size()) == _plain(self.size), f"{note}: sizes diverge: product {self.size!r} vs model {m.size()!r}"

    def length(self):
        return PX.length(self.I, self.size)


def _plain(x):
    if isinstance(x, tuple):
        return tuple(_plain(y) for y in x)
    return x


def np_dtype(T):
    return {"i64": np.int64, "f32": np.float32, "u8": np.uint8}.get(T, np.uint64)


def fresh_values(rng, T, n, bound=None):
    if T == "f32":
        return np.array([rng.uniform(-4, 4) for _ in range(n)], dtype=np.float32)
    if T == "i64":
        return np.array([rng.randint(-9, 9) for _ in range(n)], dtype=np.int64)
    if T is bool:
        return np.array([rng.random() < 0.5 for _ in range(n)], dtype=np.bool_)
    return np.array([rng.randrange(bound) for _ in range(n)], dtype=np.uint64)  # usize index values


def model_items(T, values):
    if T == "f32":
        return [np.float32(v) for v in values]
    if T is bool:
        return [bool(v) for v in values]
    return [int(v) for v in values]


def fresh_array(rng, I, size, T, bound=None, note="array"):
    n = PX.length(I, size)
    vals = fresh_values(rng, T, n, bound)
    pa = P.Array.new(I, size, vals, T if T is not P.usize else P.usize)
    ma = M.Array.new(mt(I), size, model_items(T, vals))
    return Both(ma, pa, T, f"{note}<{I!r},{size!r}>")


LEAF_KINDS = ["usize"] * 10 + ["bool", "fixed", "fixed", "reversed", "option"]


def random_leaves(rng, max_axes, max_elems):
    k = rng.randint(1, max_axes)
    items, total = [], 1
    for _ in range(k):
        kind = rng.choice(LEAF_KINDS)
        if kind == "usize":
            s = rng.choice([1, 2, 3, 3, 4, 5, 6, 7, 8, 16, 17]) if rng.random() < 0.97 else 0
            t = (P.usize, s)
        elif kind == "bool":
            t = (bool, ())
        elif kind == "fixed":
            t = (P.Fixed(rng.randint(1, 4)), ())
        elif kind == "reversed":
            t = (P.Reversed, rng.randint(1, 5))
        else:
            t = (P.Option(P.usize), rng.randint(1, 4))
        n = leaf_len(*t)
        if total * max(n, 1) > max_elems:
            break
        total *= max(n, 1)
        items.append(t)
    return items or [(P.usize, 3)]


def start_view(rng, T=None):
    T = T or rng.choice(["i64", "i64", "f32"])
    items = random_leaves(rng, 4, 600)
    I, size = random_tree(rng, items)
    return fresh_array(rng, I, size, T)


# ---- the operations ----------------------------------------------------------------------------------------------
class Skip(Exception):
    """The operation does not apply to this view (not an error)."""


def op_transpose(rng, v):
    L = leaf_list(v.I, v.size)
    n = len(L)
    p = sorted(rng.randint(0, n) for _ in range(3))
    (I, _), (Y, _), (X_, _), (J, _) = (random_tree(rng, L[:p[0]]), random_tree(rng, L[p[0]:p[1]]), random_tree(rng, L[p[1]:p[2]]), random_tree(rng, L[p[2]:]))
    return Both(v.m.transpose(mt(I), mt(X_), mt(Y), mt(J)), v.p.transpose(I, X_, Y, J), v.T, f"{v.note}.transpose<{I!r},{X_!r},{Y!r},{J!r}>")


def op_iso(rng, v):
    J, _ = random_tree(rng, leaf_list(v.I, v.size))
    return Both(v.m.iso(mt(J)), v.p.iso(J), v.T, f"{v.note}.iso<{J!r}>")


def _split2(rng, v, lo=0, hi=None):
    L = leaf_list(v.I, v.size)
    p = rng.randint(lo, len(L) if hi is None else hi)
    (I, si), (J, sj) = random_tree(rng, L[:p]), random_tree(rng, L[p:])
    return I, si, J, sj


def op_row(rng, v):
    I, si, J, _ = _split2(rng, v)
    i = random_index(rng, I, si)
    if i is _EMPTY:
        raise Skip
    return Both(v.m.row(mt(I), mt(J), mv(I, i)), v.p.row(I, J, i), v.T, f"{v.note}.row<{I!r},{J!r}>({i!r})")


def op_column(rng, v):
    I, _, J, sj = _split2(rng, v)
    j = random_index(rng, J, sj)
    if j is _EMPTY:
        raise Skip
    return Both(v.m.column(mt(I), mt(J), mv(J, j)), v.p.column(I, J, j), v.T, f"{v.note}.column<{I!r},{J!r}>({j!r})")


def op_diagonal(rng, v):
    if isinstance(v.T, tuple) or 2 * n_axes(v.I, v.size) > MAX_AXES or v.length() ** 2 > MAX_ELEMS:
        raise Skip
    zero = np.float32(-7.5) if v.T == "f32" else (True if v.T is bool else 77)
    mzero = zero if v.T == "f32" else (bool(zero) if v.T is bool else int(zero))
    return Both(v.m.diagonal(mzero), v.p.diagonal(zero), v.T, f"{v.note}.diagonal")


def _broadcast_partner_type(rng, I, size, budget):
    """An index type / size that broadcasts against (I, size): leaves kept or replaced by `()`, `()` components
    of I answered by a fresh axis (`()` against `()` does not implement Broadcast, src/broadcast.rs:4-9)."""
    if I == ():
        s = rng.randint(1, 3)
        if budget[0] * s > MAX_ELEMS or budget[1] >= MAX_AXES:
            raise Skip
        budget[0] *= s
        budget[1] += 1
        return P.usize, s
    if isinstance(I, tuple):
        if rng.random() < 0.2:
            return (), ()
        sub = [_broadcast_partner_type(rng, t, s, budget) for t, s in zip(I, size)]
        return tuple(t for t, _ in sub), tuple(s for _, s in sub)
    if rng.random() < 0.3:
        return (), ()
    return I, size


def _is_unit(I):
    return I == ()


BIN_OPS = [("Add", lambda a, b: a + b), ("Sub", lambda a, b: a - b), ("Mul", lambda a, b: a * b)]


def op_binary(rng, v):
    if isinstance(v.T, tuple) or v.T is bool or v.T is P.usize:
        raise Skip
    budget = [max(v.length(), 1), n_axes(v.I, v.size)]
    J, sj = _broadcast_partner_type(rng, v.I, v.size, budget)
    if _is_unit(J) and _is_unit(v.I):
        raise Skip
    name, _ = rng.choice(BIN_OPS)
    if _is_unit(J) and rng.random() < 0.5:  # Scalar operand (src/view.rs:1399-1408)
        c = np.float32(rng.uniform(-2, 2)) if v.T == "f32" else rng.randint(-3, 3)
        w = Both(M.Scalar(c if v.T == "f32" else int(c)), P.Scalar(c, v.T), v.T, "Scalar")
    else:
        w = fresh_array(rng, J, sj, v.T, note="rhs")
    if rng.random() < 0.3:
        v, w = w, v
    return Both(v.m.binary(w.m, getattr(M, name)), v.p.binary(w.p, getattr(P, name)), v.T, f"({v.note} {name} {w.note})")


def op_zip_map(rng, v):
    if isinstance(v.T, tuple) or v.T is bool or v.T is P.usize:
        raise Skip
    budget = [max(v.length(), 1), n_axes(v.I, v.size)]
    J, sj = _broadcast_partner_type(rng, v.I, v.size, budget)
    if _is_unit(J) and _is_unit(v.I):
        raise Skip
    w = fresh_array(rng, J, sj, v.T, note="rhs")
    one = np.float32(1) if v.T == "f32" else 1
    f = lambda p: p[0] * p[1] + one  # noqa: E731  (traced by the product, called per element by the model)
    return Both(v.m.zip(w.m).map(f), v.p.zip(w.p).map(f), v.T, f"zip({v.note}, {w.note}).map(x*y+1)")


def op_zip_pair(rng, v):
    if isinstance(v.T, tuple):
        raise Skip
    budget = [max(v.length(), 1), n_axes(v.I, v.size)]
    J, sj = _broadcast_partner_type(rng, v.I, v.size, budget)
    if _is_unit(J) and _is_unit(v.I):
        raise Skip
    w = fresh_array(rng, J, sj, rng.choice(["i64", "f32"]), note="rhs")
    return Both(v.m.zip(w.m), v.p.zip(w.p), (v.T, w.T), f"zip({v.note}, {w.note})")


def op_map(rng, v):
    if isinstance(v.T, tuple) or v.T is bool or v.T is P.usize:
        raise Skip
    c = np.float32(0.5) if v.T == "f32" else 3
    f = lambda x: (x * x - c) * x  # noqa: E731
    return Both(v.m.map(f), v.p.map(f), v.T, f"{v.note}.map((x*x-c)*x)")


def _pairs_only(items):
    """Right-nested pairs over the leaves: the only tuple shape `zip` can build (ops::Pair is binary)."""
    if len(items) == 1:
        return items[0]
    t, s = _pairs_only(items[1:])
    return (items[0][0], t), (items[0][1], s)


def _index_view(rng, T, size, I_out, s_out):
    """A view indexed by something that broadcasts to (I_out, s_out) whose ELEMENTS are in-range indices of
    type T (a tree of pairs over usize / bool leaves): zip of fresh index Arrays."""
    if isinstance(T, tuple):
        a = _index_view(rng, T[0], size[0], I_out, s_out)
        b = _index_view(rng, T[1], size[1], I_out, s_out)
        return Both(a.m.zip(b.m), a.p.zip(b.p), (a.T, b.T), f"zip({a.note},{b.note})")
    return fresh_array(rng, I_out, s_out, T, bound=size if T is P.usize else None, note="idx")


def op_compose(rng, v):
    """idx.compose(v): v is the SOURCE; idx is a fresh index view (src/view.rs:897-912)."""
    L = leaf_list(v.I, v.size)
    if not L or len(L) > 3 or any(t not in (P.usize, bool) for t, _ in L) or any(leaf_len(t, s) == 0 for t, s in L):
        raise Skip
    src = v
    WI, wsize = _pairs_only(L)
    if WI != v.I:
        src = Both(v.m.iso(mt(WI)), v.p.iso(WI), v.T, f"{v.note}.iso<{WI!r}>")
    out_items = random_leaves(rng, 3, max(1, 400))
    out_items = [(t, s) for t, s in out_items if t is P.usize and s > 0] or [(P.usize, 5)]
    I_out, s_out = random_tree(rng, out_items, allow_unit=False)
    idx = _index_view(rng, WI, wsize, I_out, s_out)
    return Both(idx.m.compose(src.m), idx.p.compose(src.p), v.T, f"{idx.note}.compose({src.note})")


def op_map_axis(rng, v):
    L = leaf_list(v.I, v.size)
    cand = [k for k, (t, s) in enumerate(L) if t is P.usize and s > 0]
    if not cand:
        raise Skip
    k = rng.choice(cand)
    (I, _), (J, _) = random_tree(rng, L[:k]), random_tree(rng, L[k + 1:])
    w_items = [(P.usize, rng.randint(1, 4)) for _ in range(rng.randint(1, 2))]
    if n_axes(v.I, v.size) - 1 + len(w_items) > MAX_AXES:
        raise Skip
    WI, ws = random_tree(rng, w_items, allow_unit=False)
    if v.length() // L[k][1] * PX.length(WI, ws) > MAX_ELEMS:
        raise Skip
    w = fresh_array(rng, WI, ws, P.usize, bound=L[k][1], note="take")
    vi = Both(v.m.iso(mt((I, P.usize, J))), v.p.iso((I, P.usize, J)), v.T, v.note)
    return Both(vi.m.map_axis(w.m, mt(I), mt(J)), vi.p.map_axis(w.p, I, J), v.T, f"{v.note}.map_axis<{I!r},{J!r}>({w.note})")


def op_concat(rng, v):
    L = leaf_list(v.I, v.size)
    cand = [k for k, (t, _) in enumerate(L) if t is P.usize]
    if not cand or isinstance(v.T, tuple) or v.T is P.usize:
        raise Skip
    k = rng.choice(cand)
    (I, si), (J, sj) = random_tree(rng, L[:k]), random_tree(rng, L[k + 1:])
    extra = rng.randint(0, 4)
    if v.length() // max(L[k][1], 1) * (L[k][1] + extra) > MAX_ELEMS:
        raise Skip
    full = (I, P.usize, J)
    w = fresh_array(rng, full, (si, extra, sj), v.T, note="tail")
    vi = Both(v.m.iso(mt(full)), v.p.iso(full), v.T, v.note)
    if rng.random() < 0.3:
        vi, w = w, vi
    return Both(vi.m.concat(w.m, mt(I), mt(J)), vi.p.concat(w.p, I, J), v.T, f"{vi.note}.concat<{I!r},{J!r}>({w.note})")


def op_from_usize(rng, v):
    L = leaf_list(v.I, v.size)
    cand = [k for k, (t, s) in enumerate(L) if t is P.usize and s >= 1]
    if not cand or len(L) + 2 > MAX_AXES:
        raise Skip
    k = rng.choice(cand)
    s = L[k][1]
    splits = [(a, s // a) for a in range(1, s + 1) if s % a == 0]
    a, b = rng.choice(splits)
    choices = [((P.usize, P.usize), (a, b)), ((P.usize, (P.usize,)), (a, (b,)))]
    if s == 2:
        choices.append((bool, ()))
    if s % 2 == 0:
        choices.append(((P.usize, bool), (s // 2, ())))
    if 1 <= s <= 4:
        choices.append((P.Fixed(s), ()))
    Xt, xs = rng.choice(choices)
    (I, _), (J, _) = random_tree(rng, L[:k]), random_tree(rng, L[k + 1:])
    vi = Both(v.m.iso(mt((I, P.usize, J))), v.p.iso((I, P.usize, J)), v.T, v.note)
    return Both(vi.m.from_usize(mt(I), mt(Xt), mt(J), lambda n: xs), vi.p.from_usize(I, Xt, J, lambda n: xs), v.T, f"{v.note}.from_usize<{I!r},{Xt!r},{J!r}>")


def op_to_usize(rng, v):
    L = leaf_list(v.I, v.size)
    if not L:
        raise Skip
    lo = rng.randint(0, len(L) - 1)
    hi = rng.randint(lo + 1, len(L))
    (I, _), (Xt, _), (J, _) = random_tree(rng, L[:lo]), random_tree(rng, L[lo:hi], allow_unit=False), random_tree(rng, L[hi:])
    vi = Both(v.m.iso(mt((I, Xt, J))), v.p.iso((I, Xt, J)), v.T, v.note)
    return Both(vi.m.to_usize(mt(I), mt(Xt), mt(J)), vi.p.to_usize(I, Xt, J), v.T, f"{v.note}.to_usize<{I!r},{Xt!r},{J!r}>")


def op_insert_one(rng, v):
    if n_axes(v.I, v.size) + 1 > MAX_AXES:
        raise Skip
    I, _, K, _ = _split2(rng, v)
    J, sj = rng.choice([(P.usize, 1), (P.Fixed(1), ()), ((), ()), ((P.usize, ()), (1, ()))])
    vi = Both(v.m.iso(mt((I, K))), v.p.iso((I, K)), v.T, v.note)
    return Both(vi.m.insert_one(mt(I), mt(J), mt(K), sj), vi.p.insert_one(I, J, K, sj), v.T, f"{v.note}.insert_one<{I!r},{J!r},{K!r}>")


def op_remove_one(rng, v):
    L = leaf_list(v.I, v.size)
    cand = [k for k, (t, s) in enumerate(L) if leaf_len(t, s) == 1 and (t is P.usize or isinstance(t, P.Fixed))]
    if not cand:
        raise Skip
    k = rng.choice(cand)
    (I, _), (K, _) = random_tree(rng, L[:k]), random_tree(rng, L[k + 1:])
    J = L[k][0]
    vi = Both(v.m.iso(mt((I, J, K))), v.p.iso((I, J, K)), v.T, v.note)
    return Both(vi.m.remove_one(mt(I), mt(J), mt(K)), vi.p.remove_one(I, J, K), v.T, f"{v.note}.remove_one<{I!r},{J!r},{K!r}>")


def op_enumerate(rng, v):
    if isinstance(v.T, tuple) or any(not (t is P.usize or t is bool) for t, _ in leaf_list(v.I, v.size)):
        raise Skip
    return Both(v.m.enumerate(), v.p.enumerate(), (v.I, v.T), f"{v.note}.enumerate()")


def op_fold(rng, v):
    if isinstance(v.T, tuple) or v.T is bool or v.T is P.usize:
        raise Skip
    L = leaf_list(v.I, v.size)
    p = rng.randint(0, max(len(L) - 1, 0))
    (I, _), (J, _) = random_tree(rng, L[:p]), random_tree(rng, L[p:])
    name, fn = rng.choice(BIN_OPS[:2] + ([BIN_OPS[2]] if v.T == "f32" else []))
    def has_unit(t):
        return t == () or (isinstance(t, tuple) and any(has_unit(x) for x in t))
    if rng.random() < 0.35 and p > 0 and not has_unit(I):  # `let mut s = init.at(i)`: the initial value is a view indexed like rows()
        init = fresh_array(rng, I, PX.to_iso_size(v.size, v.I, (I, J))[0], v.T, note="init")
        return Both(M.fold_rows_from(v.m, mt(I), mt(J), fn, init.m), v.p.rows(I, J).map(P.Fold(getattr(P, name), init.p)), v.T,
                    f"fold_rows<{I!r},{J!r}>({v.note}, {name}, init={init.note})")
    init = np.float32(0.25) if v.T == "f32" else 1
    return Both(M.fold_rows(v.m, mt(I), mt(J), fn, init if v.T == "f32" else int(init)), P.fold_rows(v.p, I, J, getattr(P, name), init), v.T,
                f"fold_rows<{I!r},{J!r}>({v.note}, {name})")


STRUCTURAL = [op_transpose, op_transpose, op_iso, op_row, op_column, op_diagonal, op_compose, op_map_axis, op_concat, op_from_usize, op_to_usize,
              op_insert_one, op_remove_one]
ARITH = [op_binary, op_binary, op_zip_map, op_map]
TERMINAL = [op_zip_pair, op_enumerate, op_fold, None, None, None]


def build_chain(seed):
    rng = random.Random(seed)
    v = start_view(rng)
    declined = []
    n_ops = rng.randint(1, 4)
    applied = 0
    for _ in range(12):
        if applied >= n_ops:
            break
        op = rng.choice(STRUCTURAL * 2 + ARITH)
        try:
            w = op(rng, v)
        except Skip:
            continue
        except P.Unsupported as e:  # declined while the chain is being built (e.g. compose onto a diagonal)
            declined.append(f"{op.__name__}: {e}")
            continue
        if w.length() > MAX_ELEMS or n_axes(w.I, w.size) > MAX_AXES:
            continue
        v = w
        applied += 1
    term = rng.choice(TERMINAL)
    if term is not None:
        try:
            w = term(rng, v)
            if w.length() <= MAX_ELEMS:
                v = w
        except Skip:
            pass
        except P.Unsupported as e:
            declined.append(f"{term.__name__}: {e}")
    v.declined = declined
    shard = None
    if rng.random() < 0.25:
        world = rng.choice([2, 3, 4])
        shard = (rng.randrange(world), world)
    return v, shard


def model_result(v):
    return v.m.collect().items


def _to_python(x):
    if isinstance(x, tuple):
        return tuple(_to_python(y) for y in x)
    if isinstance(x, M.Rev):
        return ("Rev", x.i)
    if isinstance(x, M.Some):
        return ("Some", _to_python(x.i))
    if isinstance(x, (np.floating, float)):
        f = float(x)
        return ("nan",) if f != f else f
    if isinstance(x, (bool, np.bool_)):
        return bool(x)
    if isinstance(x, (int, np.integer)):
        return int(x)
    return x


def compare(v, got, want_items, what):
    if isinstance(v.T, tuple):
        got_l = [_to_python(g) for g in got]
        want_l = [_to_python(w) for w in want_items]
        assert got_l == want_l, f"{what}: tuple-typed result differs (first few: got {got_l[:4]} want {want_l[:4]})"
        return
    if v.T is bool:
        want = np.array([bool(w) for w in want_items], dtype=np.bool_)
    elif v.T == "i64":
        want = np.array([int(w) for w in want_items], dtype=np.int64).reshape(-1)
    elif v.T == "f32":
        want = np.array(want_items, dtype=np.float32).reshape(-1)
    else:
        want = np.array([int(w) for w in want_items], dtype=np.uint64).reshape(-1)
    assert_same_bits(np.asarray(got).reshape(-1), want, what)


def shard_of(items, v, rank, world):
    """Rows [lo, hi) of the outermost position axis of the model's collected result."""
    L = leaf_list(v.I, v.size)
    n0 = leaf_len(*L[0])
    inner = len(items) // n0 if n0 else 0
    lo, hi = sharding.shard_bounds(n0, world, rank)
    return items[lo * inner:hi * inner]


def run_chain(seed, backend, stats=None):
    v, shard = build_chain(seed)
    if stats is not None:
        stats.extend((seed, d, v.note) for d in v.declined)
    want = model_result(v)
    view = v.p
    what = f"seed {seed}: {v.note}"
    try:
        if shard is not None:
            L = leaf_list(v.I, v.size)
            if L and L[0][0] is P.usize and L[0][1] > 0:
                view = sharding.shard_view(v.p, *shard)
                want = shard_of(want, v, *shard)
                what += f" shard {shard}"
        got = product_collect(view, backend)
    except (P.MdimError, CheckerPanic) as e:  # declined loudly (MDIM_ERR_UNSUPPORTED: descriptor limits, device div/mod ...)
        if e.status != F.ERR_UNSUPPORTED:
            raise
        if stats is not None:
            stats.append((seed, str(e), v.note))
        return False
    compare(v, got, want, what)
    return True


@pytest.mark.parametrize("batch", range(N_BATCHES))
def test_random_chains_against_the_reference_model(backend, batch):
    unsupported = []
    ran = 0
    for k in range(PER_BATCH):
        ran += run_chain(1000 * batch + k, backend, unsupported)
    # the product may decline a chain (MDIM_ERR_UNSUPPORTED) but never silently: and it must not decline many
    assert len(unsupported) <= PER_BATCH // 3, f"too many chains declined: {unsupported}"
    assert ran >= PER_BATCH * 2 // 3


def test_generator_covers_every_operation():
    """The 240 seeds exercise every node kind at least a few times (a generator that always skipped an
    operation would make the test above vacuous for it)."""
    seen = {}
    for batch in range(N_BATCHES):
        for k in range(PER_BATCH):
            v, shard = build_chain(1000 * batch + k)
            for word in ("transpose", "iso<", "row<", "column<", "diagonal", "compose", "map_axis", "concat", "from_usize", "to_usize", "insert_one",
                         "remove_one", "enumerate", "fold_rows", "zip(", " Add ", " Sub ", " Mul ", "map(", "Scalar", "Reversed", "Option", "Fixed", "bool"):
                if word in v.note:
                    seen[word] = seen.get(word, 0) + 1
            if shard:
                seen["shard"] = seen.get("shard", 0) + 1
    missing = [w for w in ("transpose", "iso<", "row<", "column<", "diagonal", "compose", "map_axis", "concat", "from_usize", "to_usize", "insert_one",
                           "remove_one", "enumerate", "fold_rows", "zip(", " Add ", " Sub ", " Mul ", "map(", "Scalar", "shard", "Reversed", "Option", "Fixed", "bool")
               if seen.get(w, 0) < 3]
    assert not missing, f"operations the generator (almost) never produces: {missing}; counts: {seen}"


def test_harness_detects_a_broken_stride_rule(monkeypatch):
    """Mutation check: with a deliberately wrong row-major stride rule in the product's Array lowering (the
    two innermost strides swapped) the differential test must fail.  tests/helpers.oracle_collect could not
    notice this — it runs the oracle over the product's own (wrong) descriptor."""
    from multidimension_b200 import view as V
    from multidimension_b200 import lowering as L
    from multidimension_b200 import _ffi as F
    good = V.Array._lower

    def bad(self):
        groups, value = good(self)

        def twist(n):
            if n.kind == F.LEAF and len(n.stride) >= 2:
                axes = list(n.stride)
                s = dict(n.stride)
                s[axes[-1]], s[axes[-2]] = n.stride[axes[-2]], n.stride[axes[-1]]
                return n.clone(stride=s)
            return n
        return groups, L.map_value(value, twist)
    monkeypatch.setattr(V.Array, "_lower", bad)
    failures = 0
    for seed in range(40):
        try:
            run_chain(seed, "emu")
        except (AssertionError, CheckerPanic):
            failures += 1
    assert failures >= 10, f"only {failures} of 40 chains noticed the broken stride rule"
