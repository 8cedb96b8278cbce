"""reference_model.py — CPU ORACLE (test infrastructure, NOT product code).

A literal, element-at-a-time restatement in pure Python of the reference crate's
Index / View / Array semantics (apt1002/multidimension 0.3.3).  It works on nested-tuple
indices exactly as the Rust does, so the reference's doctests can be replayed against it
verbatim (tests/golden/doctests.json, tests/test_golden_doctests.py).  It is only suitable for
small cases (pure-Python loops); larger cases use oracle/mdim_oracle.c, which is cross-checked
against this file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product package (multidimension_b200/) never does.

Parity status: PINNED by the reference's own doctest vectors for collect/All/each/diagonal/map/
compose/zip/binary/transpose/row/column/rows/concat/from_usize/to_usize/insert_one/remove_one/
map_axis/enumerate/nested/coat/Array::new/from_fn/fn_view.  Floating point and reductions are
pinned by semantics only (the reference has no float test and no reduce API).

Index *types* are spelled almost as in Rust:
    usize, bool, (), (I,), (I, J), (I, J, K), Fixed(n), Reversed, Option(I), Coated(I)
Index *values*: int, bool, (), tuples, Rev(i), None / Some(i), CoatedV(i).
Sizes: int for usize/Reversed, () for bool/()/Fixed, tuples, CoatedV(size).
"""
from __future__ import annotations


# --------------------------------------------------------------------------- index types
class _USize:
    def __repr__(self):
        return "usize"

    # `usize::all(5)`  (src/index.rs:70)
    def all(self, size):
        return All(self, to_iso_size(size, self))


usize = _USize()


class _Reversed:
    def __repr__(self):
        return "Reversed"


Reversed = _Reversed()


class Fixed:
    def __init__(self, n):
        self.n = n

    def __eq__(self, o):
        return isinstance(o, Fixed) and o.n == self.n

    def __hash__(self):
        return hash(("Fixed", self.n))

    def __repr__(self):
        return f"Fixed<{self.n}>"


class Option:
    def __init__(self, inner):
        self.inner = inner

    def __eq__(self, o):
        return isinstance(o, Option) and o.inner == self.inner

    def __hash__(self):
        return hash(("Option", self.inner))


class Coated:
    def __init__(self, inner):
        self.inner = inner

    def __eq__(self, o):
        return isinstance(o, Coated) and o.inner == self.inner

    def __hash__(self):
        return hash(("Coated", self.inner))


# index *values* for the wrapper types
class Rev:
    def __init__(self, i):
        self.i = i

    def __eq__(self, o):
        return isinstance(o, Rev) and o.i == self.i

    def __repr__(self):
        return f"Reversed({self.i})"


class Some:
    def __init__(self, i):
        self.i = i

    def __eq__(self, o):
        return isinstance(o, Some) and o.i == self.i

    def __repr__(self):
        return f"Some({self.i})"


class CoatedV:
    def __init__(self, v):
        self.v = v

    def __eq__(self, o):
        return isinstance(o, CoatedV) and o.v == self.v

    def __repr__(self):
        return f"Coated({self.v!r})"


class Panic(Exception):
    """A Rust panic."""


def div_mod(n, d):  # src/lib.rs:38-39
    return n // d, n % d


def _is_tuple_type(t):
    return isinstance(t, tuple)


# ---- Index trait (src/index.rs:42-71) for every index type ----------------------------------
def length(I, size):
    if I is usize or I is Reversed:  # src/int.rs:13, :72
        return size
    if I is bool:  # StaticIndex, src/index.rs:236-240, :205
        return 2
    if isinstance(I, Fixed):  # src/int.rs:41
        return I.n
    if isinstance(I, Option):  # src/index.rs:248
        return 1 + length(I.inner, size)
    if isinstance(I, Coated):  # src/index.rs:159
        return length(I.inner, size.v)
    if _is_tuple_type(I):
        if len(I) == 0:  # (), src/index.rs:230-234
            return 1
        n = 1
        for t, s in zip(I, size):  # src/index.rs:79,104,131
            n *= length(t, s)
        return n
    raise TypeError(f"not an index type: {I!r}")


def to_usize(I, index, size):
    if I is usize:  # src/int.rs:16-19
        if not (index < size):
            raise Panic(f"Index {index} is out of bounds for size {size}")
        return index
    if I is Reversed:  # src/int.rs:75
        return (size - 1) - index.i
    if I is bool:  # src/index.rs:238 + the ALL[index] roundtrip assert :214
        return int(index)
    if isinstance(I, Fixed):  # src/int.rs:44 (unchecked)
        return index
    if isinstance(I, Option):  # src/index.rs:251-256
        return 0 if index is None else 1 + to_usize(I.inner, index.i, size)
    if isinstance(I, Coated):
        return to_usize(I.inner, index.v, size.v)
    if _is_tuple_type(I):
        acc = 0  # Horner, src/index.rs:83-87,109-114,136-142
        for t, i, s in zip(I, index, size):
            acc = acc * length(t, s) + to_usize(t, i, s)
        return acc
    raise TypeError(I)


def from_usize(I, size, index):
    """Returns (index / length, the I for index % length)."""
    if I is usize:  # src/int.rs:21
        return div_mod(index, size)
    if I is Reversed:  # src/int.rs:77-80
        q, r = div_mod(index, size)
        return q, Rev((size - 1) - r)
    if I is bool:  # src/index.rs:218-221
        q, r = div_mod(index, 2)
        return q, r != 0
    if isinstance(I, Fixed):
        return div_mod(index, I.n)
    if isinstance(I, Option):  # src/index.rs:258-266
        q, r = div_mod(index, length(I, size))
        if r == 0:
            return q, None
        zero, inner = from_usize(I.inner, size, r - 1)
        assert zero == 0
        return q, Some(inner)
    if isinstance(I, Coated):
        q, r = from_usize(I.inner, size.v, index)
        return q, CoatedV(r)
    if _is_tuple_type(I):
        parts = []  # peeled from the LAST component, src/index.rs:116-120,144-149
        for t, s in zip(reversed(I), reversed(size)):
            index, i = from_usize(t, s, index)
            parts.append(i)
        return index, tuple(reversed(parts))
    raise TypeError(I)


def each(I, size, f):
    """Index::each — nested loops, last axis fastest (src/index.rs:122-124,151-153)."""
    if I is usize:  # src/int.rs:23-25
        for i in range(size):
            f(i)
    elif I is Reversed:  # src/int.rs:82-84
        for i in range(size):
            f(Rev((size - 1) - i))
    elif I is bool:
        f(False)
        f(True)
    elif isinstance(I, Fixed):
        for i in range(I.n):
            f(i)
    elif isinstance(I, Option):  # src/index.rs:272-275
        f(None)
        each(I.inner, size, lambda i: f(Some(i)))
    elif isinstance(I, Coated):
        each(I.inner, size.v, lambda i: f(CoatedV(i)))
    elif _is_tuple_type(I):
        if len(I) == 0:
            f(())
        elif len(I) == 1:
            each(I[0], size[0], lambda i: f((i,)))
        elif len(I) == 2:
            each(I[0], size[0], lambda i: each(I[1], size[1], lambda j: f((i, j))))
        elif len(I) == 3:
            each(I[0], size[0], lambda i: each(I[1], size[1], lambda j: each(I[2], size[2], lambda k: f((i, j, k)))))
        else:
            raise TypeError("tuple arity is 1..3 only (src/tuple.rs:92-145)")
    else:
        raise TypeError(I)


def size_type(I):
    """<I as Index>::Size, as a type structure."""
    if I is usize or I is Reversed:
        return usize
    if I is bool or isinstance(I, Fixed):
        return ()
    if isinstance(I, Option):
        return size_type(I.inner)
    if isinstance(I, Coated):
        return Coated(size_type(I.inner))
    if _is_tuple_type(I):
        return tuple(size_type(t) for t in I)
    raise TypeError(I)


# ---- tuple isomorphism (src/tuple.rs:60-176) -------------------------------------------------
def flatten_type(T):
    """Canonical form: the ordered list of NonTuple leaves; () contributes nothing."""
    if _is_tuple_type(T):
        out = []
        for t in T:
            out.extend(flatten_type(t))
        return out
    return [T]


def flatten_value(v):
    if isinstance(v, tuple):
        out = []
        for x in v:
            out.extend(flatten_value(x))
        return out
    return [v]


def unflatten(T, leaves):
    it = iter(leaves)

    def build(t):
        if _is_tuple_type(t):
            return tuple(build(x) for x in t)
        return next(it)

    out = build(T)
    rest = list(it)
    assert not rest, "not isomorphic"
    return out


def isomorphic(T, U):
    return flatten_type(T) == flatten_type(U)


def to_iso(value, T):
    """value.to_iso() into type structure T."""
    leaves = flatten_value(value)
    assert len(leaves) == len(flatten_type(T)), f"{value!r} is not isomorphic to {T!r}"
    return unflatten(T, leaves)


def to_iso_size(size, I):
    return to_iso(size, size_type(I))


# ---- Broadcast (src/broadcast.rs:22-162) -----------------------------------------------------
def _is_nontuple(T):
    return not _is_tuple_type(T)


def broadcast_type(I, J):
    if _is_nontuple(I) and _is_nontuple(J):
        if I != J:
            raise TypeError(f"{I!r} does not implement Broadcast<{J!r}>")
        return I
    if I == () and J == ():
        raise TypeError("() does not implement Broadcast<()> (Expand excludes (), src/broadcast.rs:4-9)")
    if I == ():
        return J
    if J == ():
        return I
    if _is_tuple_type(I) and _is_tuple_type(J) and len(I) == len(J):
        return tuple(broadcast_type(a, b) for a, b in zip(I, J))
    raise TypeError(f"{I!r} does not implement Broadcast<{J!r}>")


def broadcast_size(I, J, si, sj):
    if _is_nontuple(I) and _is_nontuple(J):
        if si != sj:
            raise Panic("Unequal sizes")  # src/broadcast.rs:38
        return si
    if I == ():
        return sj
    if J == ():
        return si
    return tuple(broadcast_size(a, b, x, y) for a, b, x, y in zip(I, J, si, sj))


def broadcast_index(I, J, index):
    if _is_nontuple(I) and _is_nontuple(J):
        return index, index
    if I == ():
        return (), index
    if J == ():
        return index, ()
    pairs = [broadcast_index(a, b, x) for a, b, x in zip(I, J, index)]
    return tuple(p[0] for p in pairs), tuple(p[1] for p in pairs)


# ---- ops.rs vocabulary (src/ops.rs:23-129) ---------------------------------------------------
class BinaryOp:
    def __init__(self, name, fn):
        self.name, self.call = name, fn


Pair = BinaryOp("Pair", lambda t, u: (t, u))
Add = BinaryOp("Add", lambda t, u: t + u)
Sub = BinaryOp("Sub", lambda t, u: t - u)
Mul = BinaryOp("Mul", lambda t, u: t * u)
Div = BinaryOp("Div", lambda t, u: t // u if isinstance(t, int) else t / u)
Rem = BinaryOp("Rem", lambda t, u: t % u)
BitAnd = BinaryOp("BitAnd", lambda t, u: t & u)
BitOr = BinaryOp("BitOr", lambda t, u: t | u)
BitXor = BinaryOp("BitXor", lambda t, u: t ^ u)
Shl = BinaryOp("Shl", lambda t, u: t << u)
Shr = BinaryOp("Shr", lambda t, u: t >> u)


# --------------------------------------------------------------------------- View (src/view.rs:116-653)
class View:
    I = None  # index type structure

    def size(self):
        raise NotImplementedError

    def at(self, index):
        raise NotImplementedError

    def len(self):  # src/view.rs:127
        return length(self.I, self.size())

    # src/view.rs:146-150 — drive at() once per index in to_usize order into the sink
    def collect(self, I=None):
        size = self.size()
        buffer = []
        self.each(buffer.append)
        out_I = self.I if I is None else I
        assert isomorphic(out_I, self.I)
        return Array._new_inner(out_I, to_iso(size, size_type(out_I)), buffer)

    def nested_collect(self, size):  # src/view.rs:226-239
        buffer = []

        def one(v):
            if v.size() != size:
                raise Panic("assertion failed: v.size() == size")
            v.each(buffer.append)

        self.each(one)
        I = (self.I, None)
        inner_I = []
        self.each(lambda v: inner_I.append(v.I))
        I = (self.I, inner_I[0])
        return Array._new_inner(I, (self.size(), size), buffer)

    def each(self, f):  # src/view.rs:250-252
        each(self.I, self.size(), lambda i: f(self.at(i)))

    def enumerate(self):
        return Enumerate(self)

    def diagonal(self, zero):
        return Diagonal(self, zero)

    def map(self, f):
        return Map(self, f)

    def compose(self, other):
        return Compose(self, other)

    def concat(self, other, I, J):
        return Concat(self, other, I, J)

    def from_usize(self, I, X, J, from_length):
        return FromUsize(self, I, X, J, from_length)

    def to_usize(self, I, X, J):
        return ToUsize(self, I, X, J)

    def insert_one(self, I, J, K, size):
        return InsertOne(self, I, J, K, size)

    def remove_one(self, I, J, K):
        return RemoveOne(self, I, J, K)

    def map_axis(self, other, I, J):
        return MapAxis(self, I, other, J)

    def zip(self, other):
        return Zip(self, other, Pair)

    def binary(self, other, B):
        return Zip(self, other, B)

    def coat(self, I):
        return CoatView(self, I)

    def iso(self, J):
        return Iso(self, J)

    def transpose(self, I, X, Y, J):
        return Transpose(self, I, X, Y, J)

    def row(self, I, J, i):
        return Row(self, I, J, i)

    def rows(self, I, J):
        return Rows(self, I, J)

    def column(self, I, J, j):
        return Column(self, I, J, j)

    def columns(self, I, J):
        return Columns(self, I, J)

    def nested(self):
        return Nested(self)

    # impl_ops_for_view! (src/ops.rs:195-208)
    def __add__(self, o): return self.binary(o, Add)
    def __sub__(self, o): return self.binary(o, Sub)
    def __mul__(self, o): return self.binary(o, Mul)
    def __truediv__(self, o): return self.binary(o, Div)
    def __mod__(self, o): return self.binary(o, Rem)
    def __and__(self, o): return self.binary(o, BitAnd)
    def __or__(self, o): return self.binary(o, BitOr)
    def __xor__(self, o): return self.binary(o, BitXor)
    def __lshift__(self, o): return self.binary(o, Shl)
    def __rshift__(self, o): return self.binary(o, Shr)

    # impl_ops_for_memoryview! (src/ops.rs:238-253) — only meaningful on memory-backed views
    def __getitem__(self, index):
        return self.at(index)


class Array(View):  # src/array.rs:5-114
    def __init__(self, I, size, items):
        self.I, self._size, self.items = I, size, items

    @staticmethod
    def _new_inner(I, size, items):  # src/array.rs:11-14
        items = list(items)
        if length(I, size) != len(items):
            raise Panic(f"assertion `left == right` failed\n  left: {length(I, size)}\n right: {len(items)}")
        return Array(I, size, items)

    @staticmethod
    def new(I, size, items):  # src/array.rs:28-30
        return Array._new_inner(I, to_iso_size(size, I), items)

    @staticmethod
    def from_fn(I, size, f):  # src/array.rs:43-51
        size = to_iso_size(size, I)
        items = []
        each(I, size, lambda i: items.append(f(i)))
        return Array._new_inner(I, size, items)

    def to_raw(self):
        return self.items

    def as_ref(self):
        return self.items

    def iso(self, J):  # src/array.rs:57-62 (no data movement)
        assert isomorphic(J, self.I)
        return Array(J, to_iso(self._size, size_type(J)), self.items)

    def size(self):
        return self._size

    def len(self):
        return len(self.items)

    def at(self, index):  # src/array.rs:81,86
        k = to_usize(self.I, index, self._size)
        if not (0 <= k < len(self.items)):
            raise Panic(f"index out of bounds: the len is {len(self.items)} but the index is {k}")
        return self.items[k]

    def at_mut_set(self, index, value):  # src/array.rs:91
        self.items[to_usize(self.I, index, self._size)] = value


class All(View):  # src/index.rs:177-186
    def __init__(self, I, size):
        self.I, self._size = I, size

    def size(self): return self._size
    def at(self, index): return index


def all_(I, size):
    """`I::all(size)`, src/index.rs:70."""
    return All(I, to_iso_size(size, I))


class Scalar(View):  # src/view.rs:1399-1408
    I = ()

    def __init__(self, value):
        self.value = value

    def size(self): return ()
    def at(self, index): return self.value


class Enumerate(View):  # src/view.rs:829-838
    def __init__(self, v):
        self.v, self.I = v, v.I

    def size(self): return self.v.size()
    def at(self, index): return (index, self.v.at(index))


class Diagonal(View):  # src/view.rs:846-857
    def __init__(self, v, zero):
        self.v, self.zero, self.I = v, zero, (v.I, v.I)

    def size(self): return (self.v.size(), self.v.size())

    def at(self, index):
        return self.v.at(index[0]) if index[0] == index[1] else self.zero


class Map(View):  # src/view.rs:880-889
    def __init__(self, v, f):
        self.v, self.f, self.I = v, f, v.I

    def size(self): return self.v.size()
    def at(self, index): return self.f(self.v.at(index))


class Compose(View):  # src/view.rs:897-912
    def __init__(self, v, w):
        self.v, self.w, self.I = v, w, v.I

    def size(self): return self.v.size()
    def at(self, index): return self.w.at(self.v.at(index))


class Concat(View):  # src/view.rs:920-946, ctor asserts :334-338
    def __init__(self, v, w, I, J):
        T = (I, usize, J)
        assert isomorphic(v.I, T) and isomorphic(w.I, T)
        si, sn, sj = to_iso(v.size(), size_type(T))
        oi, _on, oj = to_iso(w.size(), size_type(T))
        if si != oi or sj != oj:
            raise Panic("assertion `left == right` failed")
        self.v, self.w, self.n, self.I = v, w, sn, T

    def size(self):
        i, wn, j = to_iso(self.w.size(), size_type(self.I))
        return (i, self.n + wn, j)

    def at(self, index):
        i, k, j = index
        if k < self.n:
            return self.v.at(to_iso((i, k, j), self.v.I))
        return self.w.at(to_iso((i, k - self.n, j), self.w.I))


class FromUsize(View):  # src/view.rs:352-363, 993-1021
    def __init__(self, v, I, X, J, from_length):
        assert isomorphic(v.I, (I, usize, J))
        _, old, _ = to_iso(v.size(), size_type((I, usize, J)))
        self.xsize = from_length(old)
        if length(X, self.xsize) != old:
            raise Panic("assertion `left == right` failed")
        self.v, self.I, self._inner = v, (I, X, J), (I, usize, J)

    def size(self):
        i, _, j = to_iso(self.v.size(), size_type(self._inner))
        return (i, self.xsize, j)

    def at(self, index):
        i, x, j = index
        return self.v.at(to_iso((i, to_usize(self.I[1], x, self.xsize), j), self.v.I))


class ToUsize(View):  # src/view.rs:373-378, 1029-1059
    def __init__(self, v, I, X, J):
        assert isomorphic(v.I, (I, X, J))
        self.v, self.I, self._inner = v, (I, usize, J), (I, X, J)

    def size(self):
        i, x, j = to_iso(self.v.size(), size_type(self._inner))
        return (i, length(self._inner[1], x), j)

    def at(self, index):
        i, k, j = index
        xs = to_iso(self.v.size(), size_type(self._inner))[1]
        q, x = from_usize(self._inner[1], xs, k)
        if q != 0:
            raise Panic("assertion `left == right` failed")  # src/view.rs:1056
        return self.v.at(to_iso((i, x, j), self.v.I))


class InsertOne(View):  # src/view.rs:391-397, 1067-1096
    def __init__(self, v, I, J, K, size):
        assert isomorphic(v.I, (I, K))
        if length(J, size) != 1:
            raise Panic("assertion `left == right` failed")
        self.v, self.jsize, self.I, self._inner = v, size, (I, J, K), (I, K)

    def size(self):
        i, k = to_iso(self.v.size(), size_type(self._inner))
        return (i, self.jsize, k)

    def at(self, index):
        i, j, k = index
        if to_usize(self.I[1], j, self.jsize) != 0:
            raise Panic("assertion `left == right` failed")
        return self.v.at(to_iso((i, k), self.v.I))


class RemoveOne(View):  # src/view.rs:408-418, 1104-1132
    def __init__(self, v, I, J, K):
        assert isomorphic(v.I, (I, J, K))
        _, js, _ = to_iso(v.size(), size_type((I, J, K)))
        if length(J, js) != 1:
            raise Panic("assertion `left == right` failed")
        q, j = from_usize(J, js, 0)
        assert q == 0 and to_usize(J, j, js) == 0
        self.v, self.j, self.I, self._inner = v, j, (I, K), (I, J, K)

    def size(self):
        i, _, k = to_iso(self.v.size(), size_type(self._inner))
        return (i, k)

    def at(self, index):
        i, k = index
        return self.v.at(to_iso((i, self.j, k), self.v.I))


class MapAxis(View):  # src/view.rs:436-442, 1140-1170
    def __init__(self, v, I, w, J, WT=usize):
        # W::T is the index type of the mapped axis; the model takes it as usize unless told.
        self.v, self.w, self._inner = v, w, (I, WT, J)
        assert isomorphic(v.I, self._inner)
        self.I = (I, w.I, J)

    def size(self):
        i, _, j = to_iso(self.v.size(), size_type(self._inner))
        return (i, self.w.size(), j)

    def at(self, index):
        i, x, j = index
        return self.v.at(to_iso((i, self.w.at(x), j), self.v.I))


class Zip(View):  # src/view.rs:1178-1198
    def __init__(self, v, w, B):
        self.v, self.w, self.B = v, w, B
        self.I = broadcast_type(v.I, w.I)

    def size(self):
        return broadcast_size(self.v.I, self.w.I, self.v.size(), self.w.size())

    def at(self, index):
        vi, wi = broadcast_index(self.v.I, self.w.I, index)
        return self.B.call(self.v.at(vi), self.w.at(wi))


def _coat_value(v, src, dst):
    """value.coat(): add or remove at most one level of Coated per position (src/coat.rs:23-69)."""
    if isinstance(dst, Coated) and not isinstance(src, Coated):
        return CoatedV(v)
    if isinstance(src, Coated) and not isinstance(dst, Coated):
        return v.v
    if _is_tuple_type(src) and _is_tuple_type(dst):
        return tuple(_coat_value(x, s, d) for x, s, d in zip(v, src, dst))
    return v


class CoatView(View):  # src/view.rs:549-556, 1206-1230
    def __init__(self, v, I):
        self.v, self.I = v, I

    def size(self):
        return _coat_value(self.v.size(), size_type(self.v.I), size_type(self.I))

    def at(self, index):
        return self.v.at(_coat_value(index, self.I, self.v.I))


class Iso(View):  # src/view.rs:559-564, 1238-1258
    def __init__(self, v, J):
        assert isomorphic(J, v.I), f"{J!r} is not isomorphic to {v.I!r}"
        self.v, self.I = v, J

    def size(self): return to_iso(self.v.size(), size_type(self.I))
    def at(self, index): return self.v.at(to_iso(index, self.v.I))


class Transpose(View):  # src/view.rs:586-592, 1266-1294
    def __init__(self, v, I, X, Y, J):
        self._inner = (I, (Y, X), J)
        assert isomorphic(self._inner, v.I)
        self.v, self.I = v, (I, (X, Y), J)

    def size(self):
        i, (y, x), j = to_iso(self.v.size(), size_type(self._inner))
        return (i, (x, y), j)

    def at(self, index):
        i, (x, y), j = index
        return self.v.at(to_iso((i, (y, x), j), self.v.I))


class Row(View):  # src/view.rs:609-614, 1302-1322
    def __init__(self, v, I, J, i):
        assert isomorphic((I, J), v.I)
        self.v, self.i, self.I, self._inner = v, i, J, (I, J)

    def size(self): return to_iso(self.v.size(), size_type(self._inner))[1]
    def at(self, index): return self.v.at(to_iso((self.i, index), self.v.I))


class Rows(View):  # src/view.rs:617-622, 1330-1342
    def __init__(self, v, I, J):
        assert isomorphic((I, J), v.I)
        self.v, self.I, self._I, self._J = v, I, I, J

    def size(self): return to_iso(self.v.size(), size_type((self._I, self._J)))[0]
    def at(self, index): return Row(self.v, self._I, self._J, index)


class Column(View):  # src/view.rs:639-644, 1350-1370
    def __init__(self, v, I, J, j):
        assert isomorphic((I, J), v.I)
        self.v, self.j, self.I, self._inner = v, j, I, (I, J)

    def size(self): return to_iso(self.v.size(), size_type(self._inner))[0]
    def at(self, index): return self.v.at(to_iso((index, self.j), self.v.I))


class Columns(View):  # src/view.rs:647-652, 1376-1390
    def __init__(self, v, I, J):
        assert isomorphic((I, J), v.I)
        self.v, self.I, self._I, self._J = v, J, I, J

    def size(self): return to_iso(self.v.size(), size_type((self._I, self._J)))[1]
    def at(self, index): return Column(self.v, self._I, self._J, index)


class Nested(View):  # src/view.rs:173-179, 788-800
    def __init__(self, v):
        self.v = v
        inner = []
        each(v.I, v.size(), lambda i: inner.append(v.at(i).I) if not inner else None)
        self.I = (v.I, inner[0])

    def size(self): return (self.v.size(), to_iso((), size_type(self.I[1])))
    def at(self, index): return self.v.at(index[0]).at(index[1])


def fn_view(I, size, f):  # src/view.rs:1436-1443
    return all_(I, size).map(f)


def fold_rows_from(v, I, J, op, init_view):
    """The same fold with a per-row initial value: `rows().zip(init).map(|(row, s0)| { let mut s = s0; row.each(..); s })`
    (Zip of src/view.rs:1178-1198 over views indexed by I; each row starts from init_view.at(i))."""
    def fold(pair):
        row, s0 = pair
        s = [s0]
        row.each(lambda x: s.__setitem__(0, op(s[0], x)))
        return s[0]
    return v.rows(I, J).zip(init_view).map(fold)


def fold_rows(v, I, J, op, init):
    """The reference's ONLY spelling of an axis fold (no reduce API exists):
        v.rows::<I,J>().map(|row| { let mut s = init; row.each(|x| s = op(s, x)); s })
    src/view.rs:617-622 (rows), :1341 (Rows::at), :250-252 (each) — sequential, index order."""
    def fold(row):
        s = [init]
        row.each(lambda x: s.__setitem__(0, op(s[0], x)))
        return s[0]
    return v.rows(I, J).map(fold)


def blocked_fold_over_sharded_axis(blocks, op, init, identity):
    """The all-reduce route of a fold over the SHARDED (outermost) axis (last row), restated with numpy rows:
    `blocks[r]` holds rank r's rows (2-D array, rows x columns) in index order.  Every rank folds ITS rows exactly as the reference
    does — `let mut s = start; row.each(|x| s = op(s, x))`, sequential, index order (src/view.rs:617-622, 250-252) — rank 0 starting
    from `init`, the others from the operator's `identity`; the partial results are then combined in RANK ORDER:
        out = ((P_0 (op) P_1) (op) P_2) ... (op) P_{N-1}.
    For an associative operator this IS the reference's result; a float sum is reassociated at the rank boundaries only."""
    import numpy as np
    total = None
    for r, rows in enumerate(blocks):
        rows = np.asarray(rows)
        part = np.full(rows.shape[1], init if r == 0 else identity, dtype=rows.dtype)
        for i in range(rows.shape[0]):
            part = op(part, rows[i])
        total = part if r == 0 else op(total, part)
    return total
