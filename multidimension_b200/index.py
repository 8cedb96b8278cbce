"""Index types of the reference (src/index.rs, src/int.rs, src/tuple.rs, src/coat.rs), host side.

Index TYPES are spelled almost as in Rust:  usize, bool, (), (I,), (I, J), (I, J, K), Fixed(n),
Reversed, Coated(I).  SIZES are `int` for usize/Reversed, `()` for bool/()/Fixed, tuples of sizes
for tuples, and the inner size for Coated.

On the device everything lives in POSITION SPACE: an index type flattens to an ordered list of
leaf axes (tuple.rs:60-176), `Index::to_usize` is row-major over that list (index.rs:109-114) and
`Index::each` walks it last-axis-fastest (index.rs:122-124), so the only run-time content of an
index type is the list of leaf lengths computed here.
"""
from __future__ import annotations


class _Leaf:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


usize = _Leaf("usize")        # src/int.rs:9-26
Reversed = _Leaf("Reversed")  # src/int.rs:58-86: position p <-> index size-1-p


class Fixed:  # src/int.rs:33-54
    def __init__(self, n):
        self.n = int(n)

    def __eq__(self, o):
        return isinstance(o, Fixed) and o.n == self.n

    def __hash__(self):
        return hash(("Fixed", self.n))

    def __repr__(self):
        return f"Fixed<{self.n}>"


class Coated:  # src/coat.rs:6-21: hides the tuple structure of I from Isomorphic
    def __init__(self, inner):
        self.inner = inner

    def __eq__(self, o):
        return isinstance(o, Coated) and o.inner == self.inner

    def __hash__(self):
        return hash(("Coated", self.inner))

    def __repr__(self):
        return f"Coated<{self.inner!r}>"


class Option:  # src/index.rs:244-276: None is position 0, Some(i) is 1 + i.to_usize()
    def __init__(self, inner):
        self.inner = inner

    def __eq__(self, o):
        return isinstance(o, Option) and o.inner == self.inner

    def __hash__(self):
        return hash(("Option", self.inner))

    def __repr__(self):
        return f"Option<{self.inner!r}>"


class Some:
    """Index VALUE `Some(i)` of an Option<I> axis (`None` is Python's None)."""
    def __init__(self, i):
        self.i = i


class IndexError_(TypeError):
    """A constraint the Rust type checker would have rejected (e.g. a failed Isomorphic bound)."""


def is_tuple_type(I):
    return isinstance(I, tuple)


def check_type(I):
    if I is usize or I is Reversed or I is bool or isinstance(I, (Fixed, Coated, Option)):
        if isinstance(I, (Coated, Option)):
            check_type(I.inner)
        return
    if isinstance(I, tuple):
        if len(I) > 3:  # src/tuple.rs:92-145: arity 0..3 only
            raise IndexError_(f"tuple index types have arity <= 3, got {len(I)}")
        for t in I:
            check_type(t)
        return
    raise IndexError_(f"not an index type: {I!r}")


def type_leaves(I):
    """Flatten (tuple.rs:60-145): the ordered NonTuple leaves.  Coated(I) is ONE leaf here."""
    if isinstance(I, tuple):
        out = []
        for t in I:
            out.extend(type_leaves(t))
        return out
    return [I]


def isomorphic(I, J):  # tuple.rs:166-176
    return type_leaves(I) == type_leaves(J)


def leaf_lengths(I, size):
    """Lengths of the position-space axes of ONE NonTuple leaf type (Coated expands)."""
    if I is usize or I is Reversed:
        if not isinstance(size, int) or isinstance(size, bool) or size < 0:
            raise IndexError_(f"size of {I!r} must be a non-negative int, got {size!r}")
        return [size]
    if I is bool:  # StaticIndex, index.rs:236-240
        _expect_unit(I, size)
        return [2]
    if isinstance(I, Fixed):
        _expect_unit(I, size)
        return [I.n]
    if isinstance(I, Coated):
        out = []
        for t, s in zip(type_leaves(I.inner), size_leaves(I.inner, size)):
            out.extend(leaf_lengths(t, s))
        return out
    if isinstance(I, Option):  # one position axis: [None, Some(0), Some(1), ...] (src/index.rs:248)
        return [1 + length(I.inner, size)]
    raise IndexError_(f"not a leaf index type: {I!r}")


def _expect_unit(I, size):
    if size != ():
        raise IndexError_(f"size of {I!r} must be (), got {size!r}")


def size_leaves(I, size):
    """The sizes of the NonTuple leaves of I, in order (the Flatten of I::Size)."""
    if isinstance(I, tuple):
        if not isinstance(size, tuple) or len(size) != len(I):
            raise IndexError_(f"size {size!r} does not fit index type {I!r}")
        out = []
        for t, s in zip(I, size):
            out.extend(size_leaves(t, s))
        return out
    return [size]


def build_size(I, leaves):
    """Inverse of size_leaves: consume `leaves` (a list, popped from the front) into I's shape."""
    if isinstance(I, tuple):
        return tuple(build_size(t, leaves) for t in I)
    return leaves.pop(0)


def to_iso_size(size, I_from, I_to):
    """Isomorphic::to_iso on sizes (tuple.rs:171-176)."""
    if not isomorphic(I_from, I_to):
        raise IndexError_(f"{I_from!r} is not isomorphic to {I_to!r}")
    leaves = size_leaves(I_from, size)
    out = build_size(I_to, list(leaves))
    return out


def coerce_size(I, size):
    """Accept any size isomorphic to I::Size, as `Array::new(size: impl Isomorphic<I::Size>)` does
    (array.rs:28): a flat size whose non-unit entries match the non-unit leaves of I."""
    try:
        size_leaves(I, size)
        return size
    except IndexError_:
        pass
    flat = []

    def walk(s):
        if isinstance(s, tuple):
            for x in s:
                walk(x)
        else:
            flat.append(s)
    walk(size)
    leaves = []
    it = iter(flat)
    for t in type_leaves(I):
        if t is usize or t is Reversed:
            try:
                leaves.append(next(it))
            except StopIteration:
                raise IndexError_(f"size {size!r} does not fit index type {I!r}")
        elif isinstance(t, (Coated, Option)):
            raise IndexError_("give Coated / Option sizes in full")
        else:
            leaves.append(())
    if any(True for _ in it):
        raise IndexError_(f"size {size!r} does not fit index type {I!r}")
    return build_size(I, leaves)


def length(I, size):  # Index::length, index.rs:104-107
    n = 1
    for t, s in zip(type_leaves(I), size_leaves(I, size)):
        for L in leaf_lengths(t, s):
            n *= L
    return n


def index_positions(I, index, size):
    """Positions (one per position-space axis) of an index VALUE, with the reference's bounds
    assert (int.rs:16-19).  Index values: int, bool, (), tuples."""
    if isinstance(I, tuple):
        if not isinstance(index, tuple) or len(index) != len(I):
            raise IndexError_(f"index {index!r} does not fit index type {I!r}")
        out = []
        for t, x, s in zip(I, index, size):
            out.extend(index_positions(t, x, s))
        return out
    if I is usize or I is Reversed:
        if not (0 <= index < size):
            from ._ffi import Panic, ERR_OOB
            raise Panic(ERR_OOB, f"Index {index} is out of bounds for size {size}")
        return [index if I is usize else size - 1 - index]
    if I is bool:
        return [1 if index else 0]
    if isinstance(I, Fixed):
        # Fixed::to_usize is unchecked (src/int.rs:44) but the slice access items[to_usize] panics (src/array.rs:86)
        if not (0 <= int(index) < I.n):
            from ._ffi import Panic, ERR_OOB
            raise Panic(ERR_OOB, f"Index {index} is out of bounds for size {I.n}")
        return [int(index)]
    if isinstance(I, Coated):
        return index_positions(I.inner, index, size)
    if isinstance(I, Option):  # src/index.rs:250-256
        if index is None:
            return [0]
        k = 0
        inner = index.i if isinstance(index, Some) else index
        lens = [n for t, s in zip(type_leaves(I.inner), size_leaves(I.inner, size)) for n in leaf_lengths(t, s)]
        for p, n in zip(index_positions(I.inner, inner, size), lens):
            k = k * n + p
        return [1 + k]
    raise IndexError_(f"not an index type: {I!r}")


def index_leaves(I, index):
    """The components of an index VALUE, one per NonTuple leaf of I (parallel to type_leaves)."""
    if isinstance(I, tuple):
        if not isinstance(index, tuple) or len(index) != len(I):
            raise IndexError_(f"index {index!r} does not fit index type {I!r}")
        out = []
        for t, x in zip(I, index):
            out.extend(index_leaves(t, x))
        return out
    return [index]
