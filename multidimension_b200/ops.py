"""The lowerable operator vocabulary.

Binary: the reference's uninstantiable enums `ops::{Add,Sub,Mul,Div,Rem,BitAnd,BitOr,BitXor,Shl,
Shr,Pair}` (src/ops.rs:23-129), applied through `View::binary::<_, B>()` / the operator sugar of
`impl_ops_for_view!` (src/ops.rs:159-208).

Unary: `Map` takes an opaque closure in the reference (src/view.rs:299-303, 880-889), which no
device can run.  The device `map` accepts (a) the closed type-level set below, in the same style
as ops.rs, (b) `Fold(B, init)` over `rows()`, the reference's only spelling of an axis reduction
(src/view.rs:617-622, 250-252), and (c) a Python callable that is TRACED once with symbolic
values, so closures written with the operators above (`lambda p: p[0] * p[1] + 1.0`) lower to
the same fused node tree as `a * b + Scalar(1.0)`.
"""
from __future__ import annotations

from . import _ffi as F


class BinaryOp:
    def __init__(self, name, code):
        self.name, self.code = name, code

    def __repr__(self):
        return f"ops::{self.name}"


Pair = BinaryOp("Pair", None)        # src/ops.rs:25-29
Add = BinaryOp("Add", F.ADD)         # src/ops.rs:33
Sub = BinaryOp("Sub", F.SUB)         # src/ops.rs:43
Mul = BinaryOp("Mul", F.MUL)         # src/ops.rs:53
Div = BinaryOp("Div", F.DIV)         # src/ops.rs:63
Rem = BinaryOp("Rem", F.REM)         # src/ops.rs:73
BitAnd = BinaryOp("BitAnd", F.AND)   # src/ops.rs:83
BitOr = BinaryOp("BitOr", F.OR)      # src/ops.rs:93
BitXor = BinaryOp("BitXor", F.XOR)   # src/ops.rs:103
Shl = BinaryOp("Shl", F.SHL)         # src/ops.rs:113
Shr = BinaryOp("Shr", F.SHR)         # src/ops.rs:123


class UnaryOp:
    def __init__(self, name, code):
        self.name, self.code = name, code

    def __repr__(self):
        return f"ops::{self.name}"


Neg = UnaryOp("Neg", F.NEG)
Not = UnaryOp("Not", F.NOT)
Abs = UnaryOp("Abs", F.ABS)
Sqrt = UnaryOp("Sqrt", F.SQRT)


class Cast:
    """Rust `x as T`."""

    def __init__(self, T):
        self.T = T


class Fold:
    """`|row| { let mut s = init; row.each(|x| s = B::call(s, x)); s }` — sequential, index order.
    `init` is a scalar, or a View indexed like `rows()` (then row i starts from `init.at(i)`: what lets a fold be continued
    from another fold's result, e.g. rank r of a sharded axis continuing rank r-1's running values)."""

    def __init__(self, B, init):
        self.B, self.init = B, init
