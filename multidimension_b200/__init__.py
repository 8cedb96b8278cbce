"""multidimension_b200 — B200-native View -> collect() hot path of apt1002/multidimension.

Host-side mirror of the reference's View/Array/Index API (view.py, index.py, ops.py) that lowers a
lazy View chain to the position-space descriptor of include/mdim.h (lowering.py) and runs it as one
fused sm_100a kernel through the C ABI in libmdim_b200.so (csrc/).  No CPU fallback exists.
"""
from ._ffi import MdimError, Panic, LIB_PATH
from .index import usize, Reversed, Fixed, Coated, Option, Some
from .lowering import Unsupported
from .ops import Pair, Add, Sub, Mul, Div, Rem, BitAnd, BitOr, BitXor, Shl, Shr, Neg, Not, Abs, Sqrt, Cast, Fold
from .runtime import Context, Storage, default_context, set_default_context
from .view import View, Array, All, all_, Scalar, fold_rows
from . import sharding

__all__ = [
    "MdimError", "Panic", "Unsupported", "LIB_PATH", "usize", "Reversed", "Fixed", "Coated", "Option", "Some",
    "Pair", "Add", "Sub", "Mul", "Div", "Rem", "BitAnd", "BitOr", "BitXor", "Shl", "Shr",
    "Neg", "Not", "Abs", "Sqrt", "Cast", "Fold", "Context", "Storage", "default_context", "set_default_context",
    "View", "Array", "All", "all_", "Scalar", "fold_rows", "sharding",
]
