"""Expressions beyond what ONE fused kernel takes (48 nodes, 12 array operands, one fold: include/mdim.h) are collected through dense
temporaries (lowering.split_for_limits, used by View.collect and by the CPU runners of tests/helpers.py).  The reference has no such
bound, so the result must still be the reference's: every case below is built twice — with the index-tuple-level model
(oracle/reference_model.py) and with the product — and compared bit for bit, on the host build of the evaluator (`[emu]`) and through
the C ABI on the GPU (`[gpu]`)."""
import random

import numpy as np
import pytest

import multidimension_b200 as P
from multidimension_b200 import _ffi as F
from multidimension_b200 import lowering as L
from oracle import reference_model as M

from helpers import emu_collect
from test_differential_model import Both, fresh_array, compare, model_result, mt

_gpu_ctx = []
usize = P.usize


@pytest.fixture(params=[pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "gpu" and not _gpu_ctx:
        _gpu_ctx.append(P.Context(0))
        P.set_default_context(_gpu_ctx[0])
    return request.param


def product(view, backend):
    if backend == "emu":
        return emu_collect(view)
    return view.collect(location="device", ctx=_gpu_ctx[0]).as_ref()


def binary(a, b, name):
    fn = {"Add": lambda x, y: x + y, "Sub": lambda x, y: x - y, "Mul": lambda x, y: x * y}[name]
    return Both(a.m.binary(b.m, getattr(M, name)), a.p.binary(b.p, getattr(P, name)), a.T, f"({a.note} {name} {b.note})")


def fold(v, I, J, name, init):
    fn = {"Add": lambda x, y: x + y, "Mul": lambda x, y: x * y}[name]
    return Both(M.fold_rows(v.m, mt(I), mt(J), fn, init), P.fold_rows(v.p, I, J, getattr(P, name), init), v.T, f"fold({v.note})")


def n_nodes(view):
    return L._tree_cost(L.flatten_value(view._lower()[1])[0])


def test_a_chain_of_sixty_operators(backend):
    rng = random.Random(1)
    v = fresh_array(rng, (usize, usize), (7, 9), "f32")
    for k in range(60):
        w = fresh_array(rng, (usize, usize), (7, 9), "f32") if k % 9 == 0 else Both(M.Scalar(np.float32(k % 5 - 2)), P.Scalar(np.float32(k % 5 - 2), "f32"), "f32", "Scalar")
        v = binary(v, w, ("Add", "Sub", "Mul")[k % 3] if k % 7 else "Add")
    assert n_nodes(v.p)[0] > F.MAX_NODES
    with pytest.raises(P.Unsupported):
        v.p.describe()                                   # ONE kernel cannot take it ...
    compare(v, product(v.p, backend), model_result(v), "60 operators")   # ... collect() does, through temporaries


def test_twenty_array_operands_with_broadcasting(backend):
    rng = random.Random(2)
    v = fresh_array(rng, (usize, usize), (6, 10), "i64")
    for k in range(19):
        shape = ((usize, usize), (6, 10)) if k % 3 else (((usize, ()), (6, ())) if k % 2 else (((), usize), ((), 10)))
        v = binary(v, fresh_array(rng, shape[0], shape[1], "i64"), ("Add", "Mul", "Sub")[k % 3])
    assert n_nodes(v.p)[1] > 12
    compare(v, product(v.p, backend), model_result(v), "20 operands")


def test_two_folds_in_one_expression(backend):
    """(x - mean(x)) * (x - mean(x)) summed per row, minus a second, different fold: three FOLD nodes where one expression takes one."""
    rng = random.Random(3)
    x = fresh_array(rng, (usize, usize), (11, 16), "f32")
    y = fresh_array(rng, (usize, usize), (11, 16), "f32")
    sx = fold(x, usize, usize, "Add", np.float32(0))
    sy = fold(y, usize, usize, "Mul", np.float32(1))
    both = binary(sx, sy, "Sub")                                            # two folds side by side
    compare(both, product(both.p, backend), model_result(both), "fold - fold")
    mean_like = Both(sx.m.iso(mt((usize, ()))), sx.p.iso((usize, ())), "f32", "sums as a column")
    centred = binary(x, mean_like, "Sub")                                   # x - sums (broadcast over the row): a fold under an operator ...
    var_like = fold(binary(centred, centred, "Mul"), usize, usize, "Add", np.float32(0))   # ... inside another fold's body
    with pytest.raises(P.Unsupported):
        var_like.p.describe()
    with pytest.raises(Exception):
        product(var_like.p, backend)     # the inner fold sits INSIDE a fold body (not on the unconditional spine): declined, never wrong
    tail = binary(fold(x, usize, usize, "Add", np.float32(0.5)), binary(sx, sy, "Mul"), "Add")   # three folds, all on the spine
    compare(tail, product(tail.p, backend), model_result(tail), "fold + fold * fold")


def test_lazy_regions_are_never_cut(backend):
    """A sub-expression inside a Diagonal is evaluated only on the diagonal (src/view.rs:846-857): an over-limit chain in there stays
    in one piece — and is declined — rather than being evaluated (and possibly panicking) off the diagonal through a temporary."""
    rng = random.Random(4)
    v = fresh_array(rng, usize, 6, "i64")
    for k in range(30):
        v = binary(v, Both(M.Scalar(k % 3), P.Scalar(k % 3, "i64"), "i64", "Scalar"), "Add")
    d = Both(v.m.diagonal(0), v.p.diagonal(0), "i64", "diagonal(long chain)")
    with pytest.raises(Exception) as e:
        product(d.p, backend)
    assert getattr(e.value, "status", None) == F.ERR_UNSUPPORTED
    short = fresh_array(rng, usize, 6, "i64")                              # the same chain OUTSIDE the diagonal is fine
    dd = Both(short.m.diagonal(0), short.p.diagonal(0), "i64", "diagonal")
    w = dd
    for k in range(30):
        w = binary(w, Both(M.Scalar(k % 3), P.Scalar(k % 3, "i64"), "i64", "Scalar"), "Add")
    compare(w, product(w.p, backend), model_result(w), "long chain over a diagonal")


def _long_chain(seed):
    """test_differential_model.build_chain with 10-16 operations instead of 1-4, arithmetic-heavy: many of them exceed one expression."""
    import test_differential_model as D
    rng = random.Random(seed)
    v = D.start_view(rng, "i64")   # integers: the model's unbounded ints reduced mod 2^64 ARE Rust's wrapping + - * (a ring homomorphism); floats would drown in inf / nan
    applied = 0
    for _ in range(60):
        if applied >= 10 + seed % 7:
            break
        op = rng.choice(D.STRUCTURAL + D.ARITH * 6)
        try:
            w = op(rng, v)
        except (D.Skip, P.Unsupported):
            continue
        if w.length() > 3000 or D.n_axes(w.I, w.size) > D.MAX_AXES or isinstance(w.T, tuple):
            continue
        v = w
        applied += 1
    return v


def test_long_random_chains_against_the_reference_model(backend):
    """40 long random chains (structure AND arithmetic): whatever does not fit one kernel is split, and the result is still the model's."""
    ran = split = 0
    for seed in range(40):
        v = _long_chain(9000 + seed)
        try:
            big = n_nodes(v.p)
            got = product(v.p, backend)
        except Exception as e:
            if getattr(e, "status", None) != F.ERR_UNSUPPORTED:
                raise
            continue  # declined loudly (e.g. an over-limit chain inside a lazy region)
        want = np.array([((int(w) + (1 << 63)) % (1 << 64)) - (1 << 63) for w in model_result(v)], dtype=np.int64)
        from helpers import assert_same_bits
        assert_same_bits(np.asarray(got).reshape(-1), want, f"seed {9000 + seed}: {v.note[:200]}")
        ran += 1
        split += int(big[0] > F.MAX_NODES or big[1] > 12 or big[2] > 1)
    assert ran >= 25 and split >= 15, (ran, split)   # the rest is declined loudly: > 8 axes, or the long part sits inside a Concat side / Diagonal
