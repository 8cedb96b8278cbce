"""Regression tests for the round-1 advisor findings, each pinned to oracle/reference_model.py (the
independent index-tuple-level model), not to the product's own lowering."""
import numpy as np
import pytest

import multidimension_b200 as P
from multidimension_b200 import usize, Array, Fixed, _ffi as F
from multidimension_b200.runtime import Storage
from oracle import reference_model as M

from helpers import emu_collect, CheckerPanic

_gpu_ctx = []


@pytest.fixture(params=[pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)])
def backend(request):
    if request.param == "gpu" and not _gpu_ctx:
        _gpu_ctx.append(P.Context(0))
        P.set_default_context(_gpu_ctx[0])
    return request.param


def collect(view, backend):
    if backend == "emu":
        return emu_collect(view)
    return view.collect(location="device", ctx=_gpu_ctx[0]).as_ref()


def test_zip_after_to_usize_unifies_axis_groups(backend):
    """ADVICE high: `a(2x3).to_usize(..) + b(6)` used to pair the merged group (2, 3) with b's single axis by
    truncating zip(): b was read with the stride of the wrong axis.  src/view.rs:1029-1059, src/broadcast.rs:33-44."""
    av, bv = np.arange(6, dtype=np.int64), np.arange(6, dtype=np.int64) * 10
    a = Array.new((usize, usize), (2, 3), av, "i64")
    b = Array.new(usize, 6, bv, "i64")
    ma = M.Array.new((M.usize, M.usize), (2, 3), [int(x) for x in av])
    mb = M.Array.new(M.usize, 6, [int(x) for x in bv])
    af = a.to_usize((), (usize, usize), ()).iso(usize)
    maf = ma.to_usize((), (M.usize, M.usize), ()).iso(M.usize)
    want = (maf + mb).collect().items
    assert want == [0, 11, 22, 33, 44, 55]
    assert collect(af + b, backend).tolist() == want
    assert collect(b + af, backend).tolist() == want          # the swapped order used to raise "axis not iterated"
    # both sides merged, split differently but refinable: (2, 6) against (4, 3)
    c = Array.new((usize, usize), (2, 6), np.arange(12, dtype=np.int64), "i64").to_usize((), (usize, usize), ()).iso(usize)
    d = Array.new((usize, usize), (4, 3), np.arange(12, dtype=np.int64) * 100, "i64").to_usize((), (usize, usize), ()).iso(usize)
    assert collect(c + d, backend).tolist() == [k + 100 * k for k in range(12)]
    # (2, 3) against (3, 2) has no common refinement: declined loudly, never a wrong value
    e = Array.new((usize, usize), (3, 2), bv, "i64").to_usize((), (usize, usize), ()).iso(usize)
    with pytest.raises(P.Unsupported):
        collect(af + e, backend)


def test_concat_after_to_usize_unifies_the_other_axes(backend):
    av = np.arange(12, dtype=np.int64)
    a = Array.new(((usize, usize), usize), ((2, 3), 2), av, "i64").to_usize((), (usize, usize), usize).iso((usize, usize))
    b = Array.new((usize, usize), (6, 1), av[:6] * 7, "i64")
    ma = M.Array.new(((M.usize, M.usize), M.usize), ((2, 3), 2), [int(x) for x in av]).to_usize((), (M.usize, M.usize), M.usize).iso((M.usize, M.usize))
    mb = M.Array.new((M.usize, M.usize), (6, 1), [int(x) * 7 for x in av[:6]])
    want = ma.concat(mb, M.usize, ()).collect().items
    assert collect(a.concat(b, usize, ()), backend).tolist() == want


def test_not_on_bool_is_logical(backend):
    """ADVICE medium: `!true` was computed as ~1 & 0xff = 0xFE, which reads back as True."""
    vals = [True, False, True, False]
    a = Array.new(usize, 4, np.array(vals), bool)
    want = M.Array.new(M.usize, 4, vals).map(lambda x: not x).collect().items
    assert want == [False, True, False, True]
    assert collect(a.map(P.Not), backend).tolist() == want
    assert collect(a.map(lambda x: ~x), backend).tolist() == want
    u = Array.new(usize, 2, np.array([1, 0xF0], dtype=np.uint8), "u8")  # integers keep the bitwise Not
    assert collect(u.map(P.Not), backend).tolist() == [0xFE, 0x0F]


def test_fixed_index_is_range_checked():
    """ADVICE medium: row(Fixed(3), ..) with index 7 used to become an offset past the buffer; the reference
    panics in the slice access (src/array.rs:86)."""
    a = Array.new((Fixed(3), usize), ((), 4), np.arange(12, dtype=np.int64), "i64")
    assert emu_collect(a.row(Fixed(3), usize, 2)).tolist() == [8, 9, 10, 11]
    with pytest.raises(P.Panic) as e:
        a.row(Fixed(3), usize, 7)._lower()
    assert e.value.status == F.ERR_OOB
    with pytest.raises(P.Panic):
        a.column(Fixed(3), usize, 4)._lower()


@pytest.mark.gpu
def test_element_aligned_output_buffer():
    """ADVICE medium: an output that is only element-aligned (a slice of a larger buffer) must not fault:
    the planner falls back to scalar stores (and the row-fold fast path is skipped)."""
    if not _gpu_ctx:
        _gpu_ctx.append(P.Context(0))
    ctx = _gpu_ctx[0]
    rng = np.random.default_rng(3)
    n = 4096
    a = Array.new(usize, n, rng.uniform(-1, 1, n).astype(np.float32)).to_device(ctx)
    b = Array.new(usize, n, rng.uniform(-1, 1, n).astype(np.float32)).to_device(ctx)
    big = Storage.device(ctx, F.F32, n + 64)
    for shift in (1, 3, 4):
        out = Storage.wrap_device(ctx, F.F32, n, big.dptr + 4 * shift, keep=big)
        got = (a * b + P.Scalar(1.0, "f32")).collect(out=out, ctx=ctx).as_ref()
        want = a.as_ref() * b.as_ref() + np.float32(1)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    m = Array.new((usize, usize), (64, 64), rng.uniform(0, 1, 4096).astype(np.float32)).to_device(ctx)
    out = Storage.wrap_device(ctx, F.F32, 64, big.dptr + 4, keep=big)
    got = P.fold_rows(m, usize, usize, P.Add, np.float32(0)).collect(out=out, ctx=ctx).as_ref()
    want = np.add.accumulate(m.as_ref().reshape(64, 64), axis=1, dtype=np.float32)[:, -1]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    out = Storage.wrap_device(ctx, F.F32, 4096, big.dptr + 4, keep=big)
    got = m.transpose((), usize, usize, ()).collect(out=out, ctx=ctx).as_ref()
    assert np.array_equal(got.reshape(64, 64), m.as_ref().reshape(64, 64).T)
    with pytest.raises(P.MdimError):  # not even element-aligned
        (a * b).collect(out=Storage.wrap_device(ctx, F.F32, n, big.dptr + 2, keep=big), ctx=ctx)
