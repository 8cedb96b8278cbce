"""The reference's own doctests (tests/golden/doctests.json, transcribed with file:line) replayed
against the PRODUCT's host mirror.  On a CPU box the lowered descriptor is executed by the two
checkers (C oracle; host build of the product's planner+evaluator); with `-m gpu` the same
expressions run on the sm_100a kernels through the C ABI.

`&str` elements have no device representation; they are replaced by their position in a
vocabulary (a usize), which preserves every index-manipulation the doctests check.
"""
import numpy as np
import pytest

from conftest import untuple
from helpers import oracle_collect, emu_collect, as_list
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, all_, All
from multidimension_b200.view import Enumerate

VOCAB = ["apple", "body", "crane", "dump", "APPLE", "BODY", "CRANE", "A", "a", "B", "b", "C", "c", "repeated",
         "d", "e", "f", "D", "E", "F"]


def enc(words):
    return [VOCAB.index(w) for w in words]


def dec(x):
    if isinstance(x, (tuple, list)):
        return tuple(dec(y) for y in x)
    return x


def words(codes):
    return [VOCAB[int(c)] for c in codes]


def exp(golden, name, key="expect"):
    return [untuple(x) for x in golden[name][key]]


def gpu_collect(view):
    a = view.collect()
    return a.as_ref()


EXECUTORS = [
    pytest.param(oracle_collect, id="oracle"),
    pytest.param(emu_collect, id="emu"),
    pytest.param(gpu_collect, id="gpu", marks=pytest.mark.gpu),
]


@pytest.fixture(params=EXECUTORS)
def run(request):
    return request.param


def test_collect_all(golden, run):  # src/view.rs:138-142
    assert as_list(run(all_(usize, 5))) == exp(golden, "collect_all")


def test_diagonal(golden, run):  # src/view.rs:276-284
    v = all_(usize, 3).map(lambda x: x + 10).diagonal(0)
    assert v.size() == (3, 3)
    assert as_list(run(v)) == exp(golden, "diagonal")


def test_map_square(golden, run):  # src/view.rs:294-298
    assert as_list(run(all_(usize, 5).map(lambda x: x * x))) == exp(golden, "map_square")


def test_compose(golden, run):  # src/view.rs:307-313
    a = Array.new(bool, (), [2, 1])
    b = Array.new(usize, 3, enc(["apple", "body", "crane"]))
    ab = a.compose(b)
    assert ab.I is bool
    assert words(run(ab)) == exp(golden, "compose")


def test_compose_out_of_bounds(run):  # src/int.rs:17 panic text
    a = Array.new(bool, (), [2, 3])
    b = Array.new(usize, 3, enc(["apple", "body", "crane"]))
    with pytest.raises(Exception, match="Index 3 is out of bounds for size 3"):
        run(a.compose(b))


def test_concat(golden, run):  # src/view.rs:320-326
    a = Array.new(usize, 2, enc(["apple", "body"]))
    b = Array.new(usize, 2, enc(["crane", "dump"]))
    ab = a.concat(b, (), ()).iso(usize)
    assert ab.size() == 4
    assert words(run(ab)) == exp(golden, "concat")
    with pytest.raises(P.Panic):  # src/view.rs:336-337: outer / inner sizes must agree
        Array.new((usize, usize), (2, 3), list(range(6))).concat(Array.new((usize, usize), (2, 4), list(range(8))), (), usize)


def test_from_usize_to_usize(golden, run):  # src/view.rs:346-351, 367-372
    g = golden["from_usize"]
    a = Array.new((usize, usize, usize), (3, 2, 1), enc(g["items"]))
    b = a.from_usize(usize, bool, usize, lambda _: ())
    assert b.I == (usize, bool, usize) and b.size() == (3, (), 1)
    assert words(run(b)) == g["items"]  # same positions: (2,true,0) <-> (2,1,0)
    a2 = Array.new((usize, bool, usize), (3, (), 1), enc(g["items"]))
    b2 = a2.to_usize(usize, bool, usize)
    assert b2.I == (usize, usize, usize) and b2.size() == (3, 2, 1)
    assert words(run(b2)) == g["items"]


def test_insert_remove_one(golden, run):  # src/view.rs:384-390, 401-407
    g = golden["insert_one"]
    a = Array.new((bool, bool), ((), ()), enc(g["expect"]))
    b = a.insert_one(bool, usize, bool, 1)
    assert b.size() == untuple(g["expect_size"])
    assert words(run(b)) == g["expect"]
    g = golden["remove_one"]
    a = Array.new((bool, usize, bool), ((), 1, ()), enc(g["expect"]))
    b = a.remove_one(bool, usize, bool)
    assert b.size() == untuple(g["expect_size"])
    assert words(run(b)) == g["expect"]
    with pytest.raises(P.Panic):  # src/view.rs:395
        a.insert_one(bool, usize, (usize, bool), 2)


def test_map_axis(golden, run):  # src/view.rs:423-435
    a = Array.new(usize, 2, [2, 1])
    b = Array.new((bool, usize), 3, enc(["apple", "body", "crane", "APPLE", "BODY", "CRANE"]))
    ab = b.map_axis(a, bool, ())
    assert ab.I == (bool, usize, ())
    assert words(run(ab)) == exp(golden, "map_axis")


def test_zip_pairs(golden, run):  # src/view.rs:451-461, 464-474, 477-487
    a = all_(usize, 3)
    b = Array.new(usize, 3, enc(["apple", "body", "crane"]))
    got = run(a.zip(b))
    assert [(i, VOCAB[w]) for i, w in got] == exp(golden, "zip_same_shape")
    got = run(a.zip(Scalar(VOCAB.index("repeated"))))
    assert [(i, VOCAB[w]) for i, w in got] == exp(golden, "zip_scalar")
    c = all_((usize, ()), 3).zip(all_(((), bool), ()))
    assert c.I == (usize, bool) and c.size() == (3, ())
    got = run(c)  # elements are ((usize, ()), ((), bool))
    assert [(x[0][0], x[1][1]) for x in got] == exp(golden, "zip_broadcast")


def test_zip_unequal_sizes():  # src/broadcast.rs:38
    with pytest.raises(P.Panic, match="Unequal sizes"):
        all_(usize, 3).zip(all_(usize, 4))
    with pytest.raises(TypeError):  # () does not implement Broadcast<()>
        Scalar(1).zip(Scalar(2))


def test_binary_add(golden, run):  # src/view.rs:499-506
    a = Array.new(usize, 3, [9, 8, 7])
    b = Array.new(usize, 3, [10, 20, 30])
    assert as_list(run(a.binary(b, P.Add))) == exp(golden, "binary_add")
    assert as_list(run(a + b)) == exp(golden, "binary_add")


def test_transpose(golden, run):  # src/view.rs:572-585
    a = all_((usize, usize), (3, 2))
    assert as_list(run(a)) == exp(golden, "transpose", "expect_a")
    b = a.transpose((), usize, usize, ())
    assert b.size() == ((), (2, 3), ())
    assert as_list(run(b)) == exp(golden, "transpose")


def test_row_column(golden, run):  # src/view.rs:596-608, 626-638
    a = all_((usize, usize), (3, 2))
    assert as_list(run(a.row(usize, usize, 1))) == exp(golden, "row")
    assert as_list(run(a.column(usize, usize, 1))) == exp(golden, "column")


def test_enumerate(golden, run):  # src/view.rs:257-266
    a = Array.new(usize, 3, enc(["apple", "body", "crane"]))
    got = run(a.enumerate())
    assert [(i, VOCAB[w]) for i, w in got] == exp(golden, "enumerate")


def test_coat_group_pairs(golden, run):  # src/view.rs:523-548
    g = golden["coat_group_pairs"]
    a = Array.new((usize, usize), (2, 6), enc(g["expect"]))
    # group_pairs: (I, usize) -> (I, (usize, bool)) coated as one axis; data stays in place
    b = a.from_usize(usize, (usize, bool), (), lambda n: (n // 2, ()))
    assert b.size() == (2, (3, ()), ())
    assert words(run(b)) == g["expect"]
    c = b.coat(b.I)
    assert isinstance(c.I, P.Coated)
    assert words(run(c)) == g["expect"]


def test_fn_view_from_fn(golden, run):  # src/view.rs:1431-1435, src/array.rs:38-42 (closure x % 3 == 0 as ops)
    v = all_(usize, 10).map(lambda x: (2 - x % 3) // 2)  # 1 iff x % 3 == 0, from the op vocabulary
    assert [bool(x) for x in run(v)] == golden["fn_view"]["expect"]


def test_array_new_indexing(golden):  # src/array.rs:18-27
    g = golden["array_new_indexing"]
    a = Array.new((usize, bool), 3, np.array(g["items"], dtype=np.float64))
    for idx, want in g["probes"]:
        assert a[untuple(idx)] == want
    with pytest.raises(P.Panic):  # src/array.rs:12
        Array.new((usize, bool), 3, np.array(g["items"][:-1]))
    with pytest.raises(P.Panic, match="Index 3 is out of bounds for size 3"):  # src/int.rs:17
        a[(3, False)]


def test_each_total(golden, run):  # src/view.rs:243-249: each() visits every element once, in order
    total = run(P.fold_rows(all_(usize, 5).map(lambda x: x + 0).iso(((), usize)), (), usize, P.Add, 0))
    assert as_list(total) == [golden["each_total"]["expect"]]


def test_remaining_leaf_index_kinds(run):  # Reversed, Fixed<N>, bool, Option<I> (src/int.rs:33-86, src/index.rs:236-276)
    from multidimension_b200 import Reversed, Fixed, Option, Some
    from oracle import reference_model as R
    items = list(range(100, 112))
    a = Array.new((Option(usize), Fixed(3)), (3, ()), items)          # Option<usize> of size 3 has 1 + 3 positions
    assert a.len() == 12 and a.size() == (3, ())
    ra = R.Array.new((R.Option(R.usize), R.Fixed(3)), (3, ()), items)  # the pinned reference model agrees on every position
    for i, ri in ((None, None), (Some(0), R.Some(0)), (Some(2), R.Some(2))):
        for j in range(3):
            assert a[(i, j)] == ra[(ri, j)]
    t = a.transpose((), Fixed(3), Option(usize), ())
    assert as_list(run(t)) == [items[p * 3 + f] for f in range(3) for p in range(4)]
    assert as_list(run(a.row(Option(usize), Fixed(3), Some(1)))) == items[6:9]
    assert as_list(run(a.column(Option(usize), Fixed(3), 2))) == items[2::3]
    r = all_(Reversed, 4)                                              # position p holds Reversed(size-1-p), src/int.rs:82-84
    assert as_list(run(r)) == [3, 2, 1, 0]
    b = Array.new((bool, Reversed), ((), 3), [1, 2, 3, 4, 5, 6])
    assert b[(True, 0)] == 6 and b[(False, 2)] == 1                   # Reversed(i).to_usize(size) = size-1-i, src/int.rs:74
    assert as_list(run(b.row(bool, Reversed, True))) == [4, 5, 6]
