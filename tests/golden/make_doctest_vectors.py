"""Writes tests/golden/doctests.json: the expected values of the reference crate's doctests.

The reference is Rust and no Rust toolchain exists in this image, so the vectors cannot be
*generated* by running the reference; they are TRANSCRIBED from the `assert_eq!` lines of the
reference's own doc comments (file:line given per vector; paths relative to the reference repo).
Run `python tests/golden/make_doctest_vectors.py` to regenerate the JSON from this table.
Pairs are JSON lists; `false/true` are JSON booleans.
"""
import json, os

V = [
    dict(name="rc_all_collect", ref="src/view.rs:91-96", expect=[0, 1, 2, 3, 4]),
    dict(name="collect_all", ref="src/view.rs:138-142", expect=[0, 1, 2, 3, 4]),
    dict(name="nested_index_mut", ref="src/view.rs:162-172", expect=["apple", "BODY", "crane"]),
    dict(name="nested_collect_behead", ref="src/view.rs:197-221",
         expect_a=[[False, 0], [False, 1], [False, 2], [False, 3], [True, 0], [True, 1], [True, 2], [True, 3]],
         expect=[[False, 1], [False, 2], [False, 3], [True, 1], [True, 2], [True, 3]]),
    dict(name="each_total", ref="src/view.rs:243-249", expect=10),
    dict(name="enumerate", ref="src/view.rs:257-266", expect=[[0, "apple"], [1, "body"], [2, "crane"]]),
    dict(name="diagonal", ref="src/view.rs:276-284", expect=[10, 0, 0, 0, 11, 0, 0, 0, 12]),
    dict(name="map_square", ref="src/view.rs:294-298", expect=[0, 1, 4, 9, 16]),
    dict(name="compose", ref="src/view.rs:307-313", expect=["crane", "body"]),
    dict(name="concat", ref="src/view.rs:320-326", expect=["apple", "body", "crane", "dump"]),
    dict(name="from_usize", ref="src/view.rs:346-351", items=["A", "a", "B", "b", "C", "c"],
         probe_b=[2, True, 0], probe_a=[2, 1, 0]),
    dict(name="to_usize", ref="src/view.rs:367-372", items=["A", "a", "B", "b", "C", "c"],
         probe_b=[2, 1, 0], probe_a=[2, True, 0]),
    dict(name="insert_one", ref="src/view.rs:384-390", expect=["A", "a", "B", "b"], expect_size=[[], 1, []]),
    dict(name="remove_one", ref="src/view.rs:401-407", expect=["A", "a", "B", "b"], expect_size=[[], []]),
    dict(name="map_axis", ref="src/view.rs:423-435", expect=["crane", "body", "CRANE", "BODY"]),
    dict(name="zip_same_shape", ref="src/view.rs:451-461", expect=[[0, "apple"], [1, "body"], [2, "crane"]]),
    dict(name="zip_scalar", ref="src/view.rs:464-474", expect=[[0, "repeated"], [1, "repeated"], [2, "repeated"]]),
    dict(name="zip_broadcast", ref="src/view.rs:477-487",
         expect=[[0, False], [0, True], [1, False], [1, True], [2, False], [2, True]]),
    dict(name="binary_add", ref="src/view.rs:499-506", expect=[19, 28, 37]),
    dict(name="coat_group_pairs", ref="src/view.rs:523-548",
         expect=["a", "b", "c", "d", "e", "f", "A", "B", "C", "D", "E", "F"], expect_size=[2, 3, []]),
    dict(name="transpose", ref="src/view.rs:572-585",
         expect_a=[[0, 0], [0, 1], [1, 0], [1, 1], [2, 0], [2, 1]],
         expect=[[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1]]),
    dict(name="row", ref="src/view.rs:596-608", expect=[[1, 0], [1, 1]]),
    dict(name="column", ref="src/view.rs:626-638", expect=[[0, 1], [1, 1], [2, 1]]),
    dict(name="fn_view", ref="src/view.rs:1431-1435",
         expect=[True, False, False, True, False, False, True, False, False, True]),
    dict(name="array_new_indexing", ref="src/array.rs:18-27", items=[0.0, 1.0, -1.0, 2.0, 3.0, -2.0],
         probes=[[[0, False], 0.0], [[0, True], 1.0], [[1, False], -1.0], [[1, True], 2.0], [[2, False], 3.0], [[2, True], -2.0]]),
    dict(name="array_from_fn", ref="src/array.rs:38-42",
         expect=[True, False, False, True, False, False, True, False, False, True]),
    dict(name="tuple_isomorphic", ref="src/tuple.rs:225-248",
         forms=[[1, [False, []], 2], [[1, []], [False, 2]], [[], [1, [False, 2]]], [[[1, False], 2], []]]),
    dict(name="tuple_push_pop", ref="src/tuple.rs:195-203",
         flat_of=[[False, [False]], [[], []], [[3], [3]], [[3, False], [3, False]], [[3, False, []], [3, False]]]),
    # README.md:43-50 repeats compose; README.md:33 (diagonal without `zero`) is stale vs src/view.rs:285.
]

if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "doctests.json")
    with open(path, "w") as f:
        json.dump({"source": "apt1002/multidimension 0.3.3 doc comments (transcribed)", "vectors": V}, f, indent=1)
    print("wrote", path, len(V), "vectors")
