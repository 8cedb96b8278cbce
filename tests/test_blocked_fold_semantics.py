"""What `mdim_fold_sharded_axis_blocked` (csrc/k_fold_xchg.cu) promises, checked on the CPU restatement the GPU tests compare it with
(oracle/reference_model.py::blocked_fold_over_sharded_axis): per-rank sequential partial folds combined in rank order are
  * the reference's sequential fold, bit for bit, for integer and bitwise operators (any world size, any split of the rows);
  * for an f32 sum: as ACCURATE as the reference's own order (error against the exact sum no worse than the sequential chain's, which is
    itself ~1e-6 away from the exact sum over 1024 terms), and within the north star's 1e-6 of the sequential chain over a few hundred terms;
  * sign-of-zero exact: the ranks after the first start from -0.0, the only x with x + y == y for EVERY y."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from reference_model import blocked_fold_over_sharded_axis  # noqa: E402

from helpers import assert_same_bits  # noqa: E402


def sequential(rows, op, init):
    s = np.full(rows.shape[1], init, dtype=rows.dtype)
    for i in range(rows.shape[0]):
        s = op(s, rows[i])
    return s


def split(rows, world):
    per = rows.shape[0] // world
    return [rows[r * per:(r + 1) * per] for r in range(world)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_integer_folds_are_the_sequential_fold(world):
    rng = np.random.default_rng(world)
    rows = rng.integers(0, 1 << 63, (24 * world, 40)).astype(np.uint64)
    with np.errstate(over="ignore"):
        for op, init, identity in ((np.add, 7, 0), (np.multiply, 3, 1), (np.bitwise_xor, 5, 0), (np.bitwise_and, (1 << 64) - 1, (1 << 64) - 1), (np.bitwise_or, 0, 0)):
            init, identity = np.uint64(init), np.uint64(identity)
            assert_same_bits(blocked_fold_over_sharded_axis(split(rows, world), op, init, identity), sequential(rows, op, init), op.__name__)
        i32 = rng.integers(-2**31, 2**31, (8 * world, 12)).astype(np.int32)
        assert_same_bits(blocked_fold_over_sharded_axis(split(i32, world), np.add, np.int32(-9), np.int32(0)), sequential(i32, np.add, np.int32(-9)), "i32 wrapping add")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_f32_sum_is_as_accurate_as_the_reference_order_and_exact_at_one_rank(world):
    rng = np.random.default_rng(10 + world)
    rows = rng.uniform(0, 1, (128 * world, 64)).astype(np.float32)
    seq = sequential(rows, np.add, np.float32(0.5))
    blk = blocked_fold_over_sharded_axis(split(rows, world), np.add, np.float32(0.5), np.float32(-0.0))
    exact = rows.astype(np.float64).sum(axis=0) + 0.5
    err_seq = np.max(np.abs(seq - exact) / exact)
    err_blk = np.max(np.abs(blk - exact) / exact)
    assert err_blk <= max(1.5 * err_seq, 2e-6), (err_blk, err_seq)   # the criterion bench.py holds the all-reduce routes to
    few = rows[:32 * world]
    assert np.max(np.abs(blocked_fold_over_sharded_axis(split(few, world), np.add, np.float32(0.5), np.float32(-0.0)) - sequential(few, np.add, np.float32(0.5)))
                  / sequential(few, np.add, np.float32(0.5))) <= 1e-6
    assert_same_bits(blocked_fold_over_sharded_axis([rows], np.add, np.float32(0.5), np.float32(-0.0)), seq, "one rank")


def test_negative_zero_is_the_identity_of_a_float_sum():
    rows = np.full((4, 6), np.float32(-0.0))
    want = sequential(rows, np.add, np.float32(-0.0))            # -0.0 + -0.0 = -0.0
    assert np.all(np.signbit(want))
    got = blocked_fold_over_sharded_axis(split(rows, 2), np.add, np.float32(-0.0), np.float32(-0.0))
    assert_same_bits(got, want, "blocks of -0.0 from a -0.0 start")
    wrong = blocked_fold_over_sharded_axis(split(rows, 2), np.add, np.float32(-0.0), np.float32(0.0))
    assert not np.any(np.signbit(wrong)), "+0.0 as the identity would flip the sign: that is why the kernel starts from -0.0"
