"""Pins oracle/ref_shaped.c (the rustc-shaped CPU baseline loops for the five BASELINE configs)
to the descriptor oracle and to the pure-Python reference model, which the doctests pin."""
import numpy as np

from helpers import oracle_collect, refshaped_lib, assert_same_bits
from multidimension_b200 import usize, Array, Scalar, Add, fold_rows
from oracle import reference_model as R


def ptr(a):
    return a.ctypes.data


def test_c2_zip_map():
    rng = np.random.default_rng(1)
    n = 10007
    a, b = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-1, 1, n).astype(np.float32)
    out = np.empty(n, np.float32)
    assert refshaped_lib().ref_c2_zip_map(ptr(a), ptr(b), n, ptr(out)) == 0
    v = Array.new(usize, n, a).zip(Array.new(usize, n, b)).map(lambda p: p[0] * p[1] + np.float32(1))
    assert_same_bits(out, oracle_collect(v))
    ra = R.Array.new(R.usize, 50, list(a[:50])).zip(R.Array.new(R.usize, 50, list(b[:50]))).map(lambda p: p[0] * p[1] + np.float32(1))
    assert_same_bits(out[:50], np.array(ra.collect().as_ref(), dtype=np.float32))


def test_c1_transpose():
    rng = np.random.default_rng(2)
    Y, X = 37, 53
    a = rng.uniform(-1, 1, Y * X).astype(np.float32)
    out = np.empty(Y * X, np.float32)
    assert refshaped_lib().ref_c1_transpose(ptr(a), Y, X, ptr(out)) == 0
    assert_same_bits(out, oracle_collect(Array.new((usize, usize), (Y, X), a).transpose((), usize, usize, ())))


def test_c3_compose():
    rng = np.random.default_rng(3)
    n, m = 5000, 777
    idx = rng.integers(0, m, n).astype(np.uint64)
    src = rng.uniform(-1, 1, m).astype(np.float32)
    out = np.empty(n, np.float32)
    assert refshaped_lib().ref_c3_compose(ptr(idx), n, ptr(src), m, ptr(out)) == 0
    assert_same_bits(out, oracle_collect(Array.new(usize, n, idx).compose(Array.new(usize, m, src))))


def test_c4_fold_and_subtract():
    rng = np.random.default_rng(4)
    I, J, K = 5, 7, 64
    a = rng.uniform(0, 1, I * J * K).astype(np.float32)
    sums = np.empty(I * J, np.float32)
    assert refshaped_lib().ref_c4_fold(ptr(a), I, J, K, ptr(sums)) == 0
    A = Array.new((usize, usize, usize), (I, J, K), a)
    s = fold_rows(A, (usize, usize), usize, Add, np.float32(0))
    assert_same_bits(sums, oracle_collect(s))
    mean = (sums / np.float32(K)).astype(np.float32)
    out = np.empty(I * J * K, np.float32)
    assert refshaped_lib().ref_c4_sub(ptr(a), ptr(mean), I, J, K, ptr(out)) == 0
    assert_same_bits(out, oracle_collect(A - (s / Scalar(float(K), "f32")).iso((usize, usize, ()))))
    rs = R.fold_rows(R.Array.new((R.usize, R.usize, R.usize), (I, J, K), list(a)), (R.usize, R.usize), R.usize,
                     lambda acc, x: acc + x, np.float32(0)).collect().as_ref()
    assert_same_bits(sums, np.array(rs, dtype=np.float32))


def test_c5_chain():
    rng = np.random.default_rng(5)
    P_, Q, Rn = 3, 4, 8
    a = rng.uniform(-1, 1, P_ * Q).astype(np.float32)
    w = rng.uniform(-1, 1, Rn).astype(np.float32)
    out = np.empty(Q * P_ * Q * P_ * Rn, np.float32)
    assert refshaped_lib().ref_c5_chain(ptr(a), P_, Q, ptr(w), Rn, ptr(out)) == 0
    v = (Array.new((usize, usize), (P_, Q), a).transpose((), usize, usize, ()).diagonal(np.float32(0))
         .iso((((usize, usize), (usize, usize)), ())).zip(Array.new(usize, Rn, w).iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))
    assert_same_bits(out, oracle_collect(v))
    u = R.usize
    rv = (R.Array.new((u, u), (P_, Q), list(a)).transpose((), u, u, ()).diagonal(np.float32(0))
          .iso((((u, u), (u, u)), ())).zip(R.Array.new(u, Rn, list(w)).iso(((), u))).map(lambda p: p[0] * p[1] + np.float32(1)))
    assert_same_bits(out, np.array(rv.collect().as_ref(), dtype=np.float32))
