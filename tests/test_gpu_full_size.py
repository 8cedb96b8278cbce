"""BASELINE.json's five configurations at FULL size on the GPU, checked through size-independent properties
(the oracle would need minutes per config at these sizes): exact closed forms, round trips, commutativity,
power-of-two linearity of the sequential fold, and element samples against numpy.  torch is plumbing here:
it allocates and fills device buffers and compares results on the device; every collect goes through the C ABI."""
import numpy as np
import pytest

import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, All, fold_rows, Add, _ffi as F
from multidimension_b200.runtime import Storage

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    torch.cuda.set_device(0)
    ctx = P.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    class Env:
        pass
    e = Env()
    e.torch, e.ctx = torch, ctx
    e.arr = lambda I, size, t, T: Array.from_device(I, size, t.data_ptr(), T, ctx=ctx, keep=t)
    e.out = lambda t, dt: Storage.wrap_device(ctx, dt, t.numel(), t.data_ptr(), keep=t)

    def run(view, t, dt=F.F32):
        view.collect(out=e.out(t, dt), ctx=ctx)
        ctx.sync()
        return t
    e.run = run
    yield e
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream())
    ctx.close()


def test_c1_transpose_4096_closed_form_and_round_trip(env):
    t, n = env.torch, 4096
    a = t.arange(n * n, device="cuda", dtype=t.float32)                        # a[k] = k, exact below 2^24
    v = env.arr((usize, usize), (n, n), a, "f32").transpose((), usize, usize, ())
    assert v.describe().startswith("transpose.tile")
    b = env.run(v, t.empty_like(a))
    k = t.arange(n * n, device="cuda", dtype=t.int64)
    assert t.equal(b, ((k % n) * n + k // n).to(t.float32))                      # out[x*n + y] = y*n + x
    back = env.run(env.arr((usize, usize), (n, n), b, "f32").transpose((), usize, usize, ()), t.empty_like(a))
    assert t.equal(back, a)                                                      # transpose is an involution
    bits = t.randint(-2**31, 2**31 - 1, (n * n,), device="cuda", dtype=t.int32)  # arbitrary bit patterns (NaNs included)
    tb = env.run(env.arr((usize, usize), (n, n), bits.view(t.float32), "f32").transpose((), usize, usize, ()), t.empty(n * n, device="cuda", dtype=t.float32))
    assert t.equal(tb.view(t.int32).view(n, n), bits.view(n, n).t())


def test_c2_zip_map_2_30_commutes_and_matches_numpy_samples(env):
    t, n = env.torch, 1 << 30
    g = t.Generator(device="cuda")
    g.manual_seed(0x5EED0001)
    a = t.empty(n, device="cuda", dtype=t.float32).uniform_(-1, 1, generator=g)
    b = t.empty(n, device="cuda", dtype=t.float32).uniform_(-1, 1, generator=g)
    A, B = env.arr(usize, n, a, "f32"), env.arr(usize, n, b, "f32")
    ab = env.run(A.zip(B).map(lambda p: p[0] * p[1] + np.float32(1)), t.empty_like(a))
    ba = env.run(B * A + Scalar(1.0, "f32"), t.empty_like(a))                    # the operator spelling, operands swapped
    assert t.equal(ab.view(t.int32), ba.view(t.int32))
    idx = t.randint(0, n, (1 << 20,), device="cuda", generator=g)
    idx[0], idx[1] = 0, n - 1
    ha, hb, hc = a[idx].cpu().numpy(), b[idx].cpu().numpy(), ab[idx].cpu().numpy()
    want = ha * hb + np.float32(1)                                               # two roundings, never fused (numpy has no FMA here)
    assert np.array_equal(hc.view(np.int32), want.view(np.int32))
    ones = env.run(A * Scalar(0.0, "f32") + Scalar(1.0, "f32"), t.empty_like(a))
    assert bool((ones == 1.0).all())                                             # finite x: x*0+1 == 1 exactly


def test_c3_compose_2_28_from_2_30_returns_its_indices_and_round_trips(env):
    t, n_src, n_idx = env.torch, 1 << 30, 1 << 28
    g = t.Generator(device="cuda")
    g.manual_seed(0x5EED0003)
    src = t.arange(n_src, device="cuda", dtype=t.int64)                          # src[k] = k  (usize source, 8 GiB)
    idx = t.randint(0, n_src, (n_idx,), device="cuda", dtype=t.int64, generator=g)
    got = env.run(env.arr(usize, n_idx, idx, usize).compose(env.arr(usize, n_src, src, usize)), t.empty_like(idx), F.U64)
    assert t.equal(got, idx)                                                     # gathering the identity returns the indices
    del src, got
    perm = t.randperm(n_idx, device="cuda", generator=g)
    inv = t.empty_like(perm)
    inv[perm] = t.arange(n_idx, device="cuda")
    x = t.empty(n_idx, device="cuda", dtype=t.float32).uniform_(-1, 1, generator=g)
    y = env.run(env.arr(usize, n_idx, perm, usize).compose(env.arr(usize, n_idx, x, "f32")), t.empty_like(x))
    z = env.run(env.arr(usize, n_idx, inv, usize).compose(env.arr(usize, n_idx, y, "f32")), t.empty_like(x))
    assert t.equal(z.view(t.int32), x.view(t.int32))                             # permute, then un-permute
    bad = idx.clone()
    bad[123456789] = n_src
    with pytest.raises(P.Panic, match=f"Index {n_src} is out of bounds for size {n_src}") as e:
        env.run(env.arr(usize, n_idx, bad, usize).compose(env.arr(usize, n_src, t.empty(n_src, device="cuda", dtype=t.float32), "f32")),
                t.empty(n_idx, device="cuda", dtype=t.float32))
    assert e.value.info.position == 123456789


def test_c4_fold_and_fused_subtract_1024x1024x256(env):
    t = env.torch
    I, J, K = 1024, 1024, 256
    g = t.Generator(device="cuda")
    g.manual_seed(0x5EED0004)
    a = t.empty(I * J * K, device="cuda", dtype=t.float32).uniform_(0, 1, generator=g)
    A = env.arr((usize, usize, usize), (I, J, K), a, "f32")

    def sums_of(arr):
        return fold_rows(arr.iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0))
    s = env.run(sums_of(A), t.empty(I * J, device="cuda", dtype=t.float32))
    # sequential order: a sample of rows against numpy's strictly sequential accumulate
    rows = [0, 1, 12345, I * J - 1]
    for r in rows:
        row = a[r * K:(r + 1) * K].cpu().numpy()
        assert np.float32(np.add.accumulate(row, dtype=np.float32)[-1]).view(np.int32) == s[r].cpu().numpy().view(np.int32)
    # power-of-two linearity: every partial sum of 4x is 4 times the partial sum of x, exactly
    a4 = a * 4
    s4 = env.run(sums_of(env.arr((usize, usize, usize), (I, J, K), a4, "f32")), t.empty_like(s))
    assert t.equal(s4.view(t.int32), (s * 4).view(t.int32))
    del a4, s4
    # closed form: rows of ones sum to exactly 256; mean-subtracting a constant row gives exactly 0
    ones = t.ones(I * J * K, device="cuda", dtype=t.float32)
    O = env.arr((usize, usize, usize), (I, J, K), ones, "f32")
    assert bool((env.run(sums_of(O), t.empty_like(s)) == float(K)).all())
    mean = (sums_of(O) / Scalar(float(K), "f32")).iso((usize, usize, ()))
    fused = O - mean
    assert "fold_rows.fused" in fused.describe()
    assert bool((env.run(fused, t.empty_like(ones)) == 0).all())
    del ones
    # the fused single pass equals the two-pass spelling (fold collected first), bit for bit
    meanA = (sums_of(A) / Scalar(float(K), "f32"))
    one_pass = env.run(A - meanA.iso((usize, usize, ())), t.empty_like(a))
    m = env.run(meanA, t.empty_like(s))
    two_pass = env.run(A - env.arr((usize, usize), (I, J), m, "f32").iso((usize, usize, ())), t.empty_like(a))
    assert t.equal(one_pass.view(t.int32), two_pass.view(t.int32))


def test_c5_rank5_chain_2_30_outputs(env):
    t = env.torch
    Pn = Qn = Rn = 64
    g = t.Generator(device="cuda")
    g.manual_seed(0x5EED0005)
    a = t.empty(Pn * Qn, device="cuda", dtype=t.float32).uniform_(0.5, 1, generator=g)   # no zeros: a*w+1 != 1 on the diagonal
    w = t.empty(Rn, device="cuda", dtype=t.float32).uniform_(0.5, 1, generator=g)
    v = (env.arr((usize, usize), (Pn, Qn), a, "f32").transpose((), usize, usize, ()).diagonal(np.float32(0))
         .iso((((usize, usize), (usize, usize)), ())).zip(env.arr(usize, Rn, w, "f32").iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))
    out = env.run(v, t.empty(1 << 30, device="cuda", dtype=t.float32))
    D = Pn * Qn
    m = out.view(D, D, Rn)                                                      # [(q,p), (q',p'), r]
    diag = m.diagonal(dim1=0, dim2=1).permute(1, 0)                            # [(q,p), r]
    want = a.view(Pn, Qn).t().reshape(D, 1) * w.view(1, Rn)
    want = want + 1.0                                                           # separate rounding steps, as the chain does
    assert t.equal(diag.contiguous().view(t.int32), want.view(t.int32))
    assert int((out != 1.0).sum().item()) == D * Rn                             # everything off the diagonal is 0*w+1 == 1
