import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, _ffi as F
from multidimension_b200.runtime import Storage
from helpers import oracle_collect, assert_same_bits
torch.cuda.set_device(0)
ctx = P.Context(0); P.set_default_context(ctx)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n = 1 << 28
ta, tb, tc = (torch.empty(n, device="cuda", dtype=torch.float32).uniform_(-1, 1) for _ in range(3))
to = torch.empty(n, device="cuda", dtype=torch.float32)
dev = lambda t: Array.from_device(usize, n, t.data_ptr(), "f32", ctx=ctx, keep=t)
v = (dev(ta) * dev(tb) - dev(tc)).map(P.Abs)   # no pre-built signature
print(v.describe())
o = Storage.wrap_device(ctx, F.F32, n, to.data_ptr(), keep=to)
for flags, name in ((0, "jit"), (F.COLLECT_NO_JIT, "interp")):
    t0 = time.perf_counter(); v.collect(out=o, flags=flags); first = time.perf_counter() - t0
    prep = v.prepare(out=o, flags=flags | F.COLLECT_ASYNC)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): prep.run()
    e1.record(stream); torch.cuda.synchronize(); ctx.sync()
    ms = e0.elapsed_time(e1) / 10
    ok = torch.equal(to, (ta * tb - tc).abs())
    print(f"{name}: first call {first*1e3:.1f} ms, steady {ms:.4f} ms = {16*n/ms/1e6:.0f} GB/s, bit-exact {ok}")
# small parity check of a few odd chains through the jit against the oracle
rng = np.random.default_rng(0)
a = Array.new((usize, usize), (37, 24), rng.integers(0, 1000, 37 * 24).astype(np.uint64))
b = Array.new(usize, 24, rng.integers(1, 9, 24).astype(np.uint64))
for view in ((a % b.iso(((), usize))) ^ Scalar(5), a.transpose((), usize, usize, ()).map(P.Cast("f32")).map(P.Sqrt), (a >> Scalar(3)).diagonal(7)):
    assert_same_bits(view.collect(location="device").as_ref(), oracle_collect(view), view.describe())
    print("ok", view.describe())
