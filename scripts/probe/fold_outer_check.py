import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage
torch.cuda.set_device(0)
ctx = P.Context(0); P.set_default_context(ctx)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for (I, J, K) in [(16, 32, 64), (64, 256, 64), (512, 1024, 256), (512, 64, 64), (8, 1024, 256)]:
    t = torch.empty(I * J * K, device="cuda", dtype=torch.float32).uniform_(0, 1)
    a = Array.from_device((usize, usize, usize), (I, J, K), t.data_ptr(), "f32", ctx=ctx, keep=t)
    v = fold_rows(a.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0))
    tp = torch.empty(J * K, device="cuda", dtype=torch.float32)
    v.collect(out=Storage.wrap_device(ctx, F.F32, J * K, tp.data_ptr(), keep=tp))
    torch.cuda.synchronize()
    ref = t.view(I, J * K).double().sum(dim=0)
    rel = ((tp.double() - ref).abs() / ref.abs())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o = Storage.wrap_device(ctx, F.F32, J * K, tp.data_ptr(), keep=tp)
    for _ in range(3):
        v.collect(out=o, flags=F.COLLECT_ASYNC)
    ctx.sync()
    e0.record(stream)
    for _ in range(5):
        v.collect(out=o, flags=F.COLLECT_ASYNC)
    e1.record(stream); torch.cuda.synchronize(); ctx.sync()
    ms = e0.elapsed_time(e1) / 5
    print(f"{ms:.4f} ms {4 * I * J * K / ms / 1e6:.0f} GB/s", end=" ")
    print((I, J, K), v.describe(), "max rel", rel.max().item(), "bad", int((rel > 1e-5).sum()), "first bad", int(torch.nonzero(rel > 1e-5)[0]) if (rel > 1e-5).any() else -1)
