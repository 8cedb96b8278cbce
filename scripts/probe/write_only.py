import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, all_, _ffi as F
from multidimension_b200.runtime import Storage
ctx = P.Context(0); P.set_default_context(ctx)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n = 1 << 29
to = torch.empty(n, device="cuda", dtype=torch.int64)
o = Storage.wrap_device(ctx, F.U64, n, to.data_ptr(), keep=to)
v = all_(usize, n)
prep = v.prepare(out=o, flags=F.COLLECT_ASYNC)
for _ in range(3): prep.run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10): prep.run()
e1.record(stream); torch.cuda.synchronize(); ctx.sync()
ms = e0.elapsed_time(e1) / 10
print("iota write-only", v.describe(), f"{ms:.4f} ms {8*n/ms/1e6:.0f} GB/s")
