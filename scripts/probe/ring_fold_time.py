#!/usr/bin/env python
"""Times mdim_fold_sharded_axis (k_fold_ring) at world = 1 — the pipelined fold kernel without peers — against the
evaluator's strided fold of the same data: (rows, 262144) f32, rows = 1024 / 512 / 128 (what 1 / 2 / 8 ranks hold of config 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage
from multidimension_b200.sharding import Comm

torch.cuda.set_device(0)
ctx = P.Context(0)
P.set_default_context(ctx)
stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=0)
comm = Comm(ctx, 0, 1, Comm.unique_id())
C_ = 1024 * 256
big = torch.empty(1024 * C_, device="cuda", dtype=torch.float32).uniform_(0, 1)
out = torch.empty(C_, device="cuda", dtype=torch.float32)
torch.cuda.synchronize()
for rows in (1024, 512, 128):
    st = Storage.wrap_device(ctx, F.F32, rows * C_, big.data_ptr(), keep=big)
    so = Storage.wrap_device(ctx, F.F32, C_, out.data_ptr(), keep=out)
    a = Array.from_device((usize, usize), (rows, C_), big.data_ptr(), "f32", ctx=ctx, keep=big)
    ev = fold_rows(a.transpose((), usize, usize, ()).iso((usize, usize)), usize, usize, Add, np.float32(0))
    ref = torch.empty(C_, device="cuda", dtype=torch.float32)
    pe = ev.prepare(out=Storage.wrap_device(ctx, F.F32, C_, ref.data_ptr(), keep=ref), flags=F.COLLECT_ASYNC)
    ring_run, _ = comm.prepare_fold_sharded_axis(st, rows, C_, Add, np.float32(0), out=so)
    for name, fn in (("k_fold_ring", ring_run), ("evaluator", pe.run)):
        for _ in range(3):
            fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            fn()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"rows {rows:5d} x {C_} f32  {name:12s} {ms:.4f} ms  {4 * rows * C_ / ms / 1e6:.0f} GB/s", flush=True)
    comm.fold_status()
    assert torch.equal(out, ref), "ring fold differs from the evaluator's sequential fold"
print("bit-exact against the evaluator at every size")
