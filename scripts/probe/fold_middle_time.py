#!/usr/bin/env python
"""Config 4's Array (1024, 1024, 256) f32 folded over each of its three axes through collect(): last axis (k_fold_rows), middle axis
(k_fold_cols, one column walk per outer coordinate), outermost axis (k_fold_cols) — and the evaluator (NO_FASTPATH) beside the two new ones."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage

torch.cuda.set_device(0)
ctx = P.Context(0)
P.set_default_context(ctx)
stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=0)
I, J, K = 1024, 1024, 256
bufs = [torch.empty(I * J * K, device="cuda", dtype=torch.float32).uniform_(0, 1) for _ in range(2)]   # two 1 GiB operands alternate: nothing is re-read from L2
out = torch.empty(I * J, device="cuda", dtype=torch.float32)
torch.cuda.synchronize()
U3 = (usize, usize, usize)


def views(buf):
    a = Array.from_device(U3, (I, J, K), buf.data_ptr(), "f32", ctx=ctx, keep=buf)
    return {"last axis": fold_rows(a, (usize, usize), usize, Add, np.float32(0)),
            "middle axis": fold_rows(a.transpose(usize, usize, usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0)),
            "outermost axis": fold_rows(a.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0))}


vs = [views(b) for b in bufs]
for name in vs[0]:
    for flags, tag in ((F.COLLECT_ASYNC, "collect()"), (F.COLLECT_ASYNC | F.COLLECT_NO_FASTPATH, "evaluator")):
        n_out = vs[0][name].len()
        runs = [v[name].prepare(out=Storage.wrap_device(ctx, F.F32, n_out, out.data_ptr(), keep=out), flags=flags).run for v in vs]
        for k in range(3):
            runs[k % 2]()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(10):
            runs[k % 2]()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"fold over the {name:15s} {tag:10s} -> {ctx.last_kernel():28s} {ms:.4f} ms  {(4 * I * J * K + 4 * n_out) / ms / 1e6:.0f} GB/s", flush=True)
