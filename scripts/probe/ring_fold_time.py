#!/usr/bin/env python
"""Times mdim_fold_sharded_axis (k_fold_ring) and mdim_fold_sharded_axis_blocked (k_fold_xchg) at world = 1 — the fold kernels
without peers — against collect() of the same fold (the planner's column walk, k_fold_cols; COLLECT_NO_FASTPATH would be the evaluator): (rows, 262144) f32, rows = 1024 / 512 / 128 (what 1 / 2 / 8
ranks hold of config 4).  Successive launches walk through different blocks of a 1 GiB buffer, so nothing is re-read from L2."""
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
    n_blk = 1024 // rows
    so = Storage.wrap_device(ctx, F.F32, C_, out.data_ptr(), keep=out)
    ref = torch.empty(C_, device="cuda", dtype=torch.float32)
    runs = {"k_fold_ring": [], "k_fold_xchg": [], "evaluator": []}
    for b in range(n_blk):
        ptr = big.data_ptr() + 4 * b * rows * C_
        st = Storage.wrap_device(ctx, F.F32, rows * C_, ptr, keep=big)
        a = Array.from_device((usize, usize), (rows, C_), ptr, "f32", ctx=ctx, keep=big)
        ev = fold_rows(a.transpose((), usize, usize, ()).iso((usize, usize)), usize, usize, Add, np.float32(0))
        runs["evaluator"].append(ev.prepare(out=Storage.wrap_device(ctx, F.F32, C_, ref.data_ptr(), keep=ref), flags=F.COLLECT_ASYNC).run)
        runs["k_fold_ring"].append(comm.prepare_fold_sharded_axis(st, rows, C_, Add, np.float32(0), out=so)[0])
        runs["k_fold_xchg"].append(comm.prepare_fold_sharded_axis(st, rows, C_, Add, np.float32(0), out=so, blocked=True)[0])
    for name, fns in runs.items():
        for k in range(3):
            fns[k % n_blk]()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 16
        e0.record(stream)
        for k in range(reps):
            fns[k % n_blk]()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        label = name if name != "evaluator" else f"collect() -> {ctx.last_kernel()}"
        print(f"rows {rows:5d} x {C_} f32  {label:26s} {ms:.4f} ms  {4 * rows * C_ / ms / 1e6:.0f} GB/s", flush=True)
    comm.fold_status()
    runs["evaluator"][0](); ctx.sync()
    for name in ("k_fold_ring", "k_fold_xchg"):
        out.zero_(); torch.cuda.synchronize()
        runs[name][0](); ctx.sync()
        assert torch.equal(out, ref), f"{name} differs from the evaluator's sequential fold"
print("bit-exact against the evaluator at every size")
