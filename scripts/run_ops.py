#!/usr/bin/env python
"""Runs each BASELINE config a few times, device-resident, for ncu captures and quick timings.
usage: python scripts/run_ops.py [--ops c1,c2,c3,c4a,c4c,c4b,c5] [--reps 3] [--time]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage

ap = argparse.ArgumentParser()
ap.add_argument("--ops", default="c1,c2,c3,c4a,c4c,c4b,c5")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--time", action="store_true")
ap.add_argument("--flags", type=int, default=0)
args = ap.parse_args()
ops = args.ops.split(",")

torch.cuda.set_device(0)
ctx = P.Context(0)
P.set_default_context(ctx)
if os.environ.get("MDIM_OPS_TORCH_STREAM") == "1":  # a caller-owned stream: every kernel waits for its predecessor
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    torch.cuda.set_stream(stream)
else:  # the context's own stream (dependency-aware launches), wrapped only so that torch events can be recorded on it;
    # torch's own kernels stay on torch's stream, separated from the collects by device-wide synchronisation
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=0)


def dev(I, size, t, T):
    return Array.from_device(I, size, t.data_ptr(), T, ctx=ctx, keep=t)


def out(t, dt=F.F32):
    return Storage.wrap_device(ctx, dt, t.numel(), t.data_ptr(), keep=t)


def run(name, view, o, alg_bytes):
    print(name, view.describe(args.flags), flush=True)
    torch.cuda.synchronize()
    view.collect(out=o, flags=args.flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    prep = view.prepare(out=o, flags=args.flags | F.COLLECT_ASYNC)
    if args.reps:
        for _ in range(3):  # warm-up: a chain seen twice is specialised for its shape (NVRTC, ~0.1 s)
            prep.run()
        ctx.sync()
        e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        prep.run()
    host_us = (time.perf_counter() - t0) / max(args.reps, 1) * 1e6
    e1.record(stream)
    torch.cuda.synchronize()
    ctx.sync()
    if args.time:
        ms = e0.elapsed_time(e1) / args.reps
        print(f"  {name}: {ms:.4f} ms  {alg_bytes / ms / 1e6:.1f} GB/s  (host {host_us:.1f} us/launch)", flush=True)


n = 1 << 30
big = torch.empty(n, device="cuda", dtype=torch.float32).uniform_(-1, 1)
tout = torch.empty(n, device="cuda", dtype=torch.float32)
if "c2" in ops:
    tb = torch.empty(n, device="cuda", dtype=torch.float32).uniform_(-1, 1)
    run("c2", dev(usize, n, big, "f32").zip(dev(usize, n, tb, "f32")).map(lambda p: p[0] * p[1] + np.float32(1)), out(tout), 12 * n)
    del tb
if "c1" in ops:
    m = 4096
    # a fresh region per call would need rotation; for ncu (cold cache, one launch) one pair is enough
    run("c1", dev((usize, usize), (m, m), big[: m * m], "f32").transpose((), usize, usize, ()), out(tout[: m * m]), 8 * m * m)
    m = 16384
    run("c1_16k", dev((usize, usize), (m, m), big[: m * m], "f32").transpose((), usize, usize, ()), out(tout[: m * m]), 8 * m * m)
if "c1rot" in ops and args.reps:
    # 4096^2 on 16 rotating buffer pairs (2 GiB touched per round, > L2), the way bench.py times config 1
    m = 4096
    preps = [dev((usize, usize), (m, m), big[k * m * m: (k + 1) * m * m], "f32").transpose((), usize, usize, ())
             .prepare(out=out(tout[k * m * m: (k + 1) * m * m]), flags=args.flags | F.COLLECT_ASYNC) for k in range(16)]
    torch.cuda.synchronize()
    for p_ in preps:
        p_.run()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for r in range(args.reps):
        for p_ in preps:
            p_.run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (16 * args.reps)
    print(f"  c1rot: {ms:.4f} ms  {8 * m * m / ms / 1e6:.1f} GB/s", flush=True)
if "c1peer" in ops:
    # the peer-mapped transpose kernel on ONE GPU: 8 "peers" that are consecutive blocks of the same buffer (what the
    # instruction overhead of the owner lookup costs when NVLink is not the limit)
    from multidimension_b200.sharding import PeerStorage
    m, world = 16384, 8
    blk = m * m // world
    peers = PeerStorage(F.F32, m * m, [big.data_ptr() + 4 * blk * p_ for p_ in range(world)], blk, keep=big, ctx=ctx)
    run("c1peer", Array((usize, usize), (m, m), peers, "f32").transpose((), usize, usize, ()), out(tout[: m * m]), 8 * m * m)
if "c1f64" in ops:
    m = 8192
    run("c1f64", dev((usize, usize), (m, m), big.view(torch.float64)[: m * m], "f64").transpose((), usize, usize, ()),
        out(tout.view(torch.float64)[: m * m], F.F64), 16 * m * m)
if "c3" in ops:
    n3 = 1 << 28
    tidx = torch.randint(0, n, (n3,), device="cuda", dtype=torch.int64)
    run("c3", dev(usize, n3, tidx, usize).compose(dev(usize, n, big, "f32")), out(tout[:n3]), 16 * n3)
    del tidx
shape = (1024, 1024, 256)
n4 = 1 << 28
a4 = dev((usize, usize, usize), shape, big[:n4], "f32")
sums = fold_rows(a4, (usize, usize), usize, Add, np.float32(0))
tsum = torch.empty(1 << 20, device="cuda", dtype=torch.float32)
if "c4a" in ops:
    run("c4a", sums, out(tsum), 4 * n4 + (4 << 20))
if "c4c" in ops:
    run("c4c", a4 - (sums / Scalar(256.0, "f32")).iso((usize, usize, ())), out(tout[:n4]), 8 * n4)
if "c4b" in ops:
    run("c4b", a4 - dev((usize, usize), shape[:2], tsum, "f32").iso((usize, usize, ())), out(tout[:n4]), 8 * n4 + (4 << 20))
if "c5" in ops:
    ta5 = torch.empty(64 * 64, device="cuda", dtype=torch.float32).uniform_(-1, 1)
    tw5 = torch.empty(64, device="cuda", dtype=torch.float32).uniform_(-1, 1)
    v5 = (dev((usize, usize), (64, 64), ta5, "f32").transpose((), usize, usize, ()).diagonal(np.float32(0))
          .iso((((usize, usize), (usize, usize)), ())).zip(dev(usize, 64, tw5, "f32").iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))
    run("c5", v5, out(tout), 4 * n)
print("done")
