#!/usr/bin/env python
"""Latency floor of the two fused fold-over-the-sharded-axis kernels on N GPUs: `python fold_exchange_probe.py <world>` spawns one
process per GPU (C-ABI communicator, no torch.distributed) and times k_fold_xchg (blocked, in-kernel all-reduce) and k_fold_ring
(bit-exact chain) for rows-per-rank = 8 ... 512 x 2^18 f32 columns: with 8 rows the fold itself is ~2 us, so the time IS the exchange."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def worker(rank, world, path):
    import numpy as np
    import torch
    import multidimension_b200 as P
    from multidimension_b200 import Add, _ffi as F
    from multidimension_b200.runtime import Storage
    from multidimension_b200.sharding import Comm
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    torch.cuda.set_device(rank)
    ctx = P.Context(rank)
    comm = Comm.from_env(ctx, path=path) if world > 1 else Comm(ctx, 0, 1, Comm.unique_id())
    stream = torch.cuda.ExternalStream(ctx.stream_handle(), device=rank)
    C_ = 1 << 18
    big = torch.empty(512 * C_, device="cuda", dtype=torch.float32).uniform_(0, 1)
    out = torch.empty(C_, device="cuda", dtype=torch.float32)
    so = Storage.wrap_device(ctx, F.F32, C_, out.data_ptr(), keep=out)
    torch.cuda.synchronize()
    row_list = tuple(int(x) for x in os.environ.get("PROBE_ROWS", "8,32,128,256,512").split(","))
    kernels = os.environ.get("PROBE_KERNELS", "xchg,ring").split(",")
    for rows in row_list:
        n_blk = max(1, 512 // rows)
        for name, blocked in (("k_fold_xchg", True), ("k_fold_ring", False)):
            if name.split("_")[-1] not in kernels:
                continue
            fns = []
            for b in range(min(n_blk, 8)):
                st = Storage.wrap_device(ctx, F.F32, rows * C_, big.data_ptr() + 4 * b * rows * C_, keep=big)
                fns.append(comm.prepare_fold_sharded_axis(st, rows, C_, Add, np.float32(0), out=so, blocked=blocked)[0])
            for k in range(3):
                fns[k % len(fns)]()
            ctx.sync(); comm.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 32
            import time
            e0.record(stream)
            t0 = time.perf_counter()
            for k in range(reps):
                fns[k % len(fns)]()
            host_us = (time.perf_counter() - t0) / reps * 1e6
            e1.record(stream)
            ctx.sync(); torch.cuda.synchronize(); comm.barrier()
            comm.fold_status()
            print(f"world {world} rank {rank} rows/rank {rows:4d} {name:12s} {e0.elapsed_time(e1) / reps * 1000:8.1f} us  (host: {host_us:.1f} us per call)\n", end="", flush=True)
    comm.close_peers(); comm.close(); ctx.close()


if __name__ == "__main__":
    if len(sys.argv) == 2:
        world = int(sys.argv[1])
        path = os.path.join(tempfile.mkdtemp(), "rdv")
        procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), str(r), str(world), path]) for r in range(world)]
        sys.exit(max(p.wait() for p in procs))
    worker(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])
