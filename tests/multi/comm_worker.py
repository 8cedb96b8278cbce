"""One rank of the multi-GPU C-ABI test (tests/test_comm_multi_gpu.py): `python comm_worker.py <rank> <world> <rendezvous file> <out dir>`.
Everything multi-GPU goes through include/mdim.h (mdim_comm_init / mdim_allgather / mdim_allreduce / mdim_peer_table);
no torch, no torch.distributed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

import multidimension_b200 as P
from multidimension_b200 import usize, Array, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage
from multidimension_b200.sharding import Comm, PeerStorage, equal_block, shard_bounds, shard_view, peer_source, all_gather_source, all_reduce_partial


def main():
    rank, world, path, out_dir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    ctx = P.Context(rank)
    P.set_default_context(ctx)
    comm = Comm.from_env(ctx, path=path)
    info = comm.info()
    assert info["rank"] == rank and info["world"] == world and info["nccl_version"] > 0
    rng = np.random.default_rng(2024)  # the same global data on every rank; each uploads only its own block
    out = {}

    # (1) all-gather of a sharded compose() source, then a local gather (what the north star names)
    n_src = 1 << 16
    src = rng.uniform(-1, 1, n_src).astype(np.float32)
    block = n_src // world
    mine = Array.new(usize, block, src[rank * block:(rank + 1) * block]).to_device(ctx)
    full = all_gather_source(mine, comm)
    idx = rng.integers(0, n_src, 5000).astype(np.uint64)
    lo, hi = shard_bounds(idx.size, world, rank)
    out["ag_gather"] = Array.new(usize, hi - lo, idx[lo:hi]).compose(Array(usize, n_src, full, "f32")).collect(location="device", ctx=ctx).as_ref()

    # (2) the same gather with the source peer-mapped: the block sits at an INTERIOR offset of a larger allocation
    #     (ADVICE low: cudaIpcGetMemHandle names the allocation, the offset must travel with it)
    slab = Storage.device(ctx, F.F32, block + 1000)
    local = Storage.wrap_device(ctx, F.F32, block, slab.dptr + 4 * 1000, keep=slab)
    ctx.upload(local.dptr, src[rank * block:(rank + 1) * block])
    comm.barrier()
    peers = peer_source(local, n_src, comm)
    out["peer_gather"] = Array.new(usize, hi - lo, idx[lo:hi]).compose(Array(usize, n_src, peers, "f32")).collect(location="device", ctx=ctx).as_ref()

    # (3) transpose of a row-sharded matrix: this rank's block of the transposed rows, tiles read from the owner
    M, N = 64 * world * 2, 384
    mat = rng.uniform(-1, 1, M * N).astype(np.float32)
    tb = M * N // world
    tlocal = Storage.device(ctx, F.F32, tb)
    ctx.upload(tlocal.dptr, mat[rank * tb:(rank + 1) * tb])
    comm.barrier()
    tpeers = PeerStorage(F.F32, M * N, comm.peer_table(tlocal.dptr, tlocal.nbytes), tb, keep=tlocal, ctx=ctx)
    tview = shard_view(Array((usize, usize), (M, N), tpeers, "f32").transpose((), usize, usize, ()), rank, world)
    out["peer_transpose"] = tview.collect(location="device", ctx=ctx).as_ref()
    out["peer_transpose_kernel"] = np.array([ord(c) for c in str(tview.describe())], dtype=np.uint8)
    out["peer_transpose_notma"] = tview.collect(location="device", ctx=ctx, flags=F.COLLECT_NO_FASTPATH).as_ref()

    # (4) fold over the SHARDED axis: per-rank partial + mdim_allreduce (1e-6), and bit-exact through the peer table
    I, J, K = 16 * world, 24, 32
    a = rng.uniform(0, 1, I * J * K).astype(np.float32)
    ib = I // world
    A = Array.new((usize, usize, usize), (ib, J, K), a[rank * ib * J * K:(rank + 1) * ib * J * K]).to_device(ctx)
    part = fold_rows(A.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0)).collect(location="device", ctx=ctx)
    all_reduce_partial(part.storage, "sum", comm)
    out["fold_allreduce"] = part.as_ref()
    fpeers = PeerStorage(F.F32, I * J * K, comm.peer_table(A.storage.dptr, A.storage.nbytes), ib * J * K, keep=A, ctx=ctx)
    whole = Array((usize, usize, usize), (I, J, K), fpeers, "f32")
    exact = shard_view(fold_rows(whole.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0)), rank, world)
    out["fold_exact"] = exact.collect(location="device", ctx=ctx).as_ref()

    # (4') the same fold as ONE fused kernel per GPU, pipelined through the ranks: bit-exact AND replicated on every rank
    ring = comm.fold_sharded_axis(A.storage, ib, J * K, Add, np.float32(0))
    comm.fold_status()
    out["fold_ring"] = ring.to_numpy()
    a64 = (a.astype(np.float64) * 3.0)[rank * ib * J * K:(rank + 1) * ib * J * K]
    A64 = Array.new((usize, usize), (ib, J * K), a64, "f64").to_device(ctx)
    out["fold_ring_f64_mul_init"] = comm.fold_sharded_axis(A64.storage, ib, J * K, P.Mul, 0.5).to_numpy()
    comm.fold_status()
    odd_vals = np.random.default_rng([7, rank]).integers(0, 1 << 40, ib * 1028).astype(np.uint64)   # 1028 columns: a ragged last slice
    odd = Array.new((usize, usize), (ib, 1028), odd_vals, usize).to_device(ctx)
    out["fold_ring_u64_xor"] = comm.fold_sharded_axis(odd.storage, ib, 1028, P.BitXor, 0).to_numpy()
    comm.fold_status()

    # (4a) collect() of the WHOLE fold over a peer-mapped Array that knows its communicator is routed to the same fused kernel by the planner rule
    fp2 = PeerStorage(F.F32, I * J * K, fpeers.peers, ib * J * K, keep=A, ctx=ctx, comm=comm)
    whole2 = Array((usize, usize, usize), (I, J, K), fp2, "f32")
    auto = fold_rows(whole2.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0)).collect(location="device", ctx=ctx)
    out["fold_auto"] = auto.as_ref()
    out["fold_auto_kernel"] = np.array([ord(c) for c in ctx.last_kernel()], dtype=np.uint8)

    # (4'') the all-reduce route as ONE fused kernel (k_fold_xchg): per-rank sequential partials combined in rank order
    out["fold_blocked"] = comm.fold_sharded_axis(A.storage, ib, J * K, Add, np.float32(0.125), blocked=True).to_numpy()
    out["fold_blocked_f64_mul_init"] = comm.fold_sharded_axis(A64.storage, ib, J * K, P.Mul, 0.5, blocked=True).to_numpy()
    out["fold_blocked_u64_xor"] = comm.fold_sharded_axis(odd.storage, ib, 1028, P.BitXor, 0, blocked=True).to_numpy()
    i32_vals = np.random.default_rng([9, rank]).integers(-2**31, 2**31, ib * 4124).astype(np.int32)   # 16496-byte rows: 16-byte aligned only, half group at the end
    i32 = Array.new((usize, usize), (ib, 4124), i32_vals, "i32").to_device(ctx)
    for _ in range(3):  # back to back: the packet areas alternate
        out["fold_blocked_i32_add"] = comm.fold_sharded_axis(i32.storage, ib, 4124, Add, np.int32(-7), blocked=True).to_numpy()
    comm.fold_status()

    # wider than one packet area (two launches: the areas alternate) and a DIFFERENT number of rows on every rank
    wide_cols, my_rows = (1 << 20) + 72, 3 + rank
    wide_vals = np.random.default_rng([11, rank]).uniform(0, 1, my_rows * wide_cols).astype(np.float32)
    wide = Array.new((usize, usize), (my_rows, wide_cols), wide_vals).to_device(ctx)
    out["fold_blocked_wide_ragged"] = comm.fold_sharded_axis(wide.storage, my_rows, wide_cols, Add, np.float32(1.5), blocked=True).to_numpy()
    comm.fold_status()

    # (5) all-reduce with the other operators / dtypes
    mine_vals = np.array([rank + 1, 10 - rank, 7], dtype=np.int64)
    for op in ("sum", "prod", "min", "max"):
        w = Storage.device(ctx, F.I64, 3)
        ctx.upload(w.dptr, mine_vals)
        all_reduce_partial(w, op, comm)
        out["ar_" + op] = w.to_numpy()

    comm.close_peers()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    comm.close()
    ctx.close()
    print(f"rank {rank} ok", flush=True)


if __name__ == "__main__":
    main()
