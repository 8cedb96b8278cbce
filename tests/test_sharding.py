"""Multi-GPU partitioning: host-side logic on CPU with world_size-2 gloo, the
peer-mapped gather kernel path on one GPU (two "peers" that are buffers of the same device)."""
import os
import socket

import numpy as np
import pytest

from helpers import oracle_collect, emu_collect, assert_same_bits
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, Add, fold_rows, _ffi as F
from multidimension_b200.runtime import Storage
from multidimension_b200.sharding import shard_bounds, shard_view, equal_block, PeerStorage


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            blocks = [shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
            assert equal_block(n, w) * w >= n


def _peer_case(rng, n_src, n_idx, world):
    src = rng.uniform(-1, 1, n_src).astype(np.float32)
    block = equal_block(n_src, world)
    padded = np.zeros(block * world, np.float32)
    padded[:n_src] = src
    shards = [np.ascontiguousarray(padded[p * block:(p + 1) * block]) for p in range(world)]
    idx = rng.integers(0, n_src, n_idx).astype(np.uint64)
    return src, shards, block, idx


@pytest.mark.parametrize("world", [2, 3, 8])
def test_peer_sharded_compose_matches_replicated(world):
    rng = np.random.default_rng(world)
    src, shards, block, idx = _peer_case(rng, 1000, 5000, world)
    peers = PeerStorage(F.F32, src.size, [s.ctypes.data for s in shards], block, keep=shards)
    sharded = Array.new(usize, idx.size, idx).compose(Array(usize, src.size, peers, "f32"))
    replicated = Array.new(usize, idx.size, idx).compose(Array.new(usize, src.size, src))
    want = oracle_collect(replicated)
    assert_same_bits(oracle_collect(sharded), want)
    assert_same_bits(emu_collect(sharded), want)
    # ... and a sharded Array is readable by any other chain as well (element-wise peer lookup)
    want2 = oracle_collect(Array.new(usize, src.size, src) * Scalar(2.0, "f32"))
    assert_same_bits(oracle_collect(Array(usize, src.size, peers, "f32") * Scalar(2.0, "f32")), want2)
    assert_same_bits(emu_collect(Array(usize, src.size, peers, "f32") * Scalar(2.0, "f32")), want2)


def _sharded_matrix(rng, rows, cols, world, T="f32"):
    """A (rows, cols) matrix cut into `world` equal blocks of its flat storage (the last block padded)."""
    dt = {"f32": np.float32, "f64": np.float64}[T]
    full = rng.uniform(-1, 1, rows * cols).astype(dt)
    block = equal_block(full.size, world)
    padded = np.zeros(block * world, dt)
    padded[:full.size] = full
    shards = [np.ascontiguousarray(padded[p * block:(p + 1) * block]) for p in range(world)]
    return full, shards, block


@pytest.mark.parametrize("world,shape", [(2, (64, 48)), (3, (50, 36)), (8, (128, 64))])
def test_peer_sharded_transpose_matches_replicated(world, shape):
    """SURVEY.md §8e transpose row: source sharded on its outermost axis, every rank collects a column block of
    the transpose.  The exchange is the kernel's own loads from the owning peer (here: host arrays)."""
    rng = np.random.default_rng(world * 100 + shape[0])
    full, shards, block = _sharded_matrix(rng, shape[0], shape[1], world)
    peers = PeerStorage(F.F32, full.size, [s.ctypes.data for s in shards], block, keep=shards)
    sharded = Array((usize, usize), shape, peers, "f32").transpose((), usize, usize, ())
    want = full.reshape(shape).T.copy().reshape(-1)
    assert_same_bits(oracle_collect(sharded), want)
    assert_same_bits(emu_collect(sharded), want)
    for rank in range(world):  # each rank's block of the output (outermost axis of the transpose = columns)
        lo, hi = shard_bounds(shape[1], world, rank)
        blockview = shard_view(sharded, rank, world)
        assert_same_bits(oracle_collect(blockview), full.reshape(shape).T[lo:hi].copy().reshape(-1))
        assert_same_bits(emu_collect(blockview), full.reshape(shape).T[lo:hi].copy().reshape(-1))


def _sharded_axis_fold(arr3, I, J, K):
    """Fold over the OUTERMOST axis of an (I, J, K) Array in index order: out[j, k] = ((0 + a[0,j,k]) + a[1,j,k]) + ..."""
    moved = arr3.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize))
    return fold_rows(moved, (usize, usize), usize, Add, np.float32(0))


@pytest.mark.parametrize("world", [2, 3])
def test_fold_over_the_sharded_axis_reading_peers_is_bit_exact(world):
    """The all-reduce route reassociates the sum (1e-6 tolerance).  Reading the peers' blocks in index order
    instead keeps the reference's add chain: bit-identical to the unsharded fold."""
    rng = np.random.default_rng(41 + world)
    I, J, K = 6 * world, 5, 8
    full = rng.uniform(0, 1, I * J * K).astype(np.float32)
    block = equal_block(full.size, world)
    shards = [np.ascontiguousarray(full[p * block:(p + 1) * block]) for p in range(world)]
    peers = PeerStorage(F.F32, full.size, [s.ctypes.data for s in shards], block, keep=shards)
    want = oracle_collect(_sharded_axis_fold(Array.new((usize, usize, usize), (I, J, K), full), I, J, K))
    view = _sharded_axis_fold(Array((usize, usize, usize), (I, J, K), peers, "f32"), I, J, K)
    assert_same_bits(oracle_collect(view), want)
    assert_same_bits(emu_collect(view), want)
    seq = np.zeros(J * K, np.float32)
    for i in range(I):
        seq = seq + full.reshape(I, J * K)[i]
    assert_same_bits(want, seq)


@pytest.mark.gpu
def test_fold_over_the_sharded_axis_reading_peers_gpu():
    ctx = P.Context(0)
    rng = np.random.default_rng(77)
    world, I, J, K = 4, 64, 48, 32
    full = rng.uniform(0, 1, I * J * K).astype(np.float32)
    block = equal_block(full.size, world)
    shards = [np.ascontiguousarray(full[p * block:(p + 1) * block]) for p in range(world)]
    dev = [Storage.from_host(F.F32, s).ensure_device(ctx) for s in shards]
    peers = PeerStorage(F.F32, full.size, [d.dptr for d in dev], block, keep=dev, ctx=ctx)
    view = _sharded_axis_fold(Array((usize, usize, usize), (I, J, K), peers, "f32"), I, J, K)
    seq = np.zeros(J * K, np.float32)
    for i in range(I):
        seq = seq + full.reshape(I, J * K)[i]
    assert_same_bits(view.collect(location="device", ctx=ctx).as_ref(), seq)
    for rank in range(world):  # each rank's block of the (J, K) result
        lo, hi = shard_bounds(J, world, rank)
        got = shard_view(view, rank, world).collect(location="device", ctx=ctx).as_ref()
        assert_same_bits(got, seq.reshape(J, K)[lo:hi].reshape(-1))
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("T,world,shape", [("f32", 2, (256, 192)), ("f32", 3, (250, 131)), ("f64", 8, (512, 320)), ("f32", 8, (1024, 1024))])
def test_peer_sharded_transpose_gpu(T, world, shape):
    ctx = P.Context(0)
    rng = np.random.default_rng(world * 1000 + shape[1])
    full, shards, block = _sharded_matrix(rng, shape[0], shape[1], world, T)
    dt = F.F32 if T == "f32" else F.F64
    dev = [Storage.from_host(dt, s).ensure_device(ctx) for s in shards]
    peers = PeerStorage(dt, full.size, [d.dptr for d in dev], block, keep=dev, ctx=ctx)
    sharded = Array((usize, usize), shape, peers, T).transpose((), usize, usize, ())
    assert ".peers" in sharded.describe()
    want = full.reshape(shape).T
    for rank in (0, world - 1):
        lo, hi = shard_bounds(shape[1], world, rank)
        got = shard_view(sharded, rank, world).collect(location="device", ctx=ctx).as_ref()
        assert_same_bits(got, want[lo:hi].copy().reshape(-1))
    assert_same_bits(sharded.collect(location="device", ctx=ctx).as_ref(), want.copy().reshape(-1))
    # the same Array through the general evaluator (per-element peer lookup)
    assert_same_bits(sharded.collect(location="device", ctx=ctx, flags=F.COLLECT_NO_FASTPATH).as_ref(), want.copy().reshape(-1))
    chain = Array((usize, usize), shape, peers, T) * Scalar(3.0, T)
    assert_same_bits(chain.collect(location="device", ctx=ctx).as_ref(), full * (np.float32(3) if T == "f32" else 3.0))
    ctx.close()


@pytest.mark.gpu
def test_peer_sharded_compose_gpu():
    ctx = P.Context(0)
    rng = np.random.default_rng(5)
    src, shards, block, idx = _peer_case(rng, 100000, 1 << 18, 2)
    dev = [Storage.from_host(F.F32, s).ensure_device(ctx) for s in shards]
    peers = PeerStorage(F.F32, src.size, [d.dptr for d in dev], block, keep=dev, ctx=ctx)
    sharded = Array.new(usize, idx.size, idx).compose(Array(usize, src.size, peers, "f32"))
    got = sharded.collect(location="device", ctx=ctx).as_ref()
    assert_same_bits(got, src[idx.astype(np.int64)])
    bad = idx.copy()
    bad[777] = src.size
    with pytest.raises(P.Panic, match=f"Index {src.size} is out of bounds for size {src.size}"):
        Array.new(usize, bad.size, bad).compose(Array(usize, src.size, peers, "f32")).collect(location="device", ctx=ctx)
    ctx.close()


def _shardable_views(rng):
    a = Array.new((usize, usize), (3, 5), rng.uniform(-1, 1, 15).astype(np.float32))
    w = Array.new(usize, 8, rng.uniform(-1, 1, 8).astype(np.float32))
    m = Array.new((usize, usize), (7, 6), rng.uniform(-1, 1, 42).astype(np.float32))
    idx = Array.new(usize, 11, rng.integers(0, 42, 11).astype(np.uint64))
    c5 = (a.transpose((), usize, usize, ()).diagonal(np.float32(0)).iso((((usize, usize), (usize, usize)), ()))
          .zip(w.iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))                 # BASELINE config 5
    return [c5, m.transpose((), usize, usize, ()), m * m + Scalar(1.0, "f32"), P.all_(usize, 6).map(lambda x: x + 10).diagonal(0),
            idx.compose(Array.new(usize, 42, rng.uniform(-1, 1, 42).astype(np.float32))), fold_rows(m, usize, usize, Add, np.float32(0)),
            m.concat(m, (), usize)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_view_blocks_tile_the_full_result(world):
    """`shard_view(v, rank, world)` is the block of v along its outermost index; the blocks of all ranks,
    concatenated, are the unsharded collect — including under a diagonal (the shard's predicate gets an offset)."""
    from multidimension_b200.sharding import shard_view
    for v in _shardable_views(np.random.default_rng(world)):
        full = oracle_collect(v)
        for run in (oracle_collect, emu_collect):
            parts = [run(shard_view(v, r, world)) for r in range(world)]
            assert_same_bits(np.concatenate(parts), full, f"{v.describe()} x{world}")


@pytest.mark.gpu
def test_shard_view_gpu():
    from multidimension_b200.sharding import shard_view
    ctx = P.Context(0)
    for v in _shardable_views(np.random.default_rng(4)):
        full = oracle_collect(v)
        parts = [shard_view(v, r, 4).collect(location="device", ctx=ctx).as_ref() for r in range(4)]
        assert_same_bits(np.concatenate(parts), full, str(v.describe()))
    ctx.close()


# ---- world_size-2 gloo: the per-rank flow of bench.py / a sharded application ---------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)  # same data on every rank; each takes its block
    I, J, K = 6, 5, 32
    a = rng.uniform(0, 1, I * J * K).astype(np.float32)
    b = rng.uniform(-1, 1, I * J * K).astype(np.float32)
    lo, hi = shard_bounds(I, world, rank)
    blk = slice(lo * J * K, hi * J * K)
    A = Array.new((usize, usize, usize), (hi - lo, J, K), a[blk])
    B = Array.new((usize, usize, usize), (hi - lo, J, K), b[blk])
    # (1) elementwise: independent shards, no collective
    ew = emu_collect(A.zip(B).map(lambda p: p[0] * p[1] + np.float32(1)))
    # (2) fold over a NON-sharded axis + broadcast subtract: independent shards
    mean = fold_rows(A, (usize, usize), usize, Add, np.float32(0)) / Scalar(float(K), "f32")
    c4 = emu_collect(A - mean.iso((usize, usize, ())))
    # (3) fold over the SHARDED axis: per-rank partial of the output shape, then all-reduce
    At = A.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize))  # (J,K) x local I
    part = emu_collect(fold_rows(At, (usize, usize), usize, Add, np.float32(0)))
    st = Storage.from_host(F.F32, part.copy())
    dist.all_reduce(torch.from_numpy(st.host))  # gloo stands in for mdim_allreduce (NCCL, GPU only): same per-rank flow
    # (4) compose with a sharded source: all-gather the source, gather locally (index shard per rank)
    src_block = torch.from_numpy(a[rank * (a.size // world):(rank + 1) * (a.size // world)].copy())
    gathered = [torch.empty_like(src_block) for _ in range(world)]
    dist.all_gather(gathered, src_block)
    full_src = torch.cat(gathered).numpy()
    idx = rng.integers(0, a.size, 4000).astype(np.uint64)
    ilo, ihi = shard_bounds(idx.size, world, rank)
    comp = emu_collect(Array.new(usize, ihi - ilo, idx[ilo:ihi]).compose(Array.new(usize, full_src.size, full_src)))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ew=ew, c4=c4, red=st.host, comp=comp)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_sharded_flows(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    rng = np.random.default_rng(99)
    I, J, K = 6, 5, 32
    a = rng.uniform(0, 1, I * J * K).astype(np.float32)
    b = rng.uniform(-1, 1, I * J * K).astype(np.float32)
    A = Array.new((usize, usize, usize), (I, J, K), a)
    B = Array.new((usize, usize, usize), (I, J, K), b)
    assert_same_bits(np.concatenate([p["ew"] for p in parts]), oracle_collect(A.zip(B).map(lambda p: p[0] * p[1] + np.float32(1))))
    mean = fold_rows(A, (usize, usize), usize, Add, np.float32(0)) / Scalar(float(K), "f32")
    assert_same_bits(np.concatenate([p["c4"] for p in parts]), oracle_collect(A - mean.iso((usize, usize, ()))))
    # sharded-axis fold: order differs from the sequential reference -> 1e-6 relative (north star)
    At = A.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize))
    want = oracle_collect(fold_rows(At, (usize, usize), usize, Add, np.float32(0)))
    for p in parts:
        assert np.max(np.abs(p["red"] - want) / np.abs(want)) <= 1e-6
    idx = rng.integers(0, a.size, 4000).astype(np.uint64)
    assert_same_bits(np.concatenate([p["comp"] for p in parts]), a[idx.astype(np.int64)])


# ---- world_size-2 gloo: peer-mapped flows.  The peers' blocks are POSIX shared memory here (CUDA IPC on the GPU
#      box): every rank maps every block into its own address space and reads the owner's copy directly. ----------
def _peer_worker(rank, world, port, out_dir):
    import ctypes
    from multiprocessing import shared_memory
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)  # the same global data on every rank; each publishes only its own block
    M, Ncols, K = 12, 10, 4
    full = rng.uniform(0, 1, M * Ncols * K).astype(np.float32)
    block = equal_block(full.size, world)
    mine = shared_memory.SharedMemory(create=True, size=block * 4)
    local = np.ndarray(block, np.float32, buffer=mine.buf)
    local[:] = 0
    chunk = full[rank * block:(rank + 1) * block]
    local[:chunk.size] = chunk
    names = [None] * world
    dist.all_gather_object(names, mine.name)  # the "IPC handle" exchange of sharding.peer_source
    opened = [mine if r == rank else shared_memory.SharedMemory(name=names[r]) for r in range(world)]
    views = [np.ndarray(block, np.float32, buffer=s.buf) for s in opened]
    peers = PeerStorage(F.F32, full.size, [v.ctypes.data for v in views], block, keep=views)
    dist.barrier()  # every block is filled before anyone reads a peer
    # (1) transpose of the row-sharded (M, Ncols*K) matrix: this rank's block of the transposed rows
    t_view = shard_view(Array((usize, usize), (M, Ncols * K), peers, "f32").transpose((), usize, usize, ()), rank, world)
    tr = emu_collect(t_view)
    # (2) fold over the sharded axis in index order, this rank's block of the (Ncols, K) result: bit-exact
    whole = Array((usize, usize, usize), (M, Ncols, K), peers, "f32")
    f_view = shard_view(fold_rows(whole.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0)), rank, world)
    fo = emu_collect(f_view)
    # (3) gather with indices sharded, source peer-mapped
    idx = rng.integers(0, full.size, 3000).astype(np.uint64)
    ilo, ihi = shard_bounds(idx.size, world, rank)
    ga = emu_collect(Array.new(usize, ihi - ilo, idx[ilo:ihi]).compose(Array(usize, full.size, peers, "f32")))
    np.savez(os.path.join(out_dir, f"peer{rank}.npz"), tr=tr, fo=fo, ga=ga)
    dist.barrier()  # nobody unmaps while a peer may still be reading
    del peers, views, local
    for r, s in enumerate(opened):
        s.close()
    dist.barrier()
    mine.unlink()
    dist.destroy_process_group()


def test_world2_gloo_peer_mapped_flows(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, f"peer{r}.npz")) for r in range(world)]
    rng = np.random.default_rng(7)
    M, Ncols, K = 12, 10, 4
    full = rng.uniform(0, 1, M * Ncols * K).astype(np.float32)
    assert_same_bits(np.concatenate([p["tr"] for p in parts]), full.reshape(M, Ncols * K).T.copy().reshape(-1))
    seq = np.zeros(Ncols * K, np.float32)
    for i in range(M):
        seq = seq + full.reshape(M, Ncols * K)[i]
    assert_same_bits(np.concatenate([p["fo"] for p in parts]), seq)  # bit-exact: no reassociation across ranks
    idx = rng.integers(0, full.size, 3000).astype(np.uint64)
    assert_same_bits(np.concatenate([p["ga"] for p in parts]), full[idx.astype(np.int64)])


def test_collect_routes_a_whole_fold_over_the_sharded_axis_to_the_fused_kernel():
    """view.py::_fused_sharded_axis_fold — the planner rule that sends `rows().map(fold)` over the sharded axis of a whole peer-mapped
    Array to `mdim_fold_sharded_axis` — matches exactly that shape and nothing else (host logic only: the communicator is a stand-in)."""
    from multidimension_b200 import view as V

    class FakeComm:
        rank, world = 1, 2
        calls = []

        def __init__(self, ctx):
            self.ctx = ctx

        def fold_sharded_axis(self, local, rows, cols, op, init, out=None):
            FakeComm.calls.append((local.dptr, local.n, rows, cols, op.code, init))
            return "result"

        def fold_status(self):
            pass

    ctx = object()
    I, J, K = 8, 3, 4
    data = np.zeros(I * J * K // 2, np.float32)
    keep = [data, data.copy()]
    comm = FakeComm(ctx)
    st = PeerStorage(F.F32, I * J * K, [k.ctypes.data for k in keep], I * J * K // 2, keep=keep, comm=comm)
    orig_wrap = V.Storage.wrap_device
    V.Storage.wrap_device = staticmethod(lambda ctx_, dtype, n, ptr, keep=None: type("S", (), {"dptr": ptr, "n": n, "dtype": dtype})())
    try:
        whole = Array((usize, usize, usize), (I, J, K), st, "f32")
        outer = fold_rows(whole.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0.5))

        def route(view, flags=0):
            groups, value = view._lower()
            from multidimension_b200 import lowering as L
            return V._fused_sharded_axis_fold(ctx, L.flatten_value(value), [a for g in groups for a in g], flags, None)
        assert route(outer) == "result"
        assert FakeComm.calls[-1] == (keep[1].ctypes.data, I * J * K // 2, I // 2, J * K, F.ADD, np.float32(0.5))   # THIS rank's block, its rows
        assert route(outer, F.COLLECT_NO_FASTPATH) is None
        assert route(shard_view(outer, 0, 2)) is None                                              # a block of the result: the peer-mapped evaluator
        assert route(fold_rows(whole, (usize, usize), usize, Add, np.float32(0))) is None           # the LAST axis is local to every row
        middle = fold_rows(whole.transpose(usize, usize, usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0))
        assert route(middle) is None                                                                # a middle axis is not the sharded one
        assert route(outer + outer) is None                                                         # anything around the fold
        plain = PeerStorage(F.F32, I * J * K, [k.ctypes.data for k in keep], I * J * K // 2, keep=keep)
        assert route(fold_rows(Array((usize, usize, usize), (I, J, K), plain, "f32").transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)),
                               (usize, usize), usize, Add, np.float32(0))) is None                  # no communicator attached
    finally:
        V.Storage.wrap_device = orig_wrap
