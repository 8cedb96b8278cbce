"""Partitioning along the outermost index: one process per GPU.

Elementwise / broadcast / diagonal / transpose-with-replicated-source / non-sharded-axis folds need
nothing from here beyond `shard_bounds`: every rank collects its own block, no data-path collective.
The two ops with a real exchange step are

  * `compose()` onto, or a `transpose()` of, a source that is itself sharded: `peer_source` maps every
    rank's block into every other rank's address space (CUDA IPC over NVLink/NVSwitch) and the gather /
    transpose kernel reads the owning peer's HBM directly (`mdim_node.n_peers`), so the exchange is fused
    into the kernel; `all_gather_source` is the NCCL alternative the north star names (it moves the whole
    source first);
  * a fold over the sharded axis itself: every rank folds its block, then `all_reduce_partial`
    (NCCL all-reduce; f32 order differs from the sequential reference, 1e-6 relative tolerance).

`torch.distributed` is plumbing here and is imported lazily: the package itself does not need torch.
"""
from __future__ import annotations

import numpy as np

from . import _ffi as F
from . import index as X
from .runtime import Storage, default_context


def shard_bounds(length, world, rank):
    """Contiguous block [lo, hi) of an axis of `length` owned by `rank`: the first `length % world` ranks
    hold one element more (blocks differ by at most one)."""
    q, r = divmod(int(length), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def equal_block(length, world):
    """Block size when every rank must own the same number of elements (peer-mapped gather sources:
    the owner of linear index k is k // block).  The last block may be short."""
    return -(-int(length) // int(world))


def shard_view(view, rank, world):
    """The block of `view` owned by `rank` when its OUTERMOST position axis is cut into `world` contiguous
    blocks (the north star's partitioning): a View over the same index type whose outermost axis is shorter.
    Pure lowering — the outer coordinate becomes lo + k — so the block is still one fused kernel."""
    from . import lowering as L
    from .view import View, _flat

    class Shard(View):
        def __init__(self):
            groups, _ = view._lower()
            axes = _flat(groups)
            if not axes:
                raise ValueError("a rank-0 view has no axis to shard")
            self.lo, self.hi = shard_bounds(axes[0].length, world, rank)
            leaves = X.size_leaves(view.I, view._size)
            types = X.type_leaves(view.I)
            first = next(i for i, (t, s) in enumerate(zip(types, leaves)) if X.leaf_lengths(t, s))
            if len(X.leaf_lengths(types[first], leaves[first])) != 1 or types[first] not in (X.usize, X.Reversed):
                raise ValueError("the outermost axis must be a usize axis to be sharded")
            if len(groups[first]) != 1:
                raise L.Unsupported("the outermost axis has been re-split (to_usize / from_usize): shard before merging it")
            leaves = list(leaves)
            leaves[first] = self.hi - self.lo
            self.I, self.T = view.I, view.T
            self._size = X.build_size(view.I, leaves)

        def _lower(self):
            groups, value = view._lower()
            axes = _flat(groups)
            k = L.Axis(self.hi - self.lo)
            table = {axes[0]: L.Sub(self.lo, ((k, 1),))}
            memo = {}
            new_groups = [[k if a is axes[0] else a for a in g] for g in groups]
            return new_groups, L.map_value(value, lambda n: L.substitute(n, table, memo))

    return Shard()


class PeerStorage(Storage):
    """A source Array split into `world` equal blocks of `block` elements, block p living on rank p.
    Any view can read it: a transpose goes through the tiled kernel, which fetches each tile from the owning
    peer (the all-to-all of a row-sharded transpose, fused into the kernel); gathers (`compose`, `map_axis`)
    and every other chain pick the peer per element."""

    def __init__(self, dtype, n, peers, block, keep=None, ctx=None, opened=()):
        super().__init__(dtype, n, dptr=peers[0], ctx=ctx, owns_device=False, keep=keep)
        self.peers = list(peers)
        self.block = int(block)
        self.home = "device"
        self._opened = list(opened)

    def pointer(self, location):
        if location == "host":
            return self.peers[0]  # CPU checkers: the "peers" are host arrays
        return self.peers[0]

    def close(self):
        for p in self._opened:
            self.ctx.ipc_close(p)
        self._opened = []


def peer_source(local_block_storage, total_len, ctx=None, group=None):
    """Collective: exchange CUDA IPC handles of every rank's block (all blocks `equal_block` long, the
    caller pads the last one) and return a PeerStorage addressing the whole source."""
    import torch.distributed as dist
    ctx = ctx or default_context()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    handle = ctx.ipc_export(local_block_storage.dptr)
    handles = [None] * world
    dist.all_gather_object(handles, handle, group=group)
    peers, opened = [], []
    for p, h in enumerate(handles):
        if p == rank:
            peers.append(local_block_storage.dptr)
        else:
            ptr = ctx.ipc_open(h)
            peers.append(ptr)
            opened.append(ptr)
    return PeerStorage(local_block_storage.dtype, total_len, peers, equal_block(total_len, world), keep=local_block_storage, ctx=ctx, opened=opened)


def as_torch(storage):
    """Zero-copy torch view of a device-resident Storage (for torch.distributed collectives)."""
    import torch
    np_dtype = {F.U8: "|u1", F.I32: "<i4", F.U32: "<u4", F.I64: "<i8", F.U64: "<u8", F.F32: "<f4", F.F64: "<f8"}[storage.dtype]

    class _Iface:
        __cuda_array_interface__ = {"shape": (storage.n,), "typestr": np_dtype, "data": (storage.dptr, False), "version": 2}
    t = torch.as_tensor(_Iface(), device="cuda")
    if storage.dtype == F.U64:
        t = t.view(torch.int64) if t.dtype != torch.int64 else t
    return t


def all_gather_source(local_block, ctx=None, group=None):
    """NCCL all-gather of a sharded source into a full replica on every rank (equal blocks).
    `local_block` is a device-resident Array; returns the full Storage (rank-major blocks)."""
    import torch
    import torch.distributed as dist
    ctx = ctx or default_context()
    world = dist.get_world_size(group)
    st = local_block.storage
    full = Storage.device(ctx, st.dtype, st.n * world)
    dist.all_gather_into_tensor(as_torch(full), as_torch(st), group=group)
    torch.cuda.current_stream().synchronize()
    return full


def all_reduce_partial(partial_storage, op="sum", group=None):
    """Finish a fold over the sharded axis: NCCL all-reduce of the per-rank partial folds, in place."""
    import torch
    import torch.distributed as dist
    ops = {"sum": dist.ReduceOp.SUM, "prod": dist.ReduceOp.PRODUCT, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX,
           "band": dist.ReduceOp.BAND, "bor": dist.ReduceOp.BOR, "bxor": dist.ReduceOp.BXOR}
    if partial_storage.home == "device":
        t = as_torch(partial_storage)
        dist.all_reduce(t, op=ops[op], group=group)
        torch.cuda.current_stream().synchronize()
    else:  # host-resident partials (gloo): used by the CPU test-suite
        t = torch.from_numpy(partial_storage.host)
        dist.all_reduce(t, op=ops[op], group=group)
    return partial_storage
