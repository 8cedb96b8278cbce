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

Nothing here imports torch or torch.distributed: the collectives and the handle exchange go through the C ABI
(`mdim_comm_init`, `mdim_allgather`, `mdim_allreduce`, `mdim_peer_table`; NCCL inside csrc/comm.cu).
"""
from __future__ import annotations

import os

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

    def __init__(self, dtype, n, peers, block, keep=None, ctx=None, opened=(), comm=None):
        super().__init__(dtype, n, dptr=peers[0], ctx=ctx, owns_device=False, keep=keep)
        self.peers = list(peers)
        self.block = int(block)
        self.comm = comm   # with the communicator attached, collect() of a fold over the sharded axis runs the fused ring kernel (view.py)
        self.home = "device"
        self._opened = list(opened)

    def pointer(self, location):
        if location == "host":
            return self.peers[0]  # CPU checkers: the "peers" are host arrays
        return self.peers[0]

    def close(self):
        for p in self._opened:  # mappings made with mdim_ipc_open directly; a Comm's peer table is closed by Comm.close_peers()
            self.ctx.ipc_close(p)
        self._opened = []


def _prefer_bundled_nccl():
    """Point the library at the NCCL wheel of this Python environment (`nvidia/nccl/lib/libnccl.so.2`, the copy torch links against)
    unless the caller chose one (MDIM_NCCL_LIB).  The C library `dlopen`s "libnccl.so.2" by name, which finds the SYSTEM copy when torch has
    not been imported yet; a later `import torch` in the same process would then bind libtorch_cuda.so to that older copy (one SONAME, one
    mapping per process) and fail with an undefined symbol.  A C++ / Rust host has no torch: the system library is the right default there."""
    import importlib.util
    if os.environ.get("MDIM_NCCL_LIB"):
        return
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        spec = None
    for root in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
        cand = os.path.join(root, "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["MDIM_NCCL_LIB"] = cand
            return


class Comm:
    """The communicator of the C ABI (`mdim_comm_*`, csrc/comm.cu): NCCL bound to a Context, one process per GPU.
    Every method is a collective call.  Nothing here touches torch.distributed: the data path of a sharded
    collect is the library's own (peer-mapped kernels, `mdim_allgather`, `mdim_allreduce`)."""

    def __init__(self, ctx, rank, world, unique_id):
        import ctypes as C
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        _prefer_bundled_nccl()
        idb = (C.c_uint8 * F.COMM_ID_BYTES).from_buffer_copy(unique_id)
        ctx.check(ctx.lib.mdim_comm_init(ctx.handle, self.rank, self.world, idb))
        self._open = True

    @staticmethod
    def unique_id():
        """128 bytes made by rank 0 (ncclGetUniqueId); ship them to the other ranks by any means."""
        import ctypes as C
        idb = (C.c_uint8 * F.COMM_ID_BYTES)()
        _prefer_bundled_nccl()
        st = F.lib().mdim_comm_unique_id(idb)
        if st != F.OK:
            raise F.MdimError(st, "mdim_comm_unique_id: " + F.lib().mdim_status_string(st).decode())
        return bytes(idb)

    @staticmethod
    def from_env(ctx=None, path=None, timeout=300.0):
        """Rendezvous through a file on the node (every rank of ONE box sees /tmp): rank 0 writes the id, the others
        wait for it.  RANK / WORLD_SIZE as torchrun sets them; the file name is unique per launch (parent pid)."""
        import os
        import time
        ctx = ctx or default_context()
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        path = path or os.environ.get("MDIM_COMM_FILE") or f"/tmp/mdim_comm_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
        if rank == 0:
            uid = Comm.unique_id()
            with open(path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(path + ".tmp", path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > timeout:
                    raise F.MdimError(F.ERR_NCCL, f"rendezvous file {path} did not appear")
                time.sleep(0.01)
            with open(path, "rb") as f:
                uid = f.read()
        comm = Comm(ctx, rank, world, uid)
        comm.barrier()
        if rank == 0:
            try:
                os.remove(path)
            except OSError:
                pass
        return comm

    def info(self):
        import ctypes as C
        r, w, v = C.c_int(), C.c_int(), C.c_int()
        self.ctx.check(self.ctx.lib.mdim_comm_info(self.ctx.handle, C.byref(r), C.byref(w), C.byref(v)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def barrier(self):
        self.ctx.check(self.ctx.lib.mdim_barrier(self.ctx.handle))

    def all_gather_into(self, full, local):
        """ncclAllGather of every rank's `local` Storage into `full` (rank-major blocks), asynchronous on the context's stream."""
        import ctypes as C
        if full.n != local.n * self.world or full.dtype != local.dtype:
            raise F.Panic(F.ERR_SIZE, "all_gather: the result must hold world x block elements")
        self.ctx.check(self.ctx.lib.mdim_allgather(self.ctx.handle, C.c_void_p(local.dptr), C.c_void_p(full.dptr), local.nbytes))
        return full

    def all_reduce(self, storage, op="sum"):
        """ncclAllReduce in place, asynchronous on the context's stream."""
        import ctypes as C
        code = {"sum": F.ADD, "prod": F.MUL, "min": F.REDUCE_MIN, "max": F.REDUCE_MAX}[op]
        self.ctx.check(self.ctx.lib.mdim_allreduce(self.ctx.handle, C.c_void_p(storage.dptr), storage.n, storage.dtype, code))
        return storage

    def fold_sharded_axis(self, local_rows, n_rows_local, n_cols, op, init, out=None, blocked=False):
        """`blocked=True`: `mdim_fold_sharded_axis_blocked` — per-rank sequential partial folds combined in rank order by an
        in-kernel all-reduce over NVLink (csrc/k_fold_xchg.cu): bit-identical for integer folds, reassociated at the rank
        boundaries (1e-6 relative, deterministic) for float folds.  Otherwise `mdim_fold_sharded_axis`: the fold over the SHARDED (outermost) axis, bit-identical to the reference's sequential
        chain, as one fused kernel per GPU that hands the running values from rank to rank over NVLink (csrc/k_fold_ring.cu).
        `local_rows`: this rank's rows, a device-resident Storage of n_rows_local x n_cols elements; `op`: ops.Add / Sub / Mul /
        BitAnd / BitOr / BitXor; -> the Storage of the n_cols results (on every rank).  Asynchronous on the context's stream."""
        import ctypes as C
        from . import lowering as L
        if local_rows.n != n_rows_local * n_cols:
            raise F.Panic(F.ERR_SIZE, "fold_sharded_axis: the block must hold n_rows_local x n_cols elements")
        out = out or Storage.device(self.ctx, local_rows.dtype, n_cols)
        imm = F.Scalar()
        imm.u64 = L.imm_bits(local_rows.dtype, init)
        fn = self.ctx.lib.mdim_fold_sharded_axis_blocked if blocked else self.ctx.lib.mdim_fold_sharded_axis
        self.ctx.check(fn(self.ctx.handle, C.c_void_p(local_rows.dptr), n_rows_local, n_cols, local_rows.dtype, op.code, imm, C.c_void_p(out.dptr)))
        return out

    def prepare_fold_sharded_axis(self, local_rows, n_rows_local, n_cols, op, init, out=None, blocked=False):
        """The same call with its arguments bound once: -> (run, out Storage); `run()` is a single C-ABI call (a 30 us kernel
        must not wait for Python to rebuild its arguments)."""
        import ctypes as C
        from . import lowering as L
        if local_rows.n != n_rows_local * n_cols:
            raise F.Panic(F.ERR_SIZE, "fold_sharded_axis: the block must hold n_rows_local x n_cols elements")
        out = out or Storage.device(self.ctx, local_rows.dtype, n_cols)
        imm = F.Scalar()
        imm.u64 = L.imm_bits(local_rows.dtype, init)
        fn, check = (self.ctx.lib.mdim_fold_sharded_axis_blocked if blocked else self.ctx.lib.mdim_fold_sharded_axis), self.ctx.check
        args = (self.ctx.handle, C.c_void_p(local_rows.dptr), C.c_uint64(n_rows_local), C.c_uint64(n_cols), C.c_int(local_rows.dtype), C.c_int(op.code), imm, C.c_void_p(out.dptr))

        def run():
            st = fn(*args)
            if st != F.OK:
                check(st)
        return run, out

    def fold_status(self):
        self.ctx.check(self.ctx.lib.mdim_fold_sharded_axis_status(self.ctx.handle))

    def peer_table(self, dptr, nbytes):
        import ctypes as C
        peers = (C.c_void_p * F.MAX_PEERS)()
        self.ctx.check(self.ctx.lib.mdim_peer_table(self.ctx.handle, C.c_void_p(dptr), nbytes, peers))
        return [peers[p] for p in range(self.world)]

    def close_peers(self):
        self.ctx.check(self.ctx.lib.mdim_peer_table_close(self.ctx.handle))

    def close(self):
        if self._open:
            self._open = False
            self.ctx.check(self.ctx.lib.mdim_comm_destroy(self.ctx.handle))


def peer_source(local_block_storage, total_len, comm):
    """Collective: map every rank's block (all blocks `equal_block` long, the caller pads the last one) into this
    process (`mdim_peer_table`: CUDA IPC handles exchanged over the communicator) and return a PeerStorage
    addressing the whole source."""
    peers = comm.peer_table(local_block_storage.dptr, local_block_storage.nbytes)
    return PeerStorage(local_block_storage.dtype, total_len, peers, equal_block(total_len, comm.world), keep=local_block_storage, ctx=comm.ctx, comm=comm)


def all_gather_source(local_block, comm):
    """NCCL all-gather of a sharded source into a full replica on every rank (equal blocks), through the C ABI.
    `local_block` is a device-resident Array; returns the full Storage (rank-major blocks)."""
    st = local_block.storage
    full = Storage.device(comm.ctx, st.dtype, st.n * comm.world)
    comm.all_gather_into(full, st)
    comm.ctx.sync()
    return full


def all_reduce_partial(partial_storage, op, comm):
    """Finish a fold over the sharded axis: NCCL all-reduce of the per-rank partial folds, in place (f32 sums are
    reassociated: 1e-6 relative tolerance, north star)."""
    comm.all_reduce(partial_storage, op)
    comm.ctx.sync()
    return partial_storage


# Measured on one 8 x B200 NVSwitch box (profiles/r2_bench_n{2,4,8}.json, `scaling_ops`), per number of GPUs:
_REMOTE_GATHERS_PER_S = {2: 8.3e9, 4: 7.3e9, 8: 1.1e10}   # uniform-random 4-byte reads of the peers' HBM over NVLink, per GPU
_LOCAL_GATHERS_PER_S = 44.0e9                              # the same reads from the GPU's own HBM (config 3: 2^28 in 6.0 ms)
_ALLGATHER_IN_BYTES_PER_S = {2: 480.0e9, 4: 620.0e9, 8: 637.0e9}  # mdim_allgather (NCCL), bytes arriving per GPU per second


def _rate(table, world):
    return table[min(table, key=lambda n: abs(n - world))]


def choose_compose_route(n_idx_local, src_bytes, world, reuse=1):
    """How should `idx.compose(src)` run when `src` is sharded over `world` GPUs and this rank holds `n_idx_local`
    indices?  -> "peer" (the gather kernel reads every element from the GPU that owns it, mdim_node.peer[]) or
    "allgather" (mdim_allgather the source first, then gather locally).  A cost model over the measured rates
    above; `reuse` = how many collects will read the gathered source (the all-gather is paid once).
    Measured crossover (2^28 indices into a 4 GiB source): all-gather first wins at 2 GPUs (7.4 vs 9.7 ms) and at 4 (6.7 vs 7.3),
    peer-mapped at 8 (2.7 vs 6.6 ms)."""
    if world <= 1:
        return "local"
    remote = n_idx_local * (world - 1) / world
    t_peer = remote / _rate(_REMOTE_GATHERS_PER_S, world) + (n_idx_local - remote) / _LOCAL_GATHERS_PER_S
    t_ag = src_bytes * (world - 1) / world / _rate(_ALLGATHER_IN_BYTES_PER_S, world) / max(reuse, 1) + n_idx_local / _LOCAL_GATHERS_PER_S
    return "peer" if t_peer <= t_ag else "allgather"
