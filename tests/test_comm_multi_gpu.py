"""The multi-GPU surface of the C ABI on real GPUs (>= 2 needed; skipped otherwise): one process per GPU, every
exchange through include/mdim.h — no torch.distributed anywhere.  Run with `gpurun --gpus 2 -- python -m pytest
tests/test_comm_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_same_bits

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

HERE = os.path.dirname(os.path.abspath(__file__))


_GPUS = None


def _gpu_count():
    """GPUs of the box, counted ONCE in a fresh subprocess: after an in-process test has initialised CUDA through the product
    library, torch.cuda.device_count() in the same process was seen to answer 1 on a 2-GPU box (and skipped the tests below)."""
    global _GPUS
    if _GPUS is None:
        try:
            out = subprocess.run([sys.executable, "-c", "import torch; print(torch.cuda.device_count())"], capture_output=True, text=True, timeout=300).stdout
            _GPUS = int(out.strip().splitlines()[-1])
        except Exception:
            _GPUS = 0
    return _GPUS


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_c_abi_collectives_and_peer_tables(world, tmp_path):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    rdv = str(tmp_path / "rendezvous")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi", "comm_worker.py"), str(r), str(world), rdv, str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    rng = np.random.default_rng(2024)
    n_src = 1 << 16
    src = rng.uniform(-1, 1, n_src).astype(np.float32)
    idx = rng.integers(0, n_src, 5000).astype(np.uint64)
    want = src[idx.astype(np.int64)]
    assert_same_bits(np.concatenate([p["ag_gather"] for p in parts]), want, "all-gather + gather")
    assert_same_bits(np.concatenate([p["peer_gather"] for p in parts]), want, "peer-mapped gather (block at an interior offset)")
    M, N = 64 * world * 2, 384
    mat = rng.uniform(-1, 1, M * N).astype(np.float32)
    assert_same_bits(np.concatenate([p["peer_transpose"] for p in parts]), mat.reshape(M, N).T.copy().reshape(-1), "peer-mapped transpose")
    assert_same_bits(np.concatenate([p["peer_transpose_notma"] for p in parts]), mat.reshape(M, N).T.copy().reshape(-1), "peer-mapped transpose (evaluator)")
    assert "transpose" in bytes(parts[0]["peer_transpose_kernel"]).decode()
    I, J, K = 16 * world, 24, 32
    a = rng.uniform(0, 1, I * J * K).astype(np.float32)
    seq = np.zeros(J * K, np.float32)
    for i in range(I):
        seq = seq + a.reshape(I, J * K)[i]
    for p in parts:  # reassociated across ranks: the north star's 1e-6 relative tolerance
        assert np.max(np.abs(p["fold_allreduce"] - seq) / np.abs(seq)) <= 1e-6
    assert_same_bits(np.concatenate([p["fold_exact"] for p in parts]), seq, "peer-mapped fold over the sharded axis (bit-exact)")
    seq64 = np.full(J * K, 0.5)
    for i in range(I):
        seq64 = seq64 * (a.astype(np.float64) * 3.0).reshape(I, J * K)[i]
    ib = I // world
    want_xor = np.zeros(1028, np.uint64)
    for r in range(world):
        want_xor ^= np.bitwise_xor.reduce(np.random.default_rng([7, r]).integers(0, 1 << 40, ib * 1028).astype(np.uint64).reshape(ib, 1028), axis=0)
    for p in parts:  # the ring fold is bit-exact and replicated
        assert_same_bits(p["fold_ring"], seq, "ring fold over the sharded axis (f32 add)")
        assert_same_bits(p["fold_ring_f64_mul_init"], seq64, "ring fold (f64 mul, init 0.5)")
        assert_same_bits(p["fold_ring_u64_xor"], want_xor, "ring fold (u64 xor, ragged slice)")
        assert_same_bits(p["fold_auto"], seq, "collect() of the whole fold over the sharded axis (planner rule -> fused kernel)")
        assert bytes(p["fold_auto_kernel"]).decode() == "k_fold_ring"
    # the blocked route: P_r sequential (rank 0 from init, the others from the identity), combined in rank order — restated here
    rows = a.reshape(world, ib, J * K)
    blocked, blocked64 = None, None
    for r in range(world):
        pr = np.full(J * K, np.float32(0.125) if r == 0 else np.float32(-0.0))
        pr64 = np.full(J * K, 0.5 if r == 0 else 1.0)
        for i in range(ib):
            pr = pr + rows[r, i]
            pr64 = pr64 * (rows[r, i].astype(np.float64) * 3.0)
        blocked = pr if r == 0 else blocked + pr
        blocked64 = pr64 if r == 0 else blocked64 * pr64
    want_i32 = np.full(4124, -7, np.int64)
    for r in range(world):
        want_i32 += np.random.default_rng([9, r]).integers(-2**31, 2**31, ib * 4124).astype(np.int32).reshape(ib, 4124).astype(np.int64).sum(axis=0)
    want_i32 = want_i32.astype(np.int32)  # wrapping
    for p in parts:
        assert_same_bits(p["fold_blocked"], blocked, "blocked fold (f32 add, init 0.125): the rank-ordered combination, bit for bit")
        assert np.max(np.abs(p["fold_blocked"] - (seq + np.float32(0.125))) / np.abs(seq)) <= 1e-6
        assert_same_bits(p["fold_blocked_f64_mul_init"], blocked64, "blocked fold (f64 mul, init 0.5)")
        assert_same_bits(p["fold_blocked_u64_xor"], want_xor, "blocked fold (u64 xor): associative, identical to the reference")
        assert_same_bits(p["fold_blocked_i32_add"], want_i32, "blocked fold (i32 wrapping add, rows 16-byte aligned only)")
    wide_cols = (1 << 20) + 72
    from reference_model import blocked_fold_over_sharded_axis
    wide_blocks = [np.random.default_rng([11, r]).uniform(0, 1, (3 + r) * wide_cols).astype(np.float32).reshape(3 + r, wide_cols) for r in range(world)]
    want_wide = blocked_fold_over_sharded_axis(wide_blocks, np.add, np.float32(1.5), np.float32(-0.0))
    for p in parts:
        assert_same_bits(p["fold_blocked_wide_ragged"], want_wide, "blocked fold: two packet-area windows, 3 + rank rows per rank")
    ranks = np.arange(world)
    for p in parts:
        assert p["ar_sum"].tolist() == [int((ranks + 1).sum()), int((10 - ranks).sum()), 7 * world]
        assert p["ar_prod"].tolist() == [int(np.prod(ranks + 1)), int(np.prod(10 - ranks)), 7 ** world]
        assert p["ar_min"].tolist() == [1, 10 - (world - 1), 7]
        assert p["ar_max"].tolist() == [world, 10, 7]


@pytest.mark.gpu
def test_cpp_mirror_runs_the_two_gpu_transpose(tmp_path):
    """The typed C++ host mirror drives two GPUs through the C ABI alone: communicator, peer table, the peer-mapped
    transpose of a row-sharded Array, all-gather + transpose, partial fold + all-reduce (tests/cpp/test_multi_gpu.cpp)."""
    world = 2
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import multidimension_b200 as P
    exe = str(tmp_path / "test_multi_gpu")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(HERE, "cpp", "test_multi_gpu.cpp"), "-ldl"], check=True)
    rdv = str(tmp_path / "rendezvous")
    procs = [subprocess.Popen([exe, str(r), str(world), P.LIB_PATH, rdv], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and "0 failures" in o, f"rank {r}:\n{o[-2000:]}"


@pytest.mark.gpu
def test_ring_fold_single_rank_is_the_sequential_fold():
    """world = 1: the pipelined fold kernel alone (TMA row tiles through the shared-memory ring, sequential adds), no peers."""
    import multidimension_b200 as P
    from multidimension_b200 import _ffi as F
    from multidimension_b200.runtime import Storage
    from multidimension_b200.sharding import Comm
    ctx = P.Context(0)
    comm = Comm(ctx, 0, 1, Comm.unique_id())
    rng = np.random.default_rng(11)
    for rows, cols, dt, npdt in ((100, 4096, F.F32, np.float32), (33, 1028, F.F32, np.float32), (257, 516, F.F64, np.float64), (1, 8, F.F32, np.float32)):
        x = rng.uniform(0, 1, rows * cols).astype(npdt)
        st = Storage.device(ctx, dt, rows * cols)
        ctx.upload(st.dptr, x)
        got = comm.fold_sharded_axis(st, rows, cols, P.Add, npdt(0.25)).to_numpy()
        comm.fold_status()
        want = np.full(cols, npdt(0.25))
        for i in range(rows):
            want = want + x.reshape(rows, cols)[i]
        assert_same_bits(got, want, f"ring fold {rows}x{cols}")
    comm.close()
    ctx.close()


@pytest.mark.gpu
def test_blocked_fold_single_rank_is_the_sequential_fold():
    """world = 1: k_fold_xchg alone (column walk with 256-bit loads, packets through the local area), no peers."""
    import multidimension_b200 as P
    from multidimension_b200 import _ffi as F
    from multidimension_b200.runtime import Storage
    from multidimension_b200.sharding import Comm
    ctx = P.Context(0)
    comm = Comm(ctx, 0, 1, Comm.unique_id())
    rng = np.random.default_rng(12)
    for rows, cols, dt, npdt in ((100, 4096, F.F32, np.float32), (33, 1028, F.F32, np.float32), (257, 516, F.F64, np.float64), (1, 8, F.F32, np.float32),
                                 (40, (1 << 20) + 72, F.F32, np.float32)):  # the last one: wider than one packet area -> two launches
        x = rng.uniform(0, 1, rows * cols).astype(npdt)
        st = Storage.device(ctx, dt, rows * cols)
        ctx.upload(st.dptr, x)
        got = comm.fold_sharded_axis(st, rows, cols, P.Add, npdt(0.25), blocked=True).to_numpy()
        comm.fold_status()
        assert ctx.last_kernel() == "k_fold_xchg"
        want = np.full(cols, npdt(0.25))
        for i in range(rows):
            want = want + x.reshape(rows, cols)[i]
        assert_same_bits(got, want, f"blocked fold {rows}x{cols}")
    with pytest.raises(F.MdimError):
        comm.fold_sharded_axis(st, rows, cols, P.Sub, npdt(0), blocked=True)  # no identity: the ring route serves SUB
    comm.close()
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("route", ["ring", "blocked"])
def test_a_missing_peer_is_an_error_not_a_hung_gpu(route, tmp_path):
    """Both fused fold kernels wait on peers inside the kernel; their spins are bounded: a rank that calls alone gets MDIM_ERR_NCCL
    from mdim_fold_sharded_axis_status after ~2 s, the error word is cleared, and the communicator still shuts down cleanly."""
    world = 2
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    rdv = str(tmp_path / "rendezvous")
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "multi", "timeout_worker.py"), str(r), str(world), rdv, route],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
    assert "gave up after" in outs[0]
