#!/usr/bin/env python
"""bench.py — effective HBM GB/s of View -> collect() (BASELINE.json metric).

A step is ONE collect() of BASELINE config 2 — `a.zip(b).map(|(x,y)| x*y+1)` over two 2^30-element
f32 Arrays per GPU (12 GiB of algorithmic traffic per step per GPU) — through the C ABI.  With N>1
(torchrun, one process per GPU) the global Array is partitioned along its outermost index: every
rank collects its own 2^30-element block, no data-path collective, weak scaling.

  value     device-resident: K back-to-back collects, CUDA events on the launching stream, max over ranks
  e2e       the same collect through mdim_collect_host: operands and result in pinned HOST memory,
            H2D/D2H copies inside the timed region (chunked and overlapped by the library)
  roofline  the fused elementwise kernel against the measured HBM copy rate (MEASURED_PEAKS.json)
  ops       the other BASELINE configs (transpose, gather, fold+broadcast-subtract, rank-5 chain), N=1
  cpu_baseline   the CPU oracle (a restatement of the reference's single-threaded collect(); the Rust
            crate itself cannot be built here) on a bounded sample of the same workload

`--impl reference` times that CPU restatement alone, as the reference arm.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

METRIC = "effective HBM GB/s for View->collect (zip+map x*y+1 over two 2^30-element f32 Arrays per GPU)"
UNIT = "GB/s"
N_ELEMS = 1 << 30
WORKLOAD = "configs[1]: zip+map elementwise collect of two 2^30-element f32 Arrays (a.zip(b).map(|(x,y)| x*y+1))"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def known_traffic(kernel):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, device):
        self.samples, self.reasons, self.stop_flag, self.max_mhz, self.thread = [], set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self.stop_flag = False
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join()
            self.thread = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ---- the CPU restatement (reference arm and cpu_baseline) -----------------------------------------------
def cpu_oracle_pass(n, repeat, seed=1):
    """Time `repeat` single-threaded passes of config 2 over an n-element sample through
    oracle/ref_shaped.c (the collect() loop as rustc would monomorphise it: bounds asserts, push with
    capacity check, separately rounded mul/add, no SIMD).  Returns (seconds per pass, algorithmic bytes)."""
    from helpers import refshaped_lib
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, n).astype(np.float32)
    b = rng.uniform(-1, 1, n).astype(np.float32)
    out = np.empty(n, dtype=np.float32)
    lib = refshaped_lib()
    times = []
    for _ in range(repeat):
        t0 = time.perf_counter()
        st = lib.ref_c2_zip_map(a.ctypes.data, b.ctypes.data, n, out.ctypes.data)
        times.append(time.perf_counter() - t0)
        assert st == 0
    want = a * b + np.float32(1)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    return times, 12 * n


def cpu_ops_table():
    """The other configs on one host core (bounded samples), GB/s on algorithmic bytes."""
    from helpers import refshaped_lib
    lib = refshaped_lib()
    rng = np.random.default_rng(2)
    out = {}

    def timed(fn, nbytes, sample):
        t0 = time.perf_counter()
        assert fn() == 0
        dt = time.perf_counter() - t0
        return {"GB/s": round(nbytes / dt / 1e9, 3), "sample": sample}
    m = 4096
    a = rng.uniform(-1, 1, m * m).astype(np.float32)
    o = np.empty(m * m, np.float32)
    out["c1_transpose_4096x4096_f32"] = timed(lambda: lib.ref_c1_transpose(a.ctypes.data, m, m, o.ctypes.data), 8 * m * m, "full size")
    n3, m3 = 1 << 24, 1 << 28
    src = rng.uniform(-1, 1, m3).astype(np.float32)
    idx = rng.integers(0, m3, n3).astype(np.uint64)
    o3 = np.empty(n3, np.float32)
    out["c3_compose_gather"] = timed(lambda: lib.ref_c3_compose(idx.ctypes.data, n3, src.ctypes.data, m3, o3.ctypes.data), 16 * n3,
                                     "2^24 indices into a 2^28-element source (full: 2^28 into 2^30)")
    I, J, K = 256, 256, 256
    a4 = src[: I * J * K]
    sums = np.empty(I * J, np.float32)
    out["c4a_fold_sum_last_axis"] = timed(lambda: lib.ref_c4_fold(a4.ctypes.data, I, J, K, sums.ctypes.data), 4 * I * J * K + 4 * I * J,
                                          "(256,256,256) of (1024,1024,256)")
    mean = (sums / np.float32(K)).astype(np.float32)
    o4 = np.empty(I * J * K, np.float32)
    out["c4b_broadcast_subtract"] = timed(lambda: lib.ref_c4_sub(a4.ctypes.data, mean.ctypes.data, I, J, K, o4.ctypes.data), 8 * I * J * K + 4 * I * J,
                                          "(256,256,256) of (1024,1024,256)")
    P_, Q, R = 16, 16, 64
    a5 = src[: P_ * Q]
    w5 = src[1000: 1000 + R]
    o5 = np.empty(Q * P_ * Q * P_ * R, np.float32)
    out["c5_rank5_chain"] = timed(lambda: lib.ref_c5_chain(a5.ctypes.data, P_, Q, w5.ctypes.data, R, o5.ctypes.data), 4 * o5.size,
                                  "P=Q=16, R=64 (2^22 outputs) of P=Q=R=64 (2^30)")
    return out


def pin_to_one_core():
    try:
        cores = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {cores[-1]})
    except Exception:
        pass


def run_reference(args, rank):
    if rank != 0:
        return
    pin_to_one_core()
    n = 1 << 26  # bounded sample: 2^26 of the 2^30 elements per step
    times, nbytes = cpu_oracle_pass(n, args.warmup + args.steps)
    timed = times[args.warmup:]
    sec = sum(timed) / len(timed)
    v = nbytes / sec / 1e9
    sample = f"{n} of {N_ELEMS} elements per step (same expression, same generator), single thread pinned to one core"
    emit_line({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "note": "CPU restatement of the reference's single-threaded collect() as monomorphic loops (oracle/ref_shaped.c); the Rust crate cannot be built in this image"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ---- the B200 arm ------------------------------------------------------------------------------------------------------------
def host_memory_available():
    """Bytes of host RAM this process may still take: MemAvailable, capped by the cgroup limit when there is one."""
    avail = 1 << 62
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    avail = int(ln.split()[1]) * 1024
    except OSError:
        pass
    for lim, cur in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                     ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            with open(lim) as f:
                v = f.read().strip()
            if v != "max":
                with open(cur) as f:
                    avail = min(avail, int(v) - int(f.read().strip()))
        except (OSError, ValueError):
            pass
    return max(avail, 0)


_json_fd = None


def keep_stdout_for_json():
    """Libraries print to fd 1 (NCCL writes its version banner there): point fd 1 at stderr for the rest of
    the run and keep the original for the ONE JSON line the driver parses."""
    global _json_fd
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)


def emit_line(obj):
    sys.stdout.flush()
    os.write(_json_fd if _json_fd is not None else 1, (json.dumps(obj) + "\n").encode())


def main():
    keep_stdout_for_json()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-elems", type=int, default=30, help="elements per GPU (default 2^30 = the BASELINE config)")
    ap.add_argument("--no-ops", action="store_true", help="skip the per-op table (configs 1, 3, 4, 5)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import multidimension_b200 as P
    from multidimension_b200 import usize, Array, Scalar, _ffi as F
    from multidimension_b200.runtime import Storage

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = P.Context(local)
    P.set_default_context(ctx)
    stream = torch.cuda.Stream()  # an explicit stream: handle 0 would select the context's own stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    n = 1 << args.log2_elems

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def dev_array(I, size, t, T):
        return Array.from_device(I, size, t.data_ptr(), T, ctx=ctx, keep=t)

    def out_storage(t, dtype):
        return Storage.wrap_device(ctx, dtype, t.numel(), t.data_ptr(), keep=t)

    def time_launches(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ctx.sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, ctx.launch_count() - l0

    # ---- config 2, device-resident ---------------------------------------------------------------------
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0001 + rank)
    ta = torch.empty(n, device="cuda", dtype=torch.float32).uniform_(-1, 1, generator=g)
    tb = torch.empty(n, device="cuda", dtype=torch.float32).uniform_(-1, 1, generator=g)
    tout = torch.empty(n, device="cuda", dtype=torch.float32)
    a, b = dev_array(usize, n, ta, "f32"), dev_array(usize, n, tb, "f32")
    view = a.zip(b).map(lambda p: p[0] * p[1] + np.float32(1))
    plan_text = view.describe()
    st_out = out_storage(tout, F.F32)
    # parity gate before any number: bit-exact against separately rounded mul and add
    view.collect(out=st_out)
    want = ta * tb
    want += 1.0
    assert torch.equal(tout.view(torch.int32), want.view(torch.int32)), "config 2 result is not bit-exact"
    del want

    clocks = ClockSampler(local)
    clocks.start()
    ms_step, launches = time_launches(lambda: view.collect(out=st_out, flags=F.COLLECT_ASYNC), args.steps, args.warmup)
    clocks.stop()
    alg_bytes = 12 * n
    per_gpu = alg_bytes / (ms_step * 1e-3) / 1e9
    value = per_gpu * world
    peak, peak_src = measured_peak()

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "elements_per_gpu": n, "algorithmic_bytes_per_step_per_gpu": alg_bytes,
                   "parallelism": f"outermost-index shards x{world}, no collective", "kernel": plan_text,
                   "l2": "operands (8 GiB) and result (4 GiB) are far larger than the 126 MB L2; no flush needed",
                   "parity": "bit-exact vs separately rounded f32 mul/add, checked in this run"},
        "clocks": clocks.summary(),
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                     "frac_of_nominal_8000": per_gpu / 8000.0, "peak_source": peak_src,
                     "kernel": "k_eval<SigMulAddCF32, u32, V=8, R1>", "traffic": known_traffic("k_eval_SigMulAddCF32")},
    }

    # ---- e2e: host buffers through mdim_collect_host ---------------------------------------------------------------------
    if not args.no_e2e:
        # Every rank pins 12 bytes per element of HOST memory.  One GPU runs the full 2^30-element config; with N ranks
        # on one box the per-rank sample is bounded (PCIe-bound either way) so that the box's RAM is never at risk.
        ne = n if world == 1 else min(n, 1 << 28)
        while ne > (1 << 24) and 12 * ne * world * 1.5 > host_memory_available():
            ne //= 2
        ha, hb, ho = (Storage.pinned(ctx, F.F32, ne) for _ in range(3))
        ctx.download(ha.host, ta.data_ptr())
        ctx.download(hb.host, tb.data_ptr())
        hview = Array(usize, ne, ha, "f32").zip(Array(usize, ne, hb, "f32")).map(lambda p: p[0] * p[1] + np.float32(1))
        e2e_steps = args.e2e_steps or max(1, min(args.steps, 5))
        hview.collect(out=ho)  # warm-up: grows the staging arena
        barrier()
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hview.collect(out=ho)
        barrier()
        sec = (time.perf_counter() - t0) / e2e_steps
        if world > 1:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        chk = torch.from_numpy(ho.host[: 1 << 22]).cuda()
        assert torch.equal(chk.view(torch.int32), tout[: 1 << 22].view(torch.int32)), "e2e result differs from the device-resident result"
        result["e2e"] = {"value": 12 * ne * world / sec / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * ne, "d2h_bytes_per_step": 4 * ne,
                         "ms_per_step": sec * 1e3, "steps": e2e_steps, "kernels_per_step": (ctx.launch_count() - l0) // e2e_steps,
                         "pcie_gbs": 12 * ne / sec / 1e9, "elements_per_gpu": ne}
        del ha, hb, ho, hview
        import gc
        gc.collect()  # the 12 GiB of pinned memory must be released now, not inside a later timed region
        torch.cuda.synchronize()

    # ---- the other BASELINE configs, one line each (N=1 only) ------------------------------------------------------------------
    if not args.no_ops and world == 1:
        result["ops"] = bench_ops(ctx, torch, ta, tout, dev_array, out_storage, time_launches, peak, args)

    # ---- the ops with a real exchange step, N > 1 ----------------------------------------------------------------------------------
    if not args.no_ops and world > 1:
        multi = bench_multi_ops(ctx, torch, dist, rank, world, ta, tout, dev_array, out_storage, time_launches, barrier)
        if rank == 0:
            result["ops_multi"] = multi

    # ---- CPU restatement beside it (rank 0, N=1) -----------------------------------------------------------------------------------
    if not args.no_cpu and world == 1 and rank == 0:
        pin_to_one_core()
        sample_n = 1 << 28
        times, nbytes = cpu_oracle_pass(sample_n, 2)
        result["cpu_baseline"] = {"value": nbytes / min(times) / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                                  "sample": f"{sample_n} of {N_ELEMS} elements of config 2, 2 passes (best), single thread; "
                                            f"oracle/ref_shaped.c restates the reference's collect() loop (no Rust toolchain in this image; "
                                            f"the reference is single-threaded by construction)",
                                  "host_cores_available": os.cpu_count(), "ops": cpu_ops_table()}

    if rank == 0:
        emit_line(result)
    if world > 1:
        dist.destroy_process_group()


def bench_ops(ctx, torch, ta, tout, dev_array, out_storage, time_launches, peak, args):
    """Configs 1, 3, 4, 5 of BASELINE.json at full size: GB/s on algorithmic bytes, each parity-gated."""
    import multidimension_b200 as P
    from multidimension_b200 import usize, Array, Scalar, Add, fold_rows, _ffi as F
    ops = {}
    steps = max(5, min(args.steps, 20))

    def line(name, alg_bytes, ms, plan, **extra):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        ops[name] = dict({"GB/s": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 4), "frac_of_nominal_8000": round(gbs / 8000.0, 4),
                          "ms": round(ms, 5), "algorithmic_bytes": alg_bytes, "kernel": plan}, **extra)

    # C1: 4096x4096 f32 transpose; 16 rotating source/destination pairs (2 GiB) keep it out of L2
    R = 16
    n1 = 4096
    srcs = ta[: R * n1 * n1].view(R, n1 * n1)
    dsts = tout[: R * n1 * n1].view(R, n1 * n1)
    views = [dev_array((usize, usize), (n1, n1), srcs[k], "f32").transpose((), usize, usize, ()) for k in range(R)]
    outs = [out_storage(dsts[k], F.F32) for k in range(R)]
    views[0].collect(out=outs[0])
    assert torch.equal(dsts[0].view(n1, n1), srcs[0].view(n1, n1).t()), "transpose mismatch"
    state = {"k": 0}

    prepared = [views[k].prepare(out=outs[k], flags=F.COLLECT_ASYNC) for k in range(R)]  # lowered once: a 20 us kernel must not wait for Python

    def tr():
        k = state["k"] = (state["k"] + 1) % R
        prepared[k].run()
    ms, _ = time_launches(tr, steps * R, R)
    line("c1_transpose_4096x4096_f32", 2 * 4 * n1 * n1, ms, views[0].describe(), l2="16 rotating buffer pairs (2 GiB total)")

    # C3: compose — 2^28 uniform random usize indices into the 2^30-element f32 Array
    n3 = 1 << 28
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0003)
    tidx = torch.randint(0, ta.numel(), (n3,), device="cuda", dtype=torch.int64, generator=g)
    idx = dev_array(usize, n3, tidx, usize)
    src = dev_array(usize, ta.numel(), ta, "f32")
    v3 = idx.compose(src)
    o3 = out_storage(tout[:n3], F.F32)
    v3.collect(out=o3)
    assert torch.equal(tout[:n3], ta[tidx]), "gather mismatch"
    ms, _ = time_launches(lambda: v3.collect(out=o3, flags=F.COLLECT_ASYNC), steps, 3)
    line("c3_compose_gather_2^28_from_2^30", 8 * n3 + 4 * n3 + 4 * n3, ms, v3.describe(),
         sector_bytes=8 * n3 + 4 * n3 + 32 * n3, sector_GBs=round((8 + 4 + 32) * n3 / (ms * 1e-3) / 1e9, 1))
    del tidx, idx, v3

    # C4: (1024,1024,256) f32: sum over the last index (sequential order) and subtract the mean
    shape = (1024, 1024, 256)
    n4 = shape[0] * shape[1] * shape[2]
    t4 = ta[:n4]
    t4.uniform_(0, 1)
    a4 = dev_array((usize, usize, usize), shape, t4, "f32")
    sums = fold_rows(a4, (usize, usize), usize, Add, np.float32(0))
    tsum = torch.empty(shape[0] * shape[1], device="cuda", dtype=torch.float32)
    osum = out_storage(tsum, F.F32)
    sums.collect(out=osum)
    ref64 = t4.view(-1, 256).double().sum(dim=1)
    rel = ((tsum.double() - ref64).abs() / ref64.abs()).max().item()  # informational: error of the SEQUENTIAL f32 sum itself
    rows = torch.arange(0, shape[0] * shape[1], 251, device="cuda")      # parity: bit-exact vs a sequential f32 sum on the host
    seq = np.add.accumulate(t4.view(-1, 256)[rows].cpu().numpy(), axis=1, dtype=np.float32)[:, -1]
    assert np.array_equal(seq.view(np.uint32), tsum[rows].cpu().numpy().view(np.uint32)), "fold is not in sequential order"
    ms, _ = time_launches(lambda: sums.collect(out=osum, flags=F.COLLECT_ASYNC), steps, 3)
    line("c4a_fold_sum_last_axis", 4 * n4 + 4 * shape[0] * shape[1], ms, sums.describe(), max_rel_err_vs_f64=rel)
    fused = a4 - (sums / Scalar(256.0, "f32")).iso((usize, usize, ()))
    o4 = out_storage(tout[:n4], F.F32)
    fused.collect(out=o4)
    want = t4.view(-1, 256) - (tsum / 256.0).unsqueeze(1)
    assert torch.equal(tout[:n4].view(-1, 256), want), "fused fold+subtract mismatch"
    del want
    ms, _ = time_launches(lambda: fused.collect(out=o4, flags=F.COLLECT_ASYNC), steps, 3)
    line("c4c_fold_mean_subtract_fused", 8 * n4, ms, fused.describe())
    means = dev_array((usize, usize), shape[:2], tsum, "f32")
    sub = a4 - means.iso((usize, usize, ()))
    ms, _ = time_launches(lambda: sub.collect(out=o4, flags=F.COLLECT_ASYNC), steps, 3)
    line("c4b_broadcast_subtract", 8 * n4 + 4 * shape[0] * shape[1], ms, sub.describe())

    # C5: rank-5 chain transpose -> diagonal -> broadcast -> map, P=Q=R=64: 2^30 outputs, write-bound
    Pn = Qn = Rn = 64
    ta5 = torch.empty(Pn * Qn, device="cuda", dtype=torch.float32).uniform_(-1, 1)
    tw5 = torch.empty(Rn, device="cuda", dtype=torch.float32).uniform_(-1, 1)
    a5 = dev_array((usize, usize), (Pn, Qn), ta5, "f32")
    w5 = dev_array(usize, Rn, tw5, "f32")
    v5 = (a5.transpose((), usize, usize, ()).diagonal(np.float32(0)).iso((((usize, usize), (usize, usize)), ()))
          .zip(w5.iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))
    o5 = out_storage(tout, F.F32)
    v5.collect(out=o5)
    eye = torch.eye(Pn * Qn, device="cuda", dtype=torch.float32)
    t_qp = ta5.view(Pn, Qn).t().reshape(-1)
    d = (eye * t_qp.unsqueeze(0))  # [(q,p),(q',p')]
    want5 = d.reshape(-1, 1) * tw5.view(1, Rn)
    want5 += 1.0
    assert torch.equal(tout.view(-1, Rn), want5), "rank-5 chain mismatch"
    del want5, d, eye
    ms, _ = time_launches(lambda: v5.collect(out=o5, flags=F.COLLECT_ASYNC), steps, 3)
    line("c5_rank5_transpose_diagonal_broadcast_map", 4 * (1 << 30) + 4 * Pn * Qn + 4 * Rn, ms, v5.describe())
    return ops


def bench_multi_ops(ctx, torch, dist, rank, world, ta, tout, dev_array, out_storage, time_launches, barrier):
    """N > 1 only (SURVEY.md §8e rows with a collective): compose() onto a SHARDED source — peer-mapped
    reads over NVLink vs NCCL all-gather + local gather — and a fold over the sharded axis + all-reduce.
    Aggregate GB/s on algorithmic bytes, device time, max over ranks."""
    import multidimension_b200 as P
    from multidimension_b200 import usize, Array, Add, fold_rows, _ffi as F
    from multidimension_b200.runtime import Storage
    from multidimension_b200 import sharding
    out = {}
    n_src = 1 << 30
    n_idx = (1 << 28) // world          # this rank's block of the index Array
    block = n_src // world              # this rank's block of the source
    src_block = ta[:block]
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0003 + rank)
    tidx = torch.randint(0, n_src, (n_idx,), device="cuda", dtype=torch.int64, generator=g)
    idx = dev_array(usize, n_idx, tidx, usize)
    o = out_storage(tout[:n_idx], F.F32)
    # (a) peer-mapped source: every rank's block is IPC-mapped into every other rank
    local = Storage.wrap_device(ctx, F.F32, block, src_block.data_ptr(), keep=src_block)
    peers = sharding.peer_source(local, n_src, ctx=ctx)
    v = idx.compose(Array(usize, n_src, peers, "f32"))
    v.collect(out=o)
    full = [torch.empty_like(src_block) for _ in range(world)]
    dist.all_gather(full, src_block)
    full = torch.cat(full)
    assert torch.equal(tout[:n_idx], full[tidx]), "peer-sharded gather mismatch"
    ms, _ = time_launches(lambda: v.collect(out=o, flags=F.COLLECT_ASYNC), 5, 3)
    alg = 16 * n_idx * world
    out["compose_sharded_source_peer_mapped"] = {"GB/s": round(alg / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4), "kernel": v.describe(),
                                                 "note": "gather kernel reads the owning peer's HBM over NVLink; no collective"}
    # (b) NCCL all-gather of the source, then a local gather (what the north star names)
    full_buf = torch.empty(n_src, device="cuda", dtype=torch.float32)
    v2 = idx.compose(dev_array(usize, n_src, full_buf, "f32"))

    def ag_then_gather():
        dist.all_gather_into_tensor(full_buf, src_block)
        v2.collect(out=o, flags=F.COLLECT_ASYNC)
    ag_then_gather()
    torch.cuda.synchronize()
    assert torch.equal(tout[:n_idx], full[tidx]), "all-gather + gather mismatch"
    ms, _ = time_launches(ag_then_gather, 5, 3)
    out["compose_sharded_source_allgather"] = {"GB/s": round(alg / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4),
                                               "note": "ncclAllGather of the 4 GiB source + local gather"}
    del full, tidx
    # (e) transpose of a ROW-SHARDED source, every rank collecting its block of the transposed rows (SURVEY.md §8e,
    #     transpose row): the all-to-all is fused into the tiled transpose kernel — each tile is loaded from the
    #     owning peer's HBM over NVLink — vs NCCL all-gather of the source + a local transpose.
    M = 16384
    tblock = M * M // world                         # this rank's rows: the first M*M/world elements of its `ta`
    tpeers = sharding.PeerStorage(F.F32, M * M, peers.peers, tblock, keep=src_block, ctx=ctx)
    vt = sharding.shard_view(Array((usize, usize), (M, M), tpeers, "f32").transpose((), usize, usize, ()), rank, world)
    ot = out_storage(tout[:tblock], F.F32)
    prep_t = vt.prepare(out=ot, flags=F.COLLECT_ASYNC)
    prep_t.run()
    whole = full_buf[: M * M]
    dist.all_gather_into_tensor(whole, ta[:tblock])
    lo_t, hi_t = sharding.shard_bounds(M, world, rank)
    assert torch.equal(tout[:tblock].view(hi_t - lo_t, M), whole.view(M, M).t()[lo_t:hi_t]), "peer-sharded transpose mismatch"
    ms, _ = time_launches(prep_t.run, 5, 3)
    out["transpose_sharded_source_peer_mapped"] = {"GB/s": round(8 * M * M / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4), "kernel": vt.describe(),
                                                   "note": f"{M}x{M} f32 sharded by rows; each tile is read from the owning peer over NVLink; no collective"}
    vt2 = sharding.shard_view(dev_array((usize, usize), (M, M), whole, "f32").transpose((), usize, usize, ()), rank, world)
    prep_t2 = vt2.prepare(out=ot, flags=F.COLLECT_ASYNC)

    def ag_then_transpose():
        dist.all_gather_into_tensor(whole, ta[:tblock])
        prep_t2.run()
    ag_then_transpose()
    torch.cuda.synchronize()
    assert torch.equal(tout[:tblock].view(hi_t - lo_t, M), whole.view(M, M).t()[lo_t:hi_t]), "all-gather + transpose mismatch"
    ms, _ = time_launches(ag_then_transpose, 5, 3)
    out["transpose_sharded_source_allgather"] = {"GB/s": round(8 * M * M / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4),
                                                 "note": "ncclAllGather of the 1 GiB source + local transpose of this rank's block"}
    del full_buf, whole
    # (d) BASELINE config 5 — the rank-5 transpose -> diagonal -> broadcast -> map chain, 2^30 outputs — sharded
    #     over the ranks along its outermost index (strong scaling: every rank writes 2^30 / world elements)
    Pn = Qn = Rn = 64
    g5 = torch.Generator(device="cuda")
    g5.manual_seed(0x5EED0005)  # the same small operands on every rank (replicated)
    ta5 = torch.empty(Pn * Qn, device="cuda", dtype=torch.float32).uniform_(-1, 1, generator=g5)
    tw5 = torch.empty(Rn, device="cuda", dtype=torch.float32).uniform_(-1, 1, generator=g5)
    v5 = (dev_array((usize, usize), (Pn, Qn), ta5, "f32").transpose((), usize, usize, ()).diagonal(np.float32(0))
          .iso((((usize, usize), (usize, usize)), ())).zip(dev_array(usize, Rn, tw5, "f32").iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))
    mine = sharding.shard_view(v5, rank, world)
    n5 = mine.len()
    o5 = out_storage(tout[:n5], F.F32)
    prep5 = mine.prepare(out=o5, flags=F.COLLECT_ASYNC)
    prep5.run()
    torch.cuda.synchronize()
    q_lo, q_hi = sharding.shard_bounds(Qn, world, rank)
    eye = torch.eye(Pn * Qn, device="cuda", dtype=torch.float32)[q_lo * Pn:q_hi * Pn]
    want5 = (eye * ta5.view(Pn, Qn).t().reshape(-1).unsqueeze(0)).reshape(-1, 1) * tw5.view(1, Rn)
    want5 += 1.0
    assert torch.equal(tout[:n5].view(-1, Rn), want5), "sharded rank-5 chain mismatch"
    del want5, eye
    ms, _ = time_launches(prep5.run, 10, 3)
    out["c5_rank5_chain_sharded_outermost"] = {"GB/s": round(4 * (1 << 30) / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4), "kernel": mine.describe(),
                                               "note": f"BASELINE configs[4]: 2^30 outputs cut into {world} blocks of the outermost index, no collective (strong scaling)"}
    # (c) fold over the SHARDED (outermost) axis + all-reduce: a (1024,1024,256) f32 Array sharded on axis 0
    I, J, K = 1024 // world, 1024, 256
    t4 = ta[: I * J * K]
    t4.uniform_(0, 1)  # config 4's data (sums far from zero, so a relative error means something)
    a4 = dev_array((usize, usize, usize), (I, J, K), t4, "f32")
    part_view = fold_rows(a4.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0))
    tpart = torch.empty(J * K, device="cuda", dtype=torch.float32)
    opart = out_storage(tpart, F.F32)

    def fold_allreduce():
        part_view.collect(out=opart, flags=F.COLLECT_ASYNC)
        dist.all_reduce(tpart)
    fold_allreduce()
    torch.cuda.synchronize()
    ref = t4.view(I, J * K).double().sum(dim=0)
    dist.all_reduce(ref)
    rel = ((tpart.double() - ref).abs() / ref.abs()).max().item()
    assert rel < 1e-5, f"sharded-axis fold error {rel}"
    ms, _ = time_launches(fold_allreduce, 5, 3)
    alg = (4 * I * J * K) * world + 4 * J * K
    out["fold_sharded_axis_allreduce"] = {"GB/s": round(alg / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4), "kernel": part_view.describe(),
                                          "max_rel_err_vs_f64": rel, "note": "per-rank partial fold + ncclAllReduce(sum) of 1 MiB"}
    # (c') the same fold with NO reassociation: every rank folds ITS block of the (J, K) result over the whole
    #      sharded axis, reading the peers' blocks in index order -> bit-identical to the unsharded sequential fold.
    barrier()
    fpeers = sharding.PeerStorage(F.F32, I * world * J * K, peers.peers, I * J * K, keep=t4, ctx=ctx)
    whole3 = Array((usize, usize, usize), (I * world, J, K), fpeers, "f32")
    exact_view = sharding.shard_view(
        fold_rows(whole3.transpose((), (usize, usize), usize, ()).iso(((usize, usize), usize)), (usize, usize), usize, Add, np.float32(0)), rank, world)
    j_lo, j_hi = sharding.shard_bounds(J, world, rank)
    texact = torch.empty((j_hi - j_lo) * K, device="cuda", dtype=torch.float32)
    prep_x = exact_view.prepare(out=out_storage(texact, F.F32), flags=F.COLLECT_ASYNC)
    prep_x.run()
    torch.cuda.synchronize()
    gathered = [torch.empty(I * (j_hi - j_lo) * K, device="cuda", dtype=torch.float32) for _ in range(world)]
    mine_cols = [t4.view(I, J, K)[:, sharding.shard_bounds(J, world, r)[0]:sharding.shard_bounds(J, world, r)[1], :].contiguous().view(-1) for r in range(world)]
    for r in range(world):  # rank r receives every rank's rows of ITS column block
        dist.gather(mine_cols[r], gathered if rank == r else None, dst=r)
    seq = torch.zeros((j_hi - j_lo) * K, device="cuda", dtype=torch.float32)
    for blk in gathered:
        for i in range(I):
            seq += blk.view(I, -1)[i]
    assert torch.equal(texact, seq), "peer-mapped fold over the sharded axis is not bit-exact"
    del gathered, mine_cols
    ms, _ = time_launches(prep_x.run, 5, 3)
    out["fold_sharded_axis_peer_mapped_bit_exact"] = {"GB/s": round(alg / (ms * 1e-3) / 1e9, 1), "ms": round(ms, 4), "kernel": exact_view.describe(),
                                                      "note": "each rank folds its block of the result over ALL ranks' rows in index order (reads over NVLink): "
                                                              "bit-identical to the sequential reference, no collective"}
    peers.close()
    return out


if __name__ == "__main__":
    main()
