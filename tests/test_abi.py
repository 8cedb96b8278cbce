"""The C-ABI library loads on a CPU-only box, exports every symbol include/mdim.h declares, agrees with
the ctypes mirror on struct layout, refuses to compute without a device, and plans the BASELINE chains
onto the intended kernels (planning needs no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import ROOT, oracle_lib
import multidimension_b200 as P
from multidimension_b200 import _ffi as F, usize, Array, Scalar, Add, fold_rows


def test_every_declared_symbol_is_exported():
    lib = F.lib()
    header = open(f"{ROOT}/include/mdim.h").read()
    declared = set(re.findall(r"\b(mdim_[a-z_]+)\s*\(", header))
    declared -= {"mdim_node", "mdim_expr"}
    assert declared == {name for name, _r, _a in F.SYMBOLS}, declared ^ {n for n, _r, _a in F.SYMBOLS}
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mdim_abi_version() == F.ABI_VERSION


def test_struct_layout_matches_the_c_compiler():
    o = oracle_lib()
    assert C.sizeof(F.Node) == o.mdim_oracle_sizeof_node()
    assert C.sizeof(F.Expr) == o.mdim_oracle_sizeof_expr()


def test_no_cpu_fallback():
    """Without an sm_100 device the context cannot be created; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(P.MdimError) as e:
        P.Context(0)
    assert e.value.status == F.ERR_CUDA
    a = Array.new(usize, 4, np.arange(4, dtype=np.float32))
    with pytest.raises(P.MdimError):
        (a * a).collect()


def test_planner_picks_the_intended_kernels():
    rng = np.random.default_rng(0)
    f = lambda *shape: rng.uniform(-1, 1, int(np.prod(shape))).astype(np.float32)
    a, b = Array.new(usize, 1 << 12, f(1 << 12)), Array.new(usize, 1 << 12, f(1 << 12))
    assert "SigMulAddCF32 r1" in a.zip(b).map(lambda p: p[0] * p[1] + np.float32(1)).describe()
    assert "SigMulAddCF32 r1" in (a * b + Scalar(1.0, "f32")).describe()
    m = Array.new((usize, usize), (128, 256), f(128, 256))
    assert m.transpose((), usize, usize, ()).describe().startswith("transpose.tile")
    idx = Array.new(usize, 64, rng.integers(0, 1 << 12, 64).astype(np.uint64))
    assert "SigGatherF32" in idx.compose(a).describe()
    c = Array.new((usize, usize, usize), (4, 8, 64), f(4, 8, 64))
    s = fold_rows(c, (usize, usize), usize, Add, np.float32(0))
    assert s.describe().startswith("fold_rows ")
    assert (c - (s / Scalar(64.0, "f32")).iso((usize, usize, ()))).describe().startswith("fold_rows.fused")
    means = Array.new((usize, usize), (4, 8), f(4, 8))
    assert "SigSubBcastF32" in (c - means.iso((usize, usize, ()))).describe()
    w = Array.new(usize, 64, f(64))
    q = Array.new((usize, usize), (4, 4), f(4, 4))
    v5 = (q.transpose((), usize, usize, ()).diagonal(np.float32(0)).iso((((usize, usize), (usize, usize)), ()))
          .zip(w.iso(((), usize))).map(lambda p: p[0] * p[1] + np.float32(1)))
    assert "SigDiagMulAddCF32 r5" in v5.describe()
    # contiguous arrays of any rank and any Iso regrouping collapse to the rank-1 stream kernel
    assert "rank=1+0" in (c * c).describe() and "rank=1+0" in (c.iso(((usize, usize), usize)) * c.iso(((usize, usize), usize))).describe()


def test_runtime_specialisation_compiles_without_a_gpu():
    """Op trees without a pre-built signature are specialised with NVRTC on first use; the compile half
    needs no device, so the CPU suite checks that the embedded evaluator source builds for sm_100a."""
    from multidimension_b200 import lowering as L
    from multidimension_b200.view import _flat
    rng = np.random.default_rng(0)
    a = Array.new((usize, usize), (37, 24), rng.integers(0, 1000, 37 * 24).astype(np.uint64))
    b = Array.new(usize, 24, rng.integers(1, 9, 24).astype(np.uint64))
    f = Array.new(usize, 64, rng.uniform(-1, 1, 64).astype(np.float32))
    views = [(a % b.iso(((), usize))) ^ Scalar(5),                      # integer ops with a broadcast operand
             a.transpose((), usize, usize, ()).map(P.Cast("f32")).map(P.Sqrt),
             (a >> Scalar(3)).diagonal(7),                              # predicates
             (f * f - f).map(P.Abs),
             fold_rows(a, usize, usize, P.BitXor, 0),                   # a fold becomes a real loop
             f.concat(f, (), ()),                                       # masked sides
             b.compose(a.row(usize, usize, 3)) + Scalar(1)]             # gather
    for v in views:
        groups, value = v._lower()
        em = L.emit(value, _flat(groups), "any")
        log = C.create_string_buffer(8000)
        st = F.lib().mdim_jit_check_nodevice(C.byref(em.expr), 0, log, 8000)
        if st == F.ERR_UNSUPPORTED and b"NVRTC" in log.value:
            pytest.skip("NVRTC is not installed")
        assert "specialised on first use" in v.describe()
        assert st == F.OK, log.value.decode()


def test_jit_disk_cache(tmp_path):
    """Compiled cubins are kept on disk (MDIM_JIT_CACHE): a second PROCESS does not pay NVRTC again."""
    import subprocess
    import sys
    import time
    script = r'''
import ctypes as C, sys, time
import numpy as np
import multidimension_b200 as P
from multidimension_b200 import usize, Array, Scalar, _ffi as F, lowering as L
from multidimension_b200.view import _flat
a = Array.new((usize, usize), (37, 24), np.arange(37 * 24, dtype=np.uint64))
v = (a >> Scalar(3)).diagonal(7)
groups, value = v._lower()
em = L.emit(value, _flat(groups), "any")
log = C.create_string_buffer(8000)
t0 = time.perf_counter()
st = F.lib().mdim_jit_check_nodevice(C.byref(em.expr), 0, log, 8000)
print(st, time.perf_counter() - t0, log.value.decode()[:200])
'''
    env = dict(os.environ, MDIM_JIT_CACHE=str(tmp_path), PYTHONPATH=ROOT)
    runs = []
    for _ in range(2):
        p = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        st, sec = p.stdout.split()[:2]
        if int(st) == F.ERR_UNSUPPORTED and "NVRTC" in p.stdout:
            pytest.skip("NVRTC is not installed")
        assert int(st) == F.OK, p.stdout
        runs.append(float(sec))
    cubins = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cubins) == 2, cubins          # the op-sequence kernel and its shape-specialised form
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".tmp")]
    assert runs[1] < 0.5 * runs[0], runs     # the second process reads them back instead of compiling
    del env["MDIM_JIT_CACHE"]                 # opt-in: without the variable nothing is written anywhere
    home = tmp_path / "home"
    home.mkdir()
    env["HOME"] = env["XDG_CACHE_HOME"] = str(home)
    p = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0 and not os.listdir(home)


def test_descriptor_limits_are_declined_cleanly():
    """MDIM_MAX_NODES (48 nodes), the 56-instruction program, MDIM_MAX_RANK (8 axes) and the 12 array operands are hard
    limits of one fused expression: beyond them the answer is MDIM_ERR_UNSUPPORTED with a message — from the host mirror,
    from the planner given a hand-made descriptor, and from the oracle — never a crash or a silently wrong kernel."""
    from helpers import oracle_collect, emu_collect, CheckerPanic
    from multidimension_b200 import lowering as L
    from multidimension_b200.view import _flat
    a = Array.new(usize, 16, np.arange(16, dtype=np.float32))
    v = a
    for k in range(24):  # 1 + 2 * 24 = 49 nodes
        v = v + Scalar(float(k), "f32")
    with pytest.raises(P.Unsupported, match="49 nodes"):
        v.describe()
    ok = a
    for k in range(23):  # 47 nodes: the largest chain of this shape that fits
        ok = ok + Scalar(float(k), "f32")
    assert ok.describe()
    want = np.arange(16, dtype=np.float32)
    for k in range(23):
        want = want + np.float32(k)
    assert np.array_equal(emu_collect(ok), want) and np.array_equal(oracle_collect(ok), want)
    # a hand-made descriptor that lies about its size goes through the C ABI's own validation
    groups, value = ok._lower()
    em = L.emit(value, _flat(groups), "any")
    em.expr.n_nodes = F.MAX_NODES + 1
    buf = C.create_string_buffer(256)
    assert F.lib().mdim_plan_describe_nodevice(C.byref(em.expr), 0, buf, 256) in (F.ERR_UNSUPPORTED, F.ERR_INVALID)
    # rank: 9 position axes
    nine = Array.new((((usize, usize, usize), (usize, usize, usize)), (usize, usize, usize)), (((2, 2, 2), (2, 2, 2)), (2, 2, 2)), np.zeros(512, np.float32))
    with pytest.raises(P.Unsupported, match="9 position axes"):
        nine.describe()
    # operands: 13 distinct Arrays in one expression
    many = a
    for k in range(12):
        many = many + Array.new(usize, 16, np.full(16, k, np.float32))
    with pytest.raises((P.MdimError, CheckerPanic)) as e:
        emu_collect(many, split=False)
    assert e.value.status == F.ERR_UNSUPPORTED


def test_fold_over_the_sharded_axis_has_both_routes_behind_the_c_abi():
    """SURVEY.md 8(e) last row: the bit-exact ring route and the blocked (all-reduce) route are both C-ABI entry points
    with the same argument list (include/mdim.h), so a host picks the route by name, not by linking NCCL."""
    import ctypes as C
    import multidimension_b200 as P
    lib = C.CDLL(P.LIB_PATH)
    header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mdim.h")).read()
    for name in ("mdim_fold_sharded_axis", "mdim_fold_sharded_axis_blocked", "mdim_fold_sharded_axis_status"):
        assert hasattr(lib, name) and f"int {name}(" in header
