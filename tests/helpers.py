"""Shared test plumbing: three executors for one lowered product view.

  oracle_collect  — the C oracle (oracle/mdim_oracle.c) over the emitted descriptor, host pointers
  emu_collect     — the product's planner + per-thread evaluator compiled for the host (tests/emu)
  view.collect()  — the real thing: sm_100a kernels through the C ABI (GPU tests only)

The first two are checkers; nothing in the product imports them.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

import multidimension_b200 as P
from multidimension_b200 import _ffi as F
from multidimension_b200 import lowering as L
from multidimension_b200.runtime import NP_OF
from multidimension_b200.view import _flat, _T_leaves, _build_like

_libs = {}


def _build():
    need = [os.path.join(ORACLE_DIR, "_build", n) for n in ("libmdim_oracle.so", "libmdim_emu.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def oracle_lib():
    if "oracle" not in _libs:
        _build()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_build", "libmdim_oracle.so"))
        lib.mdim_oracle_collect.restype = C.c_int
        lib.mdim_oracle_collect.argtypes = [C.POINTER(F.Expr), C.c_void_p, C.POINTER(F.ErrorInfo)]
        lib.mdim_oracle_sizeof_node.restype = C.c_size_t
        lib.mdim_oracle_sizeof_expr.restype = C.c_size_t
        _libs["oracle"] = lib
    return _libs["oracle"]


def emu_lib():
    if "emu" not in _libs:
        _build()
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_build", "libmdim_emu.so"))
        lib.mdim_emu_collect.restype = C.c_int
        lib.mdim_emu_collect.argtypes = [C.POINTER(F.Expr), C.c_void_p, C.c_uint32, C.POINTER(F.ErrorInfo), C.c_char_p, C.c_size_t]
        _libs["emu"] = lib
    return _libs["emu"]


class CheckerPanic(Exception):
    def __init__(self, status, info):
        super().__init__(info.message.decode())
        self.status, self.info = status, info


def _run_host_tuple(view, runner):
    """A tuple-typed view through ONE descriptor with an MDIM_NODE_TUPLE root: runner(em, [out arrays], info)."""
    groups, value = view._lower()
    em = L.emit(list(L.flatten_value(value)), _flat(groups), "host")
    outs = [np.zeros(em.out_len, dtype=NP_OF[dt]) for dt in em.out_dtypes]
    info = F.ErrorInfo()
    st = runner(em, outs, info)
    if st != F.OK:
        raise CheckerPanic(st, info)
    return outs


def oracle_collect_tuple(view):
    lib = oracle_lib()
    lib.mdim_oracle_collect_tuple.restype = C.c_int

    def run(em, outs, info):
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        return lib.mdim_oracle_collect_tuple(C.byref(em.expr), ptrs, len(outs), C.byref(info))
    return _shape_result(view, _run_host_tuple(view, run))


def emu_collect_tuple(view, flags=0):
    lib = emu_lib()
    lib.mdim_emu_collect_tuple.restype = C.c_int

    def run(em, outs, info):
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        return lib.mdim_emu_collect_tuple(C.byref(em.expr), ptrs, len(outs), C.c_uint32(flags), C.byref(info), None, C.c_size_t(0))
    return _shape_result(view, _run_host_tuple(view, run))


def _run_host(view, runner, split=True):
    """One descriptor per scalar leaf; an expression beyond what ONE fused kernel takes (MDIM_ERR_UNSUPPORTED: nodes, operands, a second
    fold) is split through dense temporaries exactly as View.collect does (lowering.split_for_limits), unless split=False."""
    groups, value = view._lower()
    axes = _flat(groups)

    def run_node(node, node_axes):
        em = L.emit(node, node_axes, "host")
        out = np.zeros(em.out_len, dtype=NP_OF[em.out_dtype])
        info = F.ErrorInfo()
        st = runner(em, out, info)
        if st != F.OK:
            raise CheckerPanic(st, info)
        return out
    outs = []
    for node in L.flatten_value(value):
        try:
            outs.append(run_node(node, axes))
        except (F.MdimError, CheckerPanic) as e:
            if not split or e.status != F.ERR_UNSUPPORTED:
                raise
            from multidimension_b200.runtime import Storage
            node2 = L.split_for_limits(node, axes, lambda sub, sub_axes: Storage.from_host(sub.dtype, run_node(sub, sub_axes)))
            if node2 is None:
                raise
            outs.append(run_node(node2, axes))
    return outs


def _shape_result(view, outs):
    leaves_T = _T_leaves(view.T)
    outs = [o.astype(bool) if t is bool else o for o, t in zip(outs, leaves_T)]
    if isinstance(view.T, tuple):
        return [_build_like(view.T, [o[k].item() for o in outs]) for k in range(len(outs[0]))]
    return outs[0]


def oracle_collect(view, split=True):
    lib = oracle_lib()
    return _shape_result(view, _run_host(view, lambda em, out, info: lib.mdim_oracle_collect(C.byref(em.expr), out.ctypes.data, C.byref(info)), split))


def emu_collect(view, flags=0, describe=None, split=True):
    lib = emu_lib()

    def run(em, out, info):
        buf = C.create_string_buffer(256)
        st = lib.mdim_emu_collect(C.byref(em.expr), out.ctypes.data, flags, C.byref(info), buf, 256)
        if describe is not None:
            describe.append(buf.value.decode())
        return st
    return _shape_result(view, _run_host(view, run, split))


def as_list(x):
    if isinstance(x, np.ndarray):
        return x.tolist()
    return list(x)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def assert_same_bits(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == want.dtype, f"{what}: dtype {got.dtype} != {want.dtype}"
    if not np.array_equal(bits(got), bits(want)):
        differ = got.view(f"u{got.dtype.itemsize}") != want.view(f"u{want.dtype.itemsize}")
        if got.dtype.kind == "f":  # NaN payloads are unspecified in Rust (and differ between x86 and sm_100): NaN == NaN here
            differ &= ~(np.isnan(got) & np.isnan(want))
        bad = np.flatnonzero(differ)
        if bad.size == 0:
            return
        k = int(bad[0])
        raise AssertionError(f"{what}: {bad.size} of {got.size} elements differ; first at {k}: got {got[k]!r}, want {want[k]!r}")


def refshaped_lib(variant=""):
    """oracle/ref_shaped.c: the five BASELINE configs as rustc-shaped monomorphic loops (CPU baseline).
    variant "" = -O2 -fno-tree-vectorize (README: "not SIMD optimized"); "_o3" = -O3 -march=x86-64-v3, vectoriser on."""
    key = "ref" + variant
    if key not in _libs:
        path = os.path.join(ORACLE_DIR, "_build", f"libmdim_refshaped{variant}.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
        lib = C.CDLL(path)
        P_, U = C.c_void_p, C.c_uint64
        lib.ref_c2_zip_map.argtypes = [P_, P_, U, P_]
        lib.ref_c1_transpose.argtypes = [P_, U, U, P_]
        lib.ref_c3_compose.argtypes = [P_, U, P_, U, P_]
        lib.ref_c4_fold.argtypes = [P_, U, U, U, P_]
        lib.ref_c4_sub.argtypes = [P_, P_, U, U, U, P_]
        lib.ref_c5_chain.argtypes = [P_, U, U, P_, U, P_]
        for f in ("ref_c2_zip_map", "ref_c1_transpose", "ref_c3_compose", "ref_c4_fold", "ref_c4_sub", "ref_c5_chain"):
            getattr(lib, f).restype = C.c_int
        _libs[key] = lib
    return _libs[key]
