"""Compiles tests/cpp/test_view.cpp (the reference's doctests and the BASELINE configs written against
the C++ host mirror include/mdim/view.hpp) and runs it: on CPU through the C oracle, on a B200
through mdim_collect_host of the product library."""
import os
import subprocess

import pytest

from helpers import ROOT, ORACLE_DIR, oracle_lib
import multidimension_b200 as P


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "test_view"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-o", str(out), os.path.join(ROOT, "tests", "cpp", "test_view.cpp"), "-ldl"], check=True)
    return str(out)


def test_cpp_mirror_through_oracle(binary):
    oracle_lib()
    r = subprocess.run([binary, "oracle", os.path.join(ORACLE_DIR, "_build", "libmdim_oracle.so")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_through_c_abi(binary):
    r = subprocess.run([binary, "gpu", P.LIB_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
