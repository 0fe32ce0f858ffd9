"""The C++ host mirror (include/anemoi_b200.hpp): compiles against the C ABI header on CPU; on a GPU box
the built program reproduces the reference's Jive known answers through the C++ surface."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "mirror_test.cpp")


def build(tmp_path):
    exe = str(tmp_path / "mirror_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-o", exe, SRC,
                           "-L", os.path.join(ROOT, "anemoi_rust_b200"), "-lanemoi_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "anemoi_rust_b200")])
    return exe


def test_cpp_mirror_compiles_and_links(tmp_path):
    assert os.path.exists(build(tmp_path))


@pytest.mark.gpu
def test_cpp_mirror_reproduces_reference_kats(tmp_path, kat):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = dict(l.split("=") for l in out.stdout.split() if "=" in l)
    k = kat["bls12_381"]["anemoi_2_1"]["jive2"]
    assert k["in"][0] == ["0", "0"] and k["in"][1] == ["1", "1"]
    assert int(vals["jive00"], 16) == int(k["out"][0][0])
    assert int(vals["jive11"], 16) == int(k["out"][1][0])
    assert "OK" in out.stdout


def test_c_header_is_valid_c99_and_example_links(tmp_path):
    """include/anemoi_b200.h must be usable from plain C (the cgo / FFI consumers); the example program links."""
    exe = str(tmp_path / "compress_example")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "compress_example.c"), "-o", exe,
                           "-L", os.path.join(ROOT, "anemoi_rust_b200"), "-lanemoi_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "anemoi_rust_b200")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no CUDA device" in out.stdout or "compress[3]" in out.stdout


@pytest.mark.gpu
def test_c_example_reproduces_reference_kats(tmp_path, kat):
    exe = str(tmp_path / "compress_example")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "compress_example.c"), "-o", exe,
                           "-L", os.path.join(ROOT, "anemoi_rust_b200"), "-lanemoi_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "anemoi_rust_b200")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    k = kat["bls12_381"]["anemoi_2_1"]["jive2"]
    got = [int(l.split("= ")[1], 16) for l in out.stdout.splitlines() if l.startswith("compress[")]
    assert got == [int(o[0]) for o in k["out"]]
    assert "k = 4 on Anemoi-2-1 -> -4" in out.stdout
