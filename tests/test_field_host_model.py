"""CPU suite: the limb-level multiplication schedules of kzg_b200/csrc/field.cuh, compiled for the HOST with the PTX
carry flag emulated (tests/cpp/field_host_model.cpp), against an independent 64-bit CIOS Montgomery product.
Covers the interleaved product the kernels use, the dedicated square, the Karatsuba wide product, the ripple-free
word-serial reduction and the lazy a*b - c*d (Fq and Fr).  New arithmetic is proven bit-exact here before it ever runs
on a GPU; the device-side twin is tests/test_gpu_parity.py::test_field_ops_bit_exact and test ops 9-11."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_field_schedules_match_big_integer_model(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "field_host_model")
    subprocess.run([gxx, "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "field_host_model.cpp")],
                   check=True, capture_output=True, text=True)
    r = subprocess.run([exe, "40000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Fq: 40000 iterations, 0 mismatches" in r.stdout and "Fr: 40000 iterations, 0 mismatches" in r.stdout
