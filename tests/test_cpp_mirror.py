"""include/trueno.hpp — the C++17 host mirror of trueno's Vector / Matrix / TruenoError / GpuCommandBatch over the C ABI.
Builds tests/cpp/test_trueno_hpp.cpp with g++ against the in-tree library and runs it: the validation half on any
box (every check fails before the device is touched), the KAT half on the B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "test_trueno_hpp")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "test_trueno_hpp.cpp")
    lib_dir = os.path.join(ROOT, "trueno_b200")
    if (not os.path.exists(EXE)) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(ROOT, "include", "trueno.hpp"))):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                        "-L", lib_dir, "-ltrueno_cuda", f"-Wl,-rpath,{lib_dir}", "-Wl,-rpath,/usr/local/cuda/lib64"],
                       check=True, capture_output=True, text=True)
    return EXE


def test_cpp_mirror_validation_contract(trn):   # `trn` makes sure the library is built
    r = subprocess.run([build(), "validation"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_kats_on_gpu(trn):
    r = subprocess.run([build(), "all"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
