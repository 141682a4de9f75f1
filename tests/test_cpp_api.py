"""Builds tests/cpp/test_api.cpp (the reference's fixed-scene tests written against the C++ host mirror
include/bvh_b200.hpp) with g++, links libbvh_b200.so, and runs it on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "build", "test_api")


def _build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    lib_dir = os.path.join(ROOT, "bvh_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_api.cpp"),
           "-L", lib_dir, "-lbvh_b200", f"-Wl,-rpath,{lib_dir}", "-o", EXE]
    subprocess.run(cmd, check=True)


def test_cpp_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cpp_reference_fixed_scene_tests():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference fixed-scene tests passed" in r.stdout
