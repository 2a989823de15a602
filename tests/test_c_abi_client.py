"""CPU: include/cp360.h is a valid C99 and C++11 header on its own, and a plain C program (tests/c_abi/host_only.c: no
Python, no torch, no CUDA headers) links libcp360.so and gets the reference's worked CubePad example and error codes
through the host-only entry points — the boundary really is a C ABI."""
import os
import shutil
import subprocess

import pytest

import cp360_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "cp360.h")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not installed")


@pytest.mark.parametrize("cc,std,lang", [("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")])
def test_header_is_self_contained(cc, std, lang):
    r = subprocess.run([cc, std, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", lang, HDR],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_plain_c_client(tmp_path):
    lib_dir = os.path.dirname(cp360_b200.LIB_PATH)
    exe = str(tmp_path / "host_only")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_abi", "host_only.c"), "-o", exe, "-L", lib_dir, "-lcp360",
                        "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, "host_only.c failed at check %d: %s" % (r.returncode, r.stderr)
    assert "c-abi host client ok" in r.stdout
