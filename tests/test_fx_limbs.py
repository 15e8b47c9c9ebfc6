"""CPU: the arithmetic of the lean kernel's fixed-point double sums (csrc/fx.h) — classification of special values, carries and borrows
through the three limbs, order independence, head room, and the error bound on a dozen value distributions (tests/cpp/fx_check.cpp:
the limb updates of fx_add with plain adds in place of the shared-memory atomics)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fx_limb_arithmetic(tmp_path):
    exe = str(tmp_path / "fx_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "csrc"), os.path.join(ROOT, "tests", "cpp", "fx_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fx_check ok" in out.stdout, out.stdout[-3000:]
