"""CPU: radd() (swegl_b200/csrc/radd.h), the exact O(1) replay of `s = RN(s + c)` k times, against the k-step
loop on adversarial inputs.  The same header is compiled into the CUDA kernels."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_radd_bruteforce():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "radd_test")
        subprocess.check_call(["gcc", "-O2", "-msse2", "-mfpmath=sse", "-ffp-contract=off", "-I", os.path.join(ROOT, "swegl_b200", "csrc"),
                               "-o", exe, os.path.join(ROOT, "tests", "radd_bruteforce.c"), "-lm"])
        out = subprocess.check_output([exe, "150000"], text=True)
    assert "0 failures" in out, out
