"""CPU: the kernels' fp32 EXP / 2**y / LOG (cable_b200/csrc/cbm_math.cuh) compiled for the host must return exactly
(float)exp((double)x) etc. -- the values the correctly rounded oracle build uses -- on every sampled fp32 argument.
(An exhaustive run, stride 1, over all 6.7e9 arguments in range shows 0 mismatches: DESIGN.md 'Math policy'.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lean_math_matches_libm_rounded_once(tmp_path):
    exe = str(tmp_path / "test_lean_math")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "test_lean_math.cpp")])
    out = subprocess.run([exe, "211"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches=0" in out.stdout and out.stdout.strip().endswith("ok"), out.stdout
