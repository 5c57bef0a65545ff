"""The product's restatement of glibc sinf/cosf (csrc/glibc_sincosf.h), compiled for the host and compared with libm
for EVERY float in [0, 2*pi] -- the only arguments computeOrbDescriptor can produce (ORBextractor.cc:113-115)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exhaustive_against_libm(tmp_path):
    exe = str(tmp_path / "sincosf_exhaustive")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "vi-orb-slam-icra2018_b200", "csrc"),
                           "-o", exe, os.path.join(ROOT, "tests", "sincosf_exhaustive.cpp"), "-lm"])
    out = subprocess.run([exe, "0", "6.2832"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0" in out.stdout and "checked 1086918650" in out.stdout
