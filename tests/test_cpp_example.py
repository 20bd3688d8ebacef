"""The C++ host mirror (include/bhw.hpp) and the coe.dat-style dump tool compile against the shared
library; on the GPU box the dump equals the oracle."""
import os
import subprocess

import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
import harness as H

EXE = os.path.join(H.ROOT, "examples", "coe_dump")


def build():
    libdir = os.path.dirname(bhw.lib_path())
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(H.ROOT, "include"),
                    os.path.join(H.ROOT, "examples", "coe_dump.cpp"), "-L", libdir, "-lbhw",
                    "-Wl,-rpath," + libdir, "-o", EXE], check=True)


def run(*args):
    return subprocess.run([EXE, *args], capture_output=True, text=True)


def test_cpp_mirror_compiles_links_and_validates():
    build()
    r = run("--check", "BH4TERM", "16", "17", "6")
    assert r.returncode == 0 and "65536 samples of 4 bytes, AA0=47022" in r.stdout
    r = run("--check", "BH4TERM", "16", "17", "6", "TAYLOR")          # 4-term has no TAYLOR
    assert r.returncode == 1 and "sin_type" in r.stderr
    r = run("--check", "BH7TERM", "20", "40", "10")
    assert r.returncode == 0 and "8 bytes" in r.stdout


@pytest.mark.gpu
def test_cpp_dump_equals_oracle():
    build()
    for args, d in ((("BH4TERM", "12", "17", "6"), bhw.variant_desc(6, 12, 17)),
                    (("BH3TERM", "14", "24", "3", "TAYLOR", "9"), bhw.variant_desc(3, 14, 24, sin_type=bhw.SIN_TAYLOR, lut_size=9)),
                    (("BH7TERM", "10", "40", "10"), bhw.variant_desc(10, 10, 40))):
        r = run(*args)
        assert r.returncode == 0, r.stderr
        got = np.array([int(x) for x in r.stdout.split()], dtype=np.int64)
        assert np.array_equal(got, H.orc_window(d))
