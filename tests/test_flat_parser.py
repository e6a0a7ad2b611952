"""The flat slot parser (f3::parse_planes_flat, the CPU-checkable form of the warp-specialised decode kernel's parser)
must fill the same plane rows and stop at the same plane as parse_planes -- which is pinned to the oracle by
test_host_emul.py -- on random slots of every density / budget / kmin, for 32 and 64 planes, and on encoded fields."""
import os
import subprocess

from conftest import ROOT


def test_flat_parser_equals_nested_parser():
    exe = os.path.join(ROOT, "build", "flat_check")
    src = os.path.join(ROOT, "tools", "flat_check.cpp")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-DZFB_FLAT_STATS", "-o", exe, src], check=True)
    r = subprocess.run([exe, "60000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "60000 cases, 0 mismatches" in r.stdout
    # the SIMT model: in fixed-rate mode the lanes of a warp need nearly the same number of iterations
    for line in r.stdout.splitlines():
        if line.startswith("rate"):
            per_lane = float(line.split("iterations per lane")[1].split()[0])
            per_warp = float(line.split("per warp (max)")[1].split()[0])
            assert per_lane / per_warp > 0.85, line
            assert "mismatches 0" in line
