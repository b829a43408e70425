"""CPU: the product's host-side data movement, compiled on its own (no CUDA): widening of 4- and 3-byte indices into the
int64 table and the worker pool of whatsthepoint.jl_b200/csrc/host_pool.h."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_pool_and_widening(tmp_path):
    exe = tmp_path / "host_pool_check"
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-o", str(exe), os.path.join(ROOT, "tests", "native", "host_pool_check.cpp")],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr
