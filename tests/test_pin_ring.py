"""The pinned staging ring behind pb_h2d (csrc/pb_pin_ring.h): its bookkeeping on the CPU -- a place is never handed out while a
chunk that has not been retired overlaps it, whatever the sizes and however early a lap wraps (tools/pin_ring_test.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pin_ring_bookkeeping(tmp_path):
    exe = str(tmp_path / "pin_ring_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "pin_ring_test.cpp")])
    out = subprocess.check_output([exe, "60000"]).decode()
    assert out.strip().endswith("OK"), out
