"""The CPU model of the K3 kernel's arithmetic (tools/bitpar_model.c) against the oracle, cell for cell."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bitpar_model_matches_oracle(tmp_path):
    exe = str(tmp_path / "bitpar_model")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "bitpar_model.c"),
                           os.path.join(ROOT, "oracle", "pb_oracle.c"), "-lpthread"])
    out = subprocess.check_output([exe, "4000", "400"]).decode()
    assert out.startswith("OK 4000 cases"), out
    out = subprocess.check_output([exe, "60", "4000"]).decode()  # multi-word lanes (band > 1024 bits)
    assert out.startswith("OK 60 cases"), out


def test_narrow_strip_model_matches_oracle(tmp_path):
    """tools/narrow_model.c: the strip + block-stationary frame of K3's first pass, word by word as the warp runs it.  Every
    result the model CERTIFIES must equal the oracle's (return value, fail row, goal cell, cost, diagonal cost, transcript),
    and a certified path must stay inside the offsets whose parents the kernel stores."""
    exe = str(tmp_path / "narrow_model")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "narrow_model.c"),
                           os.path.join(ROOT, "oracle", "pb_oracle.c"), "-lpthread"])
    for args in (["2500", "1500", "0.75"], ["1200", "1200", "0.5"], ["1200", "1500", "1.0", "2"], ["150", "6000", "0.75"]):
        out = subprocess.check_output([exe] + args).decode()
        assert out.startswith("OK"), out
        aligned = int(out.split(" cases, ")[1].split(" aligned")[0])
        assert aligned > 40, out
