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
