"""The C++ host side of the drop-in boundary, on the GPU: the header-level classes (dna_seq / seq_accessor /
seq_aligner<> / hash_table) and the locator driver with the reference's CLI contract.

build/host/ref_dna_test and build/host/ref_aligner_test are the REFERENCE'S OWN test/*.cpp, compiled unmodified
against pacbioassembly_b200/host/include (built in the container where /root/reference exists; the binaries travel)."""
import os
import subprocess

import numpy as np
import pytest

import workload

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "build", "host")


def run(exe, cwd=None, stdin=None, args=()):
    path = os.path.join(HOST, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built")
    return subprocess.run([path, *args], cwd=cwd, input=stdin, capture_output=True, timeout=600)


def test_host_selftest():
    r = run("host_selftest")
    out = r.stdout.decode()
    assert r.returncode == 0, out + r.stderr.decode()
    assert "4 tests, 0 failures" in out


def test_reference_dna_test_unmodified():
    r = run("ref_dna_test")
    out = r.stdout.decode()
    assert r.returncode == 0, out + r.stderr.decode()
    assert "3 tests, 0 failures" in out  # dna_seq.binary, seq_accessor.forward, seq_accessor.backward


def test_reference_aligner_test_unmodified(tmp_path, golden):
    # aligner_test.cpp:101 opens "test/real_align.txt" relative to the cwd: regenerate it from the golden vectors
    pairs = {}
    for x in golden["real_align"]:
        if x["order"] == "seg,ref" and x["a_fwd"]:
            pairs[x["pair"]] = (x["b"], x["a"])  # file order: reference line, then segment line
    (tmp_path / "test").mkdir()
    with open(tmp_path / "test" / "real_align.txt", "w") as f:
        for k in sorted(pairs):
            f.write(pairs[k][0] + "\n" + pairs[k][1] + "\n")
    r = run("ref_aligner_test", cwd=str(tmp_path))
    out = r.stdout.decode()
    assert r.returncode == 0, out + r.stderr.decode()
    assert "5 tests, 0 failures" in out  # forward, backward, overlay, remove, sample


def test_locator_cli_matches_oracle(tmp_path, oracle):
    """locator contig_file pattern [R] < reads  ->  TSV identical to the reference driver's (locator.cpp:84-86)"""
    ref = workload.reference(61, 150_000)
    lens = workload.read_lengths(62, 80, mean=1500.0, sigma_log=0.5, lo=300, hi=4000)
    txt, offs, lens, _ = workload.reads(63, ref, lens, 0.02, 0.01, 0.01)
    contig = tmp_path / "contig.txt"
    contig.write_bytes(ref.tobytes() + b"\n")
    reads_in = b"\n".join(txt[offs[k]: offs[k] + lens[k]].tobytes() for k in range(len(lens))) + b"\n"
    pattern = "111**111*11*1111"
    mask = oracle.parse_pattern(pattern.encode())
    for R in (None, 0.3):
        r = run("locator", stdin=reads_in, args=[str(contig), pattern] + ([str(R)] if R else []))
        assert r.returncode == 0, r.stderr.decode()
        ix = oracle.index_build(ref, mask, 0)
        want = oracle.locate(ix, ref, txt, offs, lens, mask, R=R or 0.15, nthreads=4)
        oracle.index_free(ix)
        lines = ["%d\t%d\t%d\t%d\t%d" % (w["nseq"], w["pos"], w["cost"], w["seg_len"], w["diag_cost"]) for w in want if w["found"]]
        assert r.stdout.decode().splitlines() == lines
        assert len(lines) > 30
        assert b"totally %d sequences processed" % len(want) in r.stderr
