"""The C ABI from plain C (gcc), i.e. exactly what the reference's Rust `extern "C"` block would
bind: links libnfx.so, checks schema/key/partition, and that a missing GPU is a status code."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "nuclei-feature-extraction_b200")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-o", exe, "-L", PKG, "-lnfx", "-lm", f"-Wl,-rpath,{PKG}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_consumer_links_and_runs_without_gpu(libnfx, tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok" in r.stdout


@pytest.mark.gpu
def test_c_consumer_extracts_on_gpu(libnfx, tmp_path):
    r = subprocess.run([_build(tmp_path), "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke ok (GPU)" in r.stdout
