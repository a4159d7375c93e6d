"""CMakeLists.txt builds the same library as build.py (BASELINE north_star: "a cc/cmake build script"): configure + build in a
scratch directory, then check that the result is sm_100a only and exports every symbol include/nfx.h declares."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("ninja") is None, reason="cmake / ninja not installed")
def test_cmake_build_exports_the_c_abi(tmp_path):
    b = str(tmp_path / "b")
    r = subprocess.run(["cmake", "-S", ROOT, "-B", b, "-G", "Ninja"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run(["cmake", "--build", b, "--target", "nfx"], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    so = os.path.join(b, "libnfx.so")
    declared = set(re.findall(r"\b(nfx_[a-z0-9_]+)\s*\(", open(os.path.join(ROOT, "include", "nfx.h")).read()))
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (nfx_[a-z0-9_]+)", out))
    assert declared <= exported, sorted(declared - exported)
    elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"\.(sm_[0-9a-z]+)\.cubin", elf))
    assert archs == {"sm_100a"}, archs
    # the Python binding resolves every prototype against it
    env = dict(os.environ, NFX_LIB=so, PYTHONPATH=os.path.join(ROOT, "nuclei-feature-extraction_b200"))
    r = subprocess.run([sys.executable, "-c", "from nfx._lib import lib, SYMBOLS; lib(); print(len(SYMBOLS))"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and int(r.stdout) >= 56, r.stderr
