"""Build libnfx.so (sm_100a only) in-tree with nvcc. Used by __graft_entry__.build() and the tests."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libnfx.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["api.cu", "geom.cu", "color.cu", "glcm.cu", "texture2.cu", "staged.cu", "csv.cu", "slide_decode.cu", "ext.cu", "f32batch.cu", "schema.cpp", "geojson.cpp", "tiff.cpp", "jpeg_exact.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr"]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    if not os.path.exists(os.path.join(HERE, "nfx-cli")):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))] + [
        os.path.join(HERE, "..", "include", "nfx.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    os.makedirs(OBJ, exist_ok=True)

    def cc(src):
        obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(cc, SOURCES))
    objs = []
    for src, obj, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(f"--- {src}\n{r.stdout}{r.stderr}\n")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lnvjpeg"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    build_cli()
    return OUT


CLI = os.path.join(HERE, "nfx-cli")


def build_cli() -> str:
    """The C++ host driver (host/*.cpp) linked against libnfx.so."""
    srcs = [os.path.join(HERE, "host", f) for f in ("nfx_host.cpp", "cli.cpp")]
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(HERE, "..", "include"), *srcs, "-o", CLI,
           "-L", HERE, "-lnfx", "-lz", "-lpthread", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("building nfx-cli failed")
    return CLI


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
