"""Builds libbitdelta_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache: the .so travels with the repo).

    python bitdelta_b200/build.py [--force] [-v] [--bringup]      (run as a script: importing the package requires the built library)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT = os.path.join(PKG, "libbitdelta_b200.so")
OUT_BRINGUP = os.path.join(PKG, "libbitdelta_b200_bringup.so")  # -DBD_BRINGUP: trace + A/B knobs, for tools/ only
SOURCES = ["bd_api.cu", "bd_codec.cu", "bd_simt.cu", "bd_umma.cu", "bd_tenant.cu", "bd_tp.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC or install the CUDA toolkit")


def _stale(out: str = OUT) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "bitdelta_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, bringup: bool = False, defines=(), out: str | None = None) -> str:
    """`defines` / `out`: compile-time A/B variants of the release library for tools/ab_variants.sh (never loaded by the package)."""
    out_path = out or (OUT_BRINGUP if bringup else OUT)
    if not force and not _stale(out_path):
        return out_path
    nvcc = _nvcc()
    objdir = os.path.join(PKG, "build", "variant_" + os.path.basename(out).replace(".so", "") if out else ("bringup" if bringup else "release"))
    extra = (["-DBD_BRINGUP"] if bringup else []) + [f"-D{d}" for d in defines]
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-Xptxas", "-v" if verbose else "-warn-spills", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose or "warning" in out.lower():
            print(out, file=sys.stderr)
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", *objs, "-o", out_path, "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return out_path


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    _out = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv or bool(_out), verbose="-v" in sys.argv, bringup="--bringup" in sys.argv, defines=_defs, out=_out))
