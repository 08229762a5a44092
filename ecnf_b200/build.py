"""In-tree build of libecnf_b200.so (sm_100a) with nvcc.  No JIT cache: the .so lives next to the package so
that it travels to the GPU box with the repo snapshot."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# the per-phase cycle-counter build (ECNF_TC_PROFILE=1, tools/tc_profile.py) lives beside the product library
_PROF = bool(os.environ.get("ECNF_TC_PROFILE"))
OBJ = PKG / ("build_prof" if _PROF else "build")
LIB = PKG / ("libecnf_b200_prof.so" if _PROF else "libecnf_b200.so")
INCLUDE = PKG.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]
if _PROF:
    NVCC_FLAGS.append("-DECNF_TC_PROFILE")


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libecnf_b200.so")
    return exe


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE.glob("*.h"))
    OBJ.mkdir(exist_ok=True)
    hdr_digest = _digest(headers)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in sources:
        obj = OBJ / (src.stem + ".o")
        stamp = OBJ / (src.stem + ".sha")
        dig = hashlib.sha256((hdr_digest + _digest([src])).encode()).hexdigest()
        objs.append(obj)
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
            continue
        jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(dig)
        return src.name, r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for name, err in ex.map(compile_one, jobs):
                if verbose:
                    print(f"[ecnf_b200.build] {name}\n{err}", file=sys.stderr)
    if jobs or not LIB.exists():
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
