"""In-tree build of libgb25cuda.so (nvcc, sm_100a).

``python -m gb25_b200.build`` or ``__graft_entry__.build()``.  The shared objects stay in-tree
(git-ignored) so that they travel to the GPU box with the gpurun snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgb25cuda.so")
CU_SOURCES = ["gb25_api.cu", "gb25_kernels.cu", "gb25_tend_v2.cu", "gb25_tend_tma.cu", "gb25_exchange.cu", "gb25_baro.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgb25cuda cannot be built (there is no CPU fallback)")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_variant(suffix, defines, verbose=False):
    """Experiment build: libgb25cuda<suffix>.so compiled with extra -D flags, objects in a private directory."""
    nvcc = _nvcc()
    odir = os.path.join(CSRC, "build", suffix.strip("_") or "default")
    os.makedirs(odir, exist_ok=True)
    lib = os.path.join(CSRC, f"libgb25cuda{suffix}.so")
    objs, procs = [], []
    for s in CU_SOURCES:
        o = os.path.join(odir, s[:-3] + ".o")
        objs.append(o)
        cmd = [nvcc, *[f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")], *defines, "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return lib


F64_SOURCES = ["gb25_api.cu", "gb25_kernels.cu", "gb25_exchange.cu", "gb25_f64_stubs.cu"]
LIB_F64 = os.path.join(CSRC, "libgb25cuda_f64.so")


def build_f64(force=False, verbose=False):
    """libgb25cuda_f64.so: the same sources with -DGB25_F64 (scalar type double), operator-per-kernel generation only."""
    srcs = [os.path.join(CSRC, s) for s in F64_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "gb25cuda.h"))
    if not force and not _newer(LIB_F64, deps):
        return LIB_F64
    nvcc = _nvcc()
    odir = os.path.join(CSRC, "build", "f64")
    os.makedirs(odir, exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(odir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        cmd = [nvcc, *[f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")], "-DGB25_F64", "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-o", LIB_F64, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB_F64


def build_cuda(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "gb25cuda.h"))
    if not force and not _newer(LIB, deps):
        return LIB
    nvcc = _nvcc()
    objs, procs = [], []
    for s in srcs:
        o = s[:-3] + ".o"
        objs.append(o)
        cmd = [nvcc, *[f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")],
               *os.environ.get("GB25_NVCC_DEFINES", "").split(), "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    if "--suffix" in sys.argv:      # python -m gb25_b200.build --suffix _x -DFOO=1 -DBAR=2
        q = sys.argv.index("--suffix")
        print(build_variant(sys.argv[q + 1], [a for a in sys.argv[q + 2:] if a.startswith("-D")], verbose="-v" in sys.argv))
    else:
        print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
        print(build_f64(force="--force" in sys.argv, verbose="-v" in sys.argv))
