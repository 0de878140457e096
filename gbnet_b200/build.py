"""In-tree build of the CUDA library (libgbnet_b200.so) for sm_100a.

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  -fmad=false: the reference's x86-64 build contains no FMA contraction
(SURVEY.md §7), so products and sums must round separately here as well.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# GBN_LIB_VARIANT builds / loads a second library next to the shipped one: "chk" compiles bounds checks
# into the fast kernels, "fma" allows FMA contraction there, "s3"/"s4" set the S kernel's CTAs per SM
# (experiment knobs: the shipped default keeps -fmad=false everywhere)
VARIANT = os.environ.get("GBN_LIB_VARIANT", "")
LIB = os.path.join(HERE, "libgbnet_b200%s.so" % ("_" + VARIANT if VARIANT else ""))
SOURCES = ["capi.cu", "eval_kernels.cu", "sweep_strict.cu", "sweep_fast.cu", "sweep_small.cu", "sim_kernels.cu", "moments.cu", "fisher.cu"]
HEADERS = ["device_math.cuh", "device_model.cuh", "kernels.cuh", "topology.hpp", "fast_common.cuh",
           os.path.join("..", "..", "include", "gbnet_b200.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr"]


class CompileError(RuntimeError):
    """nvcc ran and rejected the sources (as opposed to: nvcc is not installed here)."""


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a into gbnet_b200/libgbnet_b200.so; returns its path."""
    if not force and not _stale():
        return LIB
    objdir = os.path.join(HERE, "build" + ("_" + VARIANT if VARIANT else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        flags = list(NVCC_FLAGS)
        if VARIANT == "fma" and src == "sweep_fast.cu":
            flags[flags.index("-fmad=false")] = "-fmad=true"
        if VARIANT == "nt512" and src == "sweep_fast.cu":
            flags += ["-DGBN_S_NT=512", "-DGBN_S_MINB=2"]
        if VARIANT == "nt128" and src == "sweep_fast.cu":
            flags += ["-DGBN_S_NT=128", "-DGBN_S_MINB=8"]
        xv = re.fullmatch(r"xq(\d+)u(\d+)(?:n(\d+))?", VARIANT)
        if xv and src == "sweep_fast.cu":
            flags += ["-DGBN_XQ=" + xv.group(1), "-DGBN_XU=" + xv.group(2)]
            if xv.group(3):
                flags.append("-DGBN_X_NT=" + xv.group(3))
        if VARIANT.startswith("xexp") and src == "sweep_fast.cu":
            flags.append("-DGBN_XEXP=" + VARIANT[4:])
        if VARIANT.startswith("D") and src == "sweep_fast.cu":      # D<MACRO>_<value>[+<MACRO>_<value>...]: macros of sweep_fast.cu (experiments)
            for part in VARIANT[1:].split("+"):
                dv = re.fullmatch(r"(\w+?)_(\d+)", part)
                if dv:
                    flags.append("-D%s=%s" % dv.groups())
        if VARIANT == "chk":
            flags.append("-DGBN_BOUNDS")
        if VARIANT.startswith("s") and VARIANT[1:].isdigit() and src == "sweep_fast.cu":
            flags.append("-DGBN_S_MINB=" + VARIANT[1:])
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc {src} failed ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
    if failed:
        raise CompileError("nvcc failed")
    cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))


# ---------------------------------------------------------------------------------- Cython binding
PYX = os.path.join(HERE, "ModelORNOR.pyx")


def cython_ext_path() -> str:
    import sysconfig
    return os.path.join(HERE, "ModelORNOR" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_cython(force: bool = False) -> str:
    """gbnet_b200/ModelORNOR.pyx -> in-tree extension module linked against libgbnet_b200.so
    (the reference compiles its .pyx together with libgbnet's sources, setup.py:19-23)."""
    import sysconfig
    out = cython_ext_path()
    srcs = [PYX, os.path.join(HERE, "ModelORNOR.pxd"), os.path.join(HERE, "..", "include", "gbnet_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(f) <= os.path.getmtime(out) for f in srcs):
        return out
    import numpy
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    cpp = os.path.join(objdir, "ModelORNOR.cpp")
    subprocess.run([sys.executable, "-m", "cython", "--cplus", "-3", "-o", cpp, PYX], check=True)
    inc = [sysconfig.get_paths()["include"], numpy.get_include(), os.path.join(HERE, "..", "include")]
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w"] + ["-I" + i for i in inc] + \
          [cpp, "-o", out, "-L" + HERE, "-lgbnet_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return out
