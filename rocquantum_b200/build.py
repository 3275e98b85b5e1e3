"""rocquantum_b200/build.py -- compile the engine in-tree for sm_100a.

Two shared libraries, like the reference's compile-time precision switch (hipStateVec.h:6-15):
    rocquantum_b200/lib/libhipStateVec.so       complex64
    rocquantum_b200/lib/libhipStateVec_f64.so   complex128 (-DROCQ_PRECISION_DOUBLE)
nvcc cross-compiles without a GPU.  The built .so files are git-ignored but travel with gpurun.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("ROCQ_LIB_DIR", os.path.join(HERE, "lib"))          # tuning variants build elsewhere
OBJ = os.environ.get("ROCQ_OBJ_DIR", os.path.join(HERE, "_obj"))
EXTRA = os.environ.get("ROCQ_EXTRA_DEFS", "").split()
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-ccbin", HOSTCXX, "-Xcompiler", "-fPIC,-fvisibility=default,-Wall,-Wno-unused-function",
          "-I", os.path.join(HERE, "..", "include")]
SOURCES = ["tile_sweep.cu", "tile_sweep_small.cu", "tile_sweep_small_swz.cu", "tile_sweep_large.cu", "tile_sweep_large_swz.cu",
           "block_sweep.cu", "sv_kernels.cu", "engine.cu", "dist.cu"]
HEADERS = ["sv_internal.h", "host_ops.h", "gate_convert.h", "dist.h", "dist_plan.h", "engine.h", "tile_sweep.cuh", os.path.join("..", "..", "include", "hipStateVec.h")]
VARIANTS = {"libhipStateVec.so": [], "libhipStateVec_f64.so": ["-DROCQ_PRECISION_DOUBLE"]}


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")


def build(force: bool = False, verbose: bool = False) -> list[str]:
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs, libs = [], []
    for libname, defs in VARIANTS.items():
        tag = "f64" if defs else "f32"
        objs = []
        for src in SOURCES:
            s = os.path.join(CSRC, src)
            o = os.path.join(OBJ, f"{os.path.splitext(src)[0]}.{tag}.o")
            objs.append(o)
            if force or _newer(o, [s] + hdrs):
                jobs.append([NVCC, *ARCH, *COMMON, *defs, *EXTRA, "-c", s, "-o", o])
        libs.append((os.path.join(LIB, libname), objs))
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, jobs))
    out = []
    for path, objs in libs:
        if force or _newer(path, objs):
            _run([NVCC, *ARCH, "-shared", "-ccbin", HOSTCXX, "-Xlinker", "-Bsymbolic", "-o", path, *objs, "-lcudart", "-ldl"])
        out.append(path)
        if verbose:
            print("built", path)
    return out


def build_bindings(force: bool = False, verbose: bool = False) -> list[str]:
    """The reference's two pybind11 modules + C++ facades, compiled with g++ against the engine libraries:
         rocquantum_b200/lib/rocquantum_bind*.so      (QuantumSimulator/QSim, MLIRCompiler)   -> libhipStateVec_f64.so
         rocquantum_b200/lib/_rocq_hip_backend*.so    (rocsv* free functions, GateFusion)     -> libhipStateVec.so"""
    import sysconfig
    import pybind11
    build(force=force)
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    inc = ["-I", os.path.join(HERE, "..", "include"), "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           "-I", "/usr/local/cuda/include"]
    fac = os.path.join(CSRC, "facade")
    common = [HOSTCXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wno-unused-function", *inc]
    rpath = ["-L", LIB, "-Wl,-rpath,$ORIGIN", "-L", "/usr/local/cuda/lib64", "-lcudart"]
    mods = [
        ("rocquantum_bind" + ext, ["bind_rocquantum.cpp", "QuantumSimulator.cpp", "HipStateVecBackend.cpp"], ["-DROCQ_PRECISION_DOUBLE"], "-l:libhipStateVec_f64.so"),
        ("_rocq_hip_backend" + ext, ["bind_rocq_hip_backend.cpp", "GateFusion.cpp"], [], "-l:libhipStateVec.so"),
    ]
    out = []
    hdrs = [os.path.join(HERE, "..", "include", h) for h in ("hipStateVec.h", "rocquantum/QuantumSimulator.h", "rocquantum/GateFusion.h",
                                                            "rocqCompiler/QuantumBackend.h", "rocqCompiler/HipStateVecBackend.h")]
    for name, srcs, defs, lib in mods:
        target = os.path.join(LIB, name)
        deps = [os.path.join(fac, s) for s in srcs] + hdrs + [os.path.abspath(__file__)]
        if force or _newer(target, deps):
            _run([*common, *defs, *[os.path.join(fac, s) for s in srcs], "-o", target, *rpath, lib])
        out.append(target)
        if verbose:
            print("built", target)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    build_bindings(force="--force" in sys.argv, verbose=True)
