"""Build the in-tree shared library piplib_b200/lib/libpiplib_dp.so for sm_100a.

nvcc cross-compiles without a GPU.  The library is named like the reference's int64 build
(libpiplib_dp.so, Makefile.am:24-25 of the reference) so that `-lpiplib_dp` keeps working.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpiplib_dp.so")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA, "bin", "nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fwrapv", "-Xptxas", "-v"]
SOURCES_CU = ["pip_kernels.cu", "pip_large.cu"]
SOURCES_CPP = ["pip_engine.cpp", "pip_host.cpp"]
SOURCES_CLI = ["pip_cli.cpp"]
HEADERS = ["pip_types.h", "pip_segments.h", "pip_convert.h", "simt.h", "pip_arith.h", "pip_decode.h", "pip_decode_warp.h", "pip_solver.h", "pip_warp_main.h", "pip_kernels.h",
           "pip_engine.h", "pip_large.h", os.path.join("..", "..", "include", "piplib_b200.h"),
           os.path.join("..", "..", "include", "piplib", "piplib.h")]


def _stale(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    for f in SOURCES_CU + SOURCES_CPP + SOURCES_CLI + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return os.path.getmtime(os.path.abspath(__file__)) > t


def build(force=False, verbose=False, profile=False, variant=None, defines=(), nvcc_extra=()):
    """profile=True builds libpiplib_dp_prof.so with per-phase clock64 accounting compiled in;
    variant="x" + defines builds an experimental libpiplib_dp_x.so (tuning only)."""
    global LIB
    lib = os.path.join(LIBDIR, "libpiplib_dp_prof.so") if profile else LIB
    if variant:
        lib = os.path.join(LIBDIR, "libpiplib_dp_%s.so" % variant)
    if not force and not _stale(lib):
        return lib
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    extra = (["-DPIP_PROFILE"] if profile else []) + ["-D" + d for d in defines]
    if variant:
        profile = variant          # distinct object / log names
    inc = ["-I", os.path.join(HERE, "..", "include"), "-I", CSRC]
    for f in SOURCES_CU:
        o = os.path.join(LIBDIR, f + ".o")
        o = os.path.join(LIBDIR, f + (".%s" % profile if profile else "") + ".o")
        cmd = [NVCC] + NVCC_FLAGS + list(nvcc_extra) + extra + inc + ["-c", os.path.join(CSRC, f), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on " + f)
        with open(os.path.join(LIBDIR, f + (".%s" % profile if profile else "") + ".ptxas.txt"), "w") as fh:
            fh.write(r.stdout + r.stderr)
        objs.append(o)
    for f in SOURCES_CPP:
        o = os.path.join(LIBDIR, f + (".%s" % profile if profile else "") + ".o")
        cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-fwrapv", "-Wall", "-pthread"] + extra + [
               "-I", os.path.join(CUDA, "include")] + inc + ["-c", os.path.join(CSRC, f), "-o", o]
        subprocess.check_call(cmd)
        objs.append(o)
    cmd = ["g++", "-shared", "-o", lib] + objs + ["-L", os.path.join(CUDA, "lib64"), "-lcudart",
                                                   "-pthread", "-Wl,-rpath," + os.path.join(CUDA, "lib64")]
    subprocess.check_call(cmd)
    for o in objs:
        os.remove(o)
    if not profile and not variant:
        build_cli()
    return lib


BINDIR = os.path.join(HERE, "bin")
CLI = os.path.join(BINDIR, "pip64")


def build_cli():
    """pip64: the reference's command-line tool (source/maind.c) over the batch library"""
    os.makedirs(BINDIR, exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-Wall", "-I", os.path.join(HERE, "..", "include"),
                           os.path.join(CSRC, "pip_cli.cpp"), "-o", CLI, "-L", LIBDIR, "-lpiplib_dp",
                           "-Wl,-rpath,$ORIGIN/../lib", "-Wl,-rpath," + os.path.join(CUDA, "lib64")])
    return CLI


if __name__ == "__main__":
    var, defs, nx = None, [], []
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            var = a.split("=", 1)[1]
        if a.startswith("--nvcc="):
            nx += a.split("=", 1)[1].split(",")
        if a.startswith("-D"):
            defs.append(a[2:])
    print(build(force="--force" in sys.argv, verbose=True, profile="--profile" in sys.argv, variant=var,
                defines=defs, nvcc_extra=nx))
