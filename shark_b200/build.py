"""Builds libshark_b200.so (CUDA kernels + C ABI) and the shark-b200 CLI for sm_100a, in-tree.

nvcc cross-compiles without a GPU.  The .so stays next to this file so that it travels to the GPU
box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libshark_b200.so")
CLI = os.path.join(HERE, "shark-b200")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["shk_capi.cu", "shk_index.cu", "shk_reads.cu", "shk_bulk.cu"]
CPP_SOURCES = ["shk_hostpack.cpp"]  # host code of the library (g++): read packing for the H2D link
HOST_SOURCES = ["host/shark_main.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    out = [os.path.join(HERE, "..", "include", "shark_b200.h"), os.path.abspath(__file__)]
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".cpp", ".hpp"))]
    return out


def _host_objects():
    objs = []
    for src in CPP_SOURCES:
        obj = os.path.join(CSRC, src.replace(".cpp", ".o"))
        subprocess.check_call(["g++", "-O3", "-std=c++17", "-Wall", "-fPIC", "-pthread", "-c", os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    return objs


def build_variant(name, flags):
    """Tuning only: libshark_b200_<name>.so built with extra nvcc flags (picked up via SHK_LIB)."""
    out = os.path.join(HERE, "libshark_b200_%s.so" % name)
    objs = []
    for src in CU_SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".%s.o" % name))
        subprocess.check_call([NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", *flags, "-c",
                               os.path.join(CSRC, src), "-o", obj])
        objs.append(obj)
    objs += _host_objects()
    subprocess.check_call([NVCC, *ARCH, "-shared", "-o", out, *objs, "-lcudart", "-lpthread"])
    return out


def build(force=False, verbose=False):
    deps = _deps()
    if force or _newer(LIB, deps):
        objs = []
        for src in CU_SOURCES:
            obj = os.path.join(CSRC, src.replace(".cu", ".o"))
            cmd = [NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3",
                   *os.environ.get("SHK_NVCC_FLAGS", "").split(), "-c", os.path.join(CSRC, src), "-o", obj]
            subprocess.check_call(cmd)
            objs.append(obj)
        objs += _host_objects()
        subprocess.check_call([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart", "-lpthread"])
    host = [os.path.join(CSRC, s) for s in HOST_SOURCES]
    if all(os.path.exists(h) for h in host) and (force or _newer(CLI, deps + [LIB])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", *host, "-o", CLI, "-L", HERE,
                               "-lshark_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link,/usr/local/cuda/lib64", "-lz"])
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:  # build.py --variant NAME -DFOO=1 ...
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
        sys.exit(0)
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
