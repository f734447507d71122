"""In-tree build of libmixlab_b200.so (the C-ABI shared library of include/mixlab_b200.h).

nvcc cross-compiles sm_100a without a GPU; the resulting .so sits next to this file so it travels
with the tree.  cudart is linked statically: the library loads on a box without a driver (every
compute entry point then reports MXL_ERR_NO_DEVICE) and has no runtime-library search problems.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(HERE, "libmixlab_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "mixlab_b200.h")

SOURCES = ["core.cu", "abi.cu", "modules.cu", "graph.cu", "audio_kernels.cu", "eq_three.cu", "eq_stream.cu",
           "envelope.cu", "video_kernels.cu", "comm.cu", "fused_voice.cu", "resample.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    # rustc never contracts a*b+c: separate operators stay two roundings; fma() is explicit
    "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]


def nvcc():
    path = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(path):
        raise RuntimeError("nvcc not found; libmixlab_b200.so cannot be built")
    return path


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _deps():
    deps = [HEADER, os.path.abspath(__file__)]
    for f in os.listdir(CSRC):
        deps.append(os.path.join(CSRC, f))
    return deps


def up_to_date():
    if not os.path.exists(LIB_PATH):
        return False
    t = os.path.getmtime(LIB_PATH)
    return all(os.path.getmtime(d) <= t for d in _deps())


def _compile(src):
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")] + [HEADER, os.path.abspath(__file__)]
    srcp = os.path.join(CSRC, src)
    if os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in headers + [srcp]):
        return obj, ""
    cmd = [nvcc()] + NVCC_FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stdout))
    return obj, r.stdout


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libmixlab_b200.so.  Returns the path."""
    if not force and up_to_date():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            if f.endswith(".o"):
                os.remove(os.path.join(OBJ_DIR, f))
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    log = "".join(out for _, out in results)
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "a") as f:
        f.write(log)
    if verbose:
        sys.stderr.write(log)
    cmd = [nvcc(), "-shared", "-o", LIB_PATH] + [o for o, _ in results] + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                          "-cudart", "static", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
