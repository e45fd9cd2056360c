"""Compile libpicaso_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree.

    python -m picaso_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so lands in picaso_b200/_build/ (git-ignored,
but it travels to the GPU box with the gpurun snapshot).
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libpicaso_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--fmad=true",
              "-Xptxas", "-v"]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; libpicaso_b200.so cannot be built")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "picaso_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), lib=None):
    """defines/lib: build an A/B variant (extra -D flags) into another file name."""
    if lib is None and not force and not _stale():
        return LIB
    lib = lib or LIB
    tag = os.path.basename(lib)[:-3]
    os.makedirs(OUT, exist_ok=True)
    cc = nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT, tag + "." + src[:-3] + ".o")
        cmd = [cc, *ARCH, *NVCC_FLAGS, *["-D" + d for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OUT, tag + "." + src[:-3] + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [cc, *ARCH, "-shared", "-o", lib, *objs, "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
