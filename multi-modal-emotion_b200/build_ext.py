"""Builds libtavk.so (the C-ABI kernel library, include/tavk.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the authoring container as well as on the B200 box; the built
``libtavk.so`` next to this file is what travels to the GPU box."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libtavk.so")
OBJ = os.path.join(HERE, "build")

SOURCES = ["api.cu", "gemm_tcgen05.cu", "attention.cu", "attention_tc.cu", "layernorm.cu", "pointwise.cu", "loss_optim.cu",
           "conv_frontend.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
    "-cudart", "shared", "--expt-relaxed-constexpr",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> build/*.o -> libtavk.so; no-op when up to date.  Returns the library path."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "tavk.h")]
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    print(out, file=sys.stderr)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        rpaths = ["/usr/local/cuda/lib64"]
        try:
            import nvidia.cuda_runtime  # torch's bundled runtime, preferred at load time

            rpaths.insert(0, os.path.join(list(nvidia.cuda_runtime.__path__)[0], "lib"))
        except Exception:
            pass
        link = [nvcc, "-shared", "-cudart", "shared", "-o", LIB] + objs
        for rp in rpaths:
            link += ["-Xlinker", "-rpath", "-Xlinker", rp]
        run(link)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
