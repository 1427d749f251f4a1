"""Builds libtavk.so (the C-ABI kernel library, include/tavk.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the authoring container as well as on the B200 box; the built
``libtavk.so`` next to this file is what travels to the GPU box."""
import hashlib
import json
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


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())
            h.update(f.read())
    return h.hexdigest()


STAMP = os.path.join(OBJ, "build_info.json")
last_build = None     # {"mode": "compiled" | "reused (content hash verified)", "sources_compiled": [...], "sha256": ...}


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> build/*.o -> libtavk.so.  Staleness is decided by CONTENT: every object records the sha256 of
    its source + the shared headers + the compiler flags, the library records the digest of all of them, and a stored
    object / library is reused only when its recorded digest matches what is on disk now (mtimes say nothing after a
    checkout or an rsync).  TAVK_FORCE_BUILD=1 or force=True recompiles everything.  ``last_build`` / build/build_info.json
    say what happened, so a driver can tell an exercised build from a reused binary.  Returns the library path."""
    global last_build
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    force = force or os.environ.get("TAVK_FORCE_BUILD", "") == "1"
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "tavk.h")]
    flags = " ".join(NVCC_FLAGS)
    try:
        with open(STAMP) as f:
            stamp = json.load(f)
    except Exception:  # noqa: BLE001
        stamp = {}
    recorded = stamp.get("objects", {})
    jobs, digests, compiled = [], {}, []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        digests[s] = _digest([src] + headers, flags)
        if force or not os.path.exists(obj) or recorded.get(s) != digests[s]:
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
            compiled.append(s)

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
    lib_digest = hashlib.sha256("".join(digests[s] for s in SOURCES).encode()).hexdigest()
    relink = force or bool(jobs) or not os.path.exists(LIB) or stamp.get("library") != lib_digest
    if relink:
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
    last_build = {"mode": "compiled" if relink else "reused (content hash verified)", "sources_compiled": compiled,
                  "sha256": lib_digest, "nvcc_flags": flags}
    with open(STAMP, "w") as f:
        json.dump({"objects": digests, "library": lib_digest, "last": last_build}, f, indent=1)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
