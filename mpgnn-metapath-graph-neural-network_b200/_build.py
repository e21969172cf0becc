"""In-tree build of libmpgnn_b200.so (sm_100a only) with plain nvcc.

The library has no torch dependency: it is the C ABI of include/mpgnn_b200.h.  The built
.so stays next to this file (git-ignored, but it travels to the GPU box with the tree).
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmpgnn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]
# profiling builds only (e.g. MPGNN_NVCC_EXTRA=-DMPGNN_TC_EXPERIMENT, see scripts/exp_variants.sh)
NVCC_FLAGS += os.environ.get("MPGNN_NVCC_EXTRA", "").split()


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    stamp = os.path.join(BUILD, "stamp")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _digest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link libmpgnn_b200.so.  Returns its path."""
    if not force and is_current():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    srcs = _sources()

    def compile_one(src):
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            print(log)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(os.path.join(BUILD, "stamp"), "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
