"""In-tree build of the C-ABI shared library ``libb200vsgg.so`` (nvcc, sm_100a only).

    python -m b200vsgg.build            # or: b200vsgg.build.build()

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.normpath(os.path.join(HERE, "..", "include"))
LIB_PATH = os.path.join(HERE, "libb200vsgg.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I", INCLUDE,
]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest(path):
    h = hashlib.sha1()
    for p in [path] + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu (incrementally, keyed by content hash) and link the shared library."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    objs, procs = [], []
    for src in _sources():
        name = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OBJ_DIR, name + ".o")
        stamp = obj + ".sha1"
        dig = _digest(src)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read().strip() == dig):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    relink = bool(procs) or not os.path.exists(LIB_PATH)
    for src, stamp, dig, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on %s" % src)
        if verbose:
            sys.stderr.write(out.decode())
        with open(stamp, "w") as f:
            f.write(dig)
    if relink:
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
