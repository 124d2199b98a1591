"""In-tree build of librpool_b200.so with nvcc for sm_100a (no JIT cache: the
built file sits next to this package so that it travels with the tree)."""
import hashlib
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "librpool_b200.so")
SOURCES = ["rpool_api.cu"]
HEADERS = ["rpool_device.cuh", "rpool_kernels.cuh", 
           os.path.join("..", "..", "include", "rpool_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build librpool_b200.so")


def sources_present():
    return all(os.path.exists(os.path.join(CSRC, f)) for f in SOURCES + HEADERS)


def source_hash():
    """Build id: sha1 over the sources and headers the library is compiled from."""
    h = hashlib.sha1()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


ID_PATH = LIB_PATH + ".id"


def is_stale():
    """True when the library is missing or was built from other sources (the build id of
    the last build is kept beside it: file times do not survive a copy of the tree)."""
    if not os.path.exists(LIB_PATH):
        return True
    if not sources_present():
        return False
    try:
        with open(ID_PATH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile csrc/*.cu -> librpool_b200.so.  Returns the library path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ['-DRPOOL_BUILD_ID="%s"' % source_hash()]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stdout))
    if verbose:
        print(proc.stdout)
    with open(ID_PATH, "w") as f:
        f.write(source_hash() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
