"""In-tree build of libscvae_b200.so (nvcc, sm_100a only).

The shared library is the drop-in boundary (``include/scvae_b200.h``); it is built next to
this file so that it travels with the source tree and is visible as a loaded in-tree ``.so``.
"""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB_PATH = os.path.join(HERE, "libscvae_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-cudart", "static",
]


FILE_FLAGS = {}          # per-file extra nvcc flags (none at present)


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


HASH_PATH = LIB_PATH + ".srchash"


def source_hash():
    """Content hash of every source the library is built from (mtimes do not survive the
    snapshot copy to the GPU box, contents do)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    deps.append(os.path.join(INCLUDE, "scvae_b200.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(sorted(FILE_FLAGS.items())).encode())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != source_hash()


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return nvcc if os.path.exists(nvcc) else None


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into ``libscvae_b200.so``; returns its path."""
    if not force and not _stale():
        return LIB_PATH
    # one builder at a time: under torchrun every rank may find the library stale at once
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():      # another process built it while we waited
                return LIB_PATH
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libscvae_b200.so")
    objdir = os.path.join(HERE, "build", "obj-{}".format(os.getpid()))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + FILE_FLAGS.get(os.path.basename(src), []) + (
            ["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                            text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for {}:\n{}".format(src, out))
        if verbose and out:
            print(out)
    tmp = "{}.{}.tmp".format(LIB_PATH, os.getpid())
    cmd = [nvcc, "-shared", "-cudart", "static", "-o", tmp] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as fh:
        fh.write(source_hash())
    shutil.rmtree(objdir, ignore_errors=True)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
