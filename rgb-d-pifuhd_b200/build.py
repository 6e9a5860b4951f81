"""Build libpifu_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import contextlib
import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpifu_b200.so")
SOURCES = ["api.cu", "gather.cu", "gemm_tc.cu", "chain_tc.cu", "runs.cu", "refine.cu", "meshclean.cu", "gemm_simt.cu", "pack.cu", "octree.cu", "mc.cu", "norm.cu", "obj.cu", "encoder_ops.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"] + [f for f in os.environ.get("PIFU_NVCC_FLAGS", "").split() if f]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode() + fh.read())
    with open(os.path.join(os.path.dirname(HERE), "include", "pifu_b200.h"), "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def up_to_date():
    """True when the library on disk was built from exactly the sources in the tree."""
    stamp = LIB + ".stamp"
    try:
        return os.path.exists(LIB) and open(stamp).read() == _digest()
    except OSError:
        return False


@contextlib.contextmanager
def _build_lock():
    """One builder at a time: the ranks of a torchrun launch may all find the library missing at once."""
    fd = os.open(LIB + ".lock", os.O_CREAT | os.O_RDWR, 0o644)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        fcntl.flock(fd, fcntl.LOCK_UN)
        os.close(fd)


def build(force=False, verbose=False):
    """Compile every .cu into objects (parallel) and link the shared library."""
    with _build_lock():
        return _build_locked(force, verbose)


def _build_locked(force, verbose):
    stamp = LIB + ".stamp"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB                                  # (also: another process built it while this one waited for the lock)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + FLAGS + ["-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs = []
    for src, obj, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out.strip():
            print(out)
        objs.append(obj)
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout.decode())
    os.replace(tmp, LIB)                            # atomic: a process that mapped the old file keeps it
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
