"""Builds nesvor_b200/csrc/libnesvor_b200.so with plain nvcc for sm_100a (no torch headers: seconds per file).

    python -m nesvor_b200.csrc.build [--force] [--verbose]

The shared object is built IN-TREE so that it travels to the GPU box with the repository snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libnesvor_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NSV_NVCC", "/usr/local/cuda/bin/nvcc")

COMMON_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
# per-source extra flags
SOURCES = {
    "common.cu": [],
    "transform_convert.cu": ["-fmad=false"],  # literal C arithmetic of the reference converters
    "slice_acq.cu": ["-fmad=false"],  # bit-exact flavour: gather passes reproduce the CPU oracle bit for bit
    "slice_acq_fast.cu": [],  # product flavour (FMA) + the C ABI of the family
    "hashgrid.cu": [],
    "mlp.cu": [],
    "inr_fused.cu": [],
    "inr_fused_tc.cu": [],
    "inr_fused_ws.cu": [],
    "inr_bias.cu": [],
    "adamw.cu": [],
    "umma_selftest.cu": [],
}


def _deps():
    return [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "..", "include", "nesvor_b200.h")
    ]


def _digest(path, flags):
    h = hashlib.sha1()
    for p in [path] + sorted(_deps()):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _compile(src, flags, force, verbose):
    path = os.path.join(HERE, src)
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    stamp = obj + ".sha1"
    dig = _digest(path, flags)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    cmd = [NVCC] + COMMON_FLAGS + flags + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, True, r.stderr if verbose else ""


def build(force=False, verbose=False):
    """Idempotent and safe to call from several processes at once (one rank per GPU): an exclusive file lock serialises
    the builders, objects are compiled only when their sources changed, and the library is replaced atomically."""
    import fcntl

    os.makedirs(OBJ_DIR, exist_ok=True)
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    srcs = {s: fl for s, fl in SOURCES.items() if os.path.exists(os.path.join(HERE, s))}
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda kv: _compile(kv[0], kv[1], force, verbose), srcs.items()))
    objs = [r[0] for r in results]
    for r in results:
        if r[2]:
            print(r[2])
    if any(r[1] for r in results) or not os.path.exists(LIB):
        tmp = f"{LIB}.tmp.{os.getpid()}"
        cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
