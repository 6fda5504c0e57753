"""In-tree build of ``csrc/libslode_b200.so`` with nvcc for sm_100a (no JIT cache, no torch headers)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(CSRC, "libslode_b200.so")
SOURCES = ["slode_host.cu", "slode_mlp.cu", "slode_fixed_api.cu", "slode_cvs.cu", "slode_heads.cu", "slode_dopri5_adj.cu",
           "slode_mlp_25_5.cu", "slode_mlp_25_8.cu", "slode_mlp_16_4.cu", "slode_mlp_32_5.cu", "slode_mlp_64_5.cu",
           "slode_fixed_25_5.cu", "slode_fixed_25_8.cu", "slode_fixed_16_4.cu", "slode_fixed_32_5.cu",
           "slode_fixed_64_5.cu", "slode_fixed_128_5.cu", "slode_fixed_256_5.cu", "slode_fixed_512_5.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


# Code-generation guard (tests/test_gpu_codegen_guard.py): the translation unit with the largest register-resident
# sorting network -- (64,5): 64 keys, spills at 96 / 168 registers -- is ALSO built at ptxas -O1 and linked into a second
# library; on the GPU both builds must give the same trajectories and gradients.  (Round 1 hit a ptxas -O3
# miscompile in exactly such an instantiation; its kernels are gone, the guard stays.)
GUARD_SRC = "slode_fixed_64_5.cu"
GUARD_LIB = os.path.join(CSRC, "libslode_b200_guard_O1.so")


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isfile(cand) or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def source_hash():
    """31-bit digest of every source the library is built from (the .cu / .cuh / .h files and the public header).
    It is compiled into the library (``slode_query(SLODE_Q_SOURCE_HASH)``), so a prebuilt .so that travelled with the
    tree can be told apart from one built from the sources beside it."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    for f in files:
        h.update(f.encode())
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "slode_b200.h"), "rb").read())
    return int(h.hexdigest()[:8], 16) & 0x7FFFFFFF


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "slode_b200.h"))
    nvcc = _nvcc()
    digest = source_hash()
    stamp = os.path.join(OBJ, "source_hash.txt")
    old = open(stamp).read().strip() if os.path.isfile(stamp) else ""
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        # slode_mlp.cu carries the digest: it is rebuilt whenever any source changed
        extra = [f"-DSLODE_SOURCE_HASH={digest}"] if src == "slode_mlp.cu" else []
        if force or _stale(o, [s] + headers) or (extra and old != str(digest)):
            jobs.append([nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs)))) as ex:
        logs = list(ex.map(run, jobs))
    if verbose:
        print("\n".join(logs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    with open(stamp, "w") as f:
        f.write(str(digest))
    go = os.path.join(OBJ, GUARD_SRC.replace(".cu", "_O1.o"))
    gs = os.path.join(CSRC, GUARD_SRC)
    if force or _stale(go, [gs] + headers):
        run([nvcc] + NVCC_FLAGS + ["-Xptxas", "-O1", "-c", gs, "-o", go])
    if force or jobs or _stale(GUARD_LIB, objs + [go]):
        others = [o for o in objs if not o.endswith(GUARD_SRC.replace(".cu", ".o"))]
        run([nvcc, "-shared", "-o", GUARD_LIB] + others + [go, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
