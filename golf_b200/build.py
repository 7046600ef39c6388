"""Build libgolf_b200.so (sm_100a only) in-tree with nvcc.

    python -m golf_b200.build [--force] [--verbose]

No torch headers are involved: the library is a plain C-ABI shared object
(include/golf_b200.h).  Objects are compiled in parallel (the per-tap-count
instantiations of the GOLF-ss kernels are separate translation units) and cached
by source hash under golf_b200/_lib/obj/.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
# A/B and instrumented builds (tools/): GOLF_B200_VARIANT=name GOLF_NVCC_DEFS="-DX ..." builds libgolf_b200_name.so beside
# the product library; select it at run time with GOLF_B200_SO=<path>
VARIANT = os.environ.get("GOLF_B200_VARIANT", "")
EXTRA_DEFS = os.environ.get("GOLF_NVCC_DEFS", "").split() if VARIANT else []
OBJDIR = os.path.join(LIBDIR, "obj" + ("_" + VARIANT if VARIANT else ""))
SO = os.path.join(LIBDIR, "libgolf_b200" + ("_" + VARIANT if VARIANT else "") + ".so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

SS_MPS = (4, 8, 12, 16, 20, 24, 32, 40)
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _units():
    units = []
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith(".cu"):
            continue
        if f == "lpc_ss_mp.cu":
            for mp in SS_MPS:
                units.append((f, f"lpc_ss_mp{mp}", [f"-DGOLF_MP={mp}"]))
        else:
            units.append((f, f[:-3], []))
    return units


def _digest(src, extra) -> str:
    """hash of this unit's source, every header, and the flags"""
    h = hashlib.sha256()
    files = [os.path.join(CSRC, src), os.path.join(INCLUDE, "golf_b200.h")]
    files += [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    for p in files:
        h.update(open(p, "rb").read())
    h.update(" ".join(ARCH + NVCC_FLAGS + EXTRA_DEFS + list(extra)).encode())
    return h.hexdigest()[:16]


def _compile(unit, verbose):
    src, name, defs = unit
    obj = os.path.join(OBJDIR, name + ".o")
    stamp = obj + ".stamp"
    dig = _digest(src, defs + [name])
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False, ""
    cmd = [nvcc()] + ARCH + NVCC_FLAGS + EXTRA_DEFS + defs + ["-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(dig)
    return obj, True, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    units = _units()
    objs, rebuilt = [], False
    with cf.ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        for obj, did, log in ex.map(lambda u: _compile(u, verbose), units):
            objs.append(obj)
            rebuilt |= did
            if verbose and log:
                print(log, file=sys.stderr)
    if rebuilt or not os.path.exists(SO):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", SO] + objs + ["-Xlinker", "--no-undefined", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return SO


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
