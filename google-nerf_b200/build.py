"""In-tree build of libb2n.so with nvcc for sm_100a (no torch C++ ABI, no pybind: the boundary is the C ABI
in include/b2n.h).  The .so lands in google-nerf_b200/lib/ so that it travels with the source tree."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "lib")
# tuning experiments: B2N_VARIANT=name B2N_EXTRA_FLAGS="-DAGG_RES_DEF=264" python build.py  ->  lib/libb2n_name.so,
# loaded by pointing B2N_LIB at it (_lib.py); the product build has neither variable set
VARIANT = os.environ.get("B2N_VARIANT", "")
EXTRA = os.environ.get("B2N_EXTRA_FLAGS", "").split() if VARIANT else []
OBJ_DIR = os.path.join(OUT_DIR, "obj" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(OUT_DIR, "libb2n" + ("_" + VARIANT if VARIANT else "") + ".so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# bit-exact geometry: one IEEE op per source op, no fused multiply-add (DESIGN.md "Numerics")
PER_FILE = {"geometry.cu": ["-fmad=false"], "march.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    deps = [os.path.join(CSRC, src), os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "b2n.h")]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not _stale(obj, deps):
        return obj, ""
    cmd = ["nvcc", *ARCH, *COMMON, *PER_FILE.get(src, []), *EXTRA, "-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr


def build(verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    if verbose:
        for (_, log), s in zip(res, srcs):
            if log:
                print(f"==== {s}\n{log}")
    if _stale(LIB, objs):
        cmd = ["nvcc", *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
