"""In-tree build of libdpilqr_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m dpilqr_b200.build [--force] [--verbose]

The shared library lands in dpilqr_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdpilqr_b200.so")
# (source, object stem, extra flags): the rollout kernel is compiled once per model / size class (rollout_inst.cu)
ROLLOUT_CLASSES = [0, 1, 2, 3, 4, 5, 6, 7, 8, 100, 101]
UNITS = [(src, src[:-3], []) for src in ["solver.cu", "forward.cu", "linquad.cu", "backward.cu", "backward_small.cu", "backward_warp.cu", "graph.cu", "scenario.cu", "dynamics_api.cu", "cost_api.cu"]]
UNITS += [("backward_small.cu", f"backward_small_{sc[0]}_{sc[1]}", [f"-DDPILQR_SMALL_S={sc[0]}", f"-DDPILQR_SMALL_C={sc[1]}"])
          for sc in [(12, 4), (6, 3), (4, 2), (3, 2), (5, 2)]]
UNITS += [("rollout_inst.cu", f"rollout_c{k}", [f"-DDPILQR_ROLLOUT_CLASS={k}"]) for k in ROLLOUT_CLASSES]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = ["common.cuh", "kernels.cuh", "models.cuh", "cost.cuh", "lu.cuh", "rollout.cuh", os.path.join("..", "..", "include", "dpilqr_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++20", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(name, extra_flags):
    """Experimental build with extra nvcc flags into lib/variants/<name>.so (accuracy / timing experiments)."""
    outdir = os.path.join(LIBDIR, "variants")
    objdir = os.path.join(outdir, "obj_" + name)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def one(unit):
        src, stem, flags = unit
        obj = os.path.join(objdir, stem + ".o")
        res = subprocess.run([nvcc, *NVCC_FLAGS, *flags, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(res.stdout + res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
        objs = list(pool.map(one, UNITS))
    out = os.path.join(outdir, name + ".so")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs, "-lcudart"])
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(unit):
        src, stem, flags = unit
        obj = os.path.join(objdir, stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(os.path.join(objdir, stem + ".log"), "w") as fh:
            fh.write(log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src} {flags}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, UNITS))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
