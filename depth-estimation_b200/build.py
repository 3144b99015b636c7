"""Builds libdepthmatch.so (hand-written sm_100a CUDA, C ABI in include/depthmatch.h).

    python depth-estimation_b200/build.py [--force] [--verbose]

The library is built IN-TREE (depth-estimation_b200/csrc/libdepthmatch.so) so that it
travels with the repository snapshot to the GPU box.  nvcc cross-compiles for
sm_100a without a GPU.  cudart is linked statically and the driver API is reached
through cudaGetDriverEntryPoint, so the .so loads on a machine without a driver.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libdepthmatch.so")
SOURCES = ["dm_context.cu", "match_fused.cu", "match_generic.cu", "extract.cu", "multiscale.cu",
           "radial.cu", "postprocess.cu", "filter.cu", "filter_tc.cu"]
HEADERS = ["dm_common.cuh", "match_kernels.cuh", "match_sweep2.cuh", "match_volume_px.cuh", "filter_tc.cuh", os.path.join("..", "..", "include", "depthmatch.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v",
         "--expt-relaxed-constexpr", "-DDM_BUILDING"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(CSRC, src.replace(".cu", ".o"))
    log = obj + ".ptxas.log"   # per translation unit, so an incremental build keeps the others
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    if not _stale(obj, deps) and os.path.exists(log):
        return obj, open(log).read()
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, p.stdout))
    with open(log, "w") as f:
        f.write("==== %s\n%s" % (src, p.stdout))
    return obj, "==== %s\n%s" % (src, p.stdout)


def build(force=False, verbose=False):
    if force:
        for s in SOURCES:
            o = os.path.join(CSRC, s.replace(".cu", ".o"))
            if os.path.exists(o):
                os.remove(o)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [r[0] for r in results]
    log = "".join(r[1] for r in results)
    if log:
        with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
            f.write(log)
        if verbose:
            print(log)
    if _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC"]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout)
    return OUT


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
