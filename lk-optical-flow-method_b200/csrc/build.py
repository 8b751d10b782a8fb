"""Build liblkflow_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the library is
a plain C-ABI shared object, loaded with ctypes).  Usage: python build.py [--force] [--verbose]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "liblkflow_b200.so")
SOURCES = ["abi.cu", "stencil.cu", "lk.cu", "lk_dense.cu", "gftt.cu", "vp.cu", "probe.cu", "resize.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "--extended-lambda", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "..", "include", "lkflow_b200.h"))
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [os.path.join(HERE, s)] + hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", os.path.join(HERE, s), "-o", o], capture_output=True, text=True)
        return s, r

    with ThreadPoolExecutor(max_workers=8) as ex:
        for s, r in ex.map(cc, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- nvcc {s}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
            with open(os.path.join(bdir, s + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
    if force or jobs or _stale(OUT, objs):
        r = subprocess.run([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
