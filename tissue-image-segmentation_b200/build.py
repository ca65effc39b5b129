"""Build ``libtiseg_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python tissue-image-segmentation_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  One object per .cu (parallel, incremental), then one shared library.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libtiseg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--fmad=false", "-Xptxas", "-v"]
# --fmad=false: several kernels must reproduce fp32/fp64 library arithmetic bit for bit (OpenCV Sobel
# column pass, numpy softmax); contraction is enabled explicitly (fma()) where it is wanted.


def _newer(a, deps):
    return os.path.exists(a) and all(os.path.getmtime(a) >= os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "tiseg_b200.h")]
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(cc, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % s)
            with open(os.path.join(OBJ, os.path.basename(s)[:-3] + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
    if force or jobs or not _newer(SO, objs):
        tmp = SO + ".%d.tmp" % os.getpid()
        subprocess.check_call([NVCC, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                      "-cudart", "shared"])
        os.replace(tmp, SO)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
