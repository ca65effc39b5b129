"""Build ``libtiseg_b200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python tissue-image-segmentation_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  One object per .cu (parallel, incremental), then one shared library.
"""
import concurrent.futures
import glob
import hashlib
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


def _sha(*parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(p if isinstance(p, bytes) else p.encode())
        h.update(b"\0")
    return h.hexdigest()


def _sources():
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "tiseg_b200.h")]
    return srcs, hdrs


def source_hash():
    """sha256 over the contents of csrc/*.cu, csrc/*.cuh, include/tiseg_b200.h and the compiler flags.  The library
    carries the hash it was built from (``tiseg_build_hash``) and ``_lib.load`` refuses a stale binary."""
    srcs, hdrs = _sources()
    return _sha(" ".join(FLAGS), *[os.path.basename(f) + "\n" + open(f, "rb").read().decode("utf-8", "replace")
                                   for f in srcs + hdrs])[:32]


def _stamp(path):
    try:
        return open(path).read().strip()
    except OSError:
        return None


def build(force=False, verbose=False):
    """Objects and the library are reused only when the CONTENT they were built from is unchanged (hash stamps next to
    them), never by modification time: the .so is git-ignored but travels with the repo snapshot, so a time-based
    check could leave a stale binary behind a fresh checkout."""
    os.makedirs(OBJ, exist_ok=True)
    srcs, hdrs = _sources()
    total = source_hash()
    hdr_hash = _sha(*[open(h, "rb").read() for h in hdrs])
    objs, jobs = [], []
    for s in srcs:
        name = os.path.basename(s)[:-3]
        o = os.path.join(OBJ, name + ".o")
        # ctx.cu embeds the hash of the whole source tree
        want = _sha(" ".join(FLAGS), open(s, "rb").read(), hdr_hash, total if name == "ctx" else "")
        objs.append(o)
        if force or not os.path.exists(o) or _stamp(o + ".sha") != want:
            jobs.append((s, o, want))

    def cc(job):
        s, o, want = job
        extra = ["-DTISEG_SRC_HASH=\"%s\"" % total] if os.path.basename(s) == "ctx.cu" else []
        r = subprocess.run([NVCC] + FLAGS + extra + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, o, want, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, o, want, r in ex.map(cc, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % s)
            with open(os.path.join(OBJ, os.path.basename(s)[:-3] + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
            with open(o + ".sha", "w") as f:
                f.write(want)
    if force or jobs or not os.path.exists(SO) or _stamp(os.path.join(OBJ, "so.sha")) != total:
        tmp = SO + ".%d.tmp" % os.getpid()
        subprocess.check_call([NVCC, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                      "-cudart", "shared"])
        os.replace(tmp, SO)
        with open(os.path.join(OBJ, "so.sha"), "w") as f:
            f.write(total)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
