"""Builds libsrk.so (all CUDA kernels + the C ABI of include/srk.h) for sm_100a, in-tree.

    python food101-super-resolution_b200/build.py [--force] [--probes] [--debug-knobs]

Objects are cached per source file (rebuilt when the source or a header is newer).
--probes        also compiles csrc/srk_probe_mma.cu (MMA-rate / TMEM-read micro-benchmarks behind srk_tc_probe >= 1000)
--debug-knobs   compiles the bring-up knobs in (-DSRK_DEBUG_KNOBS: SRK_TC_DBG / SRK_RGB_DBG skip stores / MMAs / loads;
                results are wrong by design).  Neither is part of the product library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libsrk.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False, probes=False, debug_knobs=False):
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and (probes or f != "srk_probe_mma.cu"))
    flags = FLAGS + (["-DSRK_WITH_PROBES"] if probes else []) + (["-DSRK_DEBUG_KNOBS"] if debug_knobs else [])
    stamp = os.path.join(OBJDIR, ".flags")
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(flags):
        force = True
        open(stamp, "w").write(" ".join(flags))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "srk.h"))
    jobs = []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJDIR, s[:-3] + ".o")
        if force or _newer(src, obj) or any(_newer(h, obj) for h in hdrs):
            jobs.append((src, obj))

    def cc(job):
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", job[0], "-o", job[1]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (job[0], r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        logs = list(ex.map(cc, jobs))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    objs = [os.path.join(OBJDIR, s[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probes="--probes" in sys.argv,
                debug_knobs="--debug-knobs" in sys.argv))
