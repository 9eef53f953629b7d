"""Build libgennerf_b200.so in-tree with plain nvcc for sm_100a (no torch headers: the
library is a C ABI).  `python -m gennerf_b200.build` or __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgennerf_b200.so")
SOURCES = ["error.cu", "lift.cu", "sample.cu", "sample_binned.cu", "planes.cu", "points.cu", "fusion.cu", "decoder_simt.cu", "decoder_tc.cu", "mlp_grad.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


TRACE_LIB = os.path.join(HERE, "libgennerf_b200_trace.so")


def build(force=False, verbose=False, trace=False):
    """Compile every CUDA source for sm_100a and link the shared library; returns its path.
    trace=True builds a second library (libgennerf_b200_trace.so, loaded when GNB_LIB_PATH points at it) with the
    decoder's per-phase clock64() tracing compiled in -- a profiling aid for tools/trace_decoder.py."""
    LIB = TRACE_LIB if trace else globals()["LIB"]
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "gennerf_b200.h")]
    if not force and not verbose and not _stale(LIB, srcs):
        return LIB                                   # up to date (e.g. the prebuilt .so on the GPU box)
    objdir = os.path.join(ROOT, "build", "obj_trace" if trace else "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "gennerf_b200.h"))
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DGNB_TC_TRACE"] if trace else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    # --trace: build the tracing variant of the library (see build())
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
