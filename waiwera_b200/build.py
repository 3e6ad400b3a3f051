"""Builds libwaiwera_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwaiwera_b200.so")
SOURCES = ["wb_core.cu", "wb_flow.cu", "wb_tracer.cu", "wb_linalg.cu", "wb_fused.cu", "wb_newton.cu"]
HEADERS = ["wb_common.cuh", "wb_linalg.cuh", "wb_state.cuh", "wb_tracer.cuh", "wb_eos.cuh", "wb_thermo.cuh", "wb_iapws_gen.cuh"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "waiwera_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu of the package for sm_100a into one shared library."""
    if not force and not _stale():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "_obj", src.replace(".cu", ".o"))
        objs.append(obj)
        # the physics translation unit is compiled without FMA contraction so that every evaluation of
        # F (residual kernel, Jacobian kernel, colouring loop) rounds identically, and identically to the
        # reference arithmetic order; the bandwidth-bound linear algebra keeps FMA
        extra = ["-fmad=false"] if src in ("wb_flow.cu", "wb_tracer.cu") else []
        cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
