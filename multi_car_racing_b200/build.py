"""Build libmcr.so (hand-written sm_100a CUDA + the C-ABI of include/mcr.h) in-tree with nvcc.

    python -m multi_car_racing_b200.build [--force] [--verbose]

The library is compiled for sm_100a only (B200).  -fmad=false keeps fp32 arithmetic
un-contracted so the rigid-body solver and the rasteriser reproduce the reference's
(Box2D's) operation-by-operation IEEE results.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcr.so")
SOURCES = ["api.cu", "sim.cu", "carcontacts.cu", "raster.cu", "reset.cu", "trackgen.cu"]
HEADERS = [os.path.join(CSRC, "mcr_internal.h"), os.path.join(CSRC, "solver.cuh"), os.path.join(HERE, "..", "include", "mcr.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "--cudart", "static", "-shared",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: an instrumented side build (e.g. out='libmcr_clk.so', extra=['-DMCR_PHASE_CLOCKS'],
    loaded through MCR_LIB_PATH) -- diagnostics only, the product library is always LIB."""
    global_lib = LIB if out is None else os.path.join(HERE, out)
    if out is None and not force and not needs_build():
        return LIB
    tmp = global_lib + ".tmp%d" % os.getpid()
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("nvcc failed building libmcr.so")
    if verbose:
        print(proc.stdout)
    os.replace(tmp, global_lib)
    return global_lib


if __name__ == "__main__":
    if "--stg-store" in sys.argv:
        path = build(out="libmcr_stg.so", extra=["-DMCR_FILL_STG_STORE"], verbose="--verbose" in sys.argv)
    elif "--sweep-check" in sys.argv:      # A/B: sweeps between two fixed-point comparisons (4 by default)
        n = sys.argv[sys.argv.index("--sweep-check") + 1]
        path = build(out="libmcr_sc%s.so" % n, extra=["-DSWEEP_CHECK=%s" % n], verbose="--verbose" in sys.argv)
    elif "--sweep-unroll" in sys.argv:     # A/B: sweeps per trip of the inner loop (1 by default: the body stays in the L0 instruction cache)
        n = sys.argv[sys.argv.index("--sweep-unroll") + 1]
        path = build(out="libmcr_su%s.so" % n, extra=["-DSWEEP_UNROLL=%s" % n], verbose="--verbose" in sys.argv)
    elif "--lane-bulk" in sys.argv:
        path = build(out="libmcr_lb.so", extra=["-DMCR_FILL_LANE_BULK"], verbose="--verbose" in sys.argv)
    elif "--phase-clocks" in sys.argv:
        path = build(out="libmcr_clk.so", extra=["-DMCR_PHASE_CLOCKS"], verbose="--verbose" in sys.argv)
    else:
        path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
