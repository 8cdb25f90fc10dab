"""build.py -- compiles the CUDA kernels + C ABI into cpuperformanceraytracer_b200/libb200pt.so
(in-tree, sm_100a only) and the host-side C++ mirror of the reference's render entry points into
libdemofox_b200.so + the render_offline CLI.

    python -m cpuperformanceraytracer_b200.build [--force]

nvcc cross-compiles without a GPU.  Two translation units hold the megakernel, one per
arithmetic policy, because the policies need different code generation flags:
  pt_kernels_parity.cu  --fmad=false -prec-div=true -prec-sqrt=true -ftz=false
  pt_kernels_fast.cu    --fmad=true
(and likewise pt_kernels_*_sorted.cu for the CTA-sorted scheduler, pt_kernels_*_v4sw.cu for the v4 kernels with the
reference's non-default shading switches compiled in).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libb200pt.so")
HOSTLIB = os.path.join(PKG, "libdemofox_b200.so")
CLI = os.path.join(PKG, "render_offline")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler",
          "-ffp-contract=off", "-I", os.path.join(ROOT, "include")]
IEEE = ["--fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]
# tuning knob for experiments: minimum resident CTAs per SM handed to __launch_bounds__
for _knob in ("B200PT_MIN_BLOCKS_CORNELL", "B200PT_MIN_BLOCKS_V4", "B200PT_MIN_BLOCKS_V3REDO", "B200PT_THREADS_CORNELL", "B200PT_THREADS_V4",
              "B200PT_THREADS_V3REDO"):
    if os.environ.get(_knob):
        COMMON = COMMON + ["-D%s=%s" % (_knob, os.environ[_knob])]

UNITS = [
    # (source, object, extra flags)
    (os.path.join(CSRC, "pt_kernels_parity.cu"), "pt_kernels_parity.o", IEEE),
    (os.path.join(CSRC, "pt_kernels_fast.cu"), "pt_kernels_fast.o", ["--fmad=true"]),
    (os.path.join(CSRC, "pt_kernels_parity_sorted.cu"), "pt_kernels_parity_sorted.o", IEEE),
    (os.path.join(CSRC, "pt_kernels_fast_sorted.cu"), "pt_kernels_fast_sorted.o", ["--fmad=true"]),
    (os.path.join(CSRC, "pt_kernels_parity_v4sw.cu"), "pt_kernels_parity_v4sw.o", IEEE),
    (os.path.join(CSRC, "pt_kernels_fast_v4sw.cu"), "pt_kernels_fast_v4sw.o", ["--fmad=true"]),
    (os.path.join(CSRC, "pt_post.cu"), "pt_post.o", IEEE),
    (os.path.join(CSRC, "b200pt_capi.cu"), "b200pt_capi.o", IEEE),
    (os.path.join(CSRC, "b200pt_group.cu"), "b200pt_group.o", IEEE),
    (os.path.join(HOST, "scene_setup.cpp"), "scene_setup.o", IEEE),
]
HEADERS = [os.path.join(CSRC, f) for f in ("pt_common.cuh", "pt_device.cuh", "pm_math.cuh", "pt_tonemap.cuh", "b200pt_context.h", "pt_wavefront.cuh")] + [
    os.path.join(HOST, "scene_setup.h"), os.path.join(ROOT, "include", "b200pt.h")]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    cc = nvcc()
    jobs = []
    for src, obj, extra in UNITS:
        o = os.path.join(OBJ, obj)
        if force or _stale(o, [src] + HEADERS + [__file__]):
            cmd = [cc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", o]
            jobs.append(cmd)
    logs = []
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            logs = list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, obj) for _, obj, _ in UNITS]
    if force or jobs or _stale(LIB, objs):
        _run([cc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    build_host(force)
    return "\n".join(logs)


def build_host(force=False):
    """C++ host mirror (reference entry-point names) + the offline CLI, linked against libb200pt.so."""
    cxx = shutil.which("g++") or "g++"
    srcs = [os.path.join(HOST, f) for f in ("demofox_render.cpp", "image_io.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        return
    hdrs = [os.path.join(HOST, f) for f in ("demofox_render.h", "image_io.h", "global_preprocessor_flags.h", "tiles.h")]
    flags = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", HOST]
    rpath = ["-Wl,-rpath,$ORIGIN"]
    if force or _stale(HOSTLIB, srcs + hdrs + [LIB]):
        _run([cxx] + flags + ["-shared", "-o", HOSTLIB] + srcs + ["-L", PKG, "-lb200pt"] + rpath)
    cli_src = os.path.join(HOST, "render_offline.cpp")
    if os.path.exists(cli_src) and (force or _stale(CLI, [cli_src, HOSTLIB] + hdrs)):
        _run([cxx] + flags + ["-o", CLI, cli_src, "-L", PKG, "-ldemofox_b200", "-lb200pt"] + rpath)


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    if out.strip():
        print(out)
    print("built", LIB)
