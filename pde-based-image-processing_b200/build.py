#!/usr/bin/env python
"""Build libpdegpu (CUDA, sm_100a) and the MEX gateways (C) in-tree.

    python pde-based-image-processing_b200/build.py [--force]

Outputs (git-ignored, they travel to the GPU box with the repo snapshot):
    pde-based-image-processing_b200/libpdegpu.so
    pde-based-image-processing_b200/gateways/pdegpu_mex.so   (13 gateways + mex shim, links libpdegpu)
nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GW = os.path.join(HERE, "gateways")
OBJ = os.path.join(HERE, "build" + os.environ.get("PDEGPU_BUILD_SUFFIX", ""))
LIB = os.path.join(HERE, os.environ.get("PDEGPU_BUILD_NAME", "libpdegpu.so"))
MEXLIB = os.path.join(GW, "pdegpu_mex.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]

GATEWAYS = [
    "Oflow_sor_elin4_2d", "Oflow_sor_llin4_2d", "Oflow_sor_llin8_2d", "Oflow_lhs_elin4_2d", "Oflow_lhs_llin4_2d",
    "Disp_sor_llin4_2d", "Disp_sor_llin_sym4_2d", "PDEsolver4", "PDEsolver8",
    "BilinInterp_2d", "FstDerivatives5", "SndDerivatives5", "DdiffWeights",
]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + cmd[0])
    return r.stdout


def build_lib(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(HERE, "..", "include", "pdegpu.h")]
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + NVCC_FLAGS + extra + os.environ.get("PDEGPU_NVCC_EXTRA", "").split() + ["-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        for o in outs:
            sys.stdout.write(o)
    if jobs or force or not _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"])
    return LIB


def build_gateways(force=False):
    srcs = [os.path.join(GW, g + ".c") for g in GATEWAYS if os.path.exists(os.path.join(GW, g + ".c"))]
    extra_c = [os.path.join(GW, "gw_ctx.c")]
    if not srcs:
        return None
    shim = os.path.join(GW, "mex_shim", "mex_shim.c")
    common = sorted(glob.glob(os.path.join(GW, "*.h"))) + sorted(glob.glob(os.path.join(GW, "mex_shim", "*.h")))
    deps = srcs + extra_c + [shim, LIB] + common
    if not force and _newer(MEXLIB, deps):
        return MEXLIB
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    cflags = ["-O2", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter", "-I" + os.path.join(GW, "mex_shim"),
              "-I" + os.path.join(HERE, "..", "include"), "-I" + GW]
    for s in srcs:
        name = os.path.basename(s)[:-2]
        o = os.path.join(OBJ, "gw_" + name + ".o")
        _run(["gcc"] + cflags + ["-DmexFunction=mex_" + name, "-c", s, "-o", o])
        objs.append(o)
    for s in extra_c + [shim]:
        o = os.path.join(OBJ, "gwx_" + os.path.basename(s)[:-2] + ".o")
        _run(["gcc"] + cflags + ["-c", s, "-o", o])
        objs.append(o)
    _run(["gcc", "-shared", "-o", MEXLIB] + objs + ["-L" + HERE, "-lpdegpu", "-Wl,-rpath,$ORIGIN/..", "-lm"])
    return MEXLIB


def build_cli(force=False):
    """pdegpu_flow_batch: the batch command-line front end (plain C on the C ABI)."""
    src = os.path.join(HERE, "cli", "pdegpu_flow_batch.c")
    exe = os.path.join(HERE, "cli", "pdegpu_flow_batch")
    if not os.path.exists(src):
        return None
    if force or not _newer(exe, [src, LIB, os.path.join(HERE, "..", "include", "pdegpu.h")]):
        _run(["gcc", "-O2", "-Wall", "-Wextra", "-I" + os.path.join(HERE, "..", "include"), src, "-L" + HERE, "-lpdegpu",
              "-Wl,-rpath,$ORIGIN/..", "-o", exe])
    return exe


def build_all(force=False, verbose=False):
    lib = build_lib(force, verbose)
    mex = build_gateways(force)
    build_cli(force)
    return lib, mex


if __name__ == "__main__":
    f = "--force" in sys.argv
    v = "--verbose" in sys.argv
    print(build_all(f, v))
