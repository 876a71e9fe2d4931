"""Builds librafting_b200.so in-tree with nvcc for sm_100a (no torch involved).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librafting_b200.so")
SOURCES = ["engine.cu"]
HEADERS = ["step_kernel.cuh", "step_body.inc", "seglog.cuh", "compact.cuh", "pair_kernel.cuh", "handlers.cuh", "tables.cuh",
           os.path.join("..", "..", "include", "rafting_b200.h")]
# the synthetic-stream generator (the simulated peers) is NOT part of the product library: tests, bench and the CPU
# reference arm load it on its own, so a process that times the CPU port never maps librafting_b200.so
WORKLOAD_LIB = os.path.join(HERE, "librafting_workload.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-Wno-deprecated-gpu-targets",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build librafting_b200.so")
    return exe


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    extra = os.environ.get("RAFTING_NVCC_EXTRA", "").split()      # e.g. -DRAFTING_MINBLOCKS=5 for tuning runs
    cmd = [nvcc()] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return LIB


def build_workload(force: bool = False) -> str:
    src = os.path.join(CSRC, "workload.cu")
    deps = [src, os.path.join(HERE, "..", "include", "rafting_workload.h"), os.path.join(HERE, "..", "include", "rafting_b200.h")]
    if not force and os.path.exists(WORKLOAD_LIB) and os.path.getmtime(WORKLOAD_LIB) > max(os.path.getmtime(d) for d in deps):
        return WORKLOAD_LIB
    res = subprocess.run([nvcc()] + NVCC_FLAGS + [src, "-o", WORKLOAD_LIB], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (workload) failed:\n" + (res.stdout + res.stderr)[-4000:])
    return WORKLOAD_LIB


FLAGS_LIB = os.path.join(HERE, "librafting_b200_flags.so")


def build_flags(force: bool = False) -> str:
    """The same engine compiled with -DRAFTING_ENABLE_CFG_FLAGS: the device branches of the two opt-in protocol fixes
    (RAFTING_CFG_STRICT_CANDIDATE_VOTE, RAFTING_CFG_LENIENT_FOLLOWER_COMMIT).  A library of its own: the default build stays
    byte-identical to the reference-faithful one and rejects a non-zero cfg.flags.  Loaded with RAFTING_B200_LIB=<path>."""
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(CSRC, h) for h in ("compact.cuh", "pair_kernel.cuh")]
    if not force and os.path.exists(FLAGS_LIB) and os.path.getmtime(FLAGS_LIB) > max(os.path.getmtime(d) for d in deps):
        return FLAGS_LIB
    cmd = [nvcc()] + NVCC_FLAGS + ["-DRAFTING_ENABLE_CFG_FLAGS"] + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", FLAGS_LIB, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc (flags build) failed:\n" + (res.stdout + res.stderr)[-4000:])
    return FLAGS_LIB


DURABLE_LIB = os.path.join(HERE, "librafting_durable.so")


def build_durable(force: bool = False) -> str:
    """Host-only durability journal (include/rafting_durable.h): plain g++, no CUDA."""
    src = os.path.join(CSRC, "durable.cpp")
    hdr = os.path.join(HERE, "..", "include", "rafting_durable.h")
    if not force and os.path.exists(DURABLE_LIB) and os.path.getmtime(DURABLE_LIB) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return DURABLE_LIB
    cxx = shutil.which("g++") or "g++"
    res = subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", src, "-o", DURABLE_LIB],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr[-4000:])
    return DURABLE_LIB


INGEST_LIB = os.path.join(HERE, "librafting_ingest.so")


def build_ingest(force: bool = False) -> str:
    """Host-only transport framing (include/rafting_ingest.h): plain g++, no CUDA."""
    src = os.path.join(CSRC, "ingest.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "rafting_ingest.h"), os.path.join(HERE, "..", "include", "rafting_b200.h")]
    if not force and os.path.exists(INGEST_LIB) and os.path.getmtime(INGEST_LIB) > max(os.path.getmtime(d) for d in deps):
        return INGEST_LIB
    cxx = shutil.which("g++") or "g++"
    res = subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wextra", src, "-o", INGEST_LIB],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stderr[-4000:])
    return INGEST_LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_flags(force=True))
    print(build_workload(force=True))
    print(build_durable(force=True))
    print(build_ingest(force=True))
