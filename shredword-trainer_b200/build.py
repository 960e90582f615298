#!/usr/bin/env python
"""Build every native artefact of the B200 BPE trainer IN-TREE (they travel to the GPU box with the snapshot).

  shredword/lib/libtrainer.so   CUDA (sm_100a) trainer + encoder, host control, C ABI   -- what shredword/cbase.py loads
  build/trainer.exe             CLI drop-in (csrc/cli_main.cpp), linked against libtrainer.so
  build/gen_corpus              synthetic corpus generator (tools/gen_corpus.c)

nvcc cross-compiles for sm_100a without a GPU.  Rebuilds only when a source is newer than the target.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "shredword", "lib", "libtrainer.so")
EXE = os.path.join(HERE, "build", "trainer.exe")
GEN = os.path.join(HERE, "build", "gen_corpus")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-fvisibility=hidden", "-cudart", "static", "-DSHRED_NVTX", "-ldl"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def build(force=False, verbose_ptxas=False):
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    srcs = [os.path.join(CSRC, "abi.cpp"), os.path.join(CSRC, "trainer_core.cpp"), os.path.join(CSRC, "cuda", "engine_cuda.cu"), os.path.join(CSRC, "cuda", "encoder_cuda.cu")]
    deps = srcs + [os.path.join(CSRC, h) for h in ("engine.hpp", "trainer_core.hpp", "exact_heap.hpp", "flat_map.hpp", "charset.hpp", "shard.hpp", "layout.hpp")] + \
        [os.path.join(CSRC, "cuda", h) for h in ("common.cuh", "kernels_tokenize.cuh", "kernels_scan.cuh", "kernels_ingest.cuh", "kernels_count.cuh", "kernels_fold.cuh", "kernels_dist.cuh", "kernels_merge.cuh", "kernels_encode.cuh")] + \
        [os.path.join(ROOT, "include", "shred_abi.h")]
    if force or _stale(LIB, deps):
        _run(["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-shared", "-o", LIB] + srcs)
    cli = os.path.join(CSRC, "cli_main.cpp")
    if force or _stale(EXE, [cli, LIB]):
        _run(["g++", "-O2", "-std=c++17", "-o", EXE, cli, "-L" + os.path.dirname(LIB), "-ltrainer", "-Wl,-rpath,$ORIGIN/../shredword/lib"])
    gen = os.path.join(HERE, "tools", "gen_corpus.c")
    if force or _stale(GEN, [gen]):
        _run(["gcc", "-O2", "-std=c11", "-Wall", "-o", GEN, gen, "-lpthread"])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
