"""Builds and drives tests/hostsim/_build/libtrainer_hostsim.so: the product's HOST control code (abi.cpp +
trainer_core.cpp) linked against a CPU stand-in of the device engine.  Test infrastructure only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "shredword-trainer_b200", "csrc")
HS_DIR = os.path.join(ROOT, "tests", "hostsim")
HS_SO = os.path.join(HS_DIR, "_build", "libtrainer_hostsim.so")


def build_hostsim():
    srcs = [os.path.join(CSRC, "abi.cpp"), os.path.join(CSRC, "trainer_core.cpp"), os.path.join(HS_DIR, "engine_hostsim.cpp"), os.path.join(HS_DIR, "engine_listsim.cpp")]
    deps = srcs + [os.path.join(CSRC, h) for h in ("engine.hpp", "trainer_core.hpp", "exact_heap.hpp", "flat_map.hpp", "charset.hpp", "shard.hpp", "layout.hpp")]
    if os.path.exists(HS_SO) and all(os.path.getmtime(d) <= os.path.getmtime(HS_SO) for d in deps):
        return HS_SO
    os.makedirs(os.path.dirname(HS_SO), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-o", HS_SO] + srcs, check=True)
    return HS_SO
