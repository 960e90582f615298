"""A compiled C++ consumer of the C ABI that reads Trainer fields directly, like the reference's test/bpe_test.cpp
(tests/c_abi/consumer_test.cpp).  CPU: linked against the hostsim test library (host mirrors are produced by the
product's trainer_core.cpp); GPU: linked against libtrainer.so."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c_abi", "consumer_test.cpp")
CORPUS = os.path.join(HERE, "golden", "kat_cpp.txt")


def _build_and_run(libpath, tmp_path):
    exe = str(tmp_path / "consumer_test")
    libdir, libname = os.path.dirname(libpath), os.path.basename(libpath)
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, SRC, "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe, CORPUS, str(tmp_path)], capture_output=True, text=True, env=dict(os.environ, SHRED_QUIET="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "[PASS] heap top is the maximum" in r.stdout and "0 failed" in r.stdout


def test_c_consumer_host_mirrors(native, tmp_path):
    from hostsim_lib import HS_SO
    _build_and_run(HS_SO, tmp_path)


@pytest.mark.gpu
def test_c_consumer_on_gpu(native, tmp_path):
    _build_and_run(native["lib"], tmp_path)


@pytest.mark.gpu
def test_plain_c_encoder_consumer_on_gpu(native, tmp_path):
    """tests/c_abi/encoder_consumer.c: the encoder entry points from plain C (gcc -std=c11), end to end on the GPU"""
    exe = str(tmp_path / "encoder_consumer")
    libpath = native["lib"]
    libdir, libname = os.path.dirname(libpath), os.path.basename(libpath)
    subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-o", exe, os.path.join(HERE, "c_abi", "encoder_consumer.c"), "-L" + libdir, "-l:" + libname,
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe, CORPUS, str(tmp_path)], capture_output=True, text=True, env=dict(os.environ, SHRED_QUIET="1"))
    assert r.returncode == 0 and "0 failed" in r.stdout and "[FAIL]" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
