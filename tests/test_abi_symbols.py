"""CPU: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/shred_abi.h
declares, including the 21 names the reference binding resolves at import (reference shredword/cbase.py:50-71)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_BOUND_SYMBOLS = [  # reference shredword/cbase.py:50-71
    "create_trainer", "bpe_trainer_destroy", "bpe_init", "bpe_count_bigrams", "bpe_load_corpus", "bpe_merge_batch", "bpe_train", "bpe_save",
    "trainerCreate", "trainerDestroy", "addTextToTrainer", "preprocessTexts", "extractInitialSubwords", "computeLoss", "computeTokenLoss",
    "pruneVocabStep", "updateTokenScores", "trainUnigram", "getVocab", "saveVocab", "loadVocab"]


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "shred_abi.h")).read()
    return re.findall(r"^SHRED_API [^;(]*?\b([A-Za-z_][A-Za-z_0-9]*)\(", src, re.M)


def test_library_exports_every_declared_symbol(native):
    lib = ctypes.CDLL(native["lib"])
    names = declared_symbols()
    assert len(names) >= 27
    for n in names + REFERENCE_BOUND_SYMBOLS:
        assert hasattr(lib, n), n


def test_only_abi_symbols_are_exported(native):
    out = subprocess.run(["nm", "-D", "--defined-only", native["lib"]], check=True, capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(declared_symbols()), exported ^ set(declared_symbols())


def test_library_carries_sm100a_code(native):
    out = subprocess.run(["cuobjdump", "-lelf", native["lib"]], check=True, capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layouts_match_reference(native):
    """BPEConfig 24 B, heap entry 24 B, Trainer prefix offsets."""
    sys.path.insert(0, os.path.join(ROOT, "shredword-trainer_b200"))
    from shredword import cbase
    assert ctypes.sizeof(cbase.BPEConfig) == 24 and ctypes.sizeof(cbase.BPEHeapEntry) == 24
    T = cbase.Trainer
    assert (T.heap.offset, T.corpus.offset, T.bigram_map.offset, T.num_merges.offset, T.merge_ops.offset, T.impl.offset) == (24, 48, 72, 96, 104, 128)
    # the ctypes mirrors of the extension structs have the C compiler's sizes
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "sz.c")
        open(src, "w").write('#include <stdio.h>\n#include "shred_abi.h"\nint main(void){printf("%zu %zu\\n", sizeof(shred_stats_t), sizeof(shred_encode_stats_t));return 0;}\n')
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "sz"), src], check=True)
        c_stats, c_enc = map(int, subprocess.run([os.path.join(d, "sz")], check=True, capture_output=True, text=True).stdout.split())
    assert ctypes.sizeof(cbase.Stats) == c_stats and ctypes.sizeof(cbase.EncodeStats) == c_enc


def test_fails_loudly_without_gpu(native):
    """No CPU fallback: creating a trainer on a machine without a CUDA device reports the error and exits non-zero."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    code = "import sys; sys.path.insert(0, %r); from shredword import BPETrainer; BPETrainer(300); print('created')" % os.path.join(ROOT, "shredword-trainer_b200")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "created" not in r.stdout and "no CPU fallback" in r.stderr


def test_encoder_fails_loudly_without_gpu(native, capfd):
    """The encoder has no CPU path either: model validation runs on the host, creation needs the device."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "shredword-trainer_b200"))
    from shredword import BPEEncoder, cbase
    assert ctypes.sizeof(cbase.EncodeStats) == 8 * 8 + 7 * 8
    with pytest.raises(ValueError):
        BPEEncoder(merges=[(97, 98, 257)])          # not what bpe_save writes: rejected before any device call
    assert "not a BPE model" in capfd.readouterr().err
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ValueError):
        BPEEncoder(merges=[(97, 98, 256)])
    assert "CUDA" in capfd.readouterr().err
    with pytest.raises(RuntimeError):
        BPEEncoder().encode("no model")


def test_generator_is_deterministic(native, tmp_path):
    import hashlib
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    subprocess.run([native["gen"], str(a), "100000", "1", "12", "zipf", "1"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([native["gen"], str(b), "100000", "1", "12", "zipf", "8"], check=True, stdout=subprocess.DEVNULL)
    assert a.read_bytes() == b.read_bytes() and len(a.read_bytes()) == 114658
    assert hashlib.md5(a.read_bytes()).hexdigest() == hashlib.md5(b.read_bytes()).hexdigest()
