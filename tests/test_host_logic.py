"""CPU: the product's HOST control code (trainer_core.cpp: push ordering, versions, phantom pairs, exact heap replay,
save) linked against a CPU stand-in of the device engine that returns its records shuffled (tests/hostsim), checked
against the golden vectors of the unmodified reference and against the oracle's heap array."""
import ctypes
import os

import pytest

from cases import GOLDEN, SMALL, case_ids, corpus_bytes
from hostsim_lib import HS_SO
from oracle_lib import Oracle, md5


class _Cfg(ctypes.Structure):
    _fields_ = [("vs", ctypes.c_size_t), ("unk", ctypes.c_int32), ("cov", ctypes.c_float), ("mf", ctypes.c_uint64)]


class _Pair(ctypes.Structure):
    _fields_ = [("a", ctypes.c_int32), ("b", ctypes.c_int32)]


class _HeapEnt(ctypes.Structure):
    _fields_ = [("key", _Pair), ("freq", ctypes.c_uint64), ("version", ctypes.c_uint32)]


class _Trainer(ctypes.Structure):
    _fields_ = [("config", _Cfg), ("heap_data", ctypes.POINTER(_HeapEnt)), ("heap_size", ctypes.c_size_t), ("heap_cap", ctypes.c_size_t),
                ("words", ctypes.c_void_p), ("word_counts", ctypes.POINTER(ctypes.c_uint64)), ("n_words", ctypes.c_size_t),
                ("bm", ctypes.c_void_p * 2), ("next_token", ctypes.c_size_t), ("num_merges", ctypes.c_size_t), ("merge_ops", ctypes.POINTER(_Pair))]


@pytest.fixture(scope="module")
def hs(native):
    os.environ["SHRED_QUIET"] = "1"
    L = ctypes.CDLL(HS_SO)
    L.create_trainer.argtypes, L.create_trainer.restype = [ctypes.POINTER(_Cfg)], ctypes.POINTER(_Trainer)
    L.bpe_b200_load_buffer.argtypes, L.bpe_b200_load_buffer.restype = [ctypes.POINTER(_Trainer), ctypes.c_char_p, ctypes.c_size_t], ctypes.c_int
    for f in ("bpe_init", "bpe_count_bigrams", "bpe_trainer_destroy"):
        getattr(L, f).argtypes, getattr(L, f).restype = [ctypes.POINTER(_Trainer)], None
    L.bpe_train.argtypes, L.bpe_train.restype = [ctypes.POINTER(_Trainer)], ctypes.c_int
    L.bpe_merge_batch.argtypes, L.bpe_merge_batch.restype = [ctypes.POINTER(_Trainer), ctypes.c_int], ctypes.c_int
    L.bpe_save.argtypes, L.bpe_save.restype = [ctypes.POINTER(_Trainer), ctypes.c_char_p, ctypes.c_char_p], None
    return L


def _merges(t):
    n = min(t.contents.num_merges, max(t.contents.config.vs, 1))
    return [(t.contents.merge_ops[i].a, t.contents.merge_ops[i].b, 256 + i) for i in range(n)]


@pytest.fixture(params=["naive", "lists"])
def engine(request, monkeypatch):
    """naive: per-word vectors rewritten by a left-to-right pass.  lists: the CUDA engine's data structures (position-stable
    symbol array, lazily validated per-pair occurrence lists) walked sequentially through csrc/layout.hpp."""
    monkeypatch.setenv("SHRED_HOSTSIM_ENGINE", request.param)
    return request.param


@pytest.mark.parametrize("case", GOLDEN, ids=case_ids())
def test_host_logic_matches_reference(case, hs, tmp_path, engine):
    import struct
    data = corpus_bytes(case)
    vs, unk, cov, mf = case["config"]
    t = hs.create_trainer(ctypes.byref(_Cfg(vs, unk, cov, mf)))
    assert hs.bpe_b200_load_buffer(t, data, len(data)) == 0
    assert t.contents.n_words == case["n_words"]
    assert hs.bpe_train(t) == case["merges"]
    mb = b"".join(struct.pack("<3i", *m) for m in _merges(t))
    assert md5(mb) == case["merges_md5"]
    model, vocab = tmp_path / "m.bin", tmp_path / "v.txt"
    hs.bpe_save(t, os.fsencode(str(model)), os.fsencode(str(vocab)))
    assert model.read_bytes() == mb
    if case["vocab_md5"] is not None:
        assert md5(vocab.read_bytes()) == case["vocab_md5"]
    hs.bpe_trainer_destroy(t)


@pytest.mark.parametrize("case", [c for c in SMALL if c["name"] in ("kat_py", "kat_cpp", "rnd003", "rnd017", "multi600k_1", "zipf2m_1")], ids=lambda c: c["name"])
def test_heap_array_identical_stepwise(case, hs, engine):
    """After bpe_init and after every single merge the replay heap (pairs and frequencies, array order) equals the
    reference-order heap of the oracle.  Versions may differ (the host bumps them on demotion), so they are not compared."""
    data = corpus_bytes(case)
    vs, unk, cov, mf = case["config"]
    o = Oracle(vs, unk, cov, mf); o.load_bytes(data); o.init()
    t = hs.create_trainer(ctypes.byref(_Cfg(vs, unk, cov, mf)))
    hs.bpe_b200_load_buffer(t, data, len(data)); hs.bpe_init(t)
    steps = 0
    while True:
        ho = [(a, b, f) for a, b, f, _ in o.heap()]
        hh = [(t.contents.heap_data[i].key.a, t.contents.heap_data[i].key.b, t.contents.heap_data[i].freq) for i in range(t.contents.heap_size)]
        assert hh == ho, f"heap differs after {steps} merges"
        if steps >= 60 or steps >= vs - 256:
            break
        mo, mh = o.merge_batch(1), hs.bpe_merge_batch(t, 1)
        assert mo == mh
        if mo <= 0:
            break
        steps += 1
    assert _merges(t) == o.merges()
    hs.bpe_trainer_destroy(t); o.destroy()


def test_load_from_file_nul_and_empty(hs, tmp_path):
    """bpe_load_corpus(path): plain file, file with NUL bytes (the reference's fgets/strlen quirk, bpe.cpp:131-147), empty
    file, missing file -- against the oracle, which replays the reference's line loop."""
    hs.bpe_load_corpus.argtypes, hs.bpe_load_corpus.restype = [ctypes.POINTER(_Trainer), ctypes.c_char_p], ctypes.c_int
    blobs = {"plain": b"abc abd abc\nxyz abc\n" * 40,
             "nul": b"abc abd\x00hidden hidden hidden\nabc abd abc\x00\x00 zz\nplain line abc\n" * 9 + b"tail\x00gone",
             "long_nul_line": (b"ab " * 3000) + b"\x00" + (b"zz " * 3000) + b"\nab ab zz\n",
             "empty": b""}
    for name, data in blobs.items():
        p = tmp_path / (name + ".txt")
        p.write_bytes(data)
        o = Oracle(300, 0, 0.995, 1); o.load_corpus(str(p)); n = o.train()
        t = hs.create_trainer(ctypes.byref(_Cfg(300, 0, 0.995, 1)))
        assert hs.bpe_load_corpus(t, os.fsencode(str(p))) == 0
        assert t.contents.n_words == o.num_words, name
        assert hs.bpe_train(t) == n
        assert _merges(t) == o.merges(), name
        hs.bpe_trainer_destroy(t); o.destroy()
    t = hs.create_trainer(ctypes.byref(_Cfg(300, 0, 0.995, 1)))
    assert hs.bpe_load_corpus(t, os.fsencode(str(tmp_path / "missing.txt"))) == -1
    hs.bpe_trainer_destroy(t)


def test_replay_heap_matches_literal_reference(tmp_path):
    """The product's replay heap (packed words, payload side array, renumbering, adaptive split, wide fallback) against a
    literal restatement of the reference's heap rules (heap.cpp:53-114) on random operation sequences; tests/heap/bench_heap.cpp"""
    import subprocess
    exe = tmp_path / "bench_heap"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(os.path.dirname(os.path.abspath(__file__)), "heap", "bench_heap.cpp")], check=True)
    for seed in ("1", "2"):
        r = subprocess.run([str(exe), "check", seed], capture_output=True, text=True)
        assert r.returncode == 0 and "heap check ok" in r.stdout, r.stdout + r.stderr
