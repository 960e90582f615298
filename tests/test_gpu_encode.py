"""GPU (-m gpu): the CUDA encoder / decoder, called through the C ABI (bpe_b200_encoder_* in include/shred_abi.h), against
the golden vectors of the unmodified Python reference (tests/golden/encode_golden.json) and the CPU oracle
(oracle/bpe_encode_oracle.c) on the same inputs."""
import os
import struct

import numpy as np
import pytest

from corpora import SM, generated_corpus, random_config, random_corpus
from encode_cases import CASES, MODELS, case_id, strip_delims, text_bytes
from oracle_lib import EncodeOracle, Oracle, md5

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(native):
    os.environ["SHRED_QUIET"] = "1"
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from shredword import BPEEncoder
    return BPEEncoder


@pytest.fixture(scope="module")
def encoders(E):
    made = {name: E(merges=merges) for name, merges in MODELS.items()}
    yield made
    for e in made.values():
        e.destroy()


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_encode_matches_python_reference_golden(case, encoders):
    e, data = encoders[case["model"]], text_bytes(case["text"])
    ids, off = e.encode_bytes(data)
    assert ids.size == case["n_ids"] and off.size == case["n_words"] + 1
    assert md5(ids.astype("<i4").tobytes()) == case["ids_md5"]
    assert md5(off.astype("<u8").tobytes()) == case["offsets_md5"]
    if "ids" in case:
        assert ids.tolist() == case["ids"] and off.tolist() == case["offsets"]
    if "decoded_md5" in case:
        dec = e.decode_bytes(ids)
        assert md5(dec) == case["decoded_md5"] and dec == strip_delims(data)
    st = e.stats()
    assert st["n_words"] == case["n_words"] and st["n_ids"] == case["n_ids"] and st["text_bytes"] == len(data)


def _seam_text(seed):
    """words that start at, end at and straddle the 16-byte thread spans, the 4 KB block units and the 128-symbol
    shared-memory limit of the per-word kernel"""
    r = SM(seed)
    out = bytearray()
    while len(out) < 20000:
        L = r.pick([1, 2, 3, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 130, 255, 256, 257, 1000, 4095, 4096, 4097])
        alpha = r.pick([b"ab", b"a", b"abc", b"the quick brown fox".replace(b" ", b"")])
        out += bytes(r.pick(alpha) for _ in range(L))
        out += r.pick([b" ", b"\n", b"  ", b"\r\n", b"\t"])
    return bytes(out)


def test_seams_long_words_and_nul_bytes(E):
    models = {"runs_ab": MODELS["runs_ab"], "handmade_dup": MODELS["handmade_dup"], "zipf2m_1000": MODELS["zipf2m_1000"]}
    texts = [_seam_text(s) for s in range(4)] + [b"a" * 5000, b"ab" * 4097, b"aaa\x00bbb a\x00 \x00", bytes(range(256)) * 40, b"x" * 4096 + b" " + b"y" * 4095 + b" z"]
    for name, merges in models.items():
        e, o = E(merges=merges), EncodeOracle(merges)
        for i, data in enumerate(texts):
            ids, off = e.encode_bytes(data)
            oi, oo = o.encode(data)
            assert off.tolist() == oo, (name, i)
            assert ids.tolist() == oi, (name, i)
            assert e.decode_bytes(ids) == o.decode(oi), (name, i)
        e.destroy(); o.destroy()


def test_fuzz_random_models_and_texts(E):
    """models trained by the CPU oracle on random corpora x other random corpora as text"""
    bad = []
    for seed in range(300, 360):
        train = random_corpus(seed)
        vs, unk, cov, mf = random_config(seed)
        t = Oracle(max(vs, 300), unk, cov, min(mf, 3)); t.load_bytes(train); t.train()
        merges = t.merges()
        t.destroy()
        e, o = E(merges=merges), EncodeOracle(merges)
        for data in (train, random_corpus(seed + 1000), random_corpus(seed + 2000)):
            ids, off = e.encode_bytes(data)
            oi, oo = o.encode(data)
            if ids.tolist() != oi or off.tolist() != oo or e.decode_bytes(ids) != strip_delims(data):
                bad.append(seed)
        e.destroy(); o.destroy()
    assert not bad


def test_api_contract(E, tmp_path):
    from shredword import BPETrainer
    from shredword.cbase import EncodeStats, lib
    import ctypes
    assert ctypes.sizeof(EncodeStats) == 8 * 8 + 7 * 8
    for badm in ([(97, 98, 257)], [(97, 98, 256), (300, 97, 257)], [(-1, 98, 256)], [(97, 256, 256)]):
        with pytest.raises(ValueError):
            E(merges=badm)
    with pytest.raises(ValueError):
        E(model_path=str(tmp_path / "missing.bin"))
    (tmp_path / "odd.bin").write_bytes(b"\x00" * 13)
    with pytest.raises(ValueError):
        E(model_path=str(tmp_path / "odd.bin"))
    with pytest.raises(RuntimeError):
        E().encode("no model")
    # the trainer's own files, end to end: train -> save -> load -> encode -> decode
    data = text_bytes("kat_py")
    t = BPETrainer(300, 0, 0.995, 2); t.load_bytes(data); assert t.train() == 44
    model, vocab = str(tmp_path / "m.bin"), str(tmp_path / "v.txt")
    t.save(model, vocab); t.destroy()
    e = E(model_path=model)
    assert e.vocab_size == 300
    text = "the quick brown fox jumps over the lazy dog"
    ids = e.encode(text)
    assert isinstance(ids, list) and e.decode(ids) == text.replace(" ", "")
    o = EncodeOracle(model_path=model)
    assert ids == o.encode(text.encode())[0]
    with pytest.raises(ValueError):      # reference utils/bpe.py:220
        e.decode([97, 300])
    with pytest.raises(ValueError):
        e.decode([-1])
    assert e.decode([]) == "" and e.encode("") == [] and e.encode(" \n\t ") == []
    ids2, off2 = e.encode_bytes(b"")
    assert ids2.size == 0 and off2.tolist() == [0]
    # fetch without a result, NULL handles
    assert lib.bpe_b200_encode_fetch(None, None, None) == -1 and lib.bpe_b200_encode(None, None, 0, None, None) == -1
    assert lib.bpe_b200_decode(None, None, 0, None, 0) == -1 and lib.bpe_b200_encoder_vocab_size(None) == 0
    lib.bpe_b200_encoder_destroy(None)
    # two encoders and a trainer side by side
    e2 = E(merges=MODELS["kat_cpp_500"])
    a = e.encode_bytes(data)[0]; b = e2.encode_bytes(data)[0]; a2 = e.encode_bytes(data)[0]
    assert np.array_equal(a, a2) and a.size != b.size
    e.destroy(); e2.destroy(); o.destroy()
    e.destroy()  # idempotent


def test_encoding_the_training_corpus_reproduces_the_trainers_words_at_scale(E, tmp_path):
    """100 MB: train on the GPU, save, load the model into the encoder, encode the same corpus.  Size-independent checks:
    decode(encode(text)) == text without delimiters; every word's ids == the trainer's final symbols of that word (the
    encoder replays the merges the trainer applied); and the whole id stream equals the CPU oracle's."""
    from shredword import BPETrainer
    p = generated_corpus(str(tmp_path / "z100.txt"), 100_000_000, 1, 20, "zipf")
    data = np.fromfile(p, dtype=np.uint8)
    t = BPETrainer(8192, -1, 0.9995, 2000); t.load_corpus(p); n = t.train()
    assert n > 1000
    model = str(tmp_path / "m.bin"); t.save(model, str(tmp_path / "v.txt"))
    e = E(model_path=model)
    ids, off = e.encode_bytes(data)
    st = e.stats()
    assert st["n_unique_words"] == t.num_words and off.size == st["n_words"] + 1 and int(off[-1]) == ids.size
    # round trip
    assert e.decode_bytes(ids) == strip_delims(data.tobytes())
    # oracle on the full text
    o = EncodeOracle(model_path=model)
    oi, oo = o.encode_bytes(data.tobytes())
    assert md5(ids.tobytes()) == md5(oi) and md5(off.tobytes()) == md5(oo)
    # the trainer's final words: encode each distinct word once more, as its own text, in the trainer's word order
    keep, _ = t.charset()
    words = t.words()
    o2 = Oracle(8192, -1, 0.9995, 2000); o2.load_corpus(p)          # word bytes in the same (reference) order
    wb = o2.words()
    assert len(wb) == len(words)
    sample = [i for i in range(0, len(wb), 7) if all(keep[b] for b in wb[i][0])]
    joined = b" ".join(wb[i][0] for i in sample)
    sid, soff = e.encode_bytes(joined)
    assert soff.size == len(sample) + 1
    for k, i in enumerate(sample):
        assert sid[int(soff[k]):int(soff[k + 1])].tolist() == words[i][0], i
    for x in (t, e, o, o2):
        x.destroy()


@pytest.mark.parametrize("piece", [1, 7, 64, 4096, 1 << 20])
def test_streamed_encode_equals_resident_encode(piece, E, monkeypatch):
    """bpe_b200_encode_to_host cuts the text at delimiters into pieces and overlaps H2D | encode | D2H; ids and (global)
    offsets must not depend on the cut.  Tiny pieces put a cut after nearly every word."""
    monkeypatch.setenv("SHRED_ENCODE_PIECE_BYTES", str(piece))
    for mname, tnames in (("kat_py_300", ["kat_py", "crlf_tabs", "empty", "delims_only", "one_byte", "all_bytes", "no_trailing"]),
                          ("zipf2m_1000", ["zipf2m", "long_words", "kat_cpp"]), ("handmade_dup", ["runs", "rnd005"])):
        e, o = E(merges=MODELS[mname]), EncodeOracle(MODELS[mname])
        for tname in tnames:
            data = text_bytes(tname)
            if piece < 64 and len(data) > 200_000:
                data = data[:200_000]
            ids, off = e.encode_bytes_streamed(data)
            oi, oo = o.encode_bytes(data)
            assert ids.tobytes() == oi and off.tobytes() == oo, (mname, tname)
            st = e.stats()
            assert st["n_words"] == off.size - 1 and st["n_ids"] == ids.size
        # capacities: exact fit works, one short fails with -3 (ValueError)
        data = text_bytes("kat_cpp")
        ids, off = e.encode_bytes_streamed(data)
        a = np.empty(ids.size, dtype=np.int32); b = np.empty(off.size, dtype=np.uint64)
        buf = np.frombuffer(data, dtype=np.uint8)
        assert e.encode_to_host_raw(buf.ctypes.data, buf.size, a.ctypes.data, a.size, b.ctypes.data, b.size) == (off.size - 1, ids.size)
        assert np.array_equal(a, ids) and np.array_equal(b, off)
        with pytest.raises(ValueError):
            e.encode_to_host_raw(buf.ctypes.data, buf.size, a.ctypes.data, a.size - 1, b.ctypes.data, b.size)
        with pytest.raises(ValueError):
            e.encode_to_host_raw(buf.ctypes.data, buf.size, a.ctypes.data, a.size, b.ctypes.data, b.size - 1)
        # ids only (offsets_out NULL)
        a2 = np.empty(ids.size, dtype=np.int32)
        assert e.encode_to_host_raw(buf.ctypes.data, buf.size, a2.ctypes.data, a2.size, None, 0) == (off.size - 1, ids.size) and np.array_equal(a2, ids)
        # the resident path still works on the same encoder afterwards
        assert e.encode_bytes(data)[0].tobytes() == ids.tobytes()
        e.destroy(); o.destroy()
