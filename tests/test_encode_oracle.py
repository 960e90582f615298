"""CPU: the encoder oracle (oracle/bpe_encode_oracle.c) against the golden vectors produced by the unmodified Python
reference (shredword/utils/bpe.py:_encode_chunk / decode, see tests/golden/make_encode_golden.py)."""
import struct

import pytest

from encode_cases import CASES, MODELS, case_id, strip_delims, text_bytes
from oracle_lib import EncodeOracle, Oracle, md5


@pytest.fixture(scope="module")
def oracles(native):
    made = {name: EncodeOracle(merges) for name, merges in MODELS.items()}
    yield made
    for o in made.values():
        o.destroy()


@pytest.mark.parametrize("case", CASES, ids=case_id)
def test_oracle_matches_python_reference(case, oracles):
    o, data = oracles[case["model"]], text_bytes(case["text"])
    idb, offb = o.encode_bytes(data)
    assert len(idb) // 4 == case["n_ids"] and len(offb) // 8 == case["n_words"] + 1
    assert md5(idb) == case["ids_md5"] and md5(offb) == case["offsets_md5"]
    if "ids" in case:
        ids, off = o.encode(data)
        assert ids == case["ids"] and off == case["offsets"]
    if "decoded_md5" in case:
        ids = list(struct.unpack("<%di" % (len(idb) // 4), idb))
        dec = o.decode(ids)
        assert md5(dec) == case["decoded_md5"] and dec == strip_delims(data)


def test_model_validation_and_bad_ids(oracles):
    for bad in ([(97, 98, 257)], [(97, 98, 256), (300, 97, 257)], [(-1, 98, 256)], [(97, 256, 256)]):
        with pytest.raises(ValueError):
            EncodeOracle(bad)
    o = oracles["kat_py_300"]
    assert o.vocab_size == 256 + len(MODELS["kat_py_300"])
    with pytest.raises(ValueError):
        o.decode([97, o.vocab_size])
    with pytest.raises(ValueError):
        o.decode([-1])
    assert o.decode([]) == b"" and o.encode(b"") == ([], [0])


def test_encoding_the_training_corpus_reproduces_the_trainers_words(native):
    """BPE property that ties the encoder to the trainer: re-encoding a training word with the learned merges gives the
    symbol sequence the trainer ends with (every merge is applied to every word, leftmost first, in learning order)."""
    data = text_bytes("kat_py")
    t = Oracle(300, -1, 0.9995, 2); t.load_bytes(data); t.train()   # unk -1 + coverage 0.9995 keep the drop to one byte
    e = EncodeOracle(t.merges())
    keep, _ = t.keep_mask()
    checked = 0
    for i, (word, _) in enumerate(t.words()):
        if all(keep[b] for b in word):
            assert e.encode_word(word) == t.word_ids(i)
            checked += 1
    assert checked > 50
    t.destroy(); e.destroy()
