"""CPU: the oracle restatement (oracle/bpe_oracle.c) against the golden vectors produced by the unmodified reference."""
import os

import pytest

from cases import GOLDEN, case_ids, corpus_bytes
from oracle_lib import Oracle, md5


@pytest.mark.parametrize("case", GOLDEN, ids=case_ids())
def test_oracle_matches_reference(case, native, tmp_path):
    data = corpus_bytes(case)
    assert md5(data) == case["corpus_md5"], "corpus generator drifted"
    vs, unk, cov, mf = case["config"]
    o = Oracle(vs, unk, cov, mf, verify=len(data) < 200_000)  # verify: recompute_freq by the reference's literal scan
    o.load_bytes(data)
    assert o.num_words == case["n_words"]
    n = o.train()
    assert n == case["merges"]
    assert md5(o.merges_bytes()) == case["merges_md5"]
    if "merge_list" in case:
        assert [list(m) for m in o.merges()] == case["merge_list"]
    model, vocab = tmp_path / "m.bin", tmp_path / "v.txt"
    o.save(str(model), str(vocab))
    assert model.read_bytes() == o.merges_bytes()
    if case["vocab_md5"] is not None:
        assert md5(vocab.read_bytes()) == case["vocab_md5"]
    o.destroy()


def test_known_answers_from_survey():
    """first merges and hashes of the two known-answer corpora."""
    by = {c["name"]: c for c in GOLDEN}
    assert by["kat_py"]["merges"] == 44 and by["kat_py"]["merges_md5"] == "f2366aa6e59bced605e986b48453749b"
    assert by["kat_py"]["vocab_md5"] == "eeea84c9d962d22beda18a9bf37e2d21"
    assert [m[:2] for m in by["kat_py"]["merge_list"][:5]] == [[105, 110], [101, 110], [116, 105], [111, 110], [256, 103]]
    assert by["kat_cpp"]["merges"] == 44 and by["kat_cpp"]["merges_md5"] == "22f542f833b00b30de92a9f4fafa379d"
    assert by["kat_cpp"]["vocab_md5"] == "346a0172e56b6ad260a9c5ea298a94c8"
    assert [m[:2] for m in by["kat_cpp"]["merge_list"][:4]] == [[104, 101], [114, 111], [110, 103], [105, 258]]


def test_oracle_behaviours(native, tmp_path):
    # second load replaces the first (reference bpe.cpp:176-183)
    a, b = b"aa bb aa cc aa\n" * 50, b"xyz xyz xy zz\n" * 50
    o1 = Oracle(300, 0, 0.995, 1); o1.load_bytes(a); o1.load_bytes(b); o1.train()
    o2 = Oracle(300, 0, 0.995, 1); o2.load_bytes(b); o2.train()
    assert o1.merges() == o2.merges() and o1.words() == o2.words()
    # empty corpus: no merges, 256-line vocab, empty model
    o = Oracle(300, 0, 0.995, 1); o.load_bytes(b""); assert o.train() == 0
    o.save(str(tmp_path / "m"), str(tmp_path / "v"))
    assert (tmp_path / "m").read_bytes() == b"" and (tmp_path / "v").read_bytes().count(b"\n") == 256 + 1  # token 10 is '\n'
    # missing file
    with pytest.raises(IOError):
        Oracle(300).load_corpus(str(tmp_path / "missing.txt"))
    # step-wise API == train (reference bpe.h:66-69)
    data = open(os.path.join(os.path.dirname(__file__), "golden", "kat_cpp.txt"), "rb").read()
    t = Oracle(300, -1, 0.99, 2); t.load_bytes(data); t.train()
    s = Oracle(300, -1, 0.99, 2); s.load_bytes(data); s.init()
    while s.num_merges < 44 and s.merge_batch(1) > 0:
        pass
    assert s.merges() == t.merges()
