#!/usr/bin/env python
"""Regenerates tests/golden/encode_golden.json from the UNMODIFIED reference (needs /root/reference: dev container only).

The reference's encoder is pure Python: BPETokenizer._encode_chunk / decode in shredword/utils/bpe.py:191-225.  This script
imports that module as it lies under /root/reference, fills `tokenizer.merges` from model files written by the pinned
reference *trainer* (oracle/_ref, same recipe as make_golden.py) -- i.e. `merges[(a, b)] = new_id` in file order -- and
records what `_encode_chunk` returns for every whitespace-delimited word (bpe.cpp:131-152) of a set of texts, plus the
bytes `vocab[id]` concatenates to (decode, :217-220).

Small cases keep the full id list; large ones keep md5s of the int32 / uint64 little-endian arrays.
"""
import hashlib
import importlib.util
import json
import os
import re
import struct
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
from corpora import SM, generated_corpus, random_corpus  # noqa: E402
from oracle_lib import run_reference  # noqa: E402

REF_MODULE = "/root/reference/shredword/utils/bpe.py"


def md5(b):
    return hashlib.md5(b).hexdigest()


def reference_tokenizer(merge_list):
    spec = importlib.util.spec_from_file_location("ref_utils_bpe", REF_MODULE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tok = mod.BPETokenizer()
    tok.merges = {}
    for a, b, c in merge_list:
        tok.merges[(a, b)] = c                       # what BaseTokenizer.load does per line (utils/bpe.py:150-153)
    try:
        tok.vocab = mod.build_vocab(tok.merges, {})
    except KeyError:                                 # a pair listed twice leaves its first id undefined in build_vocab
        tok.vocab = None
    return tok


def encode_with_reference(tok, text):
    words = [w for w in re.split(rb"[\t\n\r ]+", text) if w]
    cache, ids, off = {}, [], [0]
    for w in words:
        if w not in cache:
            cache[w] = tok._encode_chunk(w)
        ids.extend(cache[w])
        off.append(len(ids))
    return ids, off


def adversarial_texts():
    r = SM(99)
    runs = b" ".join(bytes([r.pick(b"ab")]) * r.pick([1, 2, 3, 4, 5, 7, 8, 31, 32, 33, 64, 65, 100]) for _ in range(60))
    longw = bytes(r.pick(b"abcab") for _ in range(3000)) + b"\n" + bytes(r.pick(b"the quick") for _ in range(700)).replace(b" ", b"_")
    allbytes = bytes(range(256)) * 3 + b" " + bytes(reversed(range(256)))
    return {"runs": runs, "long_words": longw, "all_bytes": allbytes, "empty": b"", "delims_only": b" \n\t\r\r\n  ", "one_byte": b"a",
            "no_trailing": b"the cat sat", "crlf_tabs": b"the\tcat\r\nsat  on\r\n\r\nthe mat\t\t"}


def main():
    tmp = tempfile.mkdtemp()
    kat_py = open(os.path.join(HERE, "kat_py.txt"), "rb").read()
    kat_cpp = open(os.path.join(HERE, "kat_cpp.txt"), "rb").read()
    z2 = open(generated_corpus(os.path.join(tmp, "z2.txt"), 2_000_000, 7, 16, "zipf"), "rb").read()
    m6 = open(generated_corpus(os.path.join(tmp, "m6.txt"), 600_000, 7, 14, "multi"), "rb").read()
    z2b = open(generated_corpus(os.path.join(tmp, "z2b.txt"), 1_000_000, 8, 17, "zipf"), "rb").read()   # unseen text, same spelling family
    adv = adversarial_texts()

    def trained(corpus, cfg):
        p = os.path.join(tmp, "c.txt")
        open(p, "wb").write(corpus)
        merges, _, _ = run_reference(p, *cfg, tmp)
        return [list(struct.unpack_from("<3i", merges, 12 * i)) for i in range(len(merges) // 12)]

    models = {
        "kat_py_300": trained(kat_py, (300, 0, 0.995, 2)),
        "kat_cpp_500": trained(kat_cpp, (500, 0, 0.995, 1)),
        "zipf2m_1000": trained(z2, (1000, 0, 0.9995, 5)),
        "zipf2m_5000": trained(z2, (5000, 0, 0.995, 40)),
        "multi600k_2000": trained(m6, (2000, 0, 0.995, 2)),
        "runs_ab": trained(adv["runs"] * 20, (300, -1, 0.9995, 1)),
        "empty_model": [],
        # hand-made: (a,a) self pairs, chains, and a pair listed twice (the later id wins in a Python dict)
        "handmade_dup": [[97, 97, 256], [256, 256, 257], [97, 98, 258], [98, 97, 259], [97, 97, 260], [260, 260, 261], [258, 259, 262]],
    }
    texts = {"kat_py": kat_py, "kat_cpp": kat_cpp, "zipf2m": z2, "zipf1m_unseen": z2b, "multi600k": m6}
    texts.update(adv)
    for s in (3, 5, 17, 23, 31, 44):
        texts["rnd%03d" % s] = random_corpus(s)
    text_desc = {"kat_py": {"file": "kat_py.txt"}, "kat_cpp": {"file": "kat_cpp.txt"}, "zipf2m": {"gen": [2_000_000, 7, 16, "zipf"]},
                 "zipf1m_unseen": {"gen": [1_000_000, 8, 17, "zipf"]}, "multi600k": {"gen": [600_000, 7, 14, "multi"]}}
    for k, v in adv.items():
        text_desc[k] = {"hex": v.hex()}
    for s in (3, 5, 17, 23, 31, 44):
        text_desc["rnd%03d" % s] = {"random_seed": s}

    plan = {
        "kat_py_300": ["kat_py", "kat_cpp", "crlf_tabs", "no_trailing", "all_bytes", "empty", "delims_only", "one_byte", "rnd003", "rnd017"],
        "kat_cpp_500": ["kat_cpp", "kat_py", "long_words", "rnd005"],
        "zipf2m_1000": ["zipf2m", "zipf1m_unseen", "kat_py", "long_words", "rnd023"],
        "zipf2m_5000": ["zipf2m", "zipf1m_unseen", "rnd031"],
        "multi600k_2000": ["multi600k", "all_bytes", "zipf1m_unseen", "rnd044"],
        "runs_ab": ["runs", "long_words", "rnd005"],
        "empty_model": ["kat_cpp", "all_bytes", "empty"],
        "handmade_dup": ["runs", "long_words", "rnd005", "one_byte"],
    }
    cases = []
    for mname, tnames in plan.items():
        tok = reference_tokenizer(models[mname])
        for tname in tnames:
            text = texts[tname]
            ids, off = encode_with_reference(tok, text)
            idb, offb = struct.pack("<%di" % len(ids), *ids), struct.pack("<%dQ" % len(off), *off)
            rec = {"model": mname, "text": tname, "n_words": len(off) - 1, "n_ids": len(ids), "ids_md5": md5(idb), "offsets_md5": md5(offb)}
            if len(ids) <= 4000:
                rec["ids"], rec["offsets"] = ids, off
            if tok.vocab is not None:
                dec = b"".join(tok.vocab[i] for i in ids)       # decode, utils/bpe.py:217-220
                assert dec == b"".join(w for w in re.split(rb"[\t\n\r ]+", text) if w)
                rec["decoded_md5"] = md5(dec)
                if tname in ("kat_py", "kat_cpp", "no_trailing"):
                    assert tok.decode(ids) == dec.decode("utf-8")  # the reference's own decode() agrees
            cases.append(rec)
            print(mname, tname, rec["n_words"], rec["n_ids"], flush=True)
    json.dump({"how": "reference shredword/utils/bpe.py BPETokenizer._encode_chunk (unmodified, imported from /root/reference) with merges from "
                      "model files of the pinned reference trainer (oracle/_ref)",
               "models": models, "texts": text_desc, "cases": cases}, open(os.path.join(HERE, "encode_golden.json"), "w"), separators=(",", ":"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
