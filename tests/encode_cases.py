"""Encoder golden cases (tests/golden/encode_golden.json, from the unmodified Python reference) -> model + text bytes."""
import json
import os
import tempfile

from corpora import generated_corpus, random_corpus

HERE = os.path.dirname(os.path.abspath(__file__))
_G = json.load(open(os.path.join(HERE, "golden", "encode_golden.json")))
MODELS, CASES = _G["models"], _G["cases"]
_cache = {}


def text_bytes(name):
    if name not in _cache:
        d = _G["texts"][name]
        if "file" in d:
            data = open(os.path.join(HERE, "golden", d["file"]), "rb").read()
        elif "hex" in d:
            data = bytes.fromhex(d["hex"])
        elif "random_seed" in d:
            data = random_corpus(d["random_seed"])
        else:
            with tempfile.TemporaryDirectory() as t:
                data = open(generated_corpus(os.path.join(t, "g.txt"), *d["gen"]), "rb").read()
        _cache[name] = data
    return _cache[name]


def case_id(c):
    return c["model"] + "-" + c["text"]


def strip_delims(data: bytes) -> bytes:
    return data.translate(None, b"\t\n\r ")
