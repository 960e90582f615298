"""Golden cases (tests/golden/golden.json) -> corpus bytes + expected hashes."""
import json
import os
import tempfile

from corpora import generated_corpus, random_corpus

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))["cases"]
_cache = {}


def corpus_bytes(case):
    c = case["corpus"]
    key = json.dumps(c, sort_keys=True)
    if key not in _cache:
        if "file" in c:
            data = open(os.path.join(HERE, "golden", c["file"]), "rb").read()
        elif "random_seed" in c:
            data = random_corpus(c["random_seed"])
        else:
            with tempfile.TemporaryDirectory() as d:
                data = open(generated_corpus(os.path.join(d, "g.txt"), *c["gen"]), "rb").read()
        _cache[key] = data
    return _cache[key]


def case_ids(cases=GOLDEN):
    return [c["name"] for c in cases]


SMALL = [c for c in GOLDEN if not c["name"].startswith(("zipf8m", "multi1m5"))]
