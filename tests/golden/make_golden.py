#!/usr/bin/env python
"""Regenerates tests/golden/golden.json and the known-answer corpora from the UNMODIFIED reference.

Needs /root/reference (dev container only).  The reference is compiled from its own sources by oracle/Makefile into
oracle/_ref/ and run under the zero-fill malloc shim (oracle/zmalloc.c) -- the pinned oracle of SURVEY.md Appendix B.
For every case the merge list and vocab file of the reference are hashed; tests compare the CPU oracle
(oracle/bpe_oracle.c), the host logic (tests/hostsim) and the CUDA library against these hashes.

  kat_py   the fixture of the reference's test/test_bpe.py:9-32 (20 sentences x 100)
  kat_cpp  the corpus of the reference's test/bpe_test.cpp:31-56
  rnd###   seeded random corpora x configs (tests/corpora.py)
  zipf*/multi*  synthetic corpora from tools/gen_corpus.c
"""
import ast
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
sys.path.insert(0, TESTS)
from corpora import GENERATED_CASES, N_RANDOM_CASES, generated_corpus, random_config, random_corpus  # noqa: E402
from oracle_lib import md5, run_reference  # noqa: E402

REF = "/root/reference"


def kat_py_text():
    src = open(os.path.join(REF, "test", "test_bpe.py")).read()
    m = re.search(r"sample_texts = (\[.*?\]) \* 100", src, re.S)
    sentences = ast.literal_eval(m.group(1))
    return ("\n".join(sentences * 100) + "\n").encode()


def kat_cpp_text():
    src = open(os.path.join(REF, "test", "bpe_test.cpp")).read()
    body = src[src.index("static int create_test_corpus"):src.index("// Test 1")]
    lines = re.findall(r'fprintf\(fp, "(.*?)\\n"\);', body)
    head, loop = lines[:-3], lines[-3:]
    return ("\n".join(head + loop * 20) + "\n").encode()


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)
    tmp = tempfile.mkdtemp()
    cases = []

    def run(name, corpus_bytes, cfg, keep_merges=False, corpus_desc=None):
        p = os.path.join(tmp, "c.txt")
        open(p, "wb").write(corpus_bytes)
        vs, unk, cov, mf = cfg
        merges, vocab, info = run_reference(p, vs, unk, cov, mf, tmp)
        T = 256 + info["num_merges"]
        rec = {"name": name, "corpus": corpus_desc, "corpus_md5": md5(corpus_bytes), "config": list(cfg), "merges": info["merges"],
               "n_words": info["n_words"], "merges_md5": md5(merges),
               # the reference's vocab file is only defined when freq[unk_id] is in range or unk_id < 0
               "vocab_md5": md5(vocab) if (unk < 0 or unk < T) else None}
        if keep_merges:
            import struct
            rec["merge_list"] = [list(struct.unpack_from("<3i", merges, 12 * i)) for i in range(len(merges) // 12)]
        cases.append(rec)
        print(name, cfg, info["merges"], flush=True)

    kp, kc = kat_py_text(), kat_cpp_text()
    open(os.path.join(HERE, "kat_py.txt"), "wb").write(kp)
    open(os.path.join(HERE, "kat_cpp.txt"), "wb").write(kc)
    run("kat_py", kp, (300, 0, 0.995, 2), True, {"file": "kat_py.txt"})
    run("kat_py_unk-1", kp, (300, -1, 0.9995, 2), True, {"file": "kat_py.txt"})
    run("kat_py_vocab50", kp, (50, 0, 0.995, 1000), True, {"file": "kat_py.txt"})
    run("kat_cpp", kc, (300, -1, 0.99, 2), True, {"file": "kat_cpp.txt"})
    run("kat_cpp_500", kc, (500, 0, 0.995, 1), True, {"file": "kat_cpp.txt"})
    for i in range(N_RANDOM_CASES):
        run("rnd%03d" % i, random_corpus(i), random_config(i), False, {"random_seed": i})
    for name, (gargs, cfgs) in GENERATED_CASES.items():
        p = generated_corpus(os.path.join(tmp, name + ".txt"), *gargs)
        data = open(p, "rb").read()
        for j, cfg in enumerate(cfgs):
            run("%s_%d" % (name, j), data, cfg, False, {"gen": list(gargs)})
    json.dump({"how": "oracle/_ref/libtrainer_ref.so (unmodified reference, g++ -O3) under LD_PRELOAD=zmalloc.so via oracle/ref_harness.cpp",
               "cases": cases}, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
