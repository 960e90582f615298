#!/usr/bin/env python
"""One-off GPU fuzz (not collected by pytest): random corpora x random configs through the CUDA library against the CPU
oracle.  usage: python tests/fuzz_gpu.py <first_seed> <n_cases>   -> prints mismatches, exit code = their number."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "shredword-trainer_b200"))
os.environ.setdefault("SHRED_QUIET", "1")

from corpora import SM, random_config, random_corpus  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
BPETrainer = None  # bound in main(): importing the CUDA binding is deferred so that tests/fuzz_cpu.py can reuse spicy_corpus


def spicy_corpus(seed):
    """extra shapes the plain generator rarely makes: runs of one byte, very long words, words crossing 512-slot tiles"""
    r = SM(seed * 7919 + 13)
    kind = r.below(4)
    if kind == 0:
        return random_corpus(seed)
    out = bytearray()
    alpha = [b"ab", b"abc", b"aab", b"xyzxy"][r.below(4)]
    for _ in range(r.pick([20, 200, 2000])):
        if kind == 1:   # runs
            out += bytes([alpha[r.below(len(alpha))]]) * r.pick([1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 33, 64, 129])
        elif kind == 2:  # long words
            out += bytes(alpha[r.below(len(alpha))] for _ in range(r.pick([3, 40, 300, 700, 1500])))
        else:           # many short words (tile boundaries fall everywhere)
            out += bytes(alpha[r.below(len(alpha))] for _ in range(r.pick([1, 2, 3])))
        out += r.pick([b" ", b"\n", b"\t", b"  "])
    return bytes(out)


def main():
    from shredword import BPETrainer
    first, n = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + n):
        data = spicy_corpus(seed)
        vs, unk, cov, mf = random_config(seed)
        if seed % 3 == 0:
            vs = SM(seed).pick([300, 600, 1500, 4000])
        o = Oracle(vs, unk, cov, mf); o.load_bytes(data); no = o.train()
        t = BPETrainer(vs, unk, cov, mf); t.load_bytes(data)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            ng = t.train()
        ok = ng == no and t.merges() == o.merges() and t.num_words == o.num_words
        if ok:
            wo = [o.word_ids(i) for i in range(min(o.num_words, 50))]
            ok = [w[0] for w in t.words()[:50]] == wo
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, (vs, unk, cov, mf), len(data), ng, no, flush=True)
            open(f"/tmp/fuzz_bad_{seed}.txt", "wb").write(data)
        t.destroy(); o.destroy()
    print(f"fuzz: {n} cases from seed {first}: {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(min(main(), 100))
