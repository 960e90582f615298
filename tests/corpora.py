"""Deterministic test corpora (integer-only PRNG, so the bytes never depend on the Python version)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "shredword-trainer_b200", "build", "gen_corpus")
M64 = (1 << 64) - 1


class SM:
    """splitmix64 stream"""

    def __init__(self, seed):
        self.s = seed & M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    def below(self, n):
        return self.next() % n

    def pick(self, seq):
        return seq[self.below(len(seq))]


ALPHABETS = [b"abcde", b"abcdefghijklmnopqrstuvwxyz", bytes(range(33, 127)), bytes(c for c in range(1, 256) if c not in b"\t\r\n "), b"ab"]
SEPS = [b" ", b" ", b" ", b"\n", b"\t", b"\r\n", b"  ", b" \n "]


def random_corpus(seed: int) -> bytes:
    """Small corpus with a skewed word distribution, mixed delimiters, optional missing final newline."""
    r = SM(seed)
    alpha = r.pick(ALPHABETS)
    n_tokens = r.pick([0, 1, 5, 50, 500, 5000, 20000])
    n_types = r.pick([3, 20, 200, 2000])
    types = []
    for _ in range(n_types):
        L = r.pick([1, 1, 2, 3, 4, 5, 8, 12, 30])
        types.append(bytes(r.pick(alpha) for _ in range(L)))
    uniform = r.below(5) == 0
    out = bytearray()
    for _ in range(n_tokens):
        if uniform:
            w = types[r.below(n_types)]
        else:  # ~1/x: pick an octave, then uniformly inside it
            k = r.below(max(n_types.bit_length(), 1))
            w = types[min((1 << k) - 1 + r.below(1 << k), n_types - 1)]
        out += w
        out += r.pick(SEPS)
    if r.below(3) == 0:
        out = bytearray(bytes(out).rstrip())
    return bytes(out)


def random_config(seed: int):
    r = SM(seed ^ 0xABCDEF)
    vs = r.pick([100, 256, 257, 300, 400, 1000, 3000])
    unk = r.pick([0, 0, -1, -1, 65, 97, 300, -5, 1000000])
    cov = r.pick([0.995, 0.9995, 0.9, 0.5, 0.0, 1.0, 0.99])
    mf = r.pick([0, 1, 2, 2, 3, 10, 40])
    return vs, unk, cov, mf


def generated_corpus(path, n_bytes, seed, w, mode="zipf"):
    """Synthetic Zipfian / multilingual corpus from tools/gen_corpus.c."""
    subprocess.run([GEN, path, str(n_bytes), str(seed), str(w), mode], check=True, stdout=subprocess.DEVNULL)
    return path


# name -> (generator args, [(vocab_size, unk_id, coverage, min_pair_freq), ...])
GENERATED_CASES = {
    "zipf2m": ((2_000_000, 7, 16, "zipf"), [(1000, 0, 0.9995, 5), (2000, -1, 0.9995, 1), (5000, 0, 0.995, 40), (3000, 101, 0.995, 2)]),
    "multi600k": ((600_000, 7, 14, "multi"), [(2000, 0, 0.995, 2), (1500, -5, 0.9, 1), (3000, 65, 0.5, 10), (1500, 300, 0.9, 2)]),
    "multi1m5": ((1_500_000, 7, 16, "multi"), [(6000, -1, 0.9995, 2)]),
    "zipf8m": ((8_000_000, 7, 18, "zipf"), [(4000, 0, 0.995, 50)]),
}
N_RANDOM_CASES = 60
