#!/usr/bin/env python
"""One-off GPU fuzz of the encoder (not collected by pytest): random models (trained by the CPU oracle on random corpora, or
hand-made with self pairs / repeated pairs) x random and adversarial texts, resident and streamed entry points with random
piece sizes, against the CPU encoder oracle.  usage: python tests/fuzz_gpu_encode.py <first_seed> <n_cases>"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "shredword-trainer_b200"))
os.environ.setdefault("SHRED_QUIET", "1")

from corpora import SM, random_config, random_corpus  # noqa: E402
from fuzz_gpu import spicy_corpus  # noqa: E402
from oracle_lib import EncodeOracle, Oracle  # noqa: E402
from shredword import BPEEncoder  # noqa: E402


def random_model(seed):
    r = SM(seed * 31 + 7)
    if r.below(4) == 0:  # hand-made: small alphabet, self pairs, repeated pairs
        alpha = [97, 98, 99][: 1 + r.below(3)]
        rows, toks = [], list(alpha)
        for m in range(r.pick([1, 3, 8, 20, 60])):
            a, b = r.pick(toks), r.pick(toks)
            rows.append((a, b, 256 + m))
            toks.append(256 + m)
        return rows
    vs, unk, cov, mf = random_config(seed)
    t = Oracle(max(vs, 300), unk, cov, min(mf, 3))
    t.load_bytes(spicy_corpus(seed) if r.below(2) else random_corpus(seed))
    t.train()
    m = t.merges()
    t.destroy()
    return m


def main():
    first, n = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + n):
        merges = random_model(seed)
        e, o = BPEEncoder(merges=merges), EncodeOracle(merges)
        r = SM(seed ^ 0x5151)
        for j in range(3):
            data = [random_corpus, spicy_corpus][r.below(2)](seed + 1000 * j)
            oi, oo = o.encode_bytes(data)
            ids, off = e.encode_bytes(data)
            ok = ids.tobytes() == oi and off.tobytes() == oo
            os.environ["SHRED_ENCODE_PIECE_BYTES"] = str(r.pick([1, 3, 16, 100, 1000, 4096, 50000]))
            if len(data) < 300000 or int(os.environ["SHRED_ENCODE_PIECE_BYTES"]) >= 100:
                sids, soff = e.encode_bytes_streamed(data)
                ok = ok and sids.tobytes() == oi and soff.tobytes() == oo
            ok = ok and e.decode_bytes(ids) == data.translate(None, b"\t\n\r ")
            if not ok:
                bad += 1
                print("MISMATCH seed", seed, "text", j, "merges", len(merges), "bytes", len(data), flush=True)
        e.destroy(); o.destroy()
    print(f"encoder fuzz: {n} models x 3 texts from seed {first}: {bad} mismatches")
    sys.exit(min(bad, 100))


if __name__ == "__main__":
    main()
