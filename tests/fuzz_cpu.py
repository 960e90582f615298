#!/usr/bin/env python
"""One-off CPU fuzz (not collected by pytest): random and adversarial corpora x random configs through the product's host
control code + the listsim engine (tests/hostsim/engine_listsim.cpp: the CUDA engine's data structures and csrc/layout.hpp
walked sequentially) against the CPU oracle.  usage: python tests/fuzz_cpu.py <first_seed> <n_cases> [naive|lists]"""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ.setdefault("SHRED_QUIET", "1")
os.environ["SHRED_HOSTSIM_ENGINE"] = sys.argv[3] if len(sys.argv) > 3 else "lists"

from corpora import SM, random_config  # noqa: E402
from fuzz_gpu import spicy_corpus  # noqa: E402  (imports the CUDA binding lazily: only its corpus generator is used here)
from hostsim_lib import build_hostsim  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from test_host_logic import _Cfg, _Trainer  # noqa: E402


def main():
    L = ctypes.CDLL(build_hostsim())
    L.create_trainer.argtypes, L.create_trainer.restype = [ctypes.POINTER(_Cfg)], ctypes.POINTER(_Trainer)
    L.bpe_b200_load_buffer.argtypes, L.bpe_b200_load_buffer.restype = [ctypes.POINTER(_Trainer), ctypes.c_char_p, ctypes.c_size_t], ctypes.c_int
    L.bpe_train.argtypes, L.bpe_train.restype = [ctypes.POINTER(_Trainer)], ctypes.c_int
    L.bpe_trainer_destroy.argtypes, L.bpe_trainer_destroy.restype = [ctypes.POINTER(_Trainer)], None
    L.bpe_b200_get_words.argtypes, L.bpe_b200_get_words.restype = [ctypes.POINTER(_Trainer), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64], ctypes.c_int
    first, n = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + n):
        data = spicy_corpus(seed)
        vs, unk, cov, mf = random_config(seed)
        if seed % 3 == 0:
            vs = SM(seed).pick([300, 600, 1500, 4000])
        o = Oracle(vs, unk, cov, mf); o.load_bytes(data); no = o.train()
        t = L.create_trainer(ctypes.byref(_Cfg(vs, unk, cov, mf)))
        L.bpe_b200_load_buffer(t, data, len(data))
        ng = L.bpe_train(t)
        m = [(t.contents.merge_ops[i].a, t.contents.merge_ops[i].b, 256 + i) for i in range(min(t.contents.num_merges, max(vs, 1)))]
        ok = ng == no and m == o.merges() and t.contents.n_words == o.num_words
        if ok and o.num_words:  # final symbols of every word
            nw = o.num_words
            off = (ctypes.c_uint64 * (nw + 1))()
            ids = (ctypes.c_int32 * (len(data) + 1))()
            L.bpe_b200_get_words(t, None, off, ids, len(data) + 1)
            ok = all(list(ids[off[i]:off[i + 1]]) == o.word_ids(i) for i in range(min(nw, 200)))
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, (vs, unk, cov, mf), len(data), ng, no, flush=True)
            open(f"/tmp/fuzz_cpu_bad_{seed}.txt", "wb").write(data)
        L.bpe_trainer_destroy(t); o.destroy()
    print(f"fuzz_cpu[{os.environ['SHRED_HOSTSIM_ENGINE']}]: {n} cases from seed {first}: {bad} mismatches")
    return bad


if __name__ == "__main__":
    sys.exit(min(main(), 100))
