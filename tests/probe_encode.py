"""One-off probe (not collected by pytest): encoder phase times on a generated corpus.  usage: python tests/probe_encode.py <bytes>"""
import sys, os, time
sys.path.insert(0, "shredword-trainer_b200"); sys.path.insert(0, "tests")
os.environ["SHRED_QUIET"] = "1"
import numpy as np, torch
from shredword import BPEEncoder, BPETrainer
from corpora import generated_corpus
p = generated_corpus("/tmp/z.txt", int(sys.argv[1]), 1, 20, "zipf")
host = torch.empty(os.path.getsize(p), dtype=torch.uint8, pin_memory=True)
open(p, "rb").readinto(host.numpy())
t = BPETrainer(8192, 0, 0.995, 2000); t.load_bytes(host); t.train(); m = t.merges(); t.destroy()
e = BPEEncoder(merges=m)
for i in range(4):
    t0 = time.perf_counter(); nw, ni = e.encode_raw(host.data_ptr(), host.numel()); dt = time.perf_counter() - t0
    s = e.stats()
    print(i, nw, ni, round(dt * 1e3, 2), {k: round(s[k], 3) for k in ("h2d_ms", "tokenize_ms", "words_ms", "expand_ms", "device_ms")})
