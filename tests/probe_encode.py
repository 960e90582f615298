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
# streamed entry point at several piece sizes
nw, ni = e.encode_raw(host.data_ptr(), host.numel())
ids = torch.empty(ni, dtype=torch.int32, pin_memory=True); off = torch.empty(nw + 1, dtype=torch.int64, pin_memory=True)
for _ in range(2):
    t0 = time.perf_counter(); e.encode_raw(host.data_ptr(), host.numel()); e.fetch_raw(ids.data_ptr(), off.data_ptr()); dt = time.perf_counter() - t0
print("resident+fetch e2e ms", round(dt * 1e3, 2))
for mb in (16, 64, 256, 4096):
    os.environ["SHRED_ENCODE_PIECE_BYTES"] = str(mb << 20)
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e.encode_to_host_raw(host.data_ptr(), host.numel(), ids.data_ptr(), ids.numel(), off.data_ptr(), off.numel())
        dt = time.perf_counter() - t0
    s = e.stats()
    print("piece MB", mb, "e2e ms", round(dt * 1e3, 2), {k: round(s[k], 3) for k in ("tokenize_ms", "words_ms", "expand_ms", "device_ms")}, "uniq sum", s["n_unique_words"], "launches", s["kernel_launches"])
