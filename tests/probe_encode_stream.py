"""One-off probe (not collected by pytest): per-piece timeline of the streamed encoder.  usage: python tests/probe_encode_stream.py <bytes> <piece_mb>"""
import sys, os, time
sys.path.insert(0, "shredword-trainer_b200"); sys.path.insert(0, "tests")
os.environ["SHRED_QUIET"] = "1"
import torch
from shredword import BPEEncoder, BPETrainer
from corpora import generated_corpus
p = generated_corpus("/tmp/z.txt", int(sys.argv[1]), 1, 20, "zipf")
host = torch.empty(os.path.getsize(p), dtype=torch.uint8, pin_memory=True)
open(p, "rb").readinto(host.numpy())
t = BPETrainer(8192, 0, 0.995, 2000); t.load_bytes(host); t.train(); m = t.merges(); t.destroy()
e = BPEEncoder(merges=m)
nw, ni = e.encode_raw(host.data_ptr(), host.numel())
ids = torch.empty(ni, dtype=torch.int32, pin_memory=True); off = torch.empty(nw + 1, dtype=torch.int64, pin_memory=True)
os.environ["SHRED_ENCODE_PIECE_BYTES"] = str(int(sys.argv[2]) << 20)
for i in range(3):
    if i == 2:
        os.environ["SHRED_ENCODE_DEBUG"] = "1"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e.encode_to_host_raw(host.data_ptr(), host.numel(), ids.data_ptr(), ids.numel(), off.data_ptr(), off.numel())
    print("e2e ms", round((time.perf_counter() - t0) * 1e3, 2))
# copy engines: H2D alone, D2H alone, both at once
a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
ha = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); hb = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return round((time.perf_counter() - t0) * 1e3, 2)
def h2d():
    with torch.cuda.stream(s1): a.copy_(ha, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hb.copy_(b, non_blocking=True)
for _ in range(2):
    print("1 GiB H2D", timed(h2d), "D2H", timed(d2h), "both", timed(lambda: (h2d(), d2h())))
