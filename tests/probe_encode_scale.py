"""One-off probe (not collected by pytest): the encoder beyond 4 GiB of text -- resident and streamed results must be identical,
the first 64 MiB must equal the CPU oracle's, decode(encode(x)) must give the text back.
usage: python tests/probe_encode_scale.py <bytes>"""
import hashlib, os, sys, time
sys.path.insert(0, "shredword-trainer_b200"); sys.path.insert(0, "tests")
os.environ["SHRED_QUIET"] = "1"
import numpy as np, torch
from shredword import BPEEncoder, BPETrainer
from corpora import generated_corpus
from oracle_lib import EncodeOracle
p = generated_corpus("/tmp/zs.txt", int(sys.argv[1]), 3, 22, "zipf")
size = os.path.getsize(p)
host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
open(p, "rb").readinto(host.numpy())
small = host[: 200 << 20]
t = BPETrainer(16000, 0, 0.995, 2000); t.load_bytes(small); n = t.train(); m = t.merges(); t.destroy()
print("model:", n, "merges (trained on the first 200 MiB)")
e = BPEEncoder(merges=m)
t0 = time.perf_counter(); nw, ni = e.encode_raw(host.data_ptr(), size); st = e.stats()
ids = torch.empty(ni, dtype=torch.int32, pin_memory=True); off = torch.empty(nw + 1, dtype=torch.int64, pin_memory=True)
e.fetch_raw(ids.data_ptr(), off.data_ptr())
print("resident: %d words %d ids, device %.1f ms, wall %.1f ms" % (nw, ni, st["device_ms"], (time.perf_counter() - t0) * 1e3))
ids2 = torch.empty(ni, dtype=torch.int32, pin_memory=True); off2 = torch.empty(nw + 1, dtype=torch.int64, pin_memory=True)
for i in range(2):
    t0 = time.perf_counter()
    r = e.encode_to_host_raw(host.data_ptr(), size, ids2.data_ptr(), ni, off2.data_ptr(), nw + 1)
    print("streamed: wall %.1f ms" % ((time.perf_counter() - t0) * 1e3), r == (nw, ni))
print("streamed == resident:", bool(torch.equal(ids, ids2)) and bool(torch.equal(off, off2)), "last offset", int(off[-1]), "==", ni)
o = EncodeOracle(m)
sample = bytes(host[: 64 << 20].numpy())
oi, oo = o.encode_bytes(sample)
k = len(oo) // 8 - 1
# the sample may end inside a word: compare all words but the last
a = np.frombuffer(oo, dtype=np.uint64)
print("first 64 MiB == oracle:", hashlib.md5(oi[: int(a[k - 1]) * 4]).digest() == hashlib.md5(ids[: int(a[k - 1])].numpy().tobytes()).digest() and bool(np.array_equal(a[:k], off[:k].numpy().astype(np.uint64))))
tail = ids[int(off[nw - 2_000_000]):]
dec = e.decode_bytes(tail.numpy())
ref = bytes(host.numpy()[-(len(dec) + 4_000_000):]).translate(None, b"\t\n\r ")
print("decode(last 2 M words) is the tail of the text:", ref.endswith(dec), len(dec))
