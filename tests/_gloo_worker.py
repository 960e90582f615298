"""Worker of tests/test_multi_gloo.py: one rank of a world_size-N gloo job that trains through the product's host
control code (abi.cpp + trainer_core.cpp) on the sharded CPU stand-in engine.  Exchanges go through torch.distributed
(gloo, 127.0.0.1).  usage: _gloo_worker.py <rank> <world> <port> <case_name> <out_json>"""
import ctypes
import json
import os
import struct
import sys

rank, world, port, case_name, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
os.environ.update(SHRED_RANK=str(rank), SHRED_WORLD=str(world), SHRED_QUIET="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
import torch.distributed as dist

from cases import GOLDEN, corpus_bytes
from hostsim_lib import HS_SO
from oracle_lib import md5

dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")

CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64))
n_exchanges = 0


def allgather(send, nbytes, recv, cap, sizes):
    global n_exchanges
    n_exchanges += 1
    mine = torch.frombuffer(bytearray(ctypes.string_at(send, nbytes)), dtype=torch.uint8) if nbytes else torch.empty(0, dtype=torch.uint8)
    sz = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sz, torch.tensor([nbytes], dtype=torch.int64))
    m = max(int(s) for s in sz)
    pad = torch.zeros(max(m, 1), dtype=torch.uint8)
    pad[:nbytes] = mine
    got = [torch.zeros(max(m, 1), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(got, pad)
    at = 0
    for r in range(world):
        n = int(sz[r])
        if at + n > cap:
            return -1
        if n:
            ctypes.memmove(recv + at, got[r].numpy().ctypes.data, n)
        sizes[r] = n
        at += n
    return 0


cb = CB(allgather)
L = ctypes.CDLL(HS_SO)
L.hostsim_set_allgather(cb)


class Cfg(ctypes.Structure):
    _fields_ = [("vs", ctypes.c_size_t), ("unk", ctypes.c_int32), ("cov", ctypes.c_float), ("mf", ctypes.c_uint64)]


class Pair(ctypes.Structure):
    _fields_ = [("a", ctypes.c_int32), ("b", ctypes.c_int32)]


class Tr(ctypes.Structure):
    _fields_ = [("config", Cfg), ("heap", ctypes.c_void_p * 3), ("words", ctypes.c_void_p), ("word_counts", ctypes.POINTER(ctypes.c_uint64)),
                ("n_words", ctypes.c_size_t), ("bm", ctypes.c_void_p * 2), ("next_token", ctypes.c_size_t), ("num_merges", ctypes.c_size_t),
                ("merge_ops", ctypes.POINTER(Pair))]


L.create_trainer.argtypes, L.create_trainer.restype = [ctypes.POINTER(Cfg)], ctypes.POINTER(Tr)
L.bpe_b200_load_buffer.argtypes = [ctypes.POINTER(Tr), ctypes.c_char_p, ctypes.c_size_t]
L.bpe_train.argtypes, L.bpe_train.restype = [ctypes.POINTER(Tr)], ctypes.c_int
L.bpe_save.argtypes = [ctypes.POINTER(Tr), ctypes.c_char_p, ctypes.c_char_p]
L.bpe_trainer_destroy.argtypes = [ctypes.POINTER(Tr)]

case = [c for c in GOLDEN if c["name"] == case_name][0]
data = corpus_bytes(case)
vs, unk, cov, mf = case["config"]
t = L.create_trainer(ctypes.byref(Cfg(vs, unk, cov, mf)))
assert L.bpe_b200_load_buffer(t, data, len(data)) == 0
n = L.bpe_train(t)
M = min(t.contents.num_merges, max(vs, 1))
mb = b"".join(struct.pack("<3i", t.contents.merge_ops[i].a, t.contents.merge_ops[i].b, 256 + i) for i in range(M))
model, vocab = out_path + ".model", out_path + ".vocab"
L.bpe_save(t, os.fsencode(model), os.fsencode(vocab))
json.dump({"rank": rank, "merges": n, "n_words": t.contents.n_words, "merges_md5": md5(mb), "vocab_md5": md5(open(vocab, "rb").read()),
           "model_ok": open(model, "rb").read() == mb, "exchanges": n_exchanges}, open(out_path, "w"))
L.bpe_trainer_destroy(t)
dist.barrier()
dist.destroy_process_group()
