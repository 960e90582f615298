"""Worker of tests/test_gpu_multi.py: one rank (= one GPU, one process) of a sharded multi-GPU training through the C ABI.
usage: _gpu_dist_worker.py <rank> <world> <rdv_dir> <out_json> <case> [<case> ...]"""
import json
import os
import struct
import sys

rank, world, rdv, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
os.environ.update(SHRED_RANK=str(rank), SHRED_WORLD=str(world), SHRED_RDV=rdv, SHRED_QUIET="1", SHRED_DEVICE=str(rank))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "shredword-trainer_b200"))

from cases import GOLDEN, corpus_bytes  # noqa: E402
from oracle_lib import md5  # noqa: E402
from shredword import BPETrainer  # noqa: E402

results = {}
for name in sys.argv[5:]:
    case = [c for c in GOLDEN if c["name"] == name][0]
    data = corpus_bytes(case)
    t = BPETrainer(*case["config"])
    t.load_bytes(data)
    slots_at_load = t.stats()["n_slots"]
    n = t.train()
    mb = b"".join(struct.pack("<3i", *m) for m in t.merges())
    model, vocab = out_path + ".model", out_path + ".vocab"
    t.save(model, vocab)
    st = t.stats()
    results[name] = {"merges": n, "n_words": t.num_words, "merges_md5": md5(mb), "vocab_md5": md5(open(vocab, "rb").read()),
                     "model_ok": open(model, "rb").read() == mb, "local_slots": slots_at_load, "occurrences": st["occurrences"]}
    t.destroy()
json.dump(results, open(out_path, "w"))
