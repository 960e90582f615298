"""GPU (-m gpu, ONE device is enough): the sharded multi-GPU data path with VIRTUAL ranks -- the W ranks of a job as W trainers
of this process on one GPU, one host thread each (SHRED_VIRTUAL_RANKS=W, engine_cuda.cu VirtualCluster).  Same device code, same
inbox protocol (delta exchange, sharded count, sharded token frequencies), same host control per rank as with one process per GPU;
the kernels in which ranks wait for one another run as ONE cooperative launch holding every rank's CTA group, which is the safe way
to run them on fewer GPUs than ranks.  Every rank must produce the reference's merge list and vocabulary.  (tests/test_gpu_multi.py
runs the same cases with one process per GPU where the box has 2/4/8 of them.)"""
import struct
import threading

import pytest

from cases import GOLDEN, corpus_bytes
from oracle_lib import md5

pytestmark = pytest.mark.gpu
CASES = ["kat_py", "kat_cpp", "kat_py_unk-1", "rnd005", "rnd022", "rnd054", "multi600k_0", "multi600k_1", "multi600k_3", "zipf2m_0", "zipf2m_1", "zipf2m_3", "multi1m5_0", "zipf8m_0"]


def _job(world, case, tmp_path, monkeypatch):
    from shredword import BPETrainer
    data = corpus_bytes(case)
    monkeypatch.setenv("SHRED_VIRTUAL_RANKS", str(world))
    trainers = [BPETrainer(*case["config"]) for _ in range(world)]   # rank = creation order
    monkeypatch.delenv("SHRED_VIRTUAL_RANKS")
    out, errs = [None] * world, []

    def run(r):
        try:
            t = trainers[r]
            t.load_bytes(data)
            slots = t.stats()["n_slots"]
            n = t.train()
            mb = b"".join(struct.pack("<3i", *m) for m in t.merges())
            model, vocab = str(tmp_path / f"m{r}"), str(tmp_path / f"v{r}")
            t.save(model, vocab)
            out[r] = {"merges": n, "n_words": t.num_words, "merges_md5": md5(mb), "vocab_md5": md5(open(vocab, "rb").read()), "model_ok": open(model, "rb").read() == mb,
                      "local_slots": slots, "occurrences": t.stats()["occurrences"]}
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for x in th:
        x.start()
    for x in th:
        x.join(timeout=300)
    alive = [x.is_alive() for x in th]
    for t in trainers:
        if not any(alive):
            t.destroy()
    assert not any(alive) and not errs, (alive, errs)
    return out


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("name", CASES)
def test_virtual_ranks_reproduce_the_reference(name, world, native, tmp_path, monkeypatch):
    monkeypatch.setenv("SHRED_QUIET", "1")
    case = [c for c in GOLDEN if c["name"] == name][0]
    res = _job(world, case, tmp_path, monkeypatch)
    for got in res:
        assert got["merges"] == case["merges"] and got["n_words"] == case["n_words"], got
        assert got["merges_md5"] == case["merges_md5"] and got["model_ok"], got
        if case["vocab_md5"] is not None:
            assert got["vocab_md5"] == case["vocab_md5"], got
    assert len({g["occurrences"] for g in res}) == 1                  # global occurrence counts agree
    if case["n_words"] > 1000:
        slots = [g["local_slots"] for g in res]
        assert max(slots) < 1.1 * (sum(slots) / len(slots)) + 64        # shards are balanced by symbol slots
