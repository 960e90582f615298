"""GPU (-m gpu, needs >= 2 devices; skipped on a 1-GPU box): the sharded multi-GPU trainer.  One process per GPU; every
rank owns a contiguous range of the unique words, the per-merge delta exchange runs inside the merge kernel over NVLink
peer memory (CUDA IPC), every rank replays the same heap.  All ranks must produce the reference's merge list and vocab."""
import json
import os
import subprocess
import sys
import tempfile

import pytest

from cases import GOLDEN

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["kat_py", "kat_cpp", "kat_py_unk-1", "rnd005", "rnd022", "rnd054", "multi600k_0", "multi600k_1", "multi600k_3", "zipf2m_0", "zipf2m_1", "zipf2m_3", "multi1m5_0", "zipf8m_0"]


def _run(world, native):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as rdv:
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_gpu_dist_worker.py"), str(r), str(world), rdv, os.path.join(rdv, f"out{r}.json")] + CASES,
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
        outs = [p.communicate(timeout=900)[0].decode() for p in procs]
        assert all(p.returncode == 0 for p in procs), "\n".join(o[-3000:] for o in outs)
        res = [json.load(open(os.path.join(rdv, f"out{r}.json"))) for r in range(world)]
    for name in CASES:
        case = [c for c in GOLDEN if c["name"] == name][0]
        for r in res:
            got = r[name]
            assert got["merges"] == case["merges"] and got["n_words"] == case["n_words"], (name, got)
            assert got["merges_md5"] == case["merges_md5"] and got["model_ok"], (name, got)
            if case["vocab_md5"] is not None:
                assert got["vocab_md5"] == case["vocab_md5"], (name, got)
        assert len({r[name]["occurrences"] for r in res}) == 1          # global occurrence counts agree
        if case["n_words"] > 1000:
            slots = [r[name]["local_slots"] for r in res]
            assert max(slots) < 1.1 * (sum(slots) / len(slots)) + 64     # shards are balanced by symbol slots (measured right after load)


def test_two_gpus(native):
    _run(2, native)


def test_four_gpus(native):
    _run(4, native)


def test_eight_gpus(native):
    _run(8, native)
