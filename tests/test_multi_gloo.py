"""CPU, world_size 2 and 3 over gloo: the sharded N>1 path.  Every rank owns a contiguous range of the unique words,
exchanges its per-merge (key, delta, sequence) aggregates with the others and replays the SAME heap; all ranks must end
with the reference's merge list and vocab file.  The host control code is the product's (trainer_core.cpp); the device
engine is the CPU stand-in (tests/hostsim), whose exchange goes through torch.distributed instead of NVLink peer memory."""
import json
import os
import socket
import subprocess
import sys

import pytest

from cases import GOLDEN

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return str(p)


@pytest.mark.parametrize("world,case_name", [(2, "kat_py"), (2, "kat_cpp"), (3, "kat_cpp_500"), (2, "rnd022"), (2, "rnd054"), (3, "rnd037")])
def test_sharded_training_matches_reference(world, case_name, native, tmp_path):
    case = [c for c in GOLDEN if c["name"] == case_name][0]
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py"), str(r), str(world), port, case_name, str(tmp_path / f"out{r}.json")],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(world)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    res = [json.load(open(tmp_path / f"out{r}.json")) for r in range(world)]
    for r in res:
        assert r["merges"] == case["merges"] and r["n_words"] == case["n_words"]
        assert r["merges_md5"] == case["merges_md5"] and r["model_ok"]
        if case["vocab_md5"] is not None:
            assert r["vocab_md5"] == case["vocab_md5"]
    assert len({r["exchanges"] for r in res}) == 1 and res[0]["exchanges"] >= case["merges"]  # ranks stay in lockstep
