#!/usr/bin/env python
"""Profiling helper (not collected by pytest): load a bench corpus and train a handful of merges, so that ncu sees the ingest,
count and list-fill kernels without tens of thousands of merge launches around them.
usage: python tests/probe_small_train.py <workload> [vocab_size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "shredword-trainer_b200"))
os.environ.setdefault("SHRED_QUIET", "1")
import bench  # noqa: E402
from shredword import BPETrainer  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "config1_1GB"
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 300
_, _, _, _, _, unk, cov, mf = bench.WORKLOADS[workload]
path = bench.make_corpus(workload)
t = BPETrainer(vocab, unk, cov, mf)
t.load_corpus(path)
n = t.train()
st = t.stats()
print(f"{workload}: {n} merges, count {st['count_device_ms']:.3f} ms ({st['count_bytes'] / st['count_device_ms'] / 1e6:.0f} GB/s of 4S+12N), "
      f"fold+fill {st['fill_device_ms']:.3f} ms, ingest {st['ingest_device_ms']:.1f} ms")
t.destroy()
