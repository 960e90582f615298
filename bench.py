#!/usr/bin/env python
"""bench.py -- BPE train() throughput on B200 (BASELINE.json: "BPE train() wall-s & merges/s, 32k vocab ...").

A "step" is one full pass of the hot path over the synthetic corpus: bpe_load_corpus (host buffer -> HBM -> unique-word
table) + bpe_train (pair count + merge loop) + reading the merge list back.
  value  merges/s of bpe_train() alone, corpus already resident in HBM, timed with CUDA events on the library's stream
  e2e    merges/s through the C ABI with HOST buffers: pinned corpus bytes -> load -> train -> merge list on the host
  roofline      the dominant kernel (k_merge, one cooperative launch per merge): algorithmic bytes / CUDA-event duration
  cpu_baseline  the unmodified reference (oracle/_ref, pinned with the zero-fill malloc shim) on a bounded sample

python bench.py --gpus N --steps K --warmup W            (N>1 under torch.distributed.run: one rank per GPU)
python bench.py --impl reference ...                      (times the reference's own CPU implementation)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "shredword-trainer_b200")
sys.path.insert(0, PKG)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (bytes, seed, w, mode, vocab_size, unk_id, coverage, min_pair_freq)
    "config0_100MB": (100_000_000, 1, 20, "zipf", 8192, 0, 0.995, 2000),
    "config1_1GB": (1_000_000_000, 1, 22, "zipf", 32000, 0, 0.995, 2000),
    "config2_10GB": (10_000_000_000, 1, 24, "zipf", 32000, 0, 0.995, 2000),
    "config3_multi": (10_000_000_000, 1, 24, "multi", 65536, 0, 0.995, 2),
    "config4_50GB": (50_000_000_000, 1, 25, "zipf", 131072, 0, 0.995, 2000),
    "tiny": (4_000_000, 1, 16, "zipf", 2000, 0, 0.995, 20),
}
REF_SAMPLE_BYTES = 16 << 20   # bounded sample for the CPU reference: first 16 MiB of the corpus ...
REF_SAMPLE_MERGES = 40        # ... and this many merges (the reference needs ~47 min just to load 1 GB)


def scratch_dir():
    for d in ("/dev/shm", "/tmp"):
        if os.path.isdir(d) and os.access(d, os.W_OK):
            p = os.path.join(d, "shred_bench")
            os.makedirs(p, exist_ok=True)
            return p
    return ROOT


def make_corpus(workload, rank=0):
    nbytes, seed, w, mode = WORKLOADS[workload][:4]
    path = os.path.join(scratch_dir(), f"{workload}_s{seed}_w{w}_{mode}.txt")
    if not os.path.exists(path):
        tmp = path + f".tmp{rank}"
        subprocess.run([os.path.join(PKG, "build", "gen_corpus"), tmp, str(nbytes), str(seed), str(w), mode, "16"], check=True, stdout=subprocess.DEVNULL)
        os.replace(tmp, path)
    return path


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                 stdout=subprocess.PIPE, text=True)
        except OSError:
            return
        self.proc = p
        for line in p.stdout:
            if self.stop_flag:
                break
            self.samples.append([x.strip() for x in line.split(",")])
        p.kill()

    def finish(self):
        self.stop_flag = True
        if hasattr(self, "proc"):
            self.proc.kill()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def reference_sample(corpus_path, cfg, steps=1, warmup=0):
    """The unmodified reference (oracle/_ref) on a bounded sample of the workload; falls back to the C oracle port."""
    from oracle_lib import REF_HARNESS, REF_SO, ZMALLOC, have_reference
    vocab, unk, cov, mf = cfg
    d = scratch_dir()
    sample = os.path.join(d, os.path.basename(corpus_path) + ".sample16M")
    if not os.path.exists(sample):
        with open(corpus_path, "rb") as f:
            data = f.read(REF_SAMPLE_BYTES)
        data = data[: data.rfind(b"\n") + 1] if b"\n" in data else data
        open(sample, "wb").write(data)
    nbytes = os.path.getsize(sample)
    runs = []
    kind = "reference" if have_reference() else "port"
    for it in range(warmup + steps):
        if kind == "reference":
            js = os.path.join(d, "ref_sample.json")
            env = dict(os.environ, LD_PRELOAD=ZMALLOC)
            subprocess.run([REF_HARNESS, REF_SO, sample, str(vocab), str(unk), repr(float(cov)), str(mf), "--max-merges", str(REF_SAMPLE_MERGES), "--json", js],
                           env=env, check=True, stdout=subprocess.DEVNULL)
            info = json.load(open(js))
            load_s, train_s, merges = info["load_s"], info["train_s"], info["merges"]
        else:
            from oracle_lib import Oracle
            o = Oracle(vocab, unk, cov, mf)
            t0 = time.perf_counter(); o.load_corpus(sample); t1 = time.perf_counter()
            o.init(); merges = 0
            while merges < REF_SAMPLE_MERGES and o.merge_batch(1) > 0:
                merges += 1
            train_s, load_s = time.perf_counter() - t1, t1 - t0
            o.destroy()
        if it >= warmup:
            runs.append((load_s, train_s, merges))
    load_s = sum(r[0] for r in runs) / len(runs)
    train_s = sum(r[1] for r in runs) / len(runs)
    merges = runs[0][2]
    return {"value": merges / train_s if train_s > 0 else 0.0, "unit": "merges/s", "cores": 1, "kind": kind,
            "sample": f"first {nbytes} bytes of the corpus, bpe_init + {merges} merges (train part only; load of the sample took {load_s:.2f} s); "
                      f"the reference is single-threaded",
            "load_s": load_s, "train_s": train_s, "merges": merges, "sample_bytes": nbytes}


def encoder_leg(host, size, merges, peak_hint=None, with_cpu=True, reps=3):
    """SURVEY 8f rank 2, informational: encode the bench corpus with the model just trained (bpe_b200_encode through the C
    ABI, pinned host text in, pinned host ids + offsets out), and the CPU oracle on a bounded sample beside it."""
    import torch
    from shredword import BPEEncoder
    enc = BPEEncoder(merges=merges)
    n_words, n_ids = enc.encode_raw(host.data_ptr(), size)   # warm-up 1; sizes of the result
    ids = torch.empty(max(n_ids, 1), dtype=torch.int32, pin_memory=True)
    off = torch.empty(n_words + 1, dtype=torch.int64, pin_memory=True)
    enc.fetch_raw(ids.data_ptr(), off.data_ptr())
    for _ in range(2):                                         # warm-ups 2, 3
        enc.encode_raw(host.data_ptr(), size)
        enc.fetch_raw(ids.data_ptr(), off.data_ptr())
    walls, sts = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc.encode_raw(host.data_ptr(), size)
        enc.fetch_raw(ids.data_ptr(), off.data_ptr())
        walls.append(time.perf_counter() - t0)
        sts.append(enc.stats())
    wall = sum(walls) / reps
    st = {k: sum(x[k] for x in sts) / reps for k in sts[0]}
    # streamed entry point: same buffers, H2D | encode | D2H overlapped piece by piece
    swalls = []
    for i in range(2 + reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sw_, si_ = enc.encode_to_host_raw(host.data_ptr(), size, ids.data_ptr(), ids.numel(), off.data_ptr(), off.numel())
        if i >= 2:
            swalls.append(time.perf_counter() - t0)
    assert (sw_, si_) == (n_words, n_ids)
    swall = sum(swalls) / reps
    sst = enc.stats()
    iwalls = []                                      # ids only (what the reference's encode() returns): no offsets come back
    for i in range(1 + reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc.encode_to_host_raw(host.data_ptr(), size, ids.data_ptr(), ids.numel(), None, 0)
        if i >= 1:
            iwalls.append(time.perf_counter() - t0)
    iwall = sum(iwalls) / reps
    # occurrence phase (count starts + lookup + scans + expand): text read, ids and offsets written
    expand_bytes = size + 4.0 * n_ids + 8.0 * (n_words + 1)
    out = {"what": "bpe_b200_encode + bpe_b200_encode_fetch on the whole bench corpus with the merges of the last timed step; 3 warm-ups, mean of %d runs" % reps,
           "text_bytes": size, "n_words": n_words, "n_unique_words": int(st["n_unique_words"]), "n_ids": n_ids, "ids_per_word": n_ids / max(n_words, 1),
           "e2e_words_per_s": n_words / swall, "e2e_text_gbs": size / swall / 1e9, "e2e_s": swall, "e2e_s_runs": [round(x, 5) for x in swalls],
           "e2e_what": "bpe_b200_encode_to_host (streamed: pieces of 64 MB, H2D | encode | D2H overlapped), pinned host text in, pinned host ids + offsets out",
           "e2e_streamed_device_ms_sum_of_pieces": sst["device_ms"], "e2e_streamed_launches": int(sst["kernel_launches"]),
           "e2e_ids_only_s": iwall, "e2e_ids_only_words_per_s": n_words / iwall, "e2e_ids_only_text_gbs": size / iwall / 1e9, "e2e_ids_only_d2h_bytes": 4 * n_ids,
           "unstreamed_e2e_words_per_s": n_words / wall, "unstreamed_e2e_text_gbs": size / wall / 1e9, "unstreamed_e2e_s": wall,
           "h2d_bytes": size, "d2h_bytes": 4 * n_ids + 8 * (n_words + 1),
           "device_words_per_s": n_words / (st["device_ms"] * 1e-3), "device_text_gbs": size / (st["device_ms"] * 1e-3) / 1e9,
           "device_ms": st["device_ms"], "device_ms_runs": [round(x["device_ms"], 3) for x in sts], "unstreamed_e2e_s_runs": [round(x, 5) for x in walls], "tokenize_ms": st["tokenize_ms"], "words_ms": st["words_ms"], "expand_ms": st["expand_ms"],
           "h2d_ms": st["h2d_ms"], "d2h_ms": st["d2h_ms"], "kernel_launches": int(st["kernel_launches"]),
           "expand_algorithmic_bytes": expand_bytes, "expand_algorithmic_gbs": expand_bytes / (st["expand_ms"] * 1e-3) / 1e9 if st["expand_ms"] else None}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import EncodeOracle
        sample = bytes(host[:min(size, REF_SAMPLE_BYTES)].numpy())
        o = EncodeOracle(merges)
        t0 = time.perf_counter()
        oi, oo = o.encode_bytes(sample)
        cpu_s = time.perf_counter() - t0
        o.destroy()
        import hashlib
        sw, si = enc.encode_raw(host.data_ptr(), len(sample))
        enc.fetch_raw(ids.data_ptr(), off.data_ptr())
        same = hashlib.md5(oi).digest() == hashlib.md5(ids[:si].numpy().tobytes()).digest() and hashlib.md5(oo).digest() == hashlib.md5(off[:sw + 1].numpy().tobytes()).digest()
        out["cpu_baseline"] = {"value": sw / cpu_s, "unit": "words/s", "cores": 1, "kind": "port",
                               "sample": "first %d bytes of the corpus through oracle/bpe_encode_oracle.c (C restatement of the reference's Python _encode_chunk, "
                                         "with a distinct-word cache the Python reference does not have)" % len(sample),
                               "bit_exact_vs_oracle_on_sample": bool(same)}
    enc.destroy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SHRED_BENCH_WORKLOAD", "config1_1GB"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--replicas", action="store_true", help="N>1: independent replicas per GPU instead of one sharded job")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nbytes, seed, w, mode, vocab, unk, cov, mf = WORKLOADS[args.workload]
    config = {"workload": f"{args.workload}: synthetic {mode} corpus {nbytes} bytes (gen_corpus seed={seed} w={w}), vocab_size={vocab} unk_id={unk} "
                          f"character_coverage={cov} min_pair_freq={mf}",
              "l2": "inputs_exceed_l2" if nbytes > 300_000_000 else "inputs_fit_l2_small_workload",
              "parallelism": "single GPU" if world == 1 else (f"{world} independent replicas (one per GPU, no data-path collective)" if args.replicas else
                              f"one job, unique-word table sharded over {world} GPUs (contiguous word ranges); per-merge delta exchange inside the merge kernel "
                              f"over NVLink peer memory (CUDA IPC); heap and pair table replicated")}

    if args.impl == "reference":
        if rank != 0:
            return
        corpus = make_corpus(args.workload)
        cb = reference_sample(corpus, (vocab, unk, cov, mf), steps=max(args.steps, 1), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "bpe_train_merges_per_s", "value": cb["value"], "unit": "merges/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * (cb["train_s"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["SHRED_QUIET"] = "1"
    os.environ["SHRED_TIMING"] = os.environ.get("SHRED_TIMING", "64")  # CUDA events around every 64th merge kernel (a timed merge cannot overlap its rewrite phase with the host)
    os.environ["SHRED_DEVICE"] = str(local_rank)
    sharded = world > 1 and not args.replicas
    if sharded:
        import tempfile
        box = [tempfile.mkdtemp(prefix="shred_rdv_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        os.environ.update(SHRED_RANK=str(rank), SHRED_WORLD=str(world), SHRED_RDV=box[0])
    import __graft_entry__ as ge
    if rank == 0:
        ge._load_build().build()
    if world > 1:
        dist.barrier()
    from shredword import BPETrainer

    corpus = make_corpus(args.workload, rank) if rank == 0 else None
    if world > 1:
        dist.barrier()
        corpus = make_corpus(args.workload, rank)
    # host buffer: pinned memory holding the corpus bytes (the "HOST buffers" of the end-to-end leg)
    size = os.path.getsize(corpus)
    host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
    with open(corpus, "rb") as f:
        f.readinto(host.numpy())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        t = BPETrainer(vocab, unk, cov, mf)
        t0 = time.perf_counter()
        t.load_bytes(host)
        t1 = time.perf_counter()
        n = t.train()
        merges = t.merges()  # merge list read on the host (Trainer.merge_ops mirror)
        t2 = time.perf_counter()
        st = t.stats()
        t.destroy()
        return n, t1 - t0, t2 - t1, st, merges

    import contextlib
    import io
    quiet = contextlib.redirect_stdout(io.StringIO())
    with quiet:
        for _ in range(args.warmup):
            one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    steps = []
    with quiet:
        for _ in range(args.steps):
            steps.append(one_step())
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.finish()

    # informational: the same load through bpe_load_corpus(path) (file in the page cache -> mmap -> HBM), outside the timed region
    load_file_s = None
    if world == 1:
        with quiet:
            tf = BPETrainer(vocab, unk, cov, mf)
            t0 = time.perf_counter(); tf.load_corpus(corpus); load_file_s = time.perf_counter() - t0
            tf.destroy()
    # informational (N>1, sharded default): the same GPUs as N independent replicas -- one full trainer per GPU, no exchange
    replicas = None
    if sharded:
        saved = {k: os.environ.pop(k) for k in ("SHRED_RANK", "SHRED_WORLD")}
        barrier()
        with quiet:
            rstep = one_step()
        barrier()
        os.environ.update(saved)
        rt = torch.tensor([rstep[3]["train_device_ms"], rstep[1] + rstep[2]], dtype=torch.float64, device="cuda")
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        replicas = {"what": f"{world} independent trainers (one per GPU) on the same workload, 1 step, max over ranks",
                    "value": world * rstep[0] / (rt[0].item() * 1e-3), "e2e": world * rstep[0] / rt[1].item(), "unit": "merges/s", "scaling": "weak"}
    encoder = None
    if world == 1:
        try:
            encoder = encoder_leg(host, size, steps[-1][4], peak_hint=None, with_cpu=not args.no_cpu_baseline)
        except Exception as e:  # informational leg: never takes the headline measurement down
            encoder = {"error": str(e)}
    merges = steps[0][0]
    train_dev_ms = sum(s[3]["train_device_ms"] for s in steps)
    e2e_s = sum(s[1] + s[2] for s in steps)
    # max over ranks
    if world > 1:
        tt = torch.tensor([train_dev_ms, e2e_s, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        train_dev_ms, e2e_s, wall = tt.tolist()
    total_merges = merges * args.steps * (1 if sharded or world == 1 else world)
    st = steps[-1][3]
    all_ms = sum(s[3]["scan_device_ms"] for s in steps)
    all_bytes = sum(s[3]["scan_bytes"] for s in steps)
    all_touched = sum(s[3]["scan_bytes_touched"] for s in steps)
    all_n = sum(s[3]["scan_launches"] for s in steps)
    scan_ms = sum(s[3]["dense_device_ms"] for s in steps)
    scan_bytes = sum(s[3]["dense_bytes"] for s in steps)
    scan_n = sum(s[3]["dense_launches"] for s in steps)
    dense_phase_ms = sum(s[3]["dense_phase_ms"] for s in steps)
    all_phase_ms = sum(s[3]["scan_phase_ms"] for s in steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_note = None, None   # DRAM bytes of one launch from the committed ncu capture of this workload (profiles/)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01", "ncu_traffic_k_merge_config1_run61.json")))
        if tr["workload"] == args.workload:
            traffic = tr["representative"]["traffic_bytes"]
            traffic_note = {"source": "profiles/r01/ncu_traffic_k_merge_config1_run61.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, merge %d: %s)" % (tr["representative"]["merge"], tr["representative"]["why"]),
                            "samples": [{"merge": x["merge"], "algorithmic_bytes": x["algorithmic_bytes"], "touched_bytes": x["touched_bytes"],
                                         "dram_bytes": x["dram_read_bytes"] + x["dram_write_bytes"]} for x in tr["samples"]]}
    except Exception:
        pass
    line = {
        "metric": "bpe_train_merges_per_s", "value": total_merges / (train_dev_ms * 1e-3), "unit": "merges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": total_merges / e2e_s, "unit": "merges/s", "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]) + 8 * merges,
                "load_s_per_step": sum(s[1] for s in steps) / args.steps, "train_s_per_step": sum(s[2] for s in steps) / args.steps},
        "gpu_launches": int(sum(s[3]["kernel_launches"] for s in steps)),
        "roofline": {"bound": "hbm", "kernel": "k_merge<4,false> (one cooperative launch per merge: scan | barrier | fold + publish | rewrite)",
                     "achieved": all_bytes / (all_ms * 1e-3) / 1e9 if all_ms else 0.0, "peak": peak, "unit": "GB/s",
                     "frac": all_bytes / (all_ms * 1e-3) / 1e9 / peak if all_ms and peak else None, "traffic": traffic, "traffic_detail": traffic_note,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "launches_timed": int(all_n), "avg_launch_us": 1e3 * all_ms / all_n if all_n else None,
                     "bytes_per_launch": all_bytes / all_n if all_n else None,
                     "note": "achieved = ALGORITHMIC bytes (4 B x symbol slots, SURVEY 8d B_merge) / CUDA-event duration of the whole kernel, averaged over every 64th "
                             "launch of the timed region. The kernel reads fewer bytes than that: a tile occurrence index skips tiles that cannot hold the pair "
                             "(touched_gbs = bytes actually scanned / duration), and its duration is mostly a latency chain (delta emission, grid barrier, "
                             "pair-table fold, host flag, rewrite), see DESIGN.md section 5. ncu --set full (profiles/): a dense launch reads 31.7 MB of DRAM for "
                             "31.3 MB of algorithmic bytes (no re-reads); the stand-alone scan kernel of run 1 streamed at 4.76 TB/s = 73 % of the measured peak.",
                     "touched_gbs": all_touched / (all_ms * 1e-3) / 1e9 if all_ms else None,
                     "tiles_scanned_frac": st["cand_tiles"] / st["tiles_total"] if st["tiles_total"] else None,
                     "dense_launches": {"what": "timed launches that scanned >= 90 % of the tiles (early, occurrence-heavy merges)", "n": int(scan_n),
                                        "avg_launch_us": 1e3 * scan_ms / scan_n if scan_n else None,
                                        "achieved": scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms else None,
                                        "scan_phase_gbs": scan_bytes / (dense_phase_ms * 1e-3) / 1e9 if dense_phase_ms else None,
                                        "scan_phase_frac": scan_bytes / (dense_phase_ms * 1e-3) / 1e9 / peak if dense_phase_ms and peak else None},
                     "scan_phase_avg_us_all_launches": 1e3 * all_phase_ms / all_n if all_n else None},
        "detail": {"merges_per_step": merges, "n_words": int(st["n_words"]), "n_symbols_initial": int(st["n_symbols_initial"]), "n_symbols_final": int(st["n_symbols_live"]),
                   "occurrences": int(st["occurrences"]), "pair_entries": int(st["pair_entries"]), "heap_pushes": int(st["heap_pushes"]), "heap_pops": int(st["heap_pops"]),
                   "compactions": int(st["compactions"]), "train_device_ms_per_step": train_dev_ms / args.steps, "host_heap_ms": st["host_heap_ms"], "wait_ms": st["wait_ms"], "launch_ms": st["launch_ms"],
                   "ingest_device_ms": st["ingest_device_ms"], "ingest_gbs": st["ingest_bytes"] / (st["ingest_device_ms"] * 1e-3) / 1e9 if st["ingest_device_ms"] else None,
                   "count_device_ms": st["count_device_ms"], "count_gbs": st["count_bytes"] / (st["count_device_ms"] * 1e-3) / 1e9 if st["count_device_ms"] else None,
                   "h2d_ms": st["h2d_ms"], "load_corpus_from_file_s": load_file_s, "load_s_steps": [round(x[1], 4) for x in steps], "train_s_steps": [round(x[2], 4) for x in steps], "merges_md5": __import__("hashlib").md5(b"".join(__import__("struct").pack("<3i", *m) for m in steps[-1][4])).hexdigest(),
                   "device": __import__("shredword").cbase.lib.bpe_b200_device_name().decode()},
    }
    if replicas:
        line["detail"]["replicas"] = replicas
    if encoder:
        if "expand_algorithmic_gbs" in encoder and peak:
            encoder["expand_frac_of_hbm_peak"] = encoder["expand_algorithmic_gbs"] / peak
        line["detail"]["encoder"] = encoder
    try:
        big = json.load(open(os.path.join(ROOT, "tests", "golden", "big.json")))["cases"].get(args.workload)
        if big:
            line["detail"]["golden_merges_md5"] = big["merges_md5"]
            line["detail"]["bit_exact_vs_golden"] = big["merges_md5"] == line["detail"]["merges_md5"] and big["merges"] == merges
    except Exception:
        pass
    if sharded and rank == 0:
        import shutil
        shutil.rmtree(os.environ["SHRED_RDV"], ignore_errors=True)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = reference_sample(corpus, (vocab, unk, cov, mf))
            except Exception as e:  # the checker must never take the measurement down
                line["cpu_baseline"] = {"value": None, "unit": "merges/s", "cores": 1, "kind": "unavailable", "sample": str(e)}
            try:  # the one full-size run of the unmodified reference on this workload, recorded beside its golden vector
                full = json.load(open(os.path.join(ROOT, "tests", "golden", "big.json")))["cases"][args.workload].get("reference_full_run")
                if full:
                    line["cpu_baseline"]["full_workload_recorded"] = dict(full, merges_per_s=full["merges"] / full["train_s"], note="not timed in this run")
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
