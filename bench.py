#!/usr/bin/env python
"""bench.py -- BPE train() throughput on B200 (BASELINE.json: "BPE train() wall-s & merges/s, 32k vocab ...").

A "step" is one full pass of the hot path over the synthetic corpus, through the reference's own call sequence
(shredword/trainer.py:12-29): bpe_load_corpus(path) -> bpe_train -> bpe_save(model, vocab).
Default workload: BASELINE.json's headline configuration (32k vocab / 10 GB, configs[2]) at every N.
  value  merges/s of bpe_train() alone, corpus already resident in HBM, timed with CUDA events on the library's stream
  e2e    merges/s through the drop-in C ABI with HOST data: corpus file in host memory (tmpfs / page cache) ->
         bpe_load_corpus(path) (pinned staging ring, H2D inside) -> bpe_train -> bpe_save (both files written)
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
ROOFLINE_NOTE = ("achieved = ALGORITHMIC bytes of the scan formulation (SURVEY 8d B_merge = 4 B x (live symbols + unique words) per merge) / CUDA-event "
                 "duration of the whole per-merge launch, averaged over every 64th launch of the timed region. It is an EFFECTIVE rate, far above the copy peak by "
                 "construction: the trainer keeps per-pair occurrence lists, so a merge reads only its pair's list and the few symbols around each entry "
                 "(touched_gbs = estimate of those bytes / duration; roofline.traffic = DRAM bytes of one launch measured by ncu) -- the bytes of the scan are never "
                 "read. What bounds a launch is a chain of ~8 dependent memory round trips plus launch latency, not bandwidth (DESIGN.md sections 4-5, "
                 "profiles/r02). The bandwidth-bound kernels of the path are the count pass and the list fill (detail.count_frac_of_hbm_peak, detail.list_fill_gbs).")
DENSE_NOTE = "timed launches whose occurrence list has >= 65536 entries (the first few hundred merges): the throughput-bound ones"
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


def reference_sample(corpus_path, cfg, steps=1, warmup=0, workload=None):
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
    full = None
    try:  # the one full-size run of the unmodified reference on this workload (dev container), recorded beside its golden vector
        full = json.load(open(os.path.join(ROOT, "tests", "golden", "big.json")))["cases"][workload].get("reference_full_run")
        if full:
            full = dict(full, merges_per_s=full["merges"] / full["train_s"], note="recorded once, not timed in this run")
    except Exception:
        pass
    return {"value": merges / train_s if train_s > 0 else 0.0, "unit": "merges/s", "cores": 1, "kind": kind, "same_work": False,
            "same_work_note": "the reference needs hours to load and train the full workload (its merges/s FALLS with corpus size: per-merge cost is two scans of all "
                              "symbols), so it is timed on a bounded sample; value / this number is therefore a LOWER BOUND on the like-for-like speed-up",
            "full_workload_recorded": full,
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
    ap.add_argument("--workload", default=os.environ.get("SHRED_BENCH_WORKLOAD", "config2_10GB"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip the informational load_buffer and encoder legs (they pin a corpus-sized host buffer)")
    ap.add_argument("--sharded", action="store_true", help="N>1: time ONE job whose unique-word table is sharded over the GPUs (default at N>1: one independent trainer per GPU; "
                    "the other mode is always run for one step and reported under detail)")
    ap.add_argument("--replicas", action="store_true", help="(default at N>1, kept for compatibility)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nbytes, seed, w, mode, vocab, unk, cov, mf = WORKLOADS[args.workload]
    config = {"workload": f"{args.workload}: synthetic {mode} corpus {nbytes} bytes (gen_corpus seed={seed} w={w}), vocab_size={vocab} unk_id={unk} "
                          f"character_coverage={cov} min_pair_freq={mf}",
              "l2": "inputs_exceed_l2" if nbytes > 300_000_000 else "inputs_fit_l2_small_workload",
              "parallelism": "single GPU" if world == 1 else (f"{world} independent trainers, one per GPU, each training the whole workload (no data-path collective; train() is a chain of "
                              f"dependent merges that sharding cannot shorten, DESIGN.md section 6); detail.sharded has the same GPUs as ONE sharded job" if not args.sharded else
                              f"one job, unique-word table sharded over {world} GPUs (contiguous word ranges); per-merge delta exchange inside the merge kernel "
                              f"over NVLink peer memory (CUDA IPC); heap and pair table replicated")}

    if args.impl == "reference":
        if rank != 0:
            return
        corpus = make_corpus(args.workload)
        cb = reference_sample(corpus, (vocab, unk, cov, mf), steps=max(min(args.steps, 3), 1), warmup=min(args.warmup, 1), workload=args.workload)
        line = {"impl": "reference", "metric": "bpe_train_merges_per_s", "value": cb["value"], "unit": "merges/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * (cb["train_s"]), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": config, "cpu_baseline": cb, "same_work": False,
                "e2e": {"value": cb["value"], "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import hashlib
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["SHRED_QUIET"] = "1"
    os.environ["SHRED_TIMING"] = os.environ.get("SHRED_TIMING", "64")  # CUDA events around every 64th merge kernel
    os.environ["SHRED_DEVICE"] = str(local_rank)
    sharded = world > 1 and args.sharded
    shard_env = {}
    if world > 1:
        import tempfile
        box = [tempfile.mkdtemp(prefix="shred_rdv_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        shard_env = dict(SHRED_RANK=str(rank), SHRED_WORLD=str(world), SHRED_RDV=box[0])  # read by the library when a trainer is created
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
    size = os.path.getsize(corpus)
    out_dir = os.path.join(scratch_dir(), f"out_rank{rank}")
    os.makedirs(out_dir, exist_ok=True)
    model_path, vocab_path = os.path.join(out_dir, "bpe.model"), os.path.join(out_dir, "bpe.vocab")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(buffer=None, as_shard=None):
        """The reference's call sequence (shredword/trainer.py:12-29) on the drop-in ABI; buffer != None: the load_buffer extension."""
        as_shard = sharded if as_shard is None else as_shard
        for k in ("SHRED_RANK", "SHRED_WORLD", "SHRED_RDV"):
            os.environ.pop(k, None)
        if as_shard:
            os.environ.update(shard_env)
        tc = time.perf_counter()
        t = BPETrainer(vocab, unk, cov, mf)
        t0 = time.perf_counter()
        if buffer is None:
            t.load_corpus(corpus)          # bpe_load_corpus(path): host file -> pinned staging ring -> HBM -> unique-word table
        else:
            t.load_bytes(buffer)
        t1 = time.perf_counter()
        n = t.train()                      # bpe_train
        t2 = time.perf_counter()
        t.save(model_path, vocab_path)     # bpe_save: token frequencies from the device, both files written
        t3 = time.perf_counter()
        st = t.stats()
        t4 = time.perf_counter()
        t.destroy()
        t5 = time.perf_counter()
        out_bytes = os.path.getsize(model_path) + os.path.getsize(vocab_path)
        return {"merges": n, "load_s": t1 - t0, "train_s": t2 - t1, "save_s": t3 - t2, "create_s": t0 - tc, "destroy_s": t5 - t4, "st": st, "out_bytes": out_bytes}

    import contextlib
    import io
    quiet = contextlib.redirect_stdout(io.StringIO())
    cold = None
    with quiet:
        for i in range(args.warmup):
            r = one_step()
            if i == 0:
                cold = r  # the first load/train of this process: allocator pool, pinned staging ring and module load included
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
    model_bytes = open(model_path, "rb").read()   # the file bpe_save wrote in the last timed step
    merges_md5 = hashlib.md5(model_bytes).hexdigest()
    vocab_md5 = hashlib.md5(open(vocab_path, "rb").read()).hexdigest()
    merges_list = [tuple(int.from_bytes(model_bytes[12 * i + 4 * j:12 * i + 4 * j + 4], "little", signed=True) for j in range(3)) for i in range(len(model_bytes) // 12)] if world == 1 else None

    # informational: the same step through the bpe_b200_load_buffer extension (pinned host buffer), outside the timed region
    host, side_buffer = None, None
    if world == 1 and not args.no_side_legs:
        host = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        with open(corpus, "rb") as f:
            f.readinto(host.numpy())
        with quiet:
            rb = one_step(buffer=host)
        side_buffer = {"what": "same step with bpe_b200_load_buffer (one cudaMemcpyAsync from a pinned host buffer) instead of bpe_load_corpus(path); 1 run",
                       "value": rb["merges"] / (rb["load_s"] + rb["train_s"] + rb["save_s"]), "unit": "merges/s", "load_s": rb["load_s"]}
    # informational (N>1): the other multi-GPU mode on the same GPUs, one step
    other = None
    if world > 1:
        barrier()
        with quiet:
            ostep = one_step(as_shard=not sharded)
        barrier()
        omd5 = hashlib.md5(open(model_path, "rb").read()).hexdigest()
        rt = torch.tensor([ostep["st"]["train_device_ms"], ostep["load_s"] + ostep["train_s"] + ostep["save_s"], ostep["load_s"], ostep["st"]["host_heap_ms"], ostep["st"]["wait_ms"]],
                          dtype=torch.float64, device="cuda")
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        mult = world if sharded else 1   # the other mode of a sharded main run is replicas: N trainers' merges
        other = {"what": (f"{world} independent trainers (one per GPU) on the same workload" if sharded else
                          f"ONE job sharded over the {world} GPUs (contiguous word ranges, in-kernel delta exchange over NVLink peer memory, replicated heap)") + ", 1 step, max over ranks",
                 "value": mult * ostep["merges"] / (rt[0].item() * 1e-3), "e2e": mult * ostep["merges"] / rt[1].item(), "unit": "merges/s", "scaling": "weak" if sharded else "strong",
                 "load_s": rt[2].item(), "host_heap_ms": rt[3].item(), "wait_ms": rt[4].item(), "merges_md5": omd5}
        os.environ.pop("SHRED_RANK", None); os.environ.pop("SHRED_WORLD", None)
    encoder = None
    if world == 1 and host is not None:
        try:
            encoder = encoder_leg(host, size, merges_list, peak_hint=None, with_cpu=not args.no_cpu_baseline)
        except Exception as e:  # informational leg: never takes the headline measurement down
            encoder = {"error": str(e)}
    merges = steps[0]["merges"]
    train_dev_ms = sum(s["st"]["train_device_ms"] for s in steps)
    e2e_s = sum(s["load_s"] + s["train_s"] + s["save_s"] for s in steps)
    phase = [sum(s[k] for s in steps) / args.steps for k in ("load_s", "train_s", "save_s")]
    # max over ranks
    if world > 1:
        tt = torch.tensor([train_dev_ms, e2e_s, wall] + phase, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        train_dev_ms, e2e_s, wall = tt.tolist()[:3]
        phase = tt.tolist()[3:]
    total_merges = merges * args.steps * (1 if sharded or world == 1 else world)
    st = steps[-1]["st"]
    S = lambda k: sum(s["st"][k] for s in steps)  # noqa: E731
    all_ms, all_bytes, all_touched, all_n = S("scan_device_ms"), S("scan_bytes"), S("scan_bytes_touched"), S("scan_launches")
    dense_ms, dense_bytes, dense_n, dense_phase_ms, all_phase_ms = S("dense_device_ms"), S("dense_bytes"), S("dense_launches"), S("dense_phase_ms"), S("scan_phase_ms")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_note = None, None   # DRAM bytes per launch from a committed ncu capture OF THIS WORKLOAD (profiles/), else null
    for cand in sorted(__import__("glob").glob(os.path.join(ROOT, "profiles", "r0*", "ncu_traffic_*.json")), reverse=True):
        try:
            tr = json.load(open(cand))
            if tr.get("workload") == args.workload:
                traffic = tr["representative"]["traffic_bytes"]
                traffic_note = {"source": os.path.relpath(cand, ROOT) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch: %s)" % tr["representative"].get("why", ""),
                                "samples": tr.get("samples")}
                break
        except Exception:
            pass
    line = {
        "metric": "bpe_train_merges_per_s", "value": total_merges / (train_dev_ms * 1e-3), "unit": "merges/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": config, "clocks": clocks,
        "e2e": {"value": total_merges / e2e_s, "unit": "merges/s", "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]) + int(steps[-1]["out_bytes"]),
                "what": "bpe_load_corpus(path) + bpe_train + bpe_save through the drop-in C ABI (the reference's call sequence, shredword/trainer.py:12-29); corpus file in host "
                        "memory (tmpfs), both output files written; max over ranks",
                "load_s_per_step": phase[0], "train_s_per_step": phase[1], "save_s_per_step": phase[2],
                "create_destroy_s_per_step": {"create_trainer": sum(x["create_s"] for x in steps) / len(steps), "bpe_trainer_destroy": sum(x["destroy_s"] for x in steps) / len(steps),
                                              "what": "rank 0's create_trainer / bpe_trainer_destroy around every step; not part of value or e2e.value, part of ms_per_step"},
                "cold_first_step": {"load_s": cold["load_s"], "train_s": cold["train_s"], "save_s": cold["save_s"],
                                    "what": "first step of the process (warm-up 0): allocator pool growth, pinned staging ring, lazy module load included"} if cold else None,
                "load_buffer_variant": side_buffer},
        "gpu_launches": int(S("kernel_launches")),
        "roofline": {"bound": "hbm", "kernel": "k_merge_small / k_merge (one launch per merge: probe the pair's occurrence list + count deltas | fold + publish | rewrite + new lists)",
                     "achieved": all_bytes / (all_ms * 1e-3) / 1e9 if all_ms else 0.0, "peak": peak, "unit": "GB/s",
                     "frac": all_bytes / (all_ms * 1e-3) / 1e9 / peak if all_ms and peak else None, "traffic": traffic, "traffic_detail": traffic_note,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                     "launches_timed": int(all_n), "avg_launch_us": 1e3 * all_ms / all_n if all_n else None,
                     "bytes_per_launch": all_bytes / all_n if all_n else None,
                     "note": ROOFLINE_NOTE,
                     "touched_gbs": all_touched / (all_ms * 1e-3) / 1e9 if all_ms else None,
                     "dense_launches": {"what": DENSE_NOTE, "n": int(dense_n),
                                        "avg_launch_us": 1e3 * dense_ms / dense_n if dense_n else None,
                                        "achieved": dense_bytes / (dense_ms * 1e-3) / 1e9 if dense_ms else None,
                                        "scan_phase_gbs": dense_bytes / (dense_phase_ms * 1e-3) / 1e9 if dense_phase_ms else None,
                                        "scan_phase_frac": dense_bytes / (dense_phase_ms * 1e-3) / 1e9 / peak if dense_phase_ms and peak else None},
                     "phase_avg_us": {"probe_and_deltas": 1e3 * all_phase_ms / all_n if all_n else None, "fold_and_publish": 1e3 * S("fold_phase_ms") / all_n if all_n else None,
                                      "rewrite_and_new_lists": 1e3 * S("rewrite_phase_ms") / all_n if all_n else None,
                                      "what": "in-kernel %globaltimer of the timed launches; the host continues as soon as phase 2 has published"},
                     "single_cta_merges_per_step": int(st["single_launches"]), "of_which_by_resident_server": int(st["server_merges"]), "server_starts_per_step": int(st["server_starts"])},
        "detail": {"merges_per_step": merges, "n_words": int(st["n_words"]), "n_symbols_initial": int(st["n_symbols_initial"]), "n_symbols_final": int(st["n_symbols_live"]),
                   "occurrences": int(st["occurrences"]), "pair_entries": int(st["pair_entries"]), "heap_pushes": int(st["heap_pushes"]), "heap_pops": int(st["heap_pops"]),
                   "tie_rate_upper": st["tie_root_equal"] / merges if merges else None, "tie_rate_lower": st["tie_same_as_prev"] / merges if merges else None,
                   "tie_note": "SURVEY A15: share of merges whose frequency equals that of the entry left at the heap root (upper bound on a tied maximum) / the previous merge's frequency (lower bound)",
                   "list_entries_probed": int(st["list_entries"]), "pool_entries": int(st["pool_entries"]), "train_device_ms_per_step": train_dev_ms / args.steps, "host_heap_ms": st["host_heap_ms"], "wait_ms": st["wait_ms"], "launch_ms": st["launch_ms"],
                   "ingest_device_ms": st["ingest_device_ms"], "ingest_gbs": st["ingest_bytes"] / (st["ingest_device_ms"] * 1e-3) / 1e9 if st["ingest_device_ms"] else None,
                   "count_device_ms": st["count_device_ms"], "count_gbs": st["count_bytes"] / (st["count_device_ms"] * 1e-3) / 1e9 if st["count_device_ms"] else None,
                   "count_frac_of_hbm_peak": st["count_bytes"] / (st["count_device_ms"] * 1e-3) / 1e9 / peak if st["count_device_ms"] and peak else None,
                   "list_fill_device_ms": st["fill_device_ms"], "list_fill_gbs": st["fill_bytes"] / (st["fill_device_ms"] * 1e-3) / 1e9 if st["fill_device_ms"] else None,
                   "h2d_ms": st["h2d_ms"], "load_s_steps": [round(x["load_s"], 4) for x in steps], "train_s_steps": [round(x["train_s"], 4) for x in steps],
                   "save_s_steps": [round(x["save_s"], 4) for x in steps], "merges_md5": merges_md5, "vocab_md5": vocab_md5,
                   "device": __import__("shredword").cbase.lib.bpe_b200_device_name().decode()},
    }
    if other:
        line["detail"]["replicas" if sharded else "sharded"] = other
    if encoder:
        if "expand_algorithmic_gbs" in encoder and peak:
            encoder["expand_frac_of_hbm_peak"] = encoder["expand_algorithmic_gbs"] / peak
        line["detail"]["encoder"] = encoder
    try:
        big = json.load(open(os.path.join(ROOT, "tests", "golden", "big.json")))["cases"].get(args.workload)
        if big:
            line["detail"]["golden_merges_md5"] = big["merges_md5"]
            line["detail"]["bit_exact_vs_golden"] = big["merges_md5"] == merges_md5 and big["merges"] == merges
            if big.get("vocab_md5"):
                line["detail"]["vocab_bit_exact_vs_golden"] = big["vocab_md5"] == vocab_md5
            if other:
                other["bit_exact_vs_golden"] = big["merges_md5"] == other["merges_md5"]
    except Exception:
        pass
    if world > 1 and rank == 0:
        import shutil
        shutil.rmtree(shard_env["SHRED_RDV"], ignore_errors=True)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = reference_sample(corpus, (vocab, unk, cov, mf), workload=args.workload)
            except Exception as e:  # the checker must never take the measurement down
                line["cpu_baseline"] = {"value": None, "unit": "merges/s", "cores": 1, "kind": "unavailable", "sample": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
