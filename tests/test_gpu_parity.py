"""GPU (-m gpu): the CUDA path, called through the C ABI (ctypes mirror shredword-trainer_b200/shredword), against
the CPU oracle on the same inputs and against the golden vectors of the unmodified reference.  Staged so that a
failure names the first stage that diverges: ingest -> charset -> pair count + initial heap -> single merges -> full
train -> saved files."""
import os
import struct
import subprocess

import pytest

from cases import GOLDEN, SMALL, case_ids, corpus_bytes
from corpora import generated_corpus
from oracle_lib import Oracle, md5

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T(native):
    os.environ["SHRED_QUIET"] = "1"
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from shredword import BPETrainer
    return BPETrainer


def _merge_bytes(merges):
    return b"".join(struct.pack("<3i", *m) for m in merges)


STAGED = [c for c in SMALL if c["name"] in ("kat_py", "kat_cpp", "kat_py_unk-1", "rnd003", "rnd005", "rnd017", "rnd023", "rnd031", "rnd044",
                                             "multi600k_0", "multi600k_1", "multi600k_2", "multi600k_3", "zipf2m_0", "zipf2m_1", "zipf2m_3")]


@pytest.mark.parametrize("case", STAGED, ids=lambda c: c["name"])
def test_ingest_matches_oracle(case, T):
    """unique words in reference order, their counts, ids after unk substitution, histogram and keep mask"""
    data = corpus_bytes(case)
    vs, unk, cov, mf = case["config"]
    o = Oracle(vs, unk, cov, mf); o.load_bytes(data)
    t = T(vs, unk, cov, mf); t.load_bytes(data)
    assert t.num_words == o.num_words == case["n_words"]
    keep, hist = t.charset()
    okeep, ohist = o.keep_mask()
    assert hist == ohist
    assert keep == okeep
    gw = t.words()
    for i, (ids, cnt) in enumerate(gw):
        assert cnt == o.L.oracle_word_count(o.h, i), f"count of word {i}"
        assert ids == o.word_ids(i), f"symbols of word {i}"
    st = t.stats()
    assert st["n_symbols_initial"] == o.num_symbols
    t.destroy(); o.destroy()


@pytest.mark.parametrize("case", STAGED, ids=lambda c: c["name"])
def test_count_and_stepwise_merges_match_oracle(case, T):
    """pair table after bpe_init, the replay heap (array order) after init and after each of the first merges"""
    data = corpus_bytes(case)
    vs, unk, cov, mf = case["config"]
    o = Oracle(vs, unk, cov, mf); o.load_bytes(data); o.init()
    t = T(vs, unk, cov, mf); t.load_bytes(data); t.init()
    gp = {k: v for k, v in t.pairs().items() if v}
    op = {k: v for k, v in o.pairs().items() if v}
    assert gp == op
    steps = 0
    while True:
        assert [(a, b, f) for a, b, f, _ in t.heap()] == [(a, b, f) for a, b, f, _ in o.heap()], f"heap differs after {steps} merges"
        if steps >= 40 or steps >= vs - 256:
            break
        mo, mg = o.merge_batch(1), t.merge_batch(1)
        assert mo == mg
        if mo <= 0:
            break
        steps += 1
        assert t.merges()[-1] == o.merges()[-1], f"merge {steps}"
    # word table after the merges
    gw = t.words()
    for i in (0, len(gw) // 2, len(gw) - 1) if gw else ():
        assert gw[i][0] == o.word_ids(i)
    gp = {k: v for k, v in t.pairs().items() if v and unk not in k}
    op = {k: v for k, v in o.pairs().items() if v and unk not in k}
    assert gp == op
    t.destroy(); o.destroy()


@pytest.mark.parametrize("case", GOLDEN, ids=case_ids())
def test_train_matches_reference_golden(case, T, tmp_path):
    """full train + save, bit-exact against the unmodified reference's merge list and vocab file"""
    data = corpus_bytes(case)
    vs, unk, cov, mf = case["config"]
    t = T(vs, unk, cov, mf)
    t.load_bytes(data)
    assert t.num_words == case["n_words"]
    assert t.train() == case["merges"]
    mb = _merge_bytes(t.merges())
    assert md5(mb) == case["merges_md5"]
    model, vocab = tmp_path / "m.bin", tmp_path / "v.txt"
    t.save(str(model), str(vocab))
    assert model.read_bytes() == mb
    if case["vocab_md5"] is not None:
        assert md5(vocab.read_bytes()) == case["vocab_md5"]
    t.destroy()


def test_load_from_file_and_list_statistics(T, tmp_path):
    """bpe_load_corpus (file path); every occurrence a merge rewrites comes out of an occurrence list, so the lists probed
    cover the occurrences, and the pool holds the initial lists plus the lists of the pairs the merges created"""
    p = generated_corpus(str(tmp_path / "z.txt"), 3_000_000, 11, 14, "zipf")
    o = Oracle(4000, 0, 0.995, 2); o.load_corpus(p); n = o.train()
    t = T(4000, 0, 0.995, 2); t.load_corpus(p)
    assert t.train() == n
    assert t.merges() == o.merges()
    st = t.stats()
    assert st["list_entries"] >= st["occurrences"] > 0
    assert 0 < st["pool_entries"] <= 3 * st["n_symbols_initial"]
    assert st["n_symbols_live"] == o.num_symbols == st["n_symbols_initial"] - st["occurrences"]
    for i, (ids, cnt) in enumerate(t.words()[:2000]):
        assert ids == o.word_ids(i)
    mo, vo, mg, vg = (str(tmp_path / x) for x in ("mo", "vo", "mg", "vg"))
    o.save(mo, vo); t.save(mg, vg)
    assert open(mo, "rb").read() == open(mg, "rb").read() and open(vo, "rb").read() == open(vg, "rb").read()
    t.destroy(); o.destroy()


def test_table_growth_paths(T):
    """the delta table (and with it the mapped record buffer) must be able to grow inside count_pairs / merge
    without the host reading records through a stale pointer: near-random bytes give ~60 k distinct pairs, more than half
    of the initial 65536-slot delta table; min_pair_freq 1 turns every one of them into a record.  Then train twice."""
    import random
    rng = random.Random(5)
    alphabet = [c for c in range(1, 256) if c not in b"\t\r\n "]
    data = b" ".join(bytes(rng.choice(alphabet) for _ in range(rng.randint(2, 40))) for _ in range(60000))
    o = Oracle(700, 0, 0.995, 1); o.load_bytes(data); n = o.train()
    t = T(700, 0, 0.995, 1); t.load_bytes(data)
    assert t.train() == n and t.merges() == o.merges()
    assert t.stats()["pair_entries"] > 40000
    n2, m2 = o.train(), t.train()   # bpe_train again on the already merged corpus: recount (tokens of every length), lists rebuilt
    assert m2 == n2 and t.merges() == o.merges()
    for i, (ids, cnt) in enumerate(t.words()[:500]):
        assert ids == o.word_ids(i)
    t.destroy(); o.destroy()


EDGE = {
    "empty": b"",
    "only_delims": b" \n\t\r\n   ",
    "single_token": b"a",
    "single_pair": b"ab",
    "no_final_newline": b"abc abc abd abd abc",
    "repeated_char_runs": b"aaaaaaa aaaa aaa aa a aaaaaaaaaaaaaaaa bbbbaaaabbbb\n" * 7,
    "long_word": (b"xy" * 6000) + b" " + (b"xyz" * 3000) + b"\n" + b"xy xyz xyxy\n" * 20,
    "long_line": b" ".join([b"tok%d" % (i % 37) for i in range(20000)]),
    "nul_bytes": b"abc abd\x00hidden hidden hidden\nabc abd abc\x00\x00 zz\nplain line abc\n" * 9,
    "high_bytes": bytes(range(128, 256)) * 3 + b" " + bytes(range(128, 256)) + b"\n",
    "tabs_crlf": b"one\ttwo\r\nthree  two\tone\r\n" * 30,
}


@pytest.mark.parametrize("name", sorted(EDGE), ids=sorted(EDGE))
@pytest.mark.parametrize("cfg", [(400, 0, 0.995, 1), (300, -1, 0.9, 2)], ids=["unk0", "unk-1"])
def test_edge_cases_match_oracle(name, cfg, T, tmp_path):
    data = EDGE[name]
    o = Oracle(*cfg); o.load_bytes(data); n = o.train()
    t = T(*cfg); t.load_bytes(data)
    assert t.num_words == o.num_words
    assert t.train() == n
    assert t.merges() == o.merges()
    mo, vo, mg, vg = (str(tmp_path / x) for x in ("mo", "vo", "mg", "vg"))
    o.save(mo, vo); t.save(mg, vg)
    assert open(mo, "rb").read() == open(mg, "rb").read() and open(vo, "rb").read() == open(vg, "rb").read()
    t.destroy(); o.destroy()


def test_call_order_semantics(T, tmp_path):
    a, b = b"aa bb aa cc aa\n" * 50, b"xyz xyz xy zz\n" * 50
    # second load replaces the first (reference bpe.cpp:176-183)
    t = T(300, 0, 0.995, 1); t.load_bytes(a); t.load_bytes(b); t.train()
    o = Oracle(300, 0, 0.995, 1); o.load_bytes(b); o.train()
    assert t.merges() == o.merges()
    # train twice keeps merging from the merged corpus (reference bpe.cpp:351)
    data = corpus_bytes([c for c in GOLDEN if c["name"] == "kat_py"][0])
    t2 = T(280, 0, 0.995, 2); t2.load_bytes(data); n1 = t2.train(); n2 = t2.train()
    o2 = Oracle(280, 0, 0.995, 2); o2.load_bytes(data); assert (o2.train(), o2.train()) == (n1, n2)
    assert t2.num_merges == o2.num_merges and t2.merges() == o2.merges()
    # vocab_size < 256 -> 0 merges (reference test_bpe.py:56-65); destroy without load; missing file -> IOError
    t3 = T(50, 0, 0.995, 1000); t3.load_bytes(data); assert t3.train() == 0; t3.destroy()
    t4 = T(10); 
    with pytest.raises(IOError):
        t4.load_corpus(str(tmp_path / "missing.txt"))
    t4.destroy()
    # bpe_count_bigrams directly after load seeds the heap (reference bpe_test.cpp:133-167)
    t5 = T(300, -1, 0.99, 2); t5.load_bytes(open(os.path.join(os.path.dirname(__file__), "golden", "kat_cpp.txt"), "rb").read())
    t5.count_bigrams()
    h = t5.heap()
    assert len(h) > 0 and h[0][2] >= h[1][2]
    assert t5.merge_batch(1) == 1 and t5.num_merges == 1
    for x in (t, t2, t5):
        x.destroy()
    o.destroy(); o2.destroy()


def test_reference_pytest_contract(T, tmp_path):
    """the three checks of the reference's own test/test_bpe.py, through the same Python API"""
    corpus = os.path.join(os.path.dirname(__file__), "golden", "kat_py.txt")
    t = T(vocab_size=300, min_pair_freq=2)
    t.load_corpus(corpus)
    assert t.train() > 0
    model, vocab = tmp_path / "out" / "bpe.model", tmp_path / "out" / "bpe.vocab"
    t.save(str(model), str(vocab)); t.destroy()
    assert model.stat().st_size > 0 and vocab.stat().st_size > 0
    assert model.stat().st_size == 44 * 12 and vocab.read_bytes().count(b"\n") == 256 + 44 + 1


def test_cli_drop_in(native, tmp_path):
    """trainer.exe key=value CLI (reference trainer.cpp): unk_id is fixed to -1, files equal the reference's"""
    corpus = os.path.join(os.path.dirname(__file__), "golden", "kat_cpp.txt")
    model, vocab = tmp_path / "m.bin", tmp_path / "v.txt"
    r = subprocess.run([native["exe"], f"input={corpus}", "model_type=bpe", f"output_model={model}", f"output_vocab={vocab}", "vocab_size=300",
                        "character_coverage=0.99", "min_pair_freq=2", "bogus", "other=1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = [c for c in GOLDEN if c["name"] == "kat_cpp"][0]
    assert md5(model.read_bytes()) == g["merges_md5"] and md5(vocab.read_bytes()) == g["vocab_md5"]
    assert subprocess.run([native["exe"]], capture_output=True).returncode == 0
    assert subprocess.run([native["exe"], "input=x"], capture_output=True).returncode == 1
    assert subprocess.run([native["exe"], f"input={corpus}", "model_type=unigram", "output_model=a", "output_vocab=b"], capture_output=True).returncode == 1


def test_size_independent_properties_at_scale(T, tmp_path):
    """100 MB (BASELINE config 1 shape): properties that hold at any size + the oracle's full answer"""
    p = generated_corpus(str(tmp_path / "c1.txt"), 100_000_000, 1, 20, "zipf")
    t = T(8192, 0, 0.995, 2000); t.load_corpus(p)
    st0 = t.stats()
    n = t.train()
    st = t.stats()
    merges = t.merges()
    assert n == len(merges) and all(m[2] == 256 + i for i, m in enumerate(merges))
    assert all(0 <= a < 256 + i and 0 <= b < 256 + i for i, (a, b, _) in enumerate(merges))  # operands exist before use
    assert len(set((a, b) for a, b, _ in merges)) == n                                          # a pair merges once
    assert st["n_symbols_live"] == st0["n_symbols_initial"] - st["occurrences"]                 # one symbol dies per occurrence
    model, vocab = tmp_path / "m", tmp_path / "v"
    t.save(str(model), str(vocab))
    # total token frequency mass = number of live symbols weighted by word counts; merges conserve characters
    lines = vocab.read_bytes().split(b"\n")
    assert model.stat().st_size == 12 * n
    o = Oracle(8192, 0, 0.995, 2000); o.load_corpus(p); assert o.train() == n
    assert o.merges() == merges
    ov = tmp_path / "ov"; o.save(str(tmp_path / "om"), str(ov))
    assert ov.read_bytes() == vocab.read_bytes() and len(lines) > 256
    t.destroy(); o.destroy()


def test_two_trainers_interleaved(T):
    """independent trainers are independent (reference: no globals): interleave their step-wise calls"""
    da = open(os.path.join(os.path.dirname(__file__), "golden", "kat_cpp.txt"), "rb").read()
    db = open(os.path.join(os.path.dirname(__file__), "golden", "kat_py.txt"), "rb").read()
    a, b = T(300, -1, 0.99, 2), T(300, 0, 0.995, 2)
    a.load_bytes(da); b.load_bytes(db); a.init(); b.init()
    for _ in range(44):
        assert a.merge_batch(1) == 1
        assert b.merge_batch(1) == 1
    ga = [c for c in GOLDEN if c["name"] == "kat_cpp"][0]
    gb = [c for c in GOLDEN if c["name"] == "kat_py"][0]
    assert md5(_merge_bytes(a.merges())) == ga["merges_md5"] and md5(_merge_bytes(b.merges())) == gb["merges_md5"]
    a.destroy(); b.destroy()


def test_trainers_in_concurrent_threads(T, tmp_path):
    """four host threads of one process, each with its own trainer on the same GPU (load, train, save, destroy, twice over):
    every trainer keeps its own resident merge server, the cooperative launches size their grids for the SMs the servers
    hold, the pinned blocks and the staging ring are shared process-wide -- and every result equals the golden one"""
    import threading
    names = ("zipf2m_0", "multi600k_1", "zipf2m_3", "multi600k_2")
    cases = [[c for c in GOLDEN if c["name"] == nm][0] for nm in names]
    errors = []

    def work(i, case):
        try:
            for rnd in range(2):
                path = str(tmp_path / ("c%d_%d.txt" % (i, rnd)))
                open(path, "wb").write(corpus_bytes(case))
                t = T(*case["config"])
                t.load_corpus(path) if rnd else t.load_bytes(corpus_bytes(case))
                assert t.train() == case["merges"]
                assert md5(_merge_bytes(t.merges())) == case["merges_md5"]
                t.save(str(tmp_path / ("m%d.bin" % i)), str(tmp_path / ("v%d.txt" % i)))
                if case["vocab_md5"] is not None:
                    assert md5(open(str(tmp_path / ("v%d.txt" % i)), "rb").read()) == case["vocab_md5"]
                t.destroy()
        except BaseException as e:  # noqa: BLE001 -- reported by the main thread
            errors.append((case["name"], repr(e)))

    threads = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(cases)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=300)
    assert not any(th.is_alive() for th in threads), "a trainer thread hangs"
    assert not errors, errors


def test_serialised_launches_fall_back_to_one_launch_per_merge(native, tmp_path):
    """under a tool that serialises kernel launches (here CUDA_LAUNCH_BLOCKING=1; ncu and compute-sanitizer do the same) the
    resident merge server can never run beside the host: the trainer notices (two servers in a row leave before their first
    command), says so once and does every merge as a launch -- same merges"""
    import sys
    case = [c for c in GOLDEN if c["name"] == "kat_cpp"][0]
    corpus = tmp_path / "c.txt"
    corpus.write_bytes(corpus_bytes(case))
    code = (
        "import sys, struct, hashlib; sys.path.insert(0, %r)\n"
        "from shredword import BPETrainer\n"
        "t = BPETrainer(*%r); t.load_corpus(%r); n = t.train()\n"
        "print(n, hashlib.md5(b''.join(struct.pack('<3i', *m) for m in t.merges())).hexdigest(), int(t.stats()['server_merges']))\n"
        "t.destroy()\n"
    ) % (os.path.join(os.path.dirname(os.path.dirname(__file__)), "shredword-trainer_b200"), tuple(case["config"]), str(corpus))
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1", SHRED_QUIET="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    n, digest, served = r.stdout.split()[-3:]
    assert int(n) == case["merges"] and digest == case["merges_md5"] and int(served) == 0
    assert "resident merge server cannot run" in r.stderr


def test_load_corpus_from_file_paths(T, tmp_path):
    """bpe_load_corpus(path) streams the file through the pinned staging ring; a file with NUL bytes falls back to the
    host-side blanking pass; an empty file loads; sizes around the 32 MiB staging chunk are covered by a 70 MB file"""
    blobs = {"nul": EDGE["nul_bytes"] + b"tail\x00gone", "empty": b"", "plain": EDGE["tabs_crlf"]}
    for name, data in blobs.items():
        p = tmp_path / (name + ".txt"); p.write_bytes(data)
        o = Oracle(300, 0, 0.995, 1); o.load_corpus(str(p)); n = o.train()
        t = T(300, 0, 0.995, 1); t.load_corpus(str(p))
        assert t.num_words == o.num_words and t.train() == n and t.merges() == o.merges(), name
        t.destroy(); o.destroy()
    p = generated_corpus(str(tmp_path / "z70.txt"), 70_000_000, 5, 18, "zipf")
    o = Oracle(600, 0, 0.995, 2000); o.load_corpus(p); n = o.train()
    t = T(600, 0, 0.995, 2000); t.load_corpus(p)
    t2 = T(600, 0, 0.995, 2000); t2.load_bytes(open(p, "rb").read())
    assert t2.words()[:500] == t.words()[:500]   # streamed file == host buffer
    assert t.num_words == o.num_words and t.train() == n and t.merges() == o.merges()
    for x in (t, t2):
        x.destroy()
    o.destroy()


MIDSIZE = [
    ("zipf30m_unk-1", (30_000_000, 21, 18, "zipf"), (3000, -1, 0.9995, 50)),
    ("multi20m_unk-5", (20_000_000, 22, 17, "multi"), (5000, -5, 0.9, 5)),
    ("multi20m_unk65", (20_000_000, 23, 17, "multi"), (4000, 65, 0.5, 20)),
    ("zipf50m_unk300", (50_000_000, 24, 20, "zipf"), (6000, 300, 0.995, 100)),
]


@pytest.mark.parametrize("name,gen,cfg", MIDSIZE, ids=[m[0] for m in MIDSIZE])
def test_midsize_configs_match_oracle(name, gen, cfg, T, tmp_path):
    """20-50 MB corpora x the awkward configurations (negative unk ids, unk colliding with a byte / a future merge id,
    low coverage, low min_pair_freq): merge list and vocab file against the oracle"""
    p = generated_corpus(str(tmp_path / "c.txt"), *gen)
    o = Oracle(*cfg); o.load_corpus(p); n = o.train()
    t = T(*cfg); t.load_corpus(p)
    assert t.num_words == o.num_words
    assert t.train() == n
    assert t.merges() == o.merges()
    mo, vo, mg, vg = (str(tmp_path / x) for x in ("mo", "vo", "mg", "vg"))
    o.save(mo, vo); t.save(mg, vg)
    assert open(mo, "rb").read() == open(mg, "rb").read()
    if cfg[1] < 0 or cfg[1] < 256 + n:
        assert open(vo, "rb").read() == open(vg, "rb").read()
    t.destroy(); o.destroy()


def test_api_corner_cases(T, tmp_path):
    data = open(os.path.join(os.path.dirname(__file__), "golden", "kat_py.txt"), "rb").read()
    # save right after load (no training): 256 byte tokens with their corpus frequencies, empty model
    o = Oracle(300, 0, 0.995, 2); o.load_bytes(data)
    t = T(300, 0, 0.995, 2); t.load_bytes(data)
    mo, vo, mg, vg = (str(tmp_path / x) for x in ("mo", "vo", "mg", "vg"))
    o.save(mo, vo); t.save(mg, vg)
    assert open(mg, "rb").read() == b"" and open(vo, "rb").read() == open(vg, "rb").read()
    # merge_batch before bpe_init: the heap is empty, nothing happens (reference bpe.cpp:237-240)
    assert t.merge_batch(5) == 0 and t.num_merges == 0
    # vocab_size 256 and 257: zero and exactly one merge
    for vs in (256, 257):
        a, b = T(vs, 0, 0.995, 2), Oracle(vs, 0, 0.995, 2)
        a.load_bytes(data); b.load_bytes(data)
        assert a.train() == b.train() == vs - 256 and a.merges() == b.merges()
        a.destroy(); b.destroy()
    # a trainer that never loads anything can train, save and be destroyed
    e = T(300)
    assert e.train() == 0
    e.save(str(tmp_path / "em"), str(tmp_path / "ev"))
    assert (tmp_path / "ev").read_bytes().count(b"\n") == 257 and (tmp_path / "em").read_bytes() == b""
    e.destroy(); t.destroy(); o.destroy()
