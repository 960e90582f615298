"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) -- test infrastructure only.

The oracle restates the reference trainer (shredword/csrc/bpe/*.cpp) on the CPU; see oracle/bpe_oracle.c for the
file:line map.  Nothing under the product package imports this module.
"""
import ctypes
import hashlib
import os
import struct
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_SO = os.path.join(REF_DIR, "libtrainer_ref.so")
REF_HARNESS = os.path.join(REF_DIR, "ref_harness")
ZMALLOC = os.path.join(REF_DIR, "zmalloc.so")

_lib = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ORACLE_SO) or any(os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("bpe_oracle.c", "bpe_encode_oracle.c")):
        build_oracle()
    L = ctypes.CDLL(ORACLE_SO)
    vp, u64, i32, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int32, ctypes.c_uint32
    L.oracle_create.argtypes, L.oracle_create.restype = [u64, i32, ctypes.c_float, u64], vp
    L.oracle_destroy.argtypes, L.oracle_destroy.restype = [vp], None
    L.oracle_set_verify.argtypes, L.oracle_set_verify.restype = [vp, ctypes.c_int], None
    L.oracle_load.argtypes, L.oracle_load.restype = [vp, ctypes.c_char_p], ctypes.c_int
    L.oracle_load_buffer.argtypes, L.oracle_load_buffer.restype = [vp, ctypes.c_char_p, ctypes.c_size_t], ctypes.c_int
    L.oracle_init.argtypes, L.oracle_init.restype = [vp], None
    L.oracle_count_bigrams.argtypes, L.oracle_count_bigrams.restype = [vp], None
    L.oracle_merge_batch.argtypes, L.oracle_merge_batch.restype = [vp, ctypes.c_int], ctypes.c_int
    L.oracle_train.argtypes, L.oracle_train.restype = [vp], ctypes.c_int
    L.oracle_save.argtypes, L.oracle_save.restype = [vp, ctypes.c_char_p, ctypes.c_char_p], None
    for name in ("oracle_num_words", "oracle_num_merges", "oracle_heap_size", "oracle_pair_entries", "oracle_num_symbols", "oracle_min_freq"):
        getattr(L, name).argtypes, getattr(L, name).restype = [vp], u64
    L.oracle_coverage.argtypes, L.oracle_coverage.restype = [vp], ctypes.c_float
    L.oracle_word_count.argtypes, L.oracle_word_count.restype = [vp, u64], u64
    L.oracle_word_bytes.argtypes, L.oracle_word_bytes.restype = [vp, u64, ctypes.c_char_p, u32], u32
    L.oracle_word_ids.argtypes, L.oracle_word_ids.restype = [vp, u64, ctypes.POINTER(i32), u32], u32
    L.oracle_keep_mask.argtypes, L.oracle_keep_mask.restype = [vp, ctypes.c_char_p, ctypes.POINTER(u64)], None
    L.oracle_get_merges.argtypes, L.oracle_get_merges.restype = [vp, ctypes.POINTER(i32)], None
    L.oracle_get_pairs.argtypes, L.oracle_get_pairs.restype = [vp, ctypes.POINTER(i32), ctypes.POINTER(u64), u64], u64
    L.oracle_get_heap.argtypes, L.oracle_get_heap.restype = [vp, ctypes.POINTER(i32), ctypes.POINTER(u64), u64], u64
    L.oracle_stats.argtypes, L.oracle_stats.restype = [vp, ctypes.POINTER(u64)], None
    # encoder oracle (oracle/bpe_encode_oracle.c)
    L.enc_oracle_create.argtypes, L.enc_oracle_create.restype = [ctypes.POINTER(i32), ctypes.c_size_t], vp
    L.enc_oracle_load.argtypes, L.enc_oracle_load.restype = [ctypes.c_char_p], vp
    L.enc_oracle_destroy.argtypes, L.enc_oracle_destroy.restype = [vp], None
    L.enc_oracle_encode.argtypes, L.enc_oracle_encode.restype = [vp, ctypes.c_char_p, u64], ctypes.c_int
    L.enc_oracle_encode_word.argtypes, L.enc_oracle_encode_word.restype = [vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(i32)], ctypes.c_size_t
    for name in ("enc_oracle_n_tokens", "enc_oracle_n_ids", "enc_oracle_vocab_size"):
        getattr(L, name).argtypes, getattr(L, name).restype = [vp], u64
    L.enc_oracle_ids.argtypes, L.enc_oracle_ids.restype = [vp], ctypes.POINTER(i32)
    L.enc_oracle_offsets.argtypes, L.enc_oracle_offsets.restype = [vp], ctypes.POINTER(u64)
    L.enc_oracle_decode.argtypes, L.enc_oracle_decode.restype = [vp, ctypes.POINTER(i32), u64, ctypes.c_char_p, u64], ctypes.c_int64
    _lib = L
    return L


class Oracle:
    """Mirror of BPETrainer (reference shredword/trainer.py:5-40) on top of the CPU oracle."""

    def __init__(self, vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000, verify=False):
        self.L = lib()
        self.h = self.L.oracle_create(vocab_size, unk_id, character_coverage, min_pair_freq)
        self.vocab_size = vocab_size
        if verify:
            self.L.oracle_set_verify(self.h, 1)

    def load_corpus(self, path):
        rc = self.L.oracle_load(self.h, os.fsencode(path))
        if rc != 0:
            raise IOError(path)

    def load_bytes(self, data: bytes):
        assert self.L.oracle_load_buffer(self.h, data, len(data)) == 0

    def init(self):
        self.L.oracle_init(self.h)

    def count_bigrams(self):
        self.L.oracle_count_bigrams(self.h)

    def merge_batch(self, n):
        return self.L.oracle_merge_batch(self.h, n)

    def train(self):
        return self.L.oracle_train(self.h)

    def save(self, model, vocab):
        self.L.oracle_save(self.h, os.fsencode(model), os.fsencode(vocab))

    def destroy(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # --- introspection
    @property
    def num_words(self):
        return self.L.oracle_num_words(self.h)

    @property
    def num_merges(self):
        return self.L.oracle_num_merges(self.h)

    @property
    def num_symbols(self):
        return self.L.oracle_num_symbols(self.h)

    @property
    def heap_size(self):
        return self.L.oracle_heap_size(self.h)

    def merges(self):
        n = min(self.num_merges, self.vocab_size)
        buf = (ctypes.c_int32 * (3 * max(n, 1)))()
        self.L.oracle_get_merges(self.h, buf)
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]

    def merges_bytes(self):
        return b"".join(struct.pack("<3i", *m) for m in self.merges())

    def words(self):
        out = []
        buf = ctypes.create_string_buffer(1 << 16)
        for i in range(self.num_words):
            n = self.L.oracle_word_bytes(self.h, i, buf, len(buf))
            if n > len(buf):
                buf = ctypes.create_string_buffer(n)
                self.L.oracle_word_bytes(self.h, i, buf, n)
            out.append((buf.raw[:n], self.L.oracle_word_count(self.h, i)))
        return out

    def word_ids(self, i):
        n = self.L.oracle_word_ids(self.h, i, None, 0)
        buf = (ctypes.c_int32 * max(n, 1))()
        self.L.oracle_word_ids(self.h, i, buf, n)
        return list(buf[:n])

    def keep_mask(self):
        keep = ctypes.create_string_buffer(256)
        hist = (ctypes.c_uint64 * 256)()
        self.L.oracle_keep_mask(self.h, keep, hist)
        return list(keep.raw), list(hist)

    def pairs(self):
        n = self.L.oracle_pair_entries(self.h)
        ab = (ctypes.c_int32 * (2 * max(n, 1)))()
        fr = (ctypes.c_uint64 * max(n, 1))()
        self.L.oracle_get_pairs(self.h, ab, fr, n)
        return {(ab[2 * i], ab[2 * i + 1]): fr[i] for i in range(n)}

    def heap(self):
        n = self.heap_size
        abv = (ctypes.c_int32 * (3 * max(n, 1)))()
        fr = (ctypes.c_uint64 * max(n, 1))()
        self.L.oracle_get_heap(self.h, abv, fr, n)
        return [(abv[3 * i], abv[3 * i + 1], fr[i], abv[3 * i + 2] & 0xFFFFFFFF) for i in range(n)]

    def stats(self):
        s = (ctypes.c_uint64 * 3)()
        self.L.oracle_stats(self.h, s)
        return {"pops": s[0], "pushes": s[1], "occ": s[2]}


def md5(b: bytes) -> str:
    return hashlib.md5(b).hexdigest()


def have_reference():
    return all(os.path.exists(p) for p in (REF_SO, REF_HARNESS, ZMALLOC))


def run_reference(corpus_path, vocab_size, unk_id, coverage, min_pair_freq, outdir, max_merges=None, tag="ref", pinned=True, timeout=None):
    """Run the UNMODIFIED reference (oracle/_ref) under the zero-fill malloc shim.  Returns (merges_bytes, vocab_bytes, info)."""
    import json

    mp, vp, mo, js = (os.path.join(outdir, f"{tag}.{x}") for x in ("merges", "vocab", "model", "json"))
    for p in (mp, vp, mo, js):
        if os.path.exists(p):
            os.remove(p)
    env = dict(os.environ)
    if pinned:
        env["LD_PRELOAD"] = ZMALLOC
    cmd = [REF_HARNESS, REF_SO, corpus_path, str(vocab_size), str(unk_id), repr(float(coverage)), str(min_pair_freq),
           "--merges", mp, "--model", mo, "--vocab", vp, "--json", js]
    if max_merges is not None:
        cmd += ["--max-merges", str(max_merges)]
    subprocess.run(cmd, env=env, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
    info = json.load(open(js))
    merges = open(mp, "rb").read()
    vocab = open(vp, "rb").read() if os.path.exists(vp) else b""
    return merges, vocab, info


class EncodeOracle:
    """The reference's Python encoder (shredword/utils/bpe.py:191-225) restated in C for the trainer's model file."""

    def __init__(self, merges=None, model_path=None):
        self.L = lib()
        if model_path is not None:
            self.h = self.L.enc_oracle_load(os.fsencode(model_path))
        else:
            flat = [x for m in merges for x in m]
            arr = (ctypes.c_int32 * max(len(flat), 1))(*flat)
            self.h = self.L.enc_oracle_create(arr, len(merges))
        if not self.h:
            raise ValueError("invalid BPE model")

    def encode(self, data: bytes):
        """-> (ids, offsets): ids of all words back to back, offsets[i] = index of word i's first id (len = words + 1)"""
        self.L.enc_oracle_encode(self.h, data, len(data))
        n, t = self.L.enc_oracle_n_ids(self.h), self.L.enc_oracle_n_tokens(self.h)
        return list(self.L.enc_oracle_ids(self.h)[:n]), list(self.L.enc_oracle_offsets(self.h)[:t + 1])

    def encode_bytes(self, data: bytes):
        """ids and offsets as little-endian byte strings (for hashing large results)"""
        self.L.enc_oracle_encode(self.h, data, len(data))
        n, t = self.L.enc_oracle_n_ids(self.h), self.L.enc_oracle_n_tokens(self.h)
        return ctypes.string_at(self.L.enc_oracle_ids(self.h), 4 * n), ctypes.string_at(self.L.enc_oracle_offsets(self.h), 8 * (t + 1))

    def encode_word(self, word: bytes):
        out = (ctypes.c_int32 * max(len(word), 1))()
        n = self.L.enc_oracle_encode_word(self.h, word, len(word), out)
        return list(out[:n])

    def decode(self, ids):
        arr = (ctypes.c_int32 * max(len(ids), 1))(*ids)
        n = self.L.enc_oracle_decode(self.h, arr, len(ids), None, 0)
        if n < 0:
            raise ValueError("invalid token id")
        buf = ctypes.create_string_buffer(max(n, 1))
        self.L.enc_oracle_decode(self.h, arr, len(ids), buf, n)
        return buf.raw[:n]

    @property
    def vocab_size(self):
        return self.L.enc_oracle_vocab_size(self.h)

    def destroy(self):
        if self.h:
            self.L.enc_oracle_destroy(self.h)
            self.h = None
