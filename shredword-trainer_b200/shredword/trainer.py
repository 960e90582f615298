"""BPETrainer -- same class, arguments, defaults and error behaviour as the reference (shredword/trainer.py:5-40),
running on the B200 library.  Extra read-only helpers (stats, words, pairs, merges) exist for tests and benchmarks."""
import ctypes
import os

from .cbase import BPEConfig, Stats, lib


class BPETrainer:
    def __init__(self, vocab_size=8192, unk_id=0, character_coverage=0.995, min_pair_freq=2000):
        self.config = BPEConfig(target_vocab_size=vocab_size, unk_id=unk_id, character_coverage=character_coverage, min_pair_freq=min_pair_freq)
        self.trainer = lib.create_trainer(ctypes.byref(self.config))
        if not self.trainer:
            raise RuntimeError("Failed to create BPE trainer")

    def load_corpus(self, path: str):
        if not os.path.exists(path):
            raise IOError(f"Corpus file does not exist: {path}")
        result = lib.bpe_load_corpus(self.trainer, os.fsencode(path))
        if result != 0:
            raise IOError(f"Failed to load corpus from {path} (code {int(result)})")

    def load_bytes(self, data):
        """Extension: ingest a corpus that is already in host memory (bytes or a buffer with .ctypes / address)."""
        if isinstance(data, (bytes, bytearray)):
            buf = (ctypes.c_char * len(data)).from_buffer_copy(data) if isinstance(data, bytes) else (ctypes.c_char * len(data)).from_buffer(data)
            rc = lib.bpe_b200_load_buffer(self.trainer, ctypes.cast(buf, ctypes.c_void_p), len(data))
        else:  # numpy array or torch tensor exposing a host pointer
            ptr = data.ctypes.data if hasattr(data, "ctypes") else data.data_ptr()
            n = data.nbytes if hasattr(data, "nbytes") else data.numel() * data.element_size()
            rc = lib.bpe_b200_load_buffer(self.trainer, ctypes.c_void_p(ptr), n)
        if rc != 0:
            raise IOError(f"Failed to load corpus from buffer (code {int(rc)})")

    def train(self) -> int:
        merges = lib.bpe_train(self.trainer)
        if merges < 0:
            raise RuntimeError("Training failed")
        print(f"Training completed: {int(merges)} merges performed.")
        return int(merges)

    def save(self, model_path: str, vocab_path: str):
        model_dir, vocab_dir = os.path.dirname(model_path), os.path.dirname(vocab_path)
        if model_dir:
            os.makedirs(model_dir, exist_ok=True)
        if vocab_dir:
            os.makedirs(vocab_dir, exist_ok=True)
        lib.bpe_save(self.trainer, os.fsencode(model_path), os.fsencode(vocab_path))
        print(f"Model saved to: {model_path}")
        print(f"Vocabulary saved to: {vocab_path}")

    def destroy(self):
        if getattr(self, "trainer", None):
            try:
                lib.bpe_trainer_destroy(self.trainer)
            finally:
                self.trainer = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.destroy()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # ---- step-wise API (reference bpe.h:66-69) and read-only helpers
    def init(self):
        lib.bpe_init(self.trainer)

    def count_bigrams(self):
        lib.bpe_count_bigrams(self.trainer)

    def merge_batch(self, n: int) -> int:
        return int(lib.bpe_merge_batch(self.trainer, n))

    @property
    def num_merges(self) -> int:
        return int(self.trainer.contents.num_merges)

    @property
    def num_words(self) -> int:
        return int(self.trainer.contents.corpus.vocab_size)

    def merges(self):
        t = self.trainer.contents
        n = min(int(t.num_merges), max(int(t.config.target_vocab_size), 1))
        return [(t.merge_ops[i].first, t.merge_ops[i].second, 256 + i) for i in range(n)]

    def heap(self):
        t = self.trainer.contents
        return [(t.heap.data[i].key.first, t.heap.data[i].key.second, t.heap.data[i].freq, t.heap.data[i].version) for i in range(int(t.heap.size))]

    def stats(self) -> dict:
        s = Stats()
        if lib.bpe_b200_get_stats(self.trainer, ctypes.byref(s)) != 0:
            raise RuntimeError("stats unavailable")
        return s.as_dict()

    def words(self):
        """[(ids, count)] in reference word order, read back from HBM."""
        n = self.num_words
        st = self.stats()
        cap = int(st["n_slots"]) + 1
        counts = (ctypes.c_uint64 * max(n, 1))()
        off = (ctypes.c_uint64 * (n + 1))()
        ids = (ctypes.c_int32 * cap)()
        rc = lib.bpe_b200_get_words(self.trainer, counts, off, ids, cap)
        if rc != 0:
            raise RuntimeError(f"get_words failed ({rc})")
        return [(list(ids[off[i]:off[i + 1]]), int(counts[i])) for i in range(n)]

    def charset(self):
        keep = (ctypes.c_uint8 * 256)()
        hist = (ctypes.c_uint64 * 256)()
        if lib.bpe_b200_get_charset(self.trainer, keep, hist) != 0:
            raise RuntimeError("charset unavailable")
        return list(keep), list(hist)

    def pairs(self):
        n = int(lib.bpe_b200_get_pairs(self.trainer, None, None, 0))
        ab = (ctypes.c_int32 * (2 * max(n, 1)))()
        fr = (ctypes.c_uint64 * max(n, 1))()
        lib.bpe_b200_get_pairs(self.trainer, ab, fr, n)
        return {(ab[2 * i], ab[2 * i + 1]): int(fr[i]) for i in range(n)}


class UnigramTrainer:
    """The Unigram trainer is outside the scope of the B200 library; constructing one fails the
    same way the reference does when trainerCreate returns NULL (reference shredword/trainer.py:46-47)."""

    def __init__(self, vocab_size=32000, character_coverage=0.9995, max_sentencepiece_length=16, seed_size=1000000):
        self.trainer = lib.trainerCreate(vocab_size, character_coverage, max_sentencepiece_length, seed_size) if hasattr(lib, "trainerCreate") else None
        if not self.trainer:
            raise RuntimeError("Failed to create Unigram trainer")
