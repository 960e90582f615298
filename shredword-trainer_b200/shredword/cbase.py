"""ctypes binding of the B200 libtrainer.so.

Mirrors what the reference binding does (reference shredword/cbase.py:4-71): locate a `trainer*`/`libtrainer*` shared
library next to the package (here: `<pkg>/lib/`), load it with RTLD_GLOBAL and declare argument/return types for the
8 BPE entry points (reference shredword/csrc/bpe/bpe.h:62-72).  The 13 Unigram symbols the reference binds eagerly are
exported by the library as stubs, so the reference's own cbase.py also imports cleanly against this library.
There is deliberately no fallback: if the CUDA library is missing, importing this module raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))


def _find_library():
    override = os.environ.get("SHRED_LIBTRAINER")
    if override:
        return override
    for d in (_PKG_DIR, os.path.join(_PKG_DIR, "lib"), os.path.join(_PKG_DIR, "..", "build")):
        if not os.path.isdir(d):
            continue
        for name in sorted(os.listdir(d)):
            if (name.startswith("libtrainer") or name.startswith("trainer")) and name.endswith(".so"):
                return os.path.abspath(os.path.join(d, name))
    raise FileNotFoundError(
        "libtrainer.so (the sm_100a CUDA library) was not found under %s; build it with "
        "`python shredword-trainer_b200/build.py` -- there is no CPU fallback" % _PKG_DIR)


_lib_path = _find_library()
lib = ctypes.CDLL(_lib_path, mode=getattr(ctypes, "RTLD_GLOBAL", 0))


class BPEConfig(Structure):  # reference bpe.h:43-48
    _fields_ = [("target_vocab_size", c_size_t), ("unk_id", c_int32), ("character_coverage", c_float), ("min_pair_freq", c_uint64)]


class PairKey(Structure):  # reference hash.h:27-29
    _fields_ = [("first", c_int32), ("second", c_int32)]


class BPEHeapEntry(Structure):  # reference heap.h:17-21
    _fields_ = [("key", PairKey), ("freq", c_uint64), ("version", c_uint32)]


class MaxHeap(Structure):  # reference heap.h:23-27
    _fields_ = [("data", POINTER(BPEHeapEntry)), ("size", c_size_t), ("cap", c_size_t)]


class Corpus(Structure):  # reference bpe.h:37-41
    _fields_ = [("words", c_void_p), ("word_counts", POINTER(c_uint64)), ("vocab_size", c_size_t)]


class BIMap(Structure):
    _fields_ = [("buckets", c_void_p), ("nbuckets", c_size_t)]


class Trainer(Structure):  # reference bpe.h:50-60 (+ impl)
    _fields_ = [("config", BPEConfig), ("heap", MaxHeap), ("corpus", Corpus), ("bigram_map", BIMap), ("next_token", c_size_t),
                ("num_merges", c_size_t), ("merge_ops", POINTER(PairKey)), ("token_strs", c_void_p), ("token_freq", c_void_p), ("impl", c_void_p)]


class Stats(Structure):  # include/shred_abi.h shred_stats_t
    _fields_ = [(n, c_uint64) for n in ("n_words", "n_symbols_initial", "n_symbols_live", "n_slots", "n_tokens", "corpus_bytes",
                                        "pair_entries", "heap_size", "heap_pushes", "heap_pops", "merges", "occurrences", "list_entries", "pool_entries")] + \
               [("scan_launches", c_uint64), ("scan_device_ms", c_double), ("scan_bytes", c_double),
                ("count_launches", c_uint64), ("count_device_ms", c_double), ("count_bytes", c_double), ("fill_device_ms", c_double), ("fill_bytes", c_double),
                ("ingest_launches", c_uint64), ("ingest_device_ms", c_double), ("ingest_bytes", c_double),
                ("kernel_launches", c_uint64)] + \
               [(n, c_double) for n in ("load_wall_ms", "h2d_ms", "train_wall_ms", "host_heap_ms", "wait_ms", "save_wall_ms", "train_device_ms", "launch_ms", "scan_bytes_touched")] + \
               [("dense_launches", c_uint64), ("dense_device_ms", c_double), ("dense_bytes", c_double), ("scan_phase_ms", c_double), ("dense_phase_ms", c_double)] + \
               [("h2d_bytes", c_uint64), ("d2h_bytes", c_uint64), ("tie_root_equal", c_uint64), ("tie_same_as_prev", c_uint64),
                ("fold_phase_ms", c_double), ("rewrite_phase_ms", c_double), ("single_launches", c_uint64), ("server_merges", c_uint64), ("server_starts", c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


lib.create_trainer.argtypes, lib.create_trainer.restype = [POINTER(BPEConfig)], POINTER(Trainer)
lib.bpe_trainer_destroy.argtypes, lib.bpe_trainer_destroy.restype = [POINTER(Trainer)], None
lib.bpe_load_corpus.argtypes, lib.bpe_load_corpus.restype = [POINTER(Trainer), c_char_p], c_int
lib.bpe_init.argtypes, lib.bpe_init.restype = [POINTER(Trainer)], None
lib.bpe_count_bigrams.argtypes, lib.bpe_count_bigrams.restype = [POINTER(Trainer)], None
lib.bpe_merge_batch.argtypes, lib.bpe_merge_batch.restype = [POINTER(Trainer), c_int], c_int
lib.bpe_train.argtypes, lib.bpe_train.restype = [POINTER(Trainer)], c_int
lib.bpe_save.argtypes, lib.bpe_save.restype = [POINTER(Trainer), c_char_p, c_char_p], None
# extensions (include/shred_abi.h)
lib.bpe_b200_load_buffer.argtypes, lib.bpe_b200_load_buffer.restype = [POINTER(Trainer), c_void_p, c_size_t], c_int
lib.bpe_b200_get_stats.argtypes, lib.bpe_b200_get_stats.restype = [POINTER(Trainer), POINTER(Stats)], c_int
lib.bpe_b200_get_words.argtypes, lib.bpe_b200_get_words.restype = [POINTER(Trainer), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_int32), c_uint64], c_int
lib.bpe_b200_get_charset.argtypes, lib.bpe_b200_get_charset.restype = [POINTER(Trainer), POINTER(c_uint8), POINTER(c_uint64)], c_int
lib.bpe_b200_get_pairs.argtypes, lib.bpe_b200_get_pairs.restype = [POINTER(Trainer), POINTER(c_int32), POINTER(c_uint64), c_uint64], c_uint64
lib.bpe_b200_device_name.argtypes, lib.bpe_b200_device_name.restype = [], c_char_p


class EncodeStats(Structure):  # include/shred_abi.h shred_encode_stats_t
    _fields_ = [(n, c_uint64) for n in ("text_bytes", "n_words", "n_unique_words", "n_ids", "pool_ids", "kernel_launches", "h2d_bytes", "d2h_bytes")] + \
               [(n, c_double) for n in ("h2d_ms", "tokenize_ms", "words_ms", "expand_ms", "device_ms", "encode_wall_ms", "d2h_ms")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# encoder (include/shred_abi.h; replaces the reference's pure-Python BPETokenizer, shredword/utils/bpe.py:157-225)
lib.bpe_b200_encoder_create.argtypes, lib.bpe_b200_encoder_create.restype = [POINTER(c_int32), c_size_t], c_void_p
lib.bpe_b200_encoder_load.argtypes, lib.bpe_b200_encoder_load.restype = [c_char_p], c_void_p
lib.bpe_b200_encoder_destroy.argtypes, lib.bpe_b200_encoder_destroy.restype = [c_void_p], None
lib.bpe_b200_encoder_vocab_size.argtypes, lib.bpe_b200_encoder_vocab_size.restype = [c_void_p], c_size_t
lib.bpe_b200_encode.argtypes, lib.bpe_b200_encode.restype = [c_void_p, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)], c_int
lib.bpe_b200_encode_to_host.argtypes, lib.bpe_b200_encode_to_host.restype = [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64, c_void_p, c_uint64, POINTER(c_uint64), POINTER(c_uint64)], c_int
lib.bpe_b200_encode_fetch.argtypes, lib.bpe_b200_encode_fetch.restype = [c_void_p, c_void_p, c_void_p], c_int
lib.bpe_b200_decode.argtypes, lib.bpe_b200_decode.restype = [c_void_p, c_void_p, c_uint64, c_void_p, c_uint64], c_int64
lib.bpe_b200_encoder_get_stats.argtypes, lib.bpe_b200_encoder_get_stats.restype = [c_void_p, POINTER(EncodeStats)], c_int
