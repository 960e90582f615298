"""BPE encoder / decoder on the GPU for the model file `BPETrainer.save` writes.

Host-side mirror of the reference's pure-Python `BPETokenizer` (reference shredword/utils/bpe.py:157-225): same method names
(`load`, `encode`, `decode`) and error behaviour (`decode` raises ValueError on an id outside the vocabulary), but the
model is the trainer's binary merge list (reference shredword/csrc/bpe/bpe.cpp:419-427) and words are the trainer's
whitespace-delimited tokens (bpe.cpp:131-152) instead of regex chunks.  All work happens in libtrainer.so (CUDA, sm_100a).

LOSSY with respect to whitespace: the delimiter bytes (tab, newline, carriage return, space) separate words and produce no ids,
so decode(encode("a b")) == "ab", whereas the reference's regex chunks keep whitespace and round-trip the text.  encode_bytes
returns CSR offsets (one row per word) for callers that want to re-insert their delimiters.
"""
import ctypes
import struct

import numpy as np

from .cbase import EncodeStats, lib


class BPEEncoder:
    def __init__(self, model_path=None, merges=None):
        self._h = None
        if model_path is not None:
            self.load(model_path)
        elif merges is not None:
            self.set_merges(merges)

    # ---- model
    def load(self, model_file):
        """reference BaseTokenizer.load (utils/bpe.py:140-155), for the trainer's model file"""
        self.destroy()
        self._h = lib.bpe_b200_encoder_load(str(model_file).encode("utf-8"))
        if not self._h:
            raise ValueError("cannot load BPE model %r: unreadable, invalid, or no usable CUDA device (see stderr)" % (model_file,))
        return self

    def set_merges(self, merges):
        """merges: sequence of (a, b, new_id) rows as written by bpe_save"""
        self.destroy()
        flat = [int(x) for m in merges for x in m]
        arr = (ctypes.c_int32 * max(len(flat), 1))(*flat)
        self._h = lib.bpe_b200_encoder_create(arr, len(flat) // 3)
        if not self._h:
            raise ValueError("encoder creation failed: invalid BPE model or no usable CUDA device (see stderr)")
        return self

    @property
    def vocab_size(self):
        return lib.bpe_b200_encoder_vocab_size(self._need())

    # ---- encode / decode
    def encode_bytes(self, data, offsets=True):
        """-> (ids int32[n_ids], offsets uint64[n_words + 1] or None): ids of all words back to back + CSR offsets"""
        h = self._need()
        if isinstance(data, np.ndarray):
            buf, n = data.ctypes.data, data.nbytes
        else:
            data = bytes(data) if not isinstance(data, bytes) else data
            buf, n = ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p), len(data)
        n_words, n_ids = ctypes.c_uint64(), ctypes.c_uint64()
        if lib.bpe_b200_encode(h, buf, n, ctypes.byref(n_words), ctypes.byref(n_ids)) != 0:
            raise RuntimeError("bpe_b200_encode failed")
        ids = np.empty(n_ids.value, dtype=np.int32)
        off = np.empty(n_words.value + 1, dtype=np.uint64) if offsets else None
        if lib.bpe_b200_encode_fetch(h, ids.ctypes.data, off.ctypes.data if offsets else None) != 0:
            raise RuntimeError("bpe_b200_encode_fetch failed")
        return ids, off

    def encode_raw(self, ptr, n_bytes):
        """encode n_bytes at host address ptr; results stay on the device -> (n_words, n_ids)"""
        n_words, n_ids = ctypes.c_uint64(), ctypes.c_uint64()
        if lib.bpe_b200_encode(self._need(), ptr, n_bytes, ctypes.byref(n_words), ctypes.byref(n_ids)) != 0:
            raise RuntimeError("bpe_b200_encode failed")
        return n_words.value, n_ids.value

    def fetch_raw(self, ids_ptr, offsets_ptr):
        """copy the last result to host addresses (int32[n_ids], uint64[n_words + 1]; either may be None)"""
        if lib.bpe_b200_encode_fetch(self._need(), ids_ptr, offsets_ptr) != 0:
            raise RuntimeError("bpe_b200_encode_fetch failed")

    def encode_to_host_raw(self, ptr, n_bytes, ids_ptr, ids_cap, offsets_ptr, offsets_cap):
        """streamed encode (H2D | encode | D2H overlapped) from host address ptr into host buffers -> (n_words, n_ids)"""
        n_words, n_ids = ctypes.c_uint64(), ctypes.c_uint64()
        rc = lib.bpe_b200_encode_to_host(self._need(), ptr, n_bytes, ids_ptr, ids_cap, offsets_ptr, offsets_cap, ctypes.byref(n_words), ctypes.byref(n_ids))
        if rc == -3:
            raise ValueError("output buffers too small")
        if rc != 0:
            raise RuntimeError("bpe_b200_encode_to_host failed")
        return n_words.value, n_ids.value

    def encode_bytes_streamed(self, data):
        """same result as encode_bytes through the streamed entry point (worst-case sized outputs)"""
        if isinstance(data, np.ndarray):
            buf, n = data.ctypes.data, data.nbytes
        else:
            data = bytes(data) if not isinstance(data, bytes) else data
            buf, n = ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p), len(data)
        ids = np.empty(max(n, 1), dtype=np.int32)
        off = np.empty(n // 2 + 2, dtype=np.uint64)
        n_words, n_ids = self.encode_to_host_raw(buf, n, ids.ctypes.data, ids.size, off.ctypes.data, off.size)
        return ids[:n_ids], off[:n_words + 1]

    def encode(self, text):
        """reference BPETokenizer.encode (utils/bpe.py:205-212): text -> list of token ids"""
        data = text.encode("utf-8") if isinstance(text, str) else text
        return self.encode_bytes(data, offsets=False)[0].tolist()

    def decode_bytes(self, ids):
        h = self._need()
        arr = np.ascontiguousarray(ids, dtype=np.int32)
        n = lib.bpe_b200_decode(h, arr.ctypes.data, arr.size, None, 0)
        if n == -2:
            raise ValueError("invalid token id")      # reference utils/bpe.py:220
        if n < 0:
            raise RuntimeError("bpe_b200_decode failed")
        out = np.empty(max(n, 1), dtype=np.uint8)
        if n and lib.bpe_b200_decode(h, arr.ctypes.data, arr.size, out.ctypes.data, n) != n:
            raise RuntimeError("bpe_b200_decode failed")
        return out[:n].tobytes()

    def decode(self, ids):
        """reference BPETokenizer.decode (utils/bpe.py:214-225)"""
        return self.decode_bytes(ids).decode("utf-8", errors="replace")

    def stats(self):
        s = EncodeStats()
        lib.bpe_b200_encoder_get_stats(self._need(), ctypes.byref(s))
        return s.as_dict()

    # ---- lifetime
    def _need(self):
        if not self._h:
            raise RuntimeError("no model loaded")
        return self._h

    def destroy(self):
        if getattr(self, "_h", None):
            lib.bpe_b200_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def read_model(path):
    """model file -> [(a, b, new_id), ...]"""
    raw = open(path, "rb").read()
    return [struct.unpack_from("<3i", raw, 12 * i) for i in range(len(raw) // 12)]
