"""ctypes binding of oracle/libkjarni_oracle.so (oracle/kjarni_oracle.c): the C restatement of the reference CPU path.

TEST INFRASTRUCTURE ONLY (see the header of kjarni_oracle.c): the timed CPU baseline of bench.py and a second checker in
tests/.  Weights come from the numpy oracle's loader (oracle/kjarni_oracle.py:load_model_dir).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkjarni_oracle.so")
_fp = C.POINTER(C.c_float)


class KoLayer(C.Structure):
    _fields_ = [(n, _fp) for n in ("wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "ln1_g", "ln1_b", "w1", "b1", "w2", "b2", "ln2_g", "ln2_b")]


class KoModel(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("hidden", "layers", "heads", "inter", "vocab", "max_pos", "type_vocab", "pos_offset")] + [
        ("eps", C.c_float), ("word", _fp), ("pos", _fp), ("type", _fp), ("emb_g", _fp), ("emb_b", _fp), ("layer", C.POINTER(KoLayer))]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C oracle`")
        l = C.CDLL(LIB_PATH)
        l.ko_num_threads.restype = C.c_int
        l.ko_set_num_threads.restype = None
        l.ko_set_num_threads.argtypes = [C.c_int]
        l.ko_embed.restype = C.c_int
        l.ko_embed.argtypes = [C.POINTER(KoModel), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.ko_encoder_forward.restype = C.c_int
        l.ko_encoder_forward.argtypes = [C.POINTER(KoModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.ko_scan_topk.restype = None
        l.ko_scan_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.ko_row_norms.restype = None
        l.ko_row_norms.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
        _lib = l
    return _lib


class CModel:
    """KoModel built from the numpy oracle's EncoderModel; keeps the arrays alive."""

    def __init__(self, m):
        self._keep = []

        def p(a):
            if a is None:
                return _fp()
            a = np.ascontiguousarray(a, np.float32)
            self._keep.append(a)
            return a.ctypes.data_as(_fp)

        layers = (KoLayer * len(m.layer))()
        for i, lw in enumerate(m.layer):
            for name, _ in KoLayer._fields_:
                setattr(layers[i], name, p(getattr(lw, name)))
        self._layers = layers
        self.hidden = m.hidden
        self.c = KoModel(m.hidden, m.layers, m.heads, m.layer[0].w1.shape[0], m.word.shape[0], m.pos.shape[0],
                         0 if m.typ is None else m.typ.shape[0], m.position_offset, float(m.eps), p(m.word), p(m.pos), p(m.typ),
                         p(m.emb_g), p(m.emb_b), layers)

    def embed(self, ids, mask, noalloc=None) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.uint32)
        mask = np.ascontiguousarray(mask, np.float32)
        b, s = ids.shape
        if noalloc is None:
            noalloc = ids.size <= 1 or ids.size >= 1000  # ComputeStrategy::select, KT/cpu/strategy.rs:29-47
        out = np.empty((b, self.hidden), np.float32)
        rc = lib().ko_embed(C.byref(self.c), ids.ctypes.data, mask.ctypes.data, b, s, int(noalloc), out.ctypes.data)
        if rc != 0:
            raise MemoryError("ko_embed failed")
        return out

    def hidden_states(self, ids, mask, type_ids=None, noalloc=False) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.uint32)
        mask = np.ascontiguousarray(mask, np.float32)
        b, s = ids.shape
        tt = None if type_ids is None else np.ascontiguousarray(type_ids, np.uint32)
        out = np.empty((b, s, self.hidden), np.float32)
        rc = lib().ko_encoder_forward(C.byref(self.c), ids.ctypes.data, mask.ctypes.data, None if tt is None else tt.ctypes.data, b, s,
                                      int(noalloc), out.ctypes.data)
        if rc != 0:
            raise MemoryError("ko_encoder_forward failed")
        return out


def scan_topk(rows, queries, k):
    rows = np.ascontiguousarray(rows, np.float32)
    q = np.ascontiguousarray(queries, np.float32)
    n, dim = rows.shape
    norms = np.empty((n,), np.float32)
    lib().ko_row_norms(rows.ctypes.data, n, dim, norms.ctypes.data)
    ids = np.empty((q.shape[0], k), np.uint64)
    sc = np.empty((q.shape[0], k), np.float32)
    lib().ko_scan_topk(rows.ctypes.data, norms.ctypes.data, n, dim, q.ctypes.data, q.shape[0], k, ids.ctypes.data, sc.ctypes.data)
    return ids, sc


def set_num_threads(n: int) -> None:
    lib().ko_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().ko_num_threads())
