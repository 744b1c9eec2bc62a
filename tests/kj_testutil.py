"""Shared helpers for the GPU parity tests (bf16 bit conversions, error metrics)."""
import ctypes as C

import numpy as np


def to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 bit pattern, round-to-nearest-even (what cvt.rn.bf16.f32 does)."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) >> 16).astype(np.uint16)


def from_bf16_bits(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, np.uint16).astype(np.uint32) << 16).view(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    return from_bf16_bits(to_bf16_bits(x))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def cosine_rows(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    num = (a * b).sum(-1)
    den = np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1)
    return num / np.maximum(den, 1e-30)


def has_gpu() -> bool:
    try:
        from kjarni_b200 import _native as N

        return N.lib().kjc_device_count() > 0
    except Exception:
        return False
