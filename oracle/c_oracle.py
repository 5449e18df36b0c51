"""ORACLE — test infrastructure only.  ctypes front-end of `oracle/rpx_oracle.c`."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from reprover_b200._build import ORACLE_LIB, build_oracle

_lib = None


def load():
    global _lib
    if _lib is None:
        if not ORACLE_LIB.exists():
            build_oracle()
        lib = C.CDLL(str(ORACLE_LIB))
        lib.rpx_oracle_tokenize.restype = None
        lib.rpx_oracle_tokenize.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        lib.rpx_oracle_relative_bucket.restype = C.c_int32
        lib.rpx_oracle_relative_bucket.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        lib.rpx_oracle_dot64.restype = C.c_double
        lib.rpx_oracle_dot64.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        lib.rpx_oracle_sim_topk.restype = None
        lib.rpx_oracle_sim_topk.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.rpx_oracle_topk_merge.restype = None
        lib.rpx_oracle_topk_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def tokenize(data: np.ndarray, offsets: np.ndarray, max_len: int) -> Tuple[np.ndarray, np.ndarray]:
    """(packed int32 ids, int32 cu_seqlens [n+1])."""
    lib = load()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    cu = np.zeros(n + 1, dtype=np.int32)
    lib.rpx_oracle_tokenize(_ptr(data), _ptr(offsets), n, max_len, None, _ptr(cu))
    ids = np.zeros(int(cu[-1]), dtype=np.int32)
    lib.rpx_oracle_tokenize(_ptr(data), _ptr(offsets), n, max_len, _ptr(ids), _ptr(cu))
    return ids, cu


def relative_bucket(rel: int, num_buckets: int = 32, max_distance: int = 128) -> int:
    return int(load().rpx_oracle_relative_bucket(rel, num_buckets, max_distance))


def bf16_bits(t) -> np.ndarray:
    """torch bf16 tensor (any device) -> contiguous uint16 numpy array of its raw bits."""
    import torch

    assert t.dtype == torch.bfloat16
    return t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)


def dot64(q_bits: np.ndarray, e_bits: np.ndarray) -> float:
    q_bits = np.ascontiguousarray(q_bits, dtype=np.uint16)
    e_bits = np.ascontiguousarray(e_bits, dtype=np.uint16)
    return float(load().rpx_oracle_dot64(_ptr(q_bits), _ptr(e_bits), len(q_bits)))


def sim_topk(Q_bits: np.ndarray, E_bits: np.ndarray, k: int, mask_words: Optional[np.ndarray] = None,
             idx_offset: int = 0):
    """(fp64 scores [nq,k], int64 idx [nq,k], int32 count [nq]) under (score desc, index asc)."""
    lib = load()
    Q_bits = np.ascontiguousarray(Q_bits, dtype=np.uint16)
    E_bits = np.ascontiguousarray(E_bits, dtype=np.uint16)
    nq, d = Q_bits.shape
    n = E_bits.shape[0]
    scores = np.zeros((nq, k), dtype=np.float64)
    idx = np.zeros((nq, k), dtype=np.int64)
    count = np.zeros(nq, dtype=np.int32)
    stride = 0
    if mask_words is not None:
        mask_words = np.ascontiguousarray(mask_words, dtype=np.uint32)
        stride = mask_words.shape[1]
    lib.rpx_oracle_sim_topk(_ptr(Q_bits), nq, _ptr(E_bits), n, d, k, _ptr(mask_words), stride, idx_offset,
                            _ptr(scores), _ptr(idx), _ptr(count))
    return scores, idx, count


def topk_merge(scores: np.ndarray, idx: np.ndarray):
    """scores/idx [n_parts, nq, k] -> merged ([nq,k] fp64, [nq,k] int64, [nq] int32)."""
    lib = load()
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    n_parts, nq, k = scores.shape
    out_s = np.zeros((nq, k), dtype=np.float64)
    out_i = np.zeros((nq, k), dtype=np.int64)
    out_c = np.zeros(nq, dtype=np.int32)
    lib.rpx_oracle_topk_merge(_ptr(scores), _ptr(idx), n_parts, nq, k, _ptr(out_s), _ptr(out_i), _ptr(out_c))
    return out_s, out_i, out_c
