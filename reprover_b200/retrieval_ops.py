"""Similarity + top-k on the GPU through the C ABI (`rpx_sim_topk`, `rpx_topk_merge`).

`nearest_premises_device` is the body of `Corpus.get_nearest_premises` (reference
common.py:299-326): the matmul, the ranking and the accessibility filter run in one
device pass; only k (index, score) pairs per query come back to the host.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _native

_ws_cache = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    ws = _ws_cache.get(device)
    if ws is None or ws.numel() < nbytes:
        _ws_cache[device] = None
        ws = torch.empty(nbytes + 4096, dtype=torch.uint8, device=device)
        _ws_cache[device] = ws
    return ws


def sim_topk(queries: torch.Tensor, index: torch.Tensor, k: int, access_mask: Optional[torch.Tensor] = None,
             idx_offset: int = 0, want_scores64: bool = False):
    """Top-k rows of `index` for every row of `queries` (both bf16, same CUDA device).

    Returns (scores fp32 [Q,k], indices int64 [Q,k], counts int32 [Q]) and, with
    `want_scores64`, the fp64 scores as a 4th element.  Order: score desc, index asc.
    `access_mask`: optional uint32 bitmask [Q, ceil(N/32)] (as int32 tensor) on the device.
    """
    lib = _native.load()
    if queries.device.type != "cuda":
        raise RuntimeError("sim_topk runs on a CUDA device only (no CPU path in this engine)")
    if queries.dtype != torch.bfloat16 or index.dtype != torch.bfloat16:
        raise TypeError("sim_topk takes bf16 queries and index (the dtype the reference's GPU path holds them in)")
    assert queries.dim() == 2 and index.dim() == 2 and queries.shape[1] == index.shape[1]
    assert index.device == queries.device
    queries = queries.contiguous()
    index = index.contiguous()
    nq, d = queries.shape
    n = index.shape[0]
    dev = queries.device
    scores = torch.empty(nq, k, dtype=torch.float32, device=dev)
    idx = torch.empty(nq, k, dtype=torch.int64, device=dev)
    counts = torch.empty(nq, dtype=torch.int32, device=dev)
    scores64 = torch.empty(nq, k, dtype=torch.float64, device=dev) if want_scores64 else None
    mask_ptr, stride = None, 0
    if access_mask is not None:
        assert access_mask.device == dev and access_mask.dtype == torch.int32 and access_mask.dim() == 2
        assert access_mask.shape[0] == nq
        access_mask = access_mask.contiguous()
        mask_ptr, stride = access_mask.data_ptr(), access_mask.shape[1]
    with torch.cuda.device(dev):
        need = lib.rpx_sim_topk_workspace_bytes(nq, k)
        if need == 0:
            raise _native.RpxError(_native.RPX_ERR_UNSUPPORTED, _native.last_error())
        ws = _workspace(dev, need)
        _native.check(lib.rpx_sim_topk(
            queries.data_ptr(), nq, index.data_ptr() if n else None, n, d, k, mask_ptr, stride, scores.data_ptr(),
            scores64.data_ptr() if want_scores64 else None, idx.data_ptr(), counts.data_ptr(), idx_offset,
            ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    if want_scores64:
        return scores, idx, counts, scores64
    return scores, idx, counts


def topk_merge(scores64: torch.Tensor, idx: torch.Tensor):
    """Merge [n_parts, Q, k] per-shard results into the global top-k (same ordering contract).

    Returns (scores fp32 [Q,k], idx int64 [Q,k], counts int32 [Q], scores fp64 [Q,k])."""
    lib = _native.load()
    assert scores64.dtype == torch.float64 and idx.dtype == torch.int64 and scores64.shape == idx.shape
    scores64 = scores64.contiguous()
    idx = idx.contiguous()
    n_parts, nq, k = scores64.shape
    dev = scores64.device
    out_s = torch.empty(nq, k, dtype=torch.float32, device=dev)
    out_s64 = torch.empty(nq, k, dtype=torch.float64, device=dev)
    out_i = torch.empty(nq, k, dtype=torch.int64, device=dev)
    out_c = torch.empty(nq, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _native.check(lib.rpx_topk_merge(scores64.data_ptr(), idx.data_ptr(), n_parts, nq, k, out_s.data_ptr(),
                                         out_s64.data_ptr(), out_i.data_ptr(), out_c.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream))
    return out_s, out_i, out_c, out_s64


def nearest_premises_device(corpus, premise_embeddings: torch.Tensor, batch_context, batch_context_emb: torch.Tensor,
                            k: int) -> Tuple[List[list], List[List[float]]]:
    """Reference `Corpus.get_nearest_premises` (common.py:299-326) with the heavy half on the GPU.

    Embeddings are used in bf16 (what the reference's GPU path holds them in,
    retrieval/model.py:59-64, 363-366); fp32 inputs are cast.  Raises ValueError when a
    context has fewer than k accessible premises (common.py:323-324).
    """
    dev = batch_context_emb.device
    if dev.type != "cuda":
        raise RuntimeError("get_nearest_premises needs the embeddings on a CUDA device")
    q = batch_context_emb.to(torch.bfloat16)
    e = premise_embeddings
    if e.device != dev or e.dtype != torch.bfloat16:
        e = e.to(device=dev, dtype=torch.bfloat16)
    assert len(batch_context) == q.shape[0] and e.shape[0] == len(corpus.all_premises)
    words = np.stack([corpus.accessible_mask_words(ctx.path, ctx.theorem_pos) for ctx in batch_context])
    mask = torch.from_numpy(words.view(np.int32)).to(dev)
    scores, idx, counts = sim_topk(q, e, k, access_mask=mask)
    counts_h = counts.cpu().tolist()
    if any(c < k for c in counts_h):
        raise ValueError
    idx_h = idx.cpu().tolist()
    scores_h = scores.cpu().tolist()
    results = [[corpus.all_premises[i] for i in row] for row in idx_h]
    return results, scores_h
