"""Row-sharded index across the GPUs of one box.

The premise corpus splits row-wise into `world_size` contiguous shards; rank r owns
rows [bounds[r], bounds[r+1]).  Re-indexing needs no communication (each rank
encodes its shard).  Retrieval = local fused sim+top-k on every rank, ONE all-gather
of the per-rank [Q, k] (fp64 score, int64 global index) pairs over NCCL/NVLink, and
a device-side k-way merge with the same (score desc, index asc) comparator, so the
result equals the single-GPU result on the concatenated index.

The reference has no counterpart (single device, retrieval/confs/*.yaml `devices: 1`).
The two compute steps are injectable so that the plumbing (bounds, offsets, gather
layout) can be exercised on CPU with the `gloo` backend and the oracle as stand-in.
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Any, Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world_size: int) -> List[int]:
    """Contiguous, near-equal row ranges: bounds[r] .. bounds[r+1]."""
    base, extra = divmod(n_rows, world_size)
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < extra else 0))
    return bounds


def _default_local_topk(queries, shard, k, idx_offset, access_mask):
    """This rank's [Q, k, 2] (fp64 score bits, global index) records, written by the top-k kernel itself."""
    from .retrieval_ops import sim_topk

    return sim_topk(queries, shard, k, access_mask=access_mask, idx_offset=idx_offset, want_packed=True)[-1]


def _default_merge(gathered):
    """[world, Q, k, 2] gathered records -> (scores fp32, idx, counts, scores fp64)."""
    from .retrieval_ops import topk_merge_packed

    return topk_merge_packed(gathered)


def sharded_topk(queries: torch.Tensor, local_shard, k: int, row_offset: int,
                 access_mask: Optional[torch.Tensor] = None, group=None,
                 local_topk: Callable = _default_local_topk, merge: Callable = _default_merge):
    """Global top-k over an index whose rows are spread over the ranks of `group`.

    `queries` [Q, D] must be identical on every rank (replicated); `local_shard` is this rank's
    rows (tensor or `IndexHandle`), `row_offset` its first global row.  `access_mask`, if given, is this
    rank's slice of the per-query bitmask (bit i of the slice <=> global row row_offset + i).  Every
    rank returns the same (scores fp32 [Q,k], global indices int64 [Q,k], counts int32 [Q], scores fp64 [Q,k]).

    Data path: the local kernel writes (fp64 score bits, int64 index) records [Q, k, 2]; ONE
    `all_gather_into_tensor` concatenates the ranks' buffers; the merge kernel reads the gathered
    [world, Q, k, 2] buffer as it is.  No staging copies on either side of the collective.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    packed = local_topk(queries, local_shard, k, row_offset, access_mask)
    if world == 1:
        return merge(packed.unsqueeze(0))
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=torch.int64, device=packed.device)
    dist.all_gather_into_tensor(gathered.view(world * packed.shape[0], *packed.shape[1:]), packed, group=group)
    return merge(gathered)


class ShardedIndex:
    """This rank's slice of a row-sharded embedding index."""

    def __init__(self, n_rows_total: int, rank: Optional[int] = None, world_size: Optional[int] = None,
                 group=None) -> None:
        """Rank / world size default to those of `group` (the default process group when None); the index
        remembers the group, so bounds, the all-gather and the merge always talk about the same ranks."""
        self.group = group
        self.world_size = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.bounds = shard_bounds(n_rows_total, self.world_size)
        self.lo, self.hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self.embeddings: Optional[torch.Tensor] = None

    def local_rows(self) -> range:
        return range(self.lo, self.hi)

    def set_embeddings(self, emb: torch.Tensor) -> None:
        assert emb.shape[0] == self.hi - self.lo
        self.embeddings = emb

    def topk(self, queries: torch.Tensor, k: int, access_mask: Optional[torch.Tensor] = None, **kw):
        assert self.embeddings is not None
        group = kw.pop("group", self.group)
        if dist.is_initialized():
            assert dist.get_world_size(group) == self.world_size, (
                f"index cut for {self.world_size} ranks but the collective runs over {dist.get_world_size(group)}")
        return sharded_topk(queries, self.embeddings, k, self.lo, access_mask=access_mask, group=group, **kw)

    # ------------------------------------------------------------------ on-disk form
    # A directory: `manifest.json` (row bounds, dtype, width), one `embeddings.<r>-of-<R>.pt` per
    # rank (its rows, in the dtype it holds them in) and, optionally, `corpus.pickle`.  Every rank
    # writes and reads only its own rows; a directory written by R ranks can be read by any other
    # number of ranks (rows are re-cut from the files that overlap the new range).  The single-file
    # format the reference's prover loads (`IndexedCorpus` pickle, retrieval/index.py:37-40) is
    # produced by `gather_indexed_corpus`.
    @staticmethod
    def _shard_file(directory: str, r: int, world: int) -> str:
        return os.path.join(directory, f"embeddings.{r:05d}-of-{world:05d}.pt")

    def save(self, directory: str, corpus: Any = None, group=None) -> None:
        assert self.embeddings is not None
        os.makedirs(directory, exist_ok=True)
        torch.save(self.embeddings.detach().cpu().contiguous(), self._shard_file(directory, self.rank, self.world_size))
        if self.rank == 0:
            with open(os.path.join(directory, "manifest.json"), "w") as fh:
                json.dump({"format": "reprover_b200.sharded_index/1", "n_rows_total": self.bounds[-1],
                           "world_size": self.world_size, "bounds": self.bounds, "width": int(self.embeddings.shape[1]),
                           "dtype": str(self.embeddings.dtype).replace("torch.", ""),
                           "has_corpus": corpus is not None}, fh, indent=1)
            if corpus is not None:
                with open(os.path.join(directory, "corpus.pickle"), "wb") as fh:
                    pickle.dump(corpus, fh)
        if dist.is_initialized() and self.world_size > 1:
            dist.barrier(group=group if group is not None else self.group)   # the directory is complete when any rank returns

    @classmethod
    def load(cls, directory: str, rank: Optional[int] = None, world_size: Optional[int] = None,
             device: Any = None) -> "ShardedIndex":
        with open(os.path.join(directory, "manifest.json")) as fh:
            man = json.load(fh)
        assert man.get("format") == "reprover_b200.sharded_index/1", "not a sharded index directory"
        index = cls(man["n_rows_total"], rank=rank, world_size=world_size)
        parts = []
        for r in range(man["world_size"]):
            lo, hi = man["bounds"][r], man["bounds"][r + 1]
            a, b = max(lo, index.lo), min(hi, index.hi)
            if a >= b:
                continue
            rows = torch.load(cls._shard_file(directory, r, man["world_size"]), map_location="cpu", weights_only=True)
            assert rows.shape == (hi - lo, man["width"]), f"shard {r} has shape {tuple(rows.shape)}"
            parts.append(rows[a - lo:b - lo])
        emb = torch.cat(parts) if parts else torch.empty(0, man["width"], dtype=getattr(torch, man["dtype"]))
        index.set_embeddings(emb.to(device) if device is not None else emb)
        return index

    @staticmethod
    def load_corpus(directory: str) -> Any:
        with open(os.path.join(directory, "corpus.pickle"), "rb") as fh:
            return pickle.load(fh)

    def gather_embeddings(self, dst: int = 0, group=None) -> Optional[torch.Tensor]:
        """All rows, in order, as an fp32 CPU tensor on rank `dst` (None elsewhere)."""
        assert self.embeddings is not None
        if self.world_size == 1:
            return self.embeddings.detach().to(torch.float32).cpu()
        # equal-sized buffers for the collective: pad every shard to the longest one
        longest = max(self.bounds[r + 1] - self.bounds[r] for r in range(self.world_size))
        buf = torch.zeros(longest, self.embeddings.shape[1], dtype=self.embeddings.dtype, device=self.embeddings.device)
        buf[: self.hi - self.lo] = self.embeddings
        out = torch.empty((self.world_size * longest, buf.shape[1]), dtype=buf.dtype, device=buf.device)
        group = group if group is not None else self.group
        assert dist.get_world_size(group) == self.world_size
        dist.all_gather_into_tensor(out, buf, group=group)
        if self.rank != dst:
            return None
        out = out.view(self.world_size, longest, -1)
        return torch.cat([out[r, : self.bounds[r + 1] - self.bounds[r]] for r in range(self.world_size)]).to(torch.float32).cpu()

    def gather_indexed_corpus(self, corpus: Any, dst: int = 0, group=None):
        """The reference's on-disk object (`IndexedCorpus(corpus, fp32 CPU embeddings)`) on rank `dst`."""
        from .corpus import IndexedCorpus

        emb = self.gather_embeddings(dst=dst, group=group)
        return None if emb is None else IndexedCorpus(corpus, emb)
