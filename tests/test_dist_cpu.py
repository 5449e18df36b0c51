"""world_size-2 `gloo` test (CPU) of the sharded-retrieval plumbing in reprover_b200/dist.py:
shard bounds, index offsets, the single all-gather and the merge hand-off.  The two compute
steps are stood in by the oracle (tests may use it); on a GPU box the same code path runs with
rpx_sim_topk / rpx_topk_merge and NCCL (tests/test_dist_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle
from reprover_b200 import synth
from reprover_b200.dist import ShardedIndex, shard_bounds, sharded_topk


def _oracle_local_topk(queries, shard, k, idx_offset, access_mask):
    words = None if access_mask is None else access_mask.numpy().view(np.uint32)
    s, i, _ = c_oracle.sim_topk(c_oracle.bf16_bits(queries), c_oracle.bf16_bits(shard), k, words, idx_offset)
    return torch.from_numpy(s), torch.from_numpy(i)


def _oracle_merge(scores64, idx):
    s, i, c = c_oracle.topk_merge(scores64.numpy(), idx.numpy())
    return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i), torch.from_numpy(c), torch.from_numpy(s)


def _worker(rank, world, port, n_rows, k, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = 128
        E = synth.random_unit_rows(n_rows, d, seed=5, device="cpu")       # same full index on every rank...
        Q = synth.random_unit_rows(6, d, seed=6, device="cpu")
        index = ShardedIndex(n_rows)                                      # ...each rank keeps only its rows
        index.set_embeddings(E[index.lo:index.hi].clone())
        s32, idx, cnt, s64 = index.topk(Q, k, local_topk=_oracle_local_topk, merge=_oracle_merge)
        out_q.put((rank, index.lo, index.hi, idx.numpy(), s64.numpy(), cnt.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,k", [(301, 10), (7, 5)])
def test_two_rank_sharded_topk_equals_single_index(n_rows, k):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, k, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda r: r[0])
    assert [(r[1], r[2]) for r in results] == list(zip(shard_bounds(n_rows, 2)[:-1], shard_bounds(n_rows, 2)[1:]))
    E = synth.random_unit_rows(n_rows, 128, seed=5, device="cpu")
    Q = synth.random_unit_rows(6, 128, seed=6, device="cpu")
    ws, wi, wc = c_oracle.sim_topk(c_oracle.bf16_bits(Q), c_oracle.bf16_bits(E), k)
    for _, _, _, idx, s64, cnt in results:   # every rank holds the same, globally correct answer
        assert np.array_equal(idx, wi) and np.array_equal(s64, ws) and np.array_equal(cnt, wc)


def test_single_process_path_needs_no_process_group():
    E = synth.random_unit_rows(50, 64, seed=1, device="cpu")
    Q = synth.random_unit_rows(3, 64, seed=2, device="cpu")
    s32, idx, cnt, s64 = sharded_topk(Q, E, 4, 0, local_topk=_oracle_local_topk, merge=_oracle_merge)
    ws, wi, wc = c_oracle.sim_topk(c_oracle.bf16_bits(Q), c_oracle.bf16_bits(E), 4)
    assert np.array_equal(idx.numpy(), wi)


def _disk_worker(rank, world, port, n_rows, directory, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        E = synth.random_unit_rows(n_rows, 64, seed=9, device="cpu")
        index = ShardedIndex(n_rows)
        index.set_embeddings(E[index.lo:index.hi].clone())
        index.save(directory, corpus={"stand-in": "corpus"} if rank == 0 else None)
        again = ShardedIndex.load(directory)                 # same world size: own file only
        ok_same = torch.equal(again.embeddings, index.embeddings)
        full = index.gather_embeddings(dst=0)
        ic = None
        if rank == 0:
            from reprover_b200.corpus import Corpus, File, Pos, Premise
            prem = [Premise("A.lean", f"A.p{i}", Pos(i + 1, 0), Pos(i + 1, 1), "c") for i in range(n_rows)]
            ic = index.gather_indexed_corpus(Corpus.from_files([(File("A.lean", prem), [])]))
            ic = (len(ic.corpus), tuple(ic.embeddings.shape), str(ic.embeddings.dtype), str(ic.embeddings.device))
        else:
            index.gather_indexed_corpus(None)
        out_q.put((rank, ok_same, None if full is None else torch.equal(full, E.float()), ic))
    finally:
        dist.destroy_process_group()


def test_sharded_index_on_disk_roundtrip_and_resharding(tmp_path):
    """Two ranks write their rows; the directory reads back per rank, re-cut for 1 and 3 ranks, and
    gathers into the reference's single-file object on rank 0."""
    n_rows, directory = 101, str(tmp_path / "index")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_disk_worker, args=(r, 2, port, n_rows, directory, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in results)                         # same-world reload == what was saved
    assert results[0][2] is True and results[1][2] is None    # gather: full matrix on rank 0 only
    assert results[0][3] == (n_rows, (n_rows, 64), "torch.float32", "cpu")
    E = synth.random_unit_rows(n_rows, 64, seed=9, device="cpu")
    one = ShardedIndex.load(directory, rank=0, world_size=1)
    assert torch.equal(one.embeddings, E) and one.embeddings.dtype == torch.bfloat16
    got = torch.cat([ShardedIndex.load(directory, rank=r, world_size=3).embeddings for r in range(3)])
    assert torch.equal(got, E)
    assert [ShardedIndex.load(directory, rank=r, world_size=3).lo for r in range(3)] == shard_bounds(n_rows, 3)[:-1]
    assert ShardedIndex.load_corpus(directory) == {"stand-in": "corpus"}
    assert sorted(os.listdir(directory)) == ["corpus.pickle", "embeddings.00000-of-00002.pt", "embeddings.00001-of-00002.pt",
                                             "manifest.json"]
