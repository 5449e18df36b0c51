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
    """Stand-in for the top-k kernel: the [Q, k, 2] (fp64 score bits, global index) records it writes."""
    words = None if access_mask is None else access_mask.numpy().view(np.uint32)
    s, i, _ = c_oracle.sim_topk(c_oracle.bf16_bits(queries), c_oracle.bf16_bits(shard), k, words, idx_offset)
    return torch.stack([torch.from_numpy(s).view(torch.int64), torch.from_numpy(i)], dim=-1).contiguous()


def _oracle_merge(gathered):
    """Stand-in for rpx_topk_merge_packed on the gathered [world, Q, k, 2] records."""
    scores64 = gathered[..., 0].contiguous().view(torch.float64)
    idx = gathered[..., 1].contiguous()
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


def _toy_corpus():
    from reprover_b200.corpus import Corpus, File, Pos, Premise
    files = []
    for f, (n, imports) in enumerate([(5, []), (4, ["T/F0.lean"]), (6, ["T/F1.lean"]), (3, [])]):
        prem = [Premise(f"T/F{f}.lean", f"T.F{f}.l{j}", Pos(10 * j + 1, 0), Pos(10 * j + 5, 0), f"theorem l{j} : x{f}_{j} = y")
                for j in range(n)]
        files.append((File(f"T/F{f}.lean", prem), imports))
    return Corpus.from_files(files)


def _retr_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from reprover_b200.corpus import Pos
        from tests.test_host_cpu import _stub_retriever
        r = _stub_retriever(10**9)
        r.load_corpus(_toy_corpus())
        index = r.reindex_corpus_sharded()
        assert r.reindex_corpus_sharded() is index and len(r.encoder.calls) == 1     # fresh: not re-encoded
        states = ["⊢ a = b", "h : p\n⊢ q", "⊢ True"]
        files, poses = ["T/F2.lean", "T/F1.lean", "T/F3.lean"], [Pos(25, 0), Pos(100, 0), Pos(100, 0)]
        prem, scores = r.retrieve_batch_sharded(states, files, ["t"] * 3, poses, 3,
                                                local_topk=_oracle_local_topk, merge=_oracle_merge)
        try:
            r.retrieve_batch_sharded(states[2:], files[2:], ["t"], poses[2:], 4,      # F3 alone has 3 premises
                                     local_topk=_oracle_local_topk, merge=_oracle_merge)
            raised = False
        except ValueError:
            raised = True
        out_q.put((rank, (index.lo, index.hi), [[p.full_name for p in row] for row in prem], scores, raised))
    finally:
        dist.destroy_process_group()


def test_retriever_sharded_mode_equals_the_unsharded_walk():
    """`reindex_corpus_sharded` / `retrieve_batch_sharded` on two gloo ranks (stub encoder, oracle compute
    steps): each rank encodes only its rows, slices the accessibility bitmask to its row range, and both
    ranks return what the reference walk returns on the whole index — including the ValueError."""
    from oracle import reference_path as ref
    from reprover_b200.corpus import Context, Pos
    from tests.test_host_cpu import _stub_retriever
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_retr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    corpus = _toy_corpus()
    assert [r[1] for r in results] == [(0, 9), (9, 18)]
    full = _stub_retriever(10**9)
    E = full.encode_texts([p.serialize() for p in corpus.all_premises]).to(torch.bfloat16).float()
    states = ["⊢ a = b", "h : p\n⊢ q", "⊢ True"]
    Q = full.encode_texts(states).to(torch.bfloat16).float()
    ctxs = [Context(f, "t", pos, s) for s, f, pos in zip(states, ["T/F2.lean", "T/F1.lean", "T/F3.lean"],
                                                          [Pos(25, 0), Pos(100, 0), Pos(100, 0)])]
    want_p, want_s = ref.get_nearest_premises(corpus, E, ctxs, Q, 3)
    for _, _, names, scores, raised in results:
        assert names == [[p.full_name for p in row] for row in want_p]
        assert np.allclose(scores, want_s, atol=1e-6)
        assert raised


def test_accessible_mask_words_range_is_a_bit_slice():
    from reprover_b200.corpus import Pos
    c = _toy_corpus()
    full = np.unpackbits(c.accessible_mask_words("T/F2.lean", Pos(25, 0)).view(np.uint8), bitorder="little")[: len(c)]
    for lo, hi in [(0, 18), (0, 9), (9, 18), (3, 4), (7, 7), (1, 17)]:
        w = c.accessible_mask_words_range("T/F2.lean", Pos(25, 0), lo, hi)
        assert w.dtype == np.uint32 and len(w) == (hi - lo + 31) // 32
        bits = np.unpackbits(w.view(np.uint8), bitorder="little") if len(w) else np.zeros(0, dtype=np.uint8)
        assert bits[: hi - lo].tolist() == full[lo:hi].tolist() and not bits[hi - lo:].any()
