"""Multi-GPU (NCCL) sharded retrieval == single-GPU retrieval on the concatenated index.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reprover_b200 import synth

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from reprover_b200.dist import ShardedIndex
        from reprover_b200.retrieval_ops import sim_topk

        n, d, k = 40_001, 1472, 100
        E = synth.random_unit_rows(n, d, seed=5, device=dev)
        Q = synth.random_unit_rows(130, d, seed=6, device=dev)
        index = ShardedIndex(n)
        index.set_embeddings(E[index.lo:index.hi].contiguous())
        s32, idx, cnt, s64 = index.topk(Q, k)
        one = sim_topk(Q, E, k, want_scores64=True)
        ok = bool(torch.equal(idx, one[1]) and torch.equal(s64, one[3]) and torch.equal(s32, one[0]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_nccl_sharded_topk_equals_single_gpu():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
