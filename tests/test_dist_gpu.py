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


def _cli_worker(rank, world, port, ckpt, jsonl, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from reprover_b200 import index_cli

    index_cli.main(["--ckpt_path", ckpt, "--corpus-path", jsonl, "--output-path", out, "--max-seq-len", "256"])


def test_index_cli_under_torchrun_equals_single_gpu(tmp_path):
    """`python -m reprover_b200.index_cli` launched one process per GPU: every rank encodes its rows, rank 0
    writes ONE reference-layout index — equal (bit for bit) to the single-GPU index."""
    import json

    from reprover_b200 import index_cli
    from reprover_b200.compat import load_reference_index

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    cfg = synth.tiny_config(num_layers=2)
    ckpt = tmp_path / "ckpt"
    synth.save_hf_checkpoint(str(ckpt), cfg, synth.random_t5_state_dict(cfg, seed=3))
    data, offsets = synth.synth_premises(301, seed=9, min_len=8, max_len=120)
    bodies = [s.decode() for s in synth.split_strings(data, offsets)]
    lines, k = [], 0
    for f in range(7):
        prem = []
        for j in range(43):
            prem.append({"full_name": f"T.F{f}.l{j}", "code": f"theorem l{j} : {bodies[k]}", "start": [5 * j + 1, 0], "end": [5 * j + 3, 0]})
            k += 1
        lines.append({"path": f"T/F{f}.lean", "imports": [f"T/F{f - 1}.lean"] if f else [], "premises": prem})
    jsonl = tmp_path / "corpus.jsonl"
    jsonl.write_text("\n".join(json.dumps(l) for l in lines))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_n = tmp_path / "sharded.pickle"
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_cli_worker, args=(r, world, port, str(ckpt), str(jsonl), str(out_n))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    out_1 = tmp_path / "single.pickle"
    index_cli.main(["--ckpt_path", str(ckpt), "--corpus-path", str(jsonl), "--output-path", str(out_1), "--max-seq-len", "256"])
    a, b = load_reference_index(str(out_n)), load_reference_index(str(out_1))
    assert torch.equal(a.embeddings, b.embeddings) and a.embeddings.shape == (301, cfg["d_model"])
    assert [p.full_name for p in a.corpus.all_premises] == [p.full_name for p in b.corpus.all_premises]
