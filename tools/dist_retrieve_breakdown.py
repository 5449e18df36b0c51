"""torchrun --nproc-per-node N tools/dist_retrieve_breakdown.py: where the time of a sharded retrieve goes
(local fused top-k / all-gather / merge), device-timed per phase with CUDA events on rank 0."""
import os, sys, json
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import synth
from reprover_b200.retrieval_ops import IndexHandle, sim_topk, topk_merge_packed
from reprover_b200.dist import sharded_topk

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
E = synth.random_unit_rows(200_000, 1472, 1000 + rank, dev); h = IndexHandle(E)
Q = synth.random_unit_rows(1024, 1472, 999, dev)
k = 100

def timed(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {}
out["local_topk"] = timed(lambda: sim_topk(Q, h, k))
out["local_topk_packed"] = timed(lambda: sim_topk(Q, h, k, want_packed=True))
packed = sim_topk(Q, h, k, want_packed=True)[-1]
if world > 1:
    gathered = torch.empty((world,) + tuple(packed.shape), dtype=torch.int64, device=dev)
    out["all_gather"] = timed(lambda: dist.all_gather_into_tensor(gathered.view(world * packed.shape[0], *packed.shape[1:]), packed))
else:
    gathered = packed.unsqueeze(0).contiguous()
out["merge"] = timed(lambda: topk_merge_packed(gathered))
out["sharded_topk"] = timed(lambda: sharded_topk(Q, h, k, row_offset=rank * 200_000))
# the same chain with events between the phases
if world > 1:
    reps = 30
    acc = [0.0, 0.0, 0.0]
    import time
    for it in range(reps + 5):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        pk = sim_topk(Q, h, k, want_packed=True, idx_offset=rank * 200_000)[-1]
        ev[1].record()
        dist.all_gather_into_tensor(gathered.view(world * pk.shape[0], *pk.shape[1:]), pk)
        ev[2].record()
        topk_merge_packed(gathered)
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 5:
            for j in range(3):
                acc[j] += ev[j].elapsed_time(ev[j + 1])
    out["chain_phases_synced_each_iter"] = [a / reps for a in acc]
    t0 = time.perf_counter()
    for _ in range(50):
        pk = sim_topk(Q, h, k, want_packed=True, idx_offset=rank * 200_000)[-1]
        dist.all_gather_into_tensor(gathered.view(world * pk.shape[0], *pk.shape[1:]), pk)
        topk_merge_packed(gathered)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out["chain_cpu_submit_ms"] = (t1 - t0) / 50 * 1e3
    out["chain_wall_ms"] = (t2 - t0) / 50 * 1e3
print(json.dumps({"rank": rank, **out}), flush=True)
if world > 1:
    dist.destroy_process_group()
