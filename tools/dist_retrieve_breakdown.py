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
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
