import sys, torch, json
sys.path.insert(0, ".")
from reprover_b200 import synth
from reprover_b200.retrieval_ops import IndexHandle, sim_topk
dev = torch.device("cuda:0")
Q = synth.random_unit_rows(1024, 1472, 999, dev)
for seed in (1000, 1001, 1002, 1007):
    E = synth.random_unit_rows(200_000, 1472, seed, dev); h = IndexHandle(E)
    for _ in range(10): sim_topk(Q, h, 100)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50): sim_topk(Q, h, 100)
    e1.record(); torch.cuda.synchronize()
    print(seed, e0.elapsed_time(e1) / 50, h.stats())
