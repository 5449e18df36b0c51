"""Small run over the latency-path encoder (one / several / long states) and every top-k path, meant to be run under
`compute-sanitizer --tool memcheck` (0 errors at the end of round 2)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import synth
from reprover_b200.engine import T5EncoderEngine
from reprover_b200.retrieval_ops import IndexHandle, sim_topk
dev = torch.device("cuda:0")
cfg = dict(synth.BYT5_SMALL); cfg["num_layers"] = 2
eng = T5EncoderEngine(cfg, synth.random_t5_state_dict(cfg, seed=1), dev)
eng.set_latency_tokens(4096)
for n, lo, hi in ((1, 224, 224), (3, 5, 90), (2, 300, 420), (1, 700, 700), (5, 1, 40)):
    data, offs = synth.synth_states(n, seed=n, min_len=lo, max_len=hi)
    out = eng.encode_bytes(data, offs, 2048, out_dtype=torch.float32)
    torch.cuda.synchronize()
    print("encode", n, lo, hi, float(out.norm(dim=1).mean()))
E = synth.random_unit_rows(20000, 1472, 1000, dev)
h = IndexHandle(E)
for nq in (1, 3, 64, 300):
    Q = synth.random_unit_rows(nq, 1472, 999, dev)
    s, i, c = sim_topk(Q, h, 100)[:3]
    torch.cuda.synchronize()
    print("topk", nq, int(i[0, 0]))
