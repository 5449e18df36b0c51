"""In-kernel timeline of one latency-path encode (debug stamps of the 1-CTA GEMM kernel, CTA 0)."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import _native, synth
from reprover_b200.engine import T5EncoderEngine

T = int(sys.argv[1]) if len(sys.argv) > 1 else 225
dev = torch.device("cuda:0"); cfg = dict(synth.BYT5_SMALL)
if len(sys.argv) > 2:
    cfg["num_layers"] = int(sys.argv[2])   # few layers: all weights stay in L2 between encodes
eng = T5EncoderEngine(cfg, synth.random_t5_state_dict(cfg, seed=synth.SEED), dev); eng.set_latency_tokens(4096)
d = torch.from_numpy(np.random.default_rng(0).integers(32, 120, size=T - 1, dtype=np.uint8)).to(dev)
o = torch.empty(1, 1472, dtype=torch.bfloat16, device=dev)
offs = np.array([0, T - 1], dtype=np.int64)
for _ in range(5):
    eng.encode_packed_bytes(d, offs, 4096, o)
torch.cuda.synchronize()
lib = _native.load()
n = min(48, 4 * cfg["num_layers"])
buf = torch.zeros(n, 8, dtype=torch.int64, device=dev)
lib.rpx_debug_set_timeline(buf.data_ptr(), n)
eng.encode_packed_bytes(d, offs, 4096, o)
torch.cuda.synchronize()
lib.rpx_debug_set_timeline(None, 0)
t = buf.cpu().numpy().astype(np.int64)
t0 = t[0, 0]
names = ["qkv", "oproj", "ffn_up", "ffn_down"]
print("slot kernel   entry  | +prologue +wait(prod) +1st stage +last mma +acc seen +epi done +exit | next entry gap")
for i in range(n):
    if t[i, 0] == 0: break
    r = t[i]
    rel = [(x - r[0]) / 1e3 if x else float("nan") for x in r]
    gap = (t[i + 1, 0] - r[7]) / 1e3 if i + 1 < n and t[i + 1, 0] else float("nan")
    print(f"{i:3d} {names[i % 4]:8s} {(r[0]-t0)/1e3:8.2f} | " + " ".join(f"{x:8.2f}" for x in rel[1:]) + f" | {gap:8.2f}")
