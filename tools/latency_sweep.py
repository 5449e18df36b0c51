"""One-sequence encode latency (device time, CUDA events) by token count: throughput tiles vs the latency path."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import synth
from reprover_b200.engine import T5EncoderEngine

dev = torch.device("cuda:0")
cfg = dict(synth.BYT5_SMALL)
eng = T5EncoderEngine(cfg, synth.random_t5_state_dict(cfg, seed=synth.SEED), dev)
rng = np.random.default_rng(0)
out = {}
for T in (32, 64, 128, 200, 256, 384, 512, 768, 1024, 2048):
    data = torch.from_numpy(rng.integers(32, 120, size=T - 1, dtype=np.uint8)).to(dev)
    offs = np.array([0, T - 1], dtype=np.int64)
    o = torch.empty(1, 1472, dtype=torch.bfloat16, device=dev)
    row = {}
    embs = {}
    for name, lat in (("throughput", 0), ("latency", 4096)):
        eng.set_latency_tokens(lat)
        for _ in range(5):
            eng.encode_packed_bytes(data, offs, 4096, o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        t0 = time.perf_counter()
        for _ in range(20):
            eng.encode_packed_bytes(data, offs, 4096, o)
        t1 = time.perf_counter()
        e1.record(); torch.cuda.synchronize()
        row[name] = e0.elapsed_time(e1) / 20
        row[name + "_cpu_submit"] = (t1 - t0) / 20 * 1e3
        embs[name] = o.float().clone()
    row["max_abs_diff"] = float((embs["throughput"] - embs["latency"]).abs().max())
    out[T] = row
    print(T, json.dumps(row), flush=True)
eng.set_latency_tokens(4096)
eng.set_profiling(True); eng.read_profile()
data = torch.from_numpy(rng.integers(32, 120, size=224, dtype=np.uint8)).to(dev)
for _ in range(10):
    eng.encode_packed_bytes(data, np.array([0, 224], dtype=np.int64), 4096, o)
prof = eng.read_profile()
print("per-class ms per call (T=225, latency path, with profiling events):", {k: v["ms"] / 10 for k, v in prof.items()})
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/latency_sweep.json").write_text(json.dumps(out, indent=1))
