"""Small, ncu-friendly driver: one encoder call and/or one retrieve call (no timing here —
numbers printed under a profiler are never bench values)."""
import argparse, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import synth
from reprover_b200.engine import T5EncoderEngine
from reprover_b200.retrieval_ops import sim_topk

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="both", choices=["encode", "retrieve", "both"])
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--premises", type=int, default=512)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--n", type=int, default=200_000)
ap.add_argument("--seed", type=int, default=1000)
args = ap.parse_args()
dev = torch.device("cuda:0")
if args.mode in ("encode", "both"):
    cfg = dict(synth.BYT5_SMALL); cfg["num_layers"] = args.layers
    eng = T5EncoderEngine(cfg, synth.random_t5_state_dict(cfg, seed=synth.SEED), dev)
    data, offsets = synth.synth_premises(args.premises, seed=synth.SEED)
    for _ in range(args.warm + 1):
        eng.encode_bytes(data, offsets, 512)
    torch.cuda.synchronize()
if args.mode in ("retrieve", "both"):
    E = synth.random_unit_rows(args.n, 1472, args.seed, dev)
    Q = synth.random_unit_rows(args.nq, 1472, 999, dev)
    for _ in range(args.warm + 1):
        sim_topk(Q, E, 100)
    torch.cuda.synchronize()
print("done")
