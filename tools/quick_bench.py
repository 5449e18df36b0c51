"""Fast device-side timing used while tuning kernels (CUDA events; not the bench contract)."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import synth
from reprover_b200.engine import T5EncoderEngine
from reprover_b200.retrieval_ops import sim_topk

ap = argparse.ArgumentParser()
ap.add_argument("--premises", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--skip-encode", action="store_true")
ap.add_argument("--skip-retrieve", action="store_true")
ap.add_argument("--tag", default="quick")
ap.add_argument("--max-tokens", type=int, default=1 << 18)
args = ap.parse_args()
dev = torch.device("cuda:0")
out = {}
if not args.skip_encode:
    cfg = dict(synth.BYT5_SMALL)
    eng = T5EncoderEngine(cfg, synth.random_t5_state_dict(cfg, seed=synth.SEED), dev, max_tokens_per_call=args.max_tokens)
    data, offsets = synth.synth_premises(args.premises, seed=synth.SEED)
    tl = np.minimum(np.diff(offsets) + 1, 512).astype(np.float64)
    flops = float((tl * (434_110_464.0 + 18_432.0 * tl)).sum())
    d = torch.from_numpy(data.copy()).to(dev)
    o = torch.empty(args.premises, 1472, dtype=torch.bfloat16, device=dev)
    def step():
        cum = np.concatenate([[0], np.cumsum(tl)]); a = 0
        while a < args.premises:
            b = int(np.searchsorted(cum, cum[a] + eng.max_tokens_per_call, side="right")) - 1
            b = min(max(b, a + 1), args.premises)
            b0, b1 = int(offsets[a]), int(offsets[b])
            eng.encode_packed_bytes(d[b0:b1], offsets[a:b + 1] - b0, 512, o[a:b]); a = b
    for _ in range(2): step()
    eng.set_profiling(True); eng.read_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(args.reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    prof = eng.read_profile()
    out["encode"] = {"premises_per_s": args.premises / ms * 1e3, "ms": ms, "tflops": flops / ms / 1e9,
                     "kernel_ms_per_step": {k: v["ms"] / args.reps for k, v in prof.items()}}
    print(json.dumps(out["encode"]), flush=True)
    del eng
if not args.skip_retrieve:
    for (nq, n) in [(1024, 200_000), (1, 200_000), (64, 200_000)]:
        E = synth.random_unit_rows(n, 1472, 1000, dev); Q = synth.random_unit_rows(nq, 1472, 999, dev)
        for _ in range(3): sim_topk(Q, E, 100)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): sim_topk(Q, E, 100)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[f"retrieve_{nq}x{n}"] = {"ms": ms, "qps": nq / ms * 1e3, "tflops": 2.0 * nq * n * 1472 / ms / 1e9, "gbs": n * 1472 * 2 / ms / 1e6}
        print(nq, n, json.dumps(out[f"retrieve_{nq}x{n}"]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/{args.tag}.json").write_text(json.dumps(out, indent=1))
