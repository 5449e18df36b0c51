#!/usr/bin/env python
"""bench.py — premises encoded/s (reindex) and retrieve() queries/s on B200s, next to the
reference's own CPU path.

    python bench.py --gpus 1 --steps K --warmup W            # engine arm
    python bench.py --impl reference --steps K --warmup W    # reference arm (HF CPU path)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU

Workload (BASELINE.json configs[1], "reindex 200k synthetic premises, seq_len <= 512, ByT5-small"):
a *step* is one re-index pass over one batch of `--premises-per-step` synthetic premises drawn from
the cfg2 distribution (byte length ~ U[16, 511] + EOS), random-init ByT5-small
weights (seed 3407).  Every step uses a fresh slice of the corpus; the per-step working set
(~19 KB of activations per token, > 10 GB) is far larger than L2, so no flush is needed.
`--full` makes one step the whole 200k corpus.

  value  whole-job premises/s with the premise bytes already resident in HBM, CUDA-event timed,
         max over ranks.
  e2e    the same pass through the public API (`B200PremiseRetriever.reindex_corpus` on a `Corpus`
         of Premise objects, then `.cpu()` of the index as retrieval/index.py:37 does): host
         strings -> pinned bytes -> H2D -> engine -> D2H inside the timed region.
  retrieve   extra leg, BASELINE configs[2]/[3]: 1024 states x 200k-premise (per GPU) bf16 index,
         k = 100, fused sim+top-k (+ all-gather + merge when N > 1): queries/s and roofline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from reprover_b200 import synth  # noqa: E402

D_MODEL = 1472
N_CORPUS = 200_000
MAX_SEQ_LEN = 512


def encoder_flops(token_lens: np.ndarray) -> float:
    """Algorithmic FLOPs of the encoder for sequences of the given token lengths."""
    l = token_lens.astype(np.float64)
    return float((l * (434_110_464.0 + 18_432.0 * l)).sum())


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # samples taken while the GPU was busy are the upper half of the clock distribution
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    # stdout carries exactly one JSON line: NCCL's own banner / debug lines (NCCL_DEBUG=VERSION|INFO) go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def max_over_ranks(ms: float, world: int, dev) -> float:
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def barrier(world: int):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------ engine arm
def run_engine(args) -> dict:
    from reprover_b200 import _native
    from reprover_b200.corpus import Corpus, File, Pos, Premise
    from reprover_b200.dist import sharded_topk
    from reprover_b200.engine import T5EncoderEngine
    from reprover_b200.retrieval_ops import sim_topk
    from reprover_b200.retriever import B200PremiseRetriever

    rank, world, local = dist_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    K, W = args.steps, args.warmup
    P = N_CORPUS if args.full else args.premises_per_step

    cfg = dict(synth.BYT5_SMALL)
    sd = synth.random_t5_state_dict(cfg, seed=synth.SEED)
    eng = T5EncoderEngine(cfg, sd, dev, max_tokens_per_call=args.max_tokens_per_call)

    # this rank's premises for all steps (rank r draws from seed 3407 + r: SURVEY §8d cfg4)
    n_steps = K + W
    data, offsets = synth.synth_premises(P * n_steps if not args.full else P, seed=synth.SEED + rank)
    if args.full:
        step_slices = [(0, P)] * n_steps
    else:
        step_slices = [(i * P, (i + 1) * P) for i in range(n_steps)]
    tok_lens = np.minimum(np.diff(offsets) + 1, MAX_SEQ_LEN)

    d_data = torch.from_numpy(data.copy()).to(dev)
    out = torch.empty(P, D_MODEL, dtype=torch.bfloat16, device=dev)

    def device_step(i):
        lo, hi = step_slices[i]
        cum = np.concatenate([[0], np.cumsum(tok_lens[lo:hi])])
        a = 0
        n = hi - lo
        while a < n:
            b = int(np.searchsorted(cum, cum[a] + eng.max_tokens_per_call, side="right")) - 1
            b = min(max(b, a + 1), n)
            b0, b1 = int(offsets[lo + a]), int(offsets[lo + b])
            eng.encode_packed_bytes(d_data[b0:b1], offsets[lo + a:lo + b + 1] - b0, MAX_SEQ_LEN, out[a:b])
            a = b

    # ---- device-resident throughput
    for i in range(W):
        device_step(i)
    eng.set_profiling(True)
    eng.read_profile()
    sampler = ClockSampler(local)
    barrier(world)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(W, W + K):
        device_step(i)
    e1.record()
    barrier(world)
    clocks = sampler.stop()
    ms_dev = max_over_ranks(e0.elapsed_time(e1), world, dev)
    prof = eng.read_profile()
    eng.set_profiling(False)
    timed_tokens = sum(int(tok_lens[lo:hi].sum()) for lo, hi in step_slices[W:W + K])
    timed_flops = sum(encoder_flops(tok_lens[lo:hi]) for lo, hi in step_slices[W:W + K])
    value = world * P * K / (ms_dev / 1e3)

    # dominant kernel: the FFN up-projection GEMM (58 % of the FLOPs)
    ffn = prof["ffn_up_gemm"]
    ffn_flops = 2.0 * timed_tokens * 7168 * 1472 * cfg["num_layers"]
    ffn_tf = ffn_flops / (ffn["ms"] / 1e3) / 1e12 if ffn["ms"] > 0 else 0.0
    traffic = None
    tpath = ROOT / "profiles" / "roofline_traffic.json"
    if tpath.exists():
        traffic = json.loads(tpath.read_text()).get("ffn_up_gemm_dram_bytes_per_launch")
    launches = sum(v["launches"] for v in prof.values())

    # ---- end to end through the public API (host strings in, host index out)
    e2e = None
    if not args.skip_e2e:
        ckpt = Path(args.tmp) / f"byt5_small_synth_rank{rank}"
        synth.save_hf_checkpoint(str(ckpt), cfg, sd)
        del eng
        torch.cuda.empty_cache()
        retr = B200PremiseRetriever.load_hf(str(ckpt), MAX_SEQ_LEN, dev)
        retr.encoder.max_tokens_per_call = args.max_tokens_per_call

        def make_corpus(i):
            lo, hi = step_slices[i]
            prem = []
            for j in range(lo, hi):
                code = data[offsets[j]:offsets[j + 1]].tobytes().decode()
                prem.append(Premise("Synth.lean", f"Synth.p{j}", Pos(j + 1, 0), Pos(j + 1, 1), code))
            return Corpus.from_files([(File("Synth.lean", prem), [])])

        corpora = [make_corpus(0)] * n_steps if args.full else [make_corpus(i) for i in range(n_steps)]
        h2d = d2h = 0
        host_index = None
        for i in range(W):
            retr.load_corpus(corpora[i])
            retr.reindex_corpus(batch_size=64)
            host_index = retr.corpus_embeddings.cpu()
        barrier(world)
        t0 = time.perf_counter()
        e0.record()
        for i in range(W, W + K):
            retr.load_corpus(corpora[i])
            retr.reindex_corpus(batch_size=64)
            host_index = retr.corpus_embeddings.cpu()   # D2H of the step's result
            lo, hi = step_slices[i]
            h2d += int(offsets[hi] - offsets[lo])
            d2h += host_index.numel() * host_index.element_size()
        e1.record()
        barrier(world)
        ms_e2e = max_over_ranks(e0.elapsed_time(e1), world, dev)
        wall_e2e = time.perf_counter() - t0
        e2e = {"value": world * P * K / (ms_e2e / 1e3), "unit": "premises/s", "h2d_bytes_per_step": h2d // K,
               "d2h_bytes_per_step": d2h // K, "ms_per_step": ms_e2e / K, "host_wall_s": wall_e2e}
        eng_for_retrieve = retr.encoder
    else:
        eng_for_retrieve = eng

    # ---- retrieve leg (cfg3 per GPU; cfg4 when world == 8)
    retrieve = None
    if not args.skip_retrieve:
        nq, n_idx, k = 1024, N_CORPUS, 100
        E = synth.random_unit_rows(n_idx, D_MODEL, 1000 + rank, dev)
        Q = synth.random_unit_rows(nq, D_MODEL, 999, dev)   # same queries on every rank
        Q_host = Q.cpu().pin_memory()
        # one retrieve is ~0.6 ms: a handful of repetitions would be timed while the SM clock is still
        # ramping after the host-side legs; 50 repetitions (~30 ms) after 10 warm-ups are past that
        reps = max(50, K)
        warm_r = max(10, W)

        def retrieve_dev():
            if world == 1:
                return sim_topk(Q, E, k)
            return sharded_topk(Q, E, k, row_offset=rank * n_idx)

        for _ in range(warm_r):
            retrieve_dev()
        barrier(world)
        e0.record()
        for _ in range(reps):
            retrieve_dev()
        e1.record()
        barrier(world)
        ms_r = max_over_ranks(e0.elapsed_time(e1), world, dev) / reps
        # e2e: queries from pinned host memory, results back to the host
        res_scores = torch.empty(nq, k, dtype=torch.float32).pin_memory()
        res_idx = torch.empty(nq, k, dtype=torch.int64).pin_memory()

        def retrieve_host():
            q = Q_host.to(dev, non_blocking=True)
            r = sim_topk(q, E, k) if world == 1 else sharded_topk(q, E, k, row_offset=rank * n_idx)
            res_scores.copy_(r[0], non_blocking=True)
            res_idx.copy_(r[1], non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller reads the host result here

        for _ in range(3):
            retrieve_host()   # warm-up (allocator, pinned staging)
        barrier(world)
        e0.record()
        for _ in range(reps):
            retrieve_host()
        e1.record()
        barrier(world)
        ms_re = max_over_ranks(e0.elapsed_time(e1), world, dev) / reps
        flops = 2.0 * nq * n_idx * D_MODEL
        bytes_alg = n_idx * D_MODEL * 2 + nq * D_MODEL * 2 + nq * k * 12
        t_mma = flops / (peaks["tf_burst"] * 1e12)
        t_hbm = bytes_alg / (peaks["hbm_gbs"] * 1e9)
        retrieve = {
            "metric": "retrieve queries/s", "config": {"queries": nq, "index_rows_per_gpu": n_idx, "index_rows_total": n_idx * world,
                                                      "k": k, "dtype": "bf16", "merge": "nccl all_gather + device merge" if world > 1 else "none",
                                                      "warmup": warm_r, "repetitions": reps},
            "value": nq / (ms_r / 1e3), "ms": ms_r,
            "e2e": {"value": nq / (ms_re / 1e3), "ms": ms_re, "h2d_bytes": nq * D_MODEL * 2, "d2h_bytes": nq * k * 12},
            "roofline": {"bound": "tensor", "achieved": flops / (ms_r / 1e3) / 1e12, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                         "frac": max(t_mma, t_hbm) / (ms_r / 1e3), "hbm_frac": t_hbm / (ms_r / 1e3),
                         "note": "whole retrieve (sim kernel + select/rescore [+ gather/merge]) vs max(t_MMA, t_HBM) of the sim kernel"},
        }
        if rank == 0 and world == 1 and not args.skip_cpu_baseline:
            retrieve["cpu_baseline"] = cpu_baseline_retrieve(E, Q, k)

    result = {
        "metric": "premises encoded/sec (reindex_corpus, ByT5-small, seq_len<=512)",
        "value": value, "unit": "premises/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: reindex synthetic premises, byte len~U[16,511]+EOS, ByT5-small random-init seed 3407",
                   "premises_per_step_per_gpu": P, "tokens_per_step_per_gpu": timed_tokens // K, "max_seq_len": MAX_SEQ_LEN,
                   "parallelism": f"row-sharded corpus x{world}, no data-path collective in reindex",
                   "l2": "inputs larger than L2 (fresh premises every step; ~19 KB activations/token)",
                   "max_tokens_per_call": args.max_tokens_per_call},
        "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel<6,EpiGeGLU> (FFN up-projection, 2-CTA tcgen05, 58% of FLOPs)",
                     "achieved": ffn_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": ffn_tf / peaks["tf_sustained"], "peak_source": peaks["source"] + " (sustained bf16)",
                     "traffic": traffic, "launches": ffn["launches"], "avg_launch_ms": ffn["ms"] / max(ffn["launches"], 1)},
        "encoder_roofline": {"achieved_tflops": timed_flops / (ms_dev / 1e3) / 1e12, "peak": peaks["tf_sustained"],
                             "frac": timed_flops / (ms_dev / 1e3) / 1e12 / peaks["tf_sustained"],
                             "note": "algorithmic FLOPs sum F(l_i) of SURVEY 8d over the whole step (per GPU)"},
        "kernel_ms": {k2: v["ms"] for k2, v in prof.items()},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "retrieve": retrieve,
    }
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline_encode(cfg, sd, data, offsets, n_premises=args.cpu_sample)
    if world > 1:
        torch.distributed.destroy_process_group()
    return result if rank == 0 else None


# ------------------------------------------------------------------------------------------ CPU reference
def cpu_baseline_encode(cfg, sd, data, offsets, n_premises: int, precision: str = "medium") -> dict:
    """The reference's own path on the host cores: HF T5EncoderModel fp32, reference batching
    (batch 64, pad to longest, corpus order), `torch.set_float32_matmul_precision("medium")` as
    retrieval/model.py:26 sets it.  Bounded sample; this is the oracle used as a stopwatch."""
    from oracle import reference_path as ref

    torch.set_float32_matmul_precision(precision)
    enc = ref.build_hf_encoder(cfg, sd)
    tok = ref.build_hf_tokenizer()
    texts = [s.decode() for s in synth.split_strings(data, offsets[: n_premises + 1])]
    ref.reindex_corpus(enc, tok, texts[:2], 64, MAX_SEQ_LEN)  # warm-up
    t0 = time.perf_counter()
    ref.reindex_corpus(enc, tok, texts, 64, MAX_SEQ_LEN)
    dt = time.perf_counter() - t0
    torch.set_float32_matmul_precision("highest")
    return {"value": n_premises / dt, "unit": "premises/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(),
            "kind": "port", "sample": f"first {n_premises} premises of the cfg2 corpus, batch 64 pad-to-longest, fp32 matmul precision '{precision}', {dt:.1f} s",
            "note": "oracle/reference_path.py = reference algorithm on HF T5EncoderModel (the reference modules need lean_dojo/lightning/deepspeed, absent here)"}


def cpu_baseline_retrieve(E_dev: torch.Tensor, Q_dev: torch.Tensor, k: int, n_queries: int = 64) -> dict:
    """Reference retrieve arithmetic on the host cores (common.py:307-324: fp32 matmul, full argsort,
    .tolist(), Python walk) for a bounded sample of the same queries against the same index."""
    from oracle import reference_path as ref

    torch.set_float32_matmul_precision("medium")
    E = E_dev.float().cpu()
    Q = Q_dev[:n_queries].float().cpu()
    ref.nearest_unfiltered_verbatim(E, Q[:1], k)  # warm-up
    t0 = time.perf_counter()
    ref.nearest_unfiltered_verbatim(E, Q, k)
    dt = time.perf_counter() - t0
    torch.set_float32_matmul_precision("highest")
    return {"value": n_queries / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{n_queries} of the 1024 states against the full {E.shape[0]}-row fp32 index, k={k}, {dt:.1f} s"}


def run_reference(args) -> dict:
    """--impl reference: the reference's CPU implementation of the path, all host threads,
    each step a bounded sample of the same workload."""
    rank, world, _ = dist_env()
    if rank != 0:
        return None
    from oracle import reference_path as ref

    K, W = args.steps, args.warmup
    S = args.reference_premises_per_step
    cfg = dict(synth.BYT5_SMALL)
    sd = synth.random_t5_state_dict(cfg, seed=synth.SEED)
    torch.set_float32_matmul_precision("medium")  # retrieval/model.py:26
    enc = ref.build_hf_encoder(cfg, sd)
    tok = ref.build_hf_tokenizer()
    data, offsets = synth.synth_premises(S * (K + W), seed=synth.SEED)
    texts = [s.decode() for s in synth.split_strings(data, offsets)]
    for i in range(W):
        ref.reindex_corpus(enc, tok, texts[i * S:(i + 1) * S], 64, MAX_SEQ_LEN)
    t0 = time.perf_counter()
    for i in range(W, W + K):
        ref.reindex_corpus(enc, tok, texts[i * S:(i + 1) * S], 64, MAX_SEQ_LEN)
    dt = time.perf_counter() - t0
    v = S * K / dt
    return {
        "impl": "reference",
        "metric": "premises encoded/sec (reindex_corpus, ByT5-small, seq_len<=512)",
        "value": v, "unit": "premises/s", "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (matmul precision 'medium')", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: reindex synthetic premises, byte len~U[16,511]+EOS, ByT5-small random-init seed 3407",
                   "premises_per_step": S, "max_seq_len": MAX_SEQ_LEN, "batching": "reference: batch 64, pad to longest, corpus order"},
        "cpu_baseline": {"value": v, "unit": "premises/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
                         "sample": f"{S} premises per step x {K} steps of the cfg2 corpus ({dt:.1f} s)"},
        "e2e": {"value": v, "unit": "premises/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--premises-per-step", type=int, default=8192)
    ap.add_argument("--full", action="store_true", help="one step = the whole 200k-premise corpus")
    ap.add_argument("--max-tokens-per-call", type=int, default=1 << 18)
    ap.add_argument("--cpu-sample", type=int, default=64, help="premises timed on the CPU baseline")
    ap.add_argument("--reference-premises-per-step", type=int, default=8)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-retrieve", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--tmp", default="/tmp/rpx_bench")
    args = ap.parse_args()
    if args.impl == "engine" and args.warmup < 3:
        print(f"[bench] --warmup {args.warmup} raised to 3 (timing rules: at least 3 warm-up steps)", file=sys.stderr)
        args.warmup = 3
    res = run_reference(args) if args.impl == "reference" else run_engine(args)
    if res is not None:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
