"""Single-query retrieve() latency (the prover's actual call pattern: one state per call)."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from reprover_b200 import synth
from reprover_b200.corpus import Corpus, File, Pos, Premise
from reprover_b200.retriever import B200PremiseRetriever

dev = torch.device("cuda:0")
cfg = dict(synth.BYT5_SMALL); sd = synth.random_t5_state_dict(cfg, seed=synth.SEED)
synth.save_hf_checkpoint("/tmp/rpx_lat_ckpt", cfg, sd)
r = B200PremiseRetriever.load_hf("/tmp/rpx_lat_ckpt", 2048, dev)
N, n_files = 200_000, 2000
files = []
k = 0
for f in range(n_files):
    prem = [Premise(f"F{f}.lean", f"F{f}.p{j}", Pos(j + 1, 0), Pos(j + 1, 5), f"theorem p{j} : True := trivial") for j in range(N // n_files)]
    files.append((File(f"F{f}.lean", prem), [f"F{f-1}.lean"] if f else []))
corpus = Corpus.from_files(files)
r.load_corpus(corpus)
r.corpus_embeddings = synth.random_unit_rows(N, 1472, 7, dev)   # synthetic index: latency does not depend on values
r.embeddings_staled = False
sdat, soff = synth.synth_states(64, seed=5, min_len=50, max_len=400)
states = [s.decode() for s in synth.split_strings(sdat, soff)]
for s in states[:5]:
    r.retrieve(s, f"F{n_files-1}.lean", "t", Pos(50, 0), 100)
torch.cuda.synchronize()
lat = []
for s in states:
    t0 = time.perf_counter()
    prem, sc = r.retrieve(s, f"F{n_files-1}.lean", "t", Pos(50, 0), 100)
    lat.append((time.perf_counter() - t0) * 1e3)
t0 = time.perf_counter(); m = corpus.accessible_mask_words(f"F{n_files-1}.lean", Pos(50, 0)); t_mask = (time.perf_counter() - t0) * 1e3
out = {"retrieve_single_ms": {"median": float(np.median(lat)), "p90": float(np.percentile(lat, 90)), "min": float(min(lat))},
       "host_access_mask_ms": t_mask, "index_rows": N, "k": 100}
print(json.dumps(out))
Path("gpurun_out/latency_single.json").write_text(json.dumps(out))
