"""cProfile of retrieve() host to host (one state per call) on a 200k-premise corpus: where the host time goes."""
import cProfile, io, pstats, sys, time
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
for f in range(n_files):
    prem = [Premise(f"F{f}.lean", f"F{f}.p{j}", Pos(j + 1, 0), Pos(j + 1, 5), f"theorem p{j} : True := trivial") for j in range(N // n_files)]
    files.append((File(f"F{f}.lean", prem), [f"F{f-1}.lean"] if f else []))
r.load_corpus(Corpus.from_files(files))
r.corpus_embeddings = synth.random_unit_rows(N, 1472, 7, dev)
r.embeddings_staled = False
sdat, soff = synth.synth_states(300, seed=5, min_len=50, max_len=400)
states = [s.decode() for s in synth.split_strings(sdat, soff)]
where = (f"F{n_files-1}.lean", "t", Pos(50, 0))
for s in states[:20]:
    r.retrieve(s, *where, 100)
torch.cuda.synchronize()
lat = []
for s in states[20:120]:
    t0 = time.perf_counter(); r.retrieve(s, *where, 100); lat.append((time.perf_counter() - t0) * 1e3)
print("retrieve() ms: median %.3f  p10 %.3f  p90 %.3f" % (np.median(lat), np.percentile(lat, 10), np.percentile(lat, 90)))
pr = cProfile.Profile(); pr.enable()
for s in states[120:]:
    r.retrieve(s, *where, 100)
pr.disable()
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(28); print(out.getvalue()[:6000])
