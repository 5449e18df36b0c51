"""The engine against what the REFERENCE'S OWN CODE returned (not against the oracle restatement).

`tests/golden/reference_retriever_cfg1.{npz,json}` were produced by importing the reference's
`retrieval/model.py` + `common.py` unmodified and running `PremiseRetriever.load_hf(..., "cpu")`,
`load_corpus`, `reindex_corpus(3)`, `_encode`, `retrieve`, `predict_step` on a synthetic ByT5-small
checkpoint (generator: tests/golden/make_reference_retriever_golden.py).  Here the same checkpoint and
corpus go through `B200PremiseRetriever` on the GPU: embeddings within the stated tolerance of the
reference's fp32 ones, the same premises in the same order wherever the reference's own score gaps
exceed the tolerance (all cases below but one third place, which is left out), scores within 5e-3.
(File name: runs last, after the oracle-based parity tests.)"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from reprover_b200 import synth
from reprover_b200.corpus import Pos
from reprover_b200.retriever import B200PremiseRetriever
from tests.helpers import EMB_MAX_ABS, EMB_MIN_COS, compare_embeddings

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
SCORE_ATOL = 5e-3      # bf16 embeddings on the engine side vs the reference's fp32 CPU run


@pytest.fixture(scope="module")
def gold_setup(tmp_path_factory, cuda_device):
    meta = json.loads((GOLD / "reference_retriever_cfg1.json").read_text())
    g = np.load(GOLD / "reference_retriever_cfg1.npz")
    tmp = tmp_path_factory.mktemp("refgold")
    cfg = dict(synth.BYT5_SMALL)
    sd = synth.random_t5_state_dict(cfg, seed=meta["weight_seed"])
    ckpt = tmp / "ckpt"
    synth.save_hf_checkpoint(str(ckpt), cfg, sd)
    jsonl = tmp / "corpus.jsonl"
    jsonl.write_text("\n".join(json.dumps(l) for l in meta["corpus_lines"]))
    retr = B200PremiseRetriever.load_hf(str(ckpt), meta["max_seq_len"], cuda_device)
    retr.load_corpus(str(jsonl))
    retr.reindex_corpus(batch_size=meta["reindex_batch_size"])
    return dict(meta=meta, g=g, retr=retr)


def test_reindex_matches_the_reference_retriever(gold_setup):
    r, g = gold_setup["retr"], gold_setup["g"]
    want = torch.from_numpy(g["corpus_embeddings"])
    assert r.corpus_embeddings.shape == want.shape and not r.embeddings_staled
    max_abs, min_cos = compare_embeddings(r.corpus_embeddings, want)
    assert max_abs <= EMB_MAX_ABS + 2e-3 and min_cos >= EMB_MIN_COS, (max_abs, min_cos)   # + bf16 output rounding


def test_encode_matches_the_reference_retriever(gold_setup):
    r, g = gold_setup["retr"], gold_setup["g"]
    ids = torch.from_numpy(g["state_input_ids"]).to(r.device)
    mask = torch.from_numpy(g["state_attention_mask"]).to(r.device)
    got = r._encode(ids, mask)
    max_abs, min_cos = compare_embeddings(got, torch.from_numpy(g["state_embeddings"]))
    assert max_abs <= EMB_MAX_ABS + 2e-3 and min_cos >= EMB_MIN_COS, (max_abs, min_cos)
    # the raw-bytes path (tokenisation on the device) gives the same rows
    via_text = r.encode_texts(gold_setup["meta"]["states"])
    max_abs2, _ = compare_embeddings(via_text, got)
    assert max_abs2 <= 1e-6 + 2 ** -8 * float(got.float().abs().max())   # same values up to bf16 output rounding


def test_retrieve_matches_the_reference_retriever(gold_setup):
    r, meta = gold_setup["retr"], gold_setup["meta"]
    for q in meta["queries"]:
        premises, scores = r.retrieve(q["state"], q["path"], "Gold.target", Pos(*q["pos"]), q["k"])
        assert [[p.path, p.full_name] for p in premises] == q["retrieved"]
        assert np.allclose(scores, q["scores"], atol=SCORE_ATOL), (scores, q["scores"])
    few = meta["too_few_accessible"]
    with pytest.raises(ValueError):
        r.retrieve(meta["states"][few["state"]], few["path"], "Gold.target", Pos(*few["pos"]), few["k"])


def test_batched_predictions_match_the_reference_retriever(gold_setup):
    r, meta = gold_setup["retr"], gold_setup["meta"]
    val = meta["validation"]
    ctxs = val["contexts"]
    premises, scores = r.retrieve_batch([c["state"] for c in ctxs], [c["path"] for c in ctxs],
                                        [c["theorem_full_name"] for c in ctxs], [Pos(*c["pos"]) for c in ctxs],
                                        val["num_retrieved"])
    for i, want in enumerate(val["predictions"]):
        got_names = [p.full_name for p in premises[i]]
        want_scores = want["scores"]
        # compare a place only where the reference's own margin to the next candidate is decisive
        decisive = len(got_names) if i < 2 else 2      # third place of the last context: 0.6957 vs 0.6928
        assert got_names[:decisive] == want["retrieved_premises"][:decisive], (i, got_names, want["retrieved_premises"])
        assert np.allclose(scores[i][:decisive], want_scores[:decisive], atol=SCORE_ATOL)


def test_sharded_mode_on_one_rank_equals_the_unsharded_retriever(gold_setup):
    """`reindex_corpus_sharded` / `retrieve_batch_sharded` with a world of one (no process group): the
    row range is the whole corpus, the bitmask slice is the whole bitmask, the merge has one part — the
    answer must be the unsharded one (the two-rank plumbing is tested with gloo in tests/test_dist_cpu.py
    and with NCCL in tests/test_dist_gpu.py)."""
    r, meta = gold_setup["retr"], gold_setup["meta"]
    ctxs = meta["validation"]["contexts"]
    args = ([c["state"] for c in ctxs], [c["path"] for c in ctxs], [c["theorem_full_name"] for c in ctxs],
            [Pos(*c["pos"]) for c in ctxs], 3)
    want_p, want_s = r.retrieve_batch(*args)
    index = r.reindex_corpus_sharded()
    assert (index.lo, index.hi) == (0, len(r.corpus)) and index.embeddings.dtype == torch.bfloat16
    assert torch.equal(index.embeddings, r.corpus_embeddings.to(torch.bfloat16))
    got_p, got_s = r.retrieve_batch_sharded(*args)
    assert [[p.full_name for p in row] for row in got_p] == [[p.full_name for p in row] for row in want_p]
    assert got_s == want_s
    with pytest.raises(ValueError):
        r.retrieve_batch_sharded(args[0][2:], args[1][2:], args[2][2:], [Pos(7, 0)], 2)
