"""The drop-in surface end to end on a B200: `B200PremiseRetriever` (load_hf / load_corpus /
reindex_corpus / retrieve) and `Corpus.get_nearest_premises` against the oracle restatement of
the reference path (oracle/reference_path.py) on the same synthetic checkpoint and corpus."""
import json
import pickle

import numpy as np
import pytest
import torch

from reprover_b200 import synth
from reprover_b200.corpus import Context, Corpus, File, Pos, Premise
from reprover_b200.retriever import B200PremiseRetriever
from tests.helpers import EMB_MAX_ABS, EMB_MIN_COS, compare_embeddings, ref

pytestmark = pytest.mark.gpu
MAX_LEN = 192


def _make_corpus(n_files=4, per_file=45, seed=0):
    rng = np.random.default_rng(seed)
    files, all_lines = [], []
    for f in range(n_files):
        path = f"Synth/F{f}.lean"
        prem = []
        for j in range(per_file):
            n = int(rng.integers(10, 160))
            body = bytes(rng.choice(synth._ALPHABET, size=n).tolist()).decode()
            name = f"Synth.F{f}.lemma_{j}"
            code = f"theorem lemma_{j} : {body}"
            if f == 1 and j == 3:
                code += " <pad> sentinel </s> tail <extra_id_5>"   # forces the host-tokenised ids path
            prem.append({"full_name": name, "code": code, "start": [10 * j + 1, 0], "end": [10 * j + 5, 0]})
        imports = [f"Synth/F{i}.lean" for i in range(f) if (f - i) <= 2]   # F2 imports F0,F1; F3 imports F1,F2 (+F0 transitively)
        all_lines.append({"path": path, "imports": imports, "premises": prem})
    return all_lines


@pytest.fixture(scope="module")
def setup(tmp_path_factory, cuda_device):
    tmp = tmp_path_factory.mktemp("retr")
    cfg = synth.tiny_config(num_layers=2)
    sd = synth.random_t5_state_dict(cfg, seed=21)
    ckpt = tmp / "ckpt"
    synth.save_hf_checkpoint(str(ckpt), cfg, sd)
    jsonl = tmp / "corpus.jsonl"
    jsonl.write_text("\n".join(json.dumps(l) for l in _make_corpus()))
    retr = B200PremiseRetriever.load_hf(str(ckpt), MAX_LEN, cuda_device)
    retr.load_corpus(str(jsonl))
    assert retr.embeddings_staled and retr.corpus_embeddings is None
    retr.reindex_corpus(batch_size=32)
    torch.set_float32_matmul_precision("highest")
    enc, tok = ref.build_hf_encoder(cfg, sd), ref.build_hf_tokenizer()
    return dict(tmp=tmp, cfg=cfg, sd=sd, ckpt=ckpt, jsonl=jsonl, retr=retr, enc=enc, tok=tok)


def test_load_hf_surface(setup):
    r = setup["retr"]
    assert r.embedding_size == 1472 and r.max_seq_len == MAX_LEN and r.num_retrieved == 100
    assert r.dtype == torch.bfloat16                      # reference policy on cc >= 8 (model.py:59-64)
    assert not r.embeddings_staled and r.corpus_embeddings.shape == (len(r.corpus), 1472)
    assert r.corpus_embeddings.device.type == "cuda" and r.corpus_embeddings.dtype == torch.bfloat16


def test_reindex_matches_oracle(setup):
    r = setup["retr"]
    texts = [p.serialize() for p in r.corpus.all_premises]
    want = ref.reindex_corpus(setup["enc"], setup["tok"], texts, 32, MAX_LEN)
    max_abs, min_cos = compare_embeddings(r.corpus_embeddings, want)
    assert max_abs <= EMB_MAX_ABS + 2e-3 and min_cos >= EMB_MIN_COS, (max_abs, min_cos)   # + bf16 output rounding
    # reindex is a no-op while the index is fresh (reference :185-186)
    before = r.corpus_embeddings
    r.reindex_corpus(batch_size=7)
    assert r.corpus_embeddings is before


def test_retrieve_matches_reference_walk(setup):
    """Top-k accessible premises, best first: the engine's ranking equals the oracle's
    `get_nearest_premises` run on the engine's own embedding matrix (SURVEY §7 parity definition)."""
    r = setup["retr"]
    corpus = r.corpus
    sdat, soff = synth.synth_states(6, seed=77, min_len=12, max_len=150)
    states = [s.decode() for s in synth.split_strings(sdat, soff)]
    cases = [("Synth/F3.lean", Pos(200, 0)), ("Synth/F2.lean", Pos(31, 0)), ("Synth/F1.lean", Pos(446, 0)),
             ("Synth/F3.lean", Pos(1, 0)), ("Synth/F0.lean", Pos(446, 0)), ("Synth/F2.lean", Pos(9999, 0))]
    k = 10
    for state, (path, pos) in zip(states, cases):
        premises, scores = r.retrieve(state, path, "Synth.thm", pos, k)
        assert len(premises) == len(scores) == k
        ctx = Context(path, "Synth.thm", pos, state)
        ctx_emb = r._encode_states([state])     # the (latency-path) encoding retrieve() itself uses
        want_p, want_s = ref.get_nearest_premises(corpus, r.corpus_embeddings.cpu(), [ctx], ctx_emb.cpu(), k)
        assert [p.full_name for p in premises] == [p.full_name for p in want_p[0]]
        assert np.allclose(scores, want_s[0], atol=1e-6)
        acc = corpus.get_accessible_premises(path, pos)
        assert all(p in acc for p in premises)
        assert all(scores[i] >= scores[i + 1] for i in range(k - 1))
    # the context embedding itself agrees with the oracle encoder
    tok = ref.tokenize(setup["tok"], states, MAX_LEN)
    want = ref.encode(setup["enc"], tok.input_ids, tok.attention_mask)
    max_abs, min_cos = compare_embeddings(r.encode_texts(states), want)
    assert max_abs <= EMB_MAX_ABS + 2e-3 and min_cos >= EMB_MIN_COS


def test_retrieve_raises_when_too_few_accessible(setup):
    r = setup["retr"]
    with pytest.raises(ValueError):        # reference common.py:323-324
        r.retrieve("⊢ True", "Synth/F0.lean", "Synth.thm", Pos(26, 0), 10)   # only 3 premises end before line 26
    got, _ = r.retrieve("⊢ True", "Synth/F0.lean", "Synth.thm", Pos(26, 0), 3)
    assert sorted(p.full_name for p in got) == ["Synth.F0.lemma_0", "Synth.F0.lemma_1", "Synth.F0.lemma_2"]


def test_batched_retrieve_equals_single(setup):
    r = setup["retr"]
    sdat, soff = synth.synth_states(5, seed=78, min_len=12, max_len=100)
    states = [s.decode() for s in synth.split_strings(sdat, soff)]
    P, S = r.retrieve_batch(states, ["Synth/F3.lean"] * 5, ["t"] * 5, [Pos(300, 0)] * 5, 20)
    for i, st in enumerate(states):
        p1, s1 = r.retrieve(st, "Synth/F3.lean", "t", Pos(300, 0), 20)
        assert [p.full_name for p in p1] == [p.full_name for p in P[i]] and s1 == S[i]


def test_encode_signature_parity(setup):
    r = setup["retr"]
    texts = ["⊢ a = b", "x y z : Nat\n⊢ x + (y + z) = x + y + z"]
    tok = ref.tokenize(setup["tok"], texts, MAX_LEN)
    got = r._encode(tok.input_ids.to(r.device), tok.attention_mask.to(r.device))
    want = ref.encode(setup["enc"], tok.input_ids, tok.attention_mask)
    max_abs, min_cos = compare_embeddings(got, want)
    assert got.dtype == torch.bfloat16 and max_abs <= EMB_MAX_ABS + 2e-3 and min_cos >= EMB_MIN_COS


def test_index_roundtrip_and_cli(setup, tmp_path, cuda_device):
    from reprover_b200 import index_cli

    out = tmp_path / "index.pickle"
    index_cli.main(["--ckpt_path", str(setup["ckpt"]), "--corpus-path", str(setup["jsonl"]),
                    "--output-path", str(out), "--batch-size", "16", "--max-seq-len", str(MAX_LEN)])
    # the CLI writes the reference's own layout (common.* / lean_dojo.Pos / networkx): readable by a stock
    # checkout; here (no `common` module) it is read back through the compat loader
    with pytest.raises((ModuleNotFoundError, AttributeError)):
        pickle.loads(out.read_bytes())
    from reprover_b200.compat import load_reference_index

    indexed = load_reference_index(str(out))
    assert indexed.embeddings.dtype == torch.float32 and indexed.embeddings.device.type == "cpu"
    assert torch.equal(indexed.embeddings, setup["retr"].corpus_embeddings.float().cpu())
    assert [p.full_name for p in indexed.corpus.all_premises] == [p.full_name for p in setup["retr"].corpus.all_premises]
    native = tmp_path / "native.pickle"
    setup["retr"].save_index(str(native), reference_layout=False)
    assert torch.equal(pickle.loads(native.read_bytes()).embeddings, indexed.embeddings)
    r2 = B200PremiseRetriever.load_hf(str(setup["ckpt"]), MAX_LEN, cuda_device)
    r2.load_corpus(str(out))                                # pickled IndexedCorpus -> fresh index (model.py:81-85)
    assert not r2.embeddings_staled
    a = r2.retrieve("⊢ p ∧ q", "Synth/F3.lean", "t", Pos(120, 0), 7)
    b = setup["retr"].retrieve("⊢ p ∧ q", "Synth/F3.lean", "t", Pos(120, 0), 7)
    assert [p.full_name for p in a[0]] == [p.full_name for p in b[0]] and a[1] == b[1]
    assert r2.corpus_embeddings.device.type == "cuda" and r2.corpus_embeddings.dtype == torch.bfloat16


def test_dtype_policy_is_loud(setup, cuda_device, tmp_path):
    """The engine computes in bf16 only: asking for an fp32 / fp16 MODEL (what `dtype` means in the
    reference, retrieval/model.py:52-66) is refused, not silently down-cast; fp32 OUTPUT tensors are a
    separate option, and retrieval then keeps the caller's fp32 index untouched."""
    for bad in (torch.float32, torch.float16):
        with pytest.raises(NotImplementedError, match="bf16"):
            B200PremiseRetriever.load_hf(str(setup["ckpt"]), MAX_LEN, cuda_device, dtype=bad)
    r = B200PremiseRetriever(str(setup["ckpt"]), max_seq_len=MAX_LEN, device=cuda_device, output_dtype=torch.float32)
    e = r.encode_texts(["⊢ x = x"])
    assert e.dtype == torch.float32 and abs(float(e.norm()) - 1.0) < 1e-5
    r.load_corpus(setup["retr"].corpus)
    r.reindex_corpus(32)
    assert r.corpus_embeddings.dtype == torch.float32
    a = r.retrieve("⊢ p ∧ q", "Synth/F3.lean", "t", Pos(120, 0), 7)
    b = setup["retr"].retrieve("⊢ p ∧ q", "Synth/F3.lean", "t", Pos(120, 0), 7)
    assert r.corpus_embeddings.dtype == torch.float32          # the fp32 index the caller sees is not replaced
    assert [p.full_name for p in a[0]] == [p.full_name for p in b[0]]
    # what generation/model.py:224-226 does with a retriever: persist encoder + tokenizer next to a generator
    out = tmp_path / "saved"
    r.encoder.save_pretrained(str(out))
    r.tokenizer.save_pretrained(str(out))
    r2 = B200PremiseRetriever.load_hf(str(out), MAX_LEN, cuda_device)
    assert torch.equal(r2.encode_texts(["⊢ x = x"]), setup["retr"].encode_texts(["⊢ x = x"]))
    ids = r.tokenizer("ab", return_tensors="pt").input_ids.tolist()
    assert ids == [[100, 101, 1]]


def test_validation_and_predict_steps(setup):
    """The batched validate / predict path (reference retrieval/model.py:215-327) on top of the engine."""
    from reprover_b200.evaluation import predict_step, recall_and_mrr, validation_step

    r = setup["retr"]
    old_k = r.num_retrieved
    r.num_retrieved = 10
    try:
        sdat, soff = synth.synth_states(4, seed=91, min_len=12, max_len=100)
        states = [s.decode() for s in synth.split_strings(sdat, soff)]
        ctxs = [Context("Synth/F3.lean", "t", Pos(400, 0), s) for s in states]
        premises, _ = r.retrieve_batch(states, ["Synth/F3.lean"] * 4, ["t"] * 4, [Pos(400, 0)] * 4, 10)
        positives = [[premises[0][0], premises[0][4]], [], [premises[2][9]], [r.corpus.all_premises[-1]]]
        batch = {"context": ctxs, "all_pos_premises": positives, "url": ["u"] * 4, "commit": ["c"] * 4,
                 "file_path": ["Synth/F3.lean"] * 4, "full_name": ["t"] * 4, "start": [Pos(400, 0)] * 4,
                 "tactic_idx": list(range(4))}
        m = validation_step(r, batch)
        want_recall, want_mrr, n = recall_and_mrr(positives, premises, 10)
        assert n == 3 and m["MRR"] == pytest.approx(want_mrr) and m["Recall@10_val"] == pytest.approx(want_recall[9])
        assert m["Recall@1_val"] == pytest.approx(100.0 * (0.5 + 0.0 + 0.0) / 3)
        recs = predict_step(r, batch)
        assert len(recs) == 4 and recs[2]["tactic_idx"] == 2
        assert [p.full_name for p in recs[0]["retrieved_premises"]] == [p.full_name for p in premises[0]]
        assert len(recs[0]["scores"]) == 10
    finally:
        r.num_retrieved = old_k


def test_reindex_does_not_depend_on_the_token_budget(setup):
    """`reindex_corpus` streams premises to the engine in groups cut by the engine's token budget.
    With a budget small enough to force a dozen groups — one of them split around the premise that
    takes the host-tokenised ids path — the index must equal the one built with a single group."""
    r = setup["retr"]
    base = r.corpus_embeddings.clone()
    saved = r.encoder.max_tokens_per_call
    try:
        for budget in (MAX_LEN, 1500, 4000):
            r.encoder.max_tokens_per_call = budget
            r.embeddings_staled = True
            r.reindex_corpus(batch_size=3)
            assert not r.embeddings_staled
            assert torch.equal(r.corpus_embeddings, base), budget
    finally:
        r.encoder.max_tokens_per_call = saved
    # a lazily produced sequence and a list are the same thing to encode_texts
    texts = [p.serialize() for p in r.corpus.all_premises]
    assert torch.equal(r.encode_texts(texts), base)
    assert torch.equal(r.encode_texts(texts[40:55]), base[40:55])
