"""ORACLE — test infrastructure only.  Never imported by the product path
(`reprover_b200/`); only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may use it.

CPU restatement of the reference's premise-retrieval hot path, function by function:

    encode()                 <- PremiseRetriever._encode           retrieval/model.py:92-114
    tokenize()               <- the tokenizer call                 retrieval/model.py:199-205, 351-357
    reindex_corpus()         <- PremiseRetriever.reindex_corpus    retrieval/model.py:183-210
    get_nearest_premises()   <- Corpus.get_nearest_premises        common.py:299-326
    retrieve()               <- PremiseRetriever.retrieve          retrieval/model.py:338-375

Pinning.  The reference has no tests or golden vectors for this path.  This
restatement is therefore pinned against OUTPUTS OF THE REFERENCE'S OWN CODE, generated in the build
container and committed under `tests/golden/` together with the generating scripts:

  * `make_reference_retriever_golden.py` imports `/root/reference/retrieval/model.py` + `common.py`
    unmodified (the three missing packages — lean_dojo, pytorch_lightning, deepspeed — are stubbed;
    the stubs contribute `Pos` and a LightningModule base, nothing else) and records
    `PremiseRetriever.reindex_corpus / _encode / retrieve` on a synthetic ByT5-small checkpoint;
  * `make_reference_goldens.py` records the host data model (`Premise.serialize`, `Corpus`,
    accessibility, `get_nearest_premises`).

`tests/test_oracle_cpu.py` replays both against this module (embeddings to 5e-6, identical premises /
order / scores / ValueError).  The arithmetic itself is third-party code that is installed and used
here directly — `transformers.T5EncoderModel` / `ByT5Tokenizer` (transformers 5.5.0; the reference pins
no version) and `torch` (2.11) — and is additionally pinned by `tests/golden/make_golden.py`
(bucket table, tokenizer probes, BASELINE config-1 embeddings).

Deviation that is a deterministic refinement, not a change: ranking uses a stable
descending sort on fp64 scores (ties -> lower index first) where the reference's
`argsort(descending=True)` leaves the tie order unspecified (common.py:308).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


def build_hf_encoder(cfg: Dict, state_dict: Dict[str, torch.Tensor]):
    """HF `T5EncoderModel` (what `AutoModelForTextEncoding` resolves to for a T5 config,
    retrieval/model.py:45) holding exactly `state_dict`, fp32, eval mode, on the CPU."""
    from transformers import T5Config, T5EncoderModel

    keys = ("vocab_size", "d_model", "d_kv", "d_ff", "num_layers", "num_decoder_layers", "num_heads",
            "relative_attention_num_buckets", "relative_attention_max_distance", "dropout_rate",
            "layer_norm_epsilon", "feed_forward_proj", "tie_word_embeddings", "pad_token_id", "eos_token_id")
    hf_cfg = T5Config(**{k: cfg[k] for k in keys if k in cfg})
    model = T5EncoderModel(hf_cfg)
    sd = {k: v.clone() for k, v in state_dict.items()}
    sd.setdefault("encoder.embed_tokens.weight", sd["shared.weight"])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, f"unexpected keys: {unexpected}"
    assert all("embed_tokens" in k or "shared" in k for k in missing), f"missing keys: {missing}"
    return model.float().eval()


def build_hf_tokenizer():
    """The slow `ByT5Tokenizer` `AutoTokenizer` yields for byt5 checkpoints (needs no vocab file)."""
    from transformers import ByT5Tokenizer

    return ByT5Tokenizer()


def tokenize(tokenizer, texts: Sequence[str], max_seq_len: int):
    """retrieval/model.py:199-205: padding="longest", truncation to max_seq_len, return_tensors="pt"."""
    return tokenizer(list(texts), padding="longest", max_length=max_seq_len, truncation=True, return_tensors="pt")


@torch.no_grad()
def encode(encoder, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
    """retrieval/model.py:92-114 — encoder, masked average, L2 normalise."""
    hidden_states = encoder(input_ids=input_ids, attention_mask=attention_mask, return_dict=True).last_hidden_state
    lens = attention_mask.sum(dim=1)
    features = (hidden_states * attention_mask.unsqueeze(2)).sum(dim=1) / lens.unsqueeze(1)
    return torch.nn.functional.normalize(features, dim=1)


@torch.no_grad()
def reindex_corpus(encoder, tokenizer, texts: Sequence[str], batch_size: int, max_seq_len: int) -> torch.Tensor:
    """retrieval/model.py:183-210 — consecutive batches, each padded to its own longest."""
    out = torch.zeros(len(texts), encoder.config.hidden_size, dtype=encoder.dtype)
    for i in range(0, len(texts), batch_size):
        tok = tokenize(tokenizer, texts[i:i + batch_size], max_seq_len)
        out[i:i + batch_size] = encode(encoder, tok.input_ids, tok.attention_mask)
    return out


def rank_scores(similarities: np.ndarray) -> np.ndarray:
    """common.py:308 `argsort(dim=1, descending=True)` made deterministic: stable, ties by index."""
    return np.argsort(-similarities, axis=1, kind="stable")


def similarities_fp64(q: torch.Tensor, e: torch.Tensor) -> np.ndarray:
    """common.py:307 `batch_context_emb @ premise_embeddings.t()` evaluated in fp64 on the
    values the engine holds (bf16 operands are exact in fp64)."""
    return (q.double() @ e.double().t()).cpu().numpy()


def get_nearest_premises(corpus, premise_embeddings: torch.Tensor, batch_context, batch_context_emb: torch.Tensor,
                         k: int) -> Tuple[List[list], List[List[float]]]:
    """common.py:299-326: rank all premises per context, keep the first k accessible ones;
    ValueError when fewer than k are accessible."""
    sims = similarities_fp64(batch_context_emb, premise_embeddings)
    order = rank_scores(sims)
    results: List[list] = [[] for _ in batch_context]
    scores: List[List[float]] = [[] for _ in batch_context]
    for j, (ctx, idxs) in enumerate(zip(batch_context, order)):
        accessible = corpus.get_accessible_premises(ctx.path, ctx.theorem_pos)
        for i in idxs:
            p = corpus.all_premises[i]
            if p in accessible:
                results[j].append(p)
                scores[j].append(float(sims[j, i]))
                if len(results[j]) >= k:
                    break
        else:
            raise ValueError
    return results, scores


def nearest_unfiltered_verbatim(premise_embeddings: torch.Tensor, batch_context_emb: torch.Tensor, k: int):
    """common.py:307-324 with the reference's own cost profile (used as the CPU stopwatch of the
    retrieve leg, not for parity): fp32 `@`, full `argsort(descending=True)` of all N scores per
    query, `.tolist()` of the whole order, Python walk taking the first k, `.item()` per score.
    Every premise is accessible here (BASELINE config 3 has no accessibility filter)."""
    similarities = batch_context_emb @ premise_embeddings.t()
    idxs_batch = similarities.argsort(dim=1, descending=True).tolist()
    results = [[] for _ in idxs_batch]
    scores = [[] for _ in idxs_batch]
    for j, idxs in enumerate(idxs_batch):
        for i in idxs:
            results[j].append(i)
            scores[j].append(similarities[j, i].item())
            if len(results[j]) >= k:
                break
    return results, scores


def topk_plain(q: torch.Tensor, e: torch.Tensor, k: int, mask: Optional[np.ndarray] = None):
    """Top-k without a corpus object: (indices [Q,k] int64, fp64 scores [Q,k]); `mask` [Q,N] bool."""
    sims = similarities_fp64(q, e)
    if mask is not None:
        sims = np.where(mask, sims, -np.inf)
    order = rank_scores(sims)[:, :k]
    return order, np.take_along_axis(sims, order, axis=1)


@torch.no_grad()
def retrieve(encoder, tokenizer, corpus, corpus_embeddings: torch.Tensor, ctx, k: int, max_seq_len: int):
    """retrieval/model.py:338-375 for one context (index assumed fresh)."""
    tok = tokenize(tokenizer, [ctx.serialize()], max_seq_len)
    context_emb = encode(encoder, tok.input_ids, tok.attention_mask)
    premises, scores = get_nearest_premises(corpus, corpus_embeddings.to(context_emb.dtype), [ctx], context_emb, k)
    assert len(premises) == len(scores) == 1
    return premises[0], scores[0]
