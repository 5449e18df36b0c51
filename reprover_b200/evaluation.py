"""Batched validate / predict path around the hot path.

Host-side mirror of the retrieval half of `PremiseRetriever.validation_step` (reference
retrieval/model.py:215-268: Recall@1..num_retrieved and MRR over a batch of contexts) and
`predict_step` (:281-327: per-example prediction records).  The encode + nearest-premise search
underneath run on the GPU through `B200PremiseRetriever`; the Lightning loop, logging and data
module are out of scope.
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence, Tuple

import numpy as np

from .corpus import Context, Premise


def recall_and_mrr(all_pos_premises: Sequence[Sequence[Premise]], retrieved: Sequence[Sequence[Premise]],
                   num_retrieved: int) -> Tuple[List[float], float, int]:
    """(Recall@1..num_retrieved in percent, MRR, number of examples with >= 1 positive premise).

    Same arithmetic as the reference: examples without positives are skipped; Recall@j is the
    fraction of an example's positives among its first j retrieved premises, averaged over
    examples; MRR uses the first hit (0 if none within num_retrieved)."""
    recall: List[List[float]] = [[] for _ in range(num_retrieved)]
    mrr: List[float] = []
    n = 0
    assert len(all_pos_premises) == len(retrieved)
    for positives, premises in zip(all_pos_premises, retrieved):
        positives = set(positives)
        if not positives:
            continue
        n += 1
        first = False
        hits = 0
        for j in range(num_retrieved):
            hits += premises[j] in positives   # premises are distinct, so a running count == |intersection|
            recall[j].append(hits / len(positives))
            if premises[j] in positives and not first:
                mrr.append(1.0 / (j + 1))
                first = True
        if not first:
            mrr.append(0.0)
    if n == 0:
        return [float("nan")] * num_retrieved, float("nan"), 0
    return [100.0 * float(np.mean(r)) for r in recall], float(np.mean(mrr)), n


def validation_step(retriever, batch: Dict[str, Any]) -> Dict[str, float]:
    """Retrieve `retriever.num_retrieved` premises for `batch["context"]` (a list of `Context`) and score
    them against `batch["all_pos_premises"]`.  Returns {"Recall@k_val": ..., "MRR": ...}."""
    ctxs: List[Context] = batch["context"]
    k = retriever.num_retrieved
    premises, _ = retriever.retrieve_batch([c.state for c in ctxs], [c.path for c in ctxs],
                                           [c.theorem_full_name for c in ctxs], [c.theorem_pos for c in ctxs], k)
    recall, mrr, n = recall_and_mrr(batch["all_pos_premises"], premises, k)
    out = {f"Recall@{j + 1}_val": recall[j] for j in range(k)}
    out["MRR"] = mrr
    out["num_with_premises"] = n
    return out


_PREDICT_KEYS = ("url", "commit", "file_path", "full_name", "start", "tactic_idx")


def predict_step(retriever, batch: Dict[str, Any]) -> List[Dict[str, Any]]:
    """Prediction records in the format the reference pickles to `predictions.pickle`."""
    ctxs: List[Context] = batch["context"]
    k = retriever.num_retrieved
    premises, scores = retriever.retrieve_batch([c.state for c in ctxs], [c.path for c in ctxs],
                                                [c.theorem_full_name for c in ctxs], [c.theorem_pos for c in ctxs], k)
    records = []
    for i, ctx in enumerate(ctxs):
        rec = {key: batch[key][i] for key in _PREDICT_KEYS if key in batch}
        rec.update(context=ctx, all_pos_premises=batch["all_pos_premises"][i], retrieved_premises=premises[i],
                   scores=scores[i])
        records.append(rec)
    return records
