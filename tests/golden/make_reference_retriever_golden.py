"""Golden vectors from the REFERENCE'S OWN `PremiseRetriever` (retrieval/model.py), run on the CPU.

    python tests/golden/make_reference_retriever_golden.py     # needs /root/reference (this container only)

Imports `/root/reference/retrieval/model.py` and `/root/reference/common.py` unmodified, builds the
reference retriever with `PremiseRetriever.load_hf(<synthetic ByT5-small checkpoint>, 512, "cpu")`,
attaches a small synthetic corpus (`load_corpus(jsonl)`), and records what the reference's own
`reindex_corpus(batch_size)`, `_encode(ids, mask)`, `retrieve(state, file, theorem, pos, k)`,
`validation_step` / `predict_step` return, plus the pickled `IndexedCorpus` that `retrieval/index.py:37-40`
writes (`reference_indexed_corpus.pickle`: classes `common.*`, `lean_dojo.Pos`, a networkx graph).
`tests/test_oracle_cpu.py` replays the committed fixture against `oracle/reference_path.py`, which the
GPU parity tests in turn use as their checker.

Three packages the reference imports are not installed here (lean_dojo, pytorch_lightning, deepspeed):
they are replaced by stubs.  What the exercised code needs from them is stated here in full, because
it is the only code on this path that is not the reference's (or HF's) own:
  * `lean_dojo.Pos` — 2-int ordered, hashable dataclass (line_nb, column_nb);
  * `pytorch_lightning.LightningModule` — a `torch.nn.Module` with `save_hyperparameters()` (no-op),
    `.device` / `.dtype` (of the first parameter), `.trainer` raising RuntimeError when the module is not
    attached to a trainer (Lightning's behaviour; it makes `cpu_checkpointing_enabled` return False) and
    `.log(name, value, **kw)` recording the metric;
  * `DeepSpeedStrategy`, `FusedAdam`, ... — names only.
The module sets `torch.set_float32_matmul_precision("medium")` at import (retrieval/model.py:26), which on
AMX hosts rounds fp32 matmuls through bf16; the fixture is generated under "highest" so that it pins the
fp32 arithmetic itself, not a host-dependent shortcut (the bench's CPU baseline keeps "medium").
"""
import json
import sys
import tempfile
import types
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from reprover_b200 import synth  # noqa: E402  (synthetic checkpoint + premise strings only)

OUT_NPZ = HERE / "reference_retriever_cfg1.npz"
OUT_JSON = HERE / "reference_retriever_cfg1.json"
OUT_PICKLE = HERE / "reference_indexed_corpus.pickle"


def _stub_modules():
    @dataclass(eq=True, unsafe_hash=True)
    class Pos:
        line_nb: int
        column_nb: int

        def __iter__(self):
            yield self.line_nb
            yield self.column_nb

        def __lt__(self, other):
            return (self.line_nb, self.column_nb) < (other.line_nb, other.column_nb)

        def __le__(self, other):
            return (self.line_nb, self.column_nb) <= (other.line_nb, other.column_nb)

    Pos.__module__, Pos.__qualname__ = "lean_dojo", "Pos"   # pickles as lean_dojo.Pos, like the real class

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        @property
        def device(self):
            return next(self.parameters()).device

        @property
        def dtype(self):
            return next(self.parameters()).dtype

        @property
        def trainer(self):
            raise RuntimeError("not attached to a Trainer")

        def log(self, name, value, **kw):   # Lightning's metric sink: record what the step reports
            self.__dict__.setdefault("logged", {})[name] = {"value": float(value), "batch_size": kw.get("batch_size")}

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    inert = type("Inert", (), {})
    mod("lean_dojo", Pos=Pos)
    mod("pytorch_lightning", LightningModule=LightningModule, Trainer=inert)
    mod("pytorch_lightning.utilities")
    mod("pytorch_lightning.utilities.deepspeed", convert_zero_checkpoint_to_fp32_state_dict=lambda *a, **k: None)
    mod("pytorch_lightning.strategies")
    mod("pytorch_lightning.strategies.deepspeed", DeepSpeedStrategy=inert)
    mod("deepspeed")
    mod("deepspeed.ops")
    mod("deepspeed.ops.adam", FusedAdam=inert, DeepSpeedCPUAdam=inert)
    try:
        import loguru  # noqa: F401
    except ImportError:
        mod("loguru", logger=types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None,
                                                  debug=lambda *a, **k: None))
    return Pos


def corpus_lines():
    """Two files (the second imports the first), premises = BASELINE config-1 style byte strings wrapped
    as Lean declarations so that `Premise.serialize` inserts its marks."""
    data, offsets = synth.synth_premises(8, seed=synth.SEED)
    bodies = [s.decode() for s in synth.split_strings(data, offsets)]
    lines = []
    for f, idxs in enumerate(([0, 1, 2, 3, 4], [5, 6, 7])):
        prem = []
        for j, i in enumerate(idxs):
            name = f"Gold.F{f}.lemma_{j}"
            prem.append({"full_name": name, "code": f"theorem lemma_{j} : {bodies[i]}",
                         "start": [10 * j + 1, 0], "end": [10 * j + 6, 0]})
        lines.append({"path": f"Gold/F{f}.lean", "imports": ["Gold/F0.lean"] if f == 1 else [], "premises": prem})
    return lines


def main():
    Pos = _stub_modules()
    sys.path.insert(0, "/root/reference")
    import common as refc  # noqa: F401  (the reference, unmodified)
    from retrieval.model import PremiseRetriever  # the reference, unmodified

    torch.set_float32_matmul_precision("highest")
    cfg = dict(synth.BYT5_SMALL)
    sd = synth.random_t5_state_dict(cfg, seed=synth.SEED)
    lines = corpus_lines()
    sdata, soff = synth.synth_states(2, seed=synth.SEED + 1)
    states = [s.decode() for s in synth.split_strings(sdata, soff)]
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = Path(tmp) / "ckpt"
        synth.save_hf_checkpoint(str(ckpt), cfg, sd)
        # the tokenizer files the reference's AutoTokenizer.from_pretrained(model_name) expects
        from transformers import ByT5Tokenizer
        ByT5Tokenizer().save_pretrained(str(ckpt))
        jsonl = Path(tmp) / "corpus.jsonl"
        jsonl.write_text("\n".join(json.dumps(l) for l in lines))

        retr = PremiseRetriever.load_hf(str(ckpt), 512, "cpu")
        assert retr.dtype == torch.float32 and retr.embeddings_staled
        retr.load_corpus(str(jsonl))
        retr.reindex_corpus(batch_size=3)                       # reference :183-210 (3 batches: 3 + 3 + 2)
        corpus_emb = retr.corpus_embeddings.clone().numpy()
        # _encode on an explicit padded batch (reference :92-114)
        tok = retr.tokenizer(states, padding="longest", max_length=512, truncation=True, return_tensors="pt")
        with torch.no_grad():
            state_emb = retr._encode(tok.input_ids, tok.attention_mask).numpy()
        # retrieve (reference :338-375): theorem in F1 after its first premise -> F0 + one F1 premise accessible
        queries = []
        for s, (path, pos, k) in zip(states, [("Gold/F1.lean", (8, 0), 3), ("Gold/F0.lean", (27, 0), 2)]):
            prem, scores = retr.retrieve(s, path, "Gold.target", Pos(*pos), k)
            queries.append({"state": s, "path": path, "pos": list(pos), "k": k,
                            "retrieved": [[p.path, p.full_name] for p in prem], "scores": scores})
        try:
            retr.retrieve(states[0], "Gold/F0.lean", "Gold.target", Pos(7, 0), 2)   # one accessible premise only
            raised = False
        except ValueError:
            raised = True

        # the on-disk index exactly as retrieval/index.py:37-40 writes it
        import pickle
        OUT_PICKLE.write_bytes(pickle.dumps(
            refc.IndexedCorpus(retr.corpus, retr.corpus_embeddings.to(torch.float32).cpu())))

        # validation_step / predict_step (reference :215-268, :281-327) on a batch of three contexts
        retr.num_retrieved = 3
        by_name = {p.full_name: p for p in retr.corpus.all_premises}
        vctx = [refc.Context("Gold/F1.lean", "Gold.t0", Pos(30, 0), states[0]),
                refc.Context("Gold/F1.lean", "Gold.t1", Pos(8, 0), states[1]),
                refc.Context("Gold/F0.lean", "Gold.t2", Pos(49, 0), states[0])]
        positives = [["Gold.F0.lemma_2", "Gold.F1.lemma_1"], [], ["Gold.F0.lemma_4"]]
        vtok = retr.tokenizer([c.serialize() for c in vctx], padding="longest", max_length=512, truncation=True,
                              return_tensors="pt")
        batch = {"context": vctx, "context_ids": vtok.input_ids, "context_mask": vtok.attention_mask,
                 "all_pos_premises": [[by_name[n] for n in names] for names in positives],
                 "url": ["u0", "u1", "u2"], "commit": ["c0", "c1", "c2"], "file_path": [c.path for c in vctx],
                 "full_name": [c.theorem_full_name for c in vctx], "start": [list(c.theorem_pos) for c in vctx],
                 "tactic_idx": [0, 1, 2]}
        with torch.no_grad():
            retr.validation_step(batch, 0)
            retr.predict_step_outputs = []
            retr.predict_step(batch, 0)
        validation = {
            "num_retrieved": 3,
            "contexts": [{"path": c.path, "theorem_full_name": c.theorem_full_name, "pos": list(c.theorem_pos), "state": c.state}
                         for c in vctx],
            "all_pos_premises": positives,
            "logged": retr.logged,
            "predictions": [{"url": r["url"], "commit": r["commit"], "file_path": r["file_path"], "full_name": r["full_name"],
                             "start": r["start"], "tactic_idx": r["tactic_idx"],
                             "retrieved_premises": [p.full_name for p in r["retrieved_premises"]], "scores": r["scores"]}
                            for r in retr.predict_step_outputs],
        }

    np.savez_compressed(OUT_NPZ, corpus_embeddings=corpus_emb, state_embeddings=state_emb,
                        state_input_ids=tok.input_ids.numpy(), state_attention_mask=tok.attention_mask.numpy())
    OUT_JSON.write_text(json.dumps({
        "generator": "tests/golden/make_reference_retriever_golden.py (reference retrieval/model.py + common.py imported unmodified; see its docstring)",
        "weight_seed": synth.SEED, "max_seq_len": 512, "reindex_batch_size": 3,
        "corpus_lines": lines, "states": states, "queries": queries,
        "validation": validation,
        "too_few_accessible": {"state": 0, "path": "Gold/F0.lean", "pos": [7, 0], "k": 2, "raised_value_error": raised},
    }, indent=1, ensure_ascii=False))
    print("wrote", OUT_NPZ.name, corpus_emb.shape, state_emb.shape, "and", OUT_JSON.name)
    print(json.dumps(queries, indent=1)[:600], "raised:", raised)


if __name__ == "__main__":
    main()
