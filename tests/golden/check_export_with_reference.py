"""An index written by THIS package (`reprover_b200.compat.dump_reference_index`, what
`B200PremiseRetriever.save_index` / the index CLI emit) loaded by the REFERENCE'S OWN code.

    python tests/golden/check_export_with_reference.py        # needs /root/reference (this container only)

Imports `/root/reference/common.py` and `retrieval/model.py` unmodified (same stubs for lean_dojo /
pytorch_lightning / deepspeed as make_reference_retriever_golden.py — see that docstring), builds the
golden corpus with this package's host classes, exports it together with the reference's own recorded
embeddings, and then lets the reference do what a stock prover does with an index file:
`PremiseRetriever.load_corpus(path)` (retrieval/model.py:68-85: `pickle.load`) followed by
`corpus.get_nearest_premises(...)` (common.py:299-326).  The objects must be the reference's classes and the
answers must equal the ones the reference gave on the index it built itself (the committed golden).
Prints one JSON line; exit code 0 = every check passed.
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

import make_reference_retriever_golden as gold  # noqa: E402


def main() -> int:
    from reprover_b200.compat import dump_reference_index
    from reprover_b200.corpus import Corpus

    meta = json.loads((HERE / "reference_retriever_cfg1.json").read_text())
    arrays = np.load(HERE / "reference_retriever_cfg1.npz")
    with tempfile.TemporaryDirectory() as tmp:
        jsonl = Path(tmp) / "corpus.jsonl"
        jsonl.write_text("\n".join(json.dumps(l) for l in meta["corpus_lines"]))
        ours = Corpus(str(jsonl))
        out = Path(tmp) / "engine_index.pickle"
        with open(out, "wb") as fh:      # written BEFORE the reference's modules exist in this process
            dump_reference_index(ours, torch.from_numpy(arrays["corpus_embeddings"]), fh)
        assert "common" not in sys.modules and "lean_dojo" not in sys.modules

        Pos = gold._stub_modules()
        sys.path.insert(0, "/root/reference")
        import common as refc
        from retrieval.model import PremiseRetriever

        retr = object.__new__(PremiseRetriever)          # load_corpus touches no model state
        PremiseRetriever.load_corpus(retr, str(out))
        checks = {
            "corpus_is_reference_class": type(retr.corpus) is refc.Corpus,
            "premises_are_reference_class": all(type(p) is refc.Premise for p in retr.corpus.all_premises),
            "positions_are_lean_dojo_pos": all(type(p.start) is Pos and type(p.end) is Pos for p in retr.corpus.all_premises),
            "embeddings_fp32_cpu": retr.corpus_embeddings.dtype == torch.float32 and retr.corpus_embeddings.device.type == "cpu",
            "not_staled": retr.embeddings_staled is False,
            "files": [f.path for f in retr.corpus.files] == [l["path"] for l in meta["corpus_lines"]],
            "deps": retr.corpus.get_dependencies("Gold/F1.lean") == ["Gold/F0.lean"],
        }
        # the reference's own nearest-premise walk on the loaded index, with the reference's recorded state embeddings
        state_emb = torch.from_numpy(arrays["state_embeddings"])
        answers = []
        for i, q in enumerate(meta["queries"]):
            ctx = refc.Context(q["path"], "Gold.target", Pos(*q["pos"]), q["state"])
            prem, scores = retr.corpus.get_nearest_premises(retr.corpus_embeddings, [ctx], state_emb[i:i + 1], q["k"])
            answers.append([[p.path, p.full_name] for p in prem[0]])
            checks[f"query{i}_premises"] = answers[-1] == q["retrieved"]
            checks[f"query{i}_scores"] = bool(np.allclose(scores[0], q["scores"], atol=1e-6))
        # and a fresh reference Corpus built from the jsonl agrees with the loaded one on accessibility
        fresh = refc.Corpus(str(jsonl))
        checks["accessible_sets"] = all(
            sorted(p.full_name for p in fresh.get_accessible_premises(path, Pos(*pos)))
            == sorted(p.full_name for p in retr.corpus.get_accessible_premises(path, Pos(*pos)))
            for path, pos in [("Gold/F1.lean", (8, 0)), ("Gold/F0.lean", (27, 0)), ("Gold/F1.lean", (999, 0))])
    ok = all(bool(v) for v in checks.values())
    print(json.dumps({"ok": ok, "checks": {k: bool(v) for k, v in checks.items()}}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
