"""Generates the golden fixtures in this directory from the REAL third-party code the reference
delegates to (transformers T5EncoderModel / ByT5Tokenizer + torch), run in the build container:

    python tests/golden/make_golden.py

The reference (lean-dojo/ReProver) has no tests or golden vectors for this path (SURVEY.md §4,
§8c) and its own modules do not import here (lean_dojo / lightning / deepspeed missing), so these
fixtures pin the oracle against the installed HF implementation instead:

  bucket_table.json      T5Attention._relative_position_bucket (HF modeling_t5.py:189-234) for
                         relative positions -130..130, bidirectional, 32 buckets, max distance 128
  tokenizer_probes.json  ByT5Tokenizer ids for strings with / without special-token literals and
                         several max_length truncations
  cfg1.npz               BASELINE config 1: ByT5-small geometry, synthetic weights seed 3407,
                         8 premises + 1 state (synthetic, seeds in the file): fp32 embeddings from
                         `_encode` semantics (oracle/reference_path.py on HF), cosine top-3
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import reference_path as ref  # noqa: E402
from reprover_b200 import synth  # noqa: E402

PROBES = [
    "theorem foo (a b : Nat) : a + b = b + a := by simp [Nat.add_comm]",
    "abc </s> <pad> <extra_id_0> é",
    "⊢ ∀ (x : ℝ), 0 ≤ x ^ 2",
    "a<pad>b", " x <unk>  y", "<extra_id_124>z", "<extra_id_125>z", "<extra_id_007>", "text</s>",
    "<a>Nat.add</a> x", "x </s>", "  </s>  y ", "<extra_id_3> <extra_id_4>", "a\t</s>\nb", "", " ", "</s>",
    "def «weird name» := 1 < 2",
]


def main():
    from transformers.models.t5.modeling_t5 import T5Attention

    rel = torch.arange(-130, 131)
    buckets = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
    (HERE / "bucket_table.json").write_text(json.dumps(
        {"relative_position": rel.tolist(), "bucket": buckets.tolist(), "num_buckets": 32, "max_distance": 128}))

    tok = ref.build_hf_tokenizer()
    probes = []
    for text in PROBES:
        for ml in (512, 16, 4, 2, 1):
            probes.append({"text": text, "max_length": ml, "ids": tok(text, max_length=ml, truncation=True).input_ids})
    (HERE / "tokenizer_probes.json").write_text(json.dumps(probes, ensure_ascii=False, indent=0))

    torch.set_float32_matmul_precision("highest")
    cfg = dict(synth.BYT5_SMALL)
    sd = synth.random_t5_state_dict(cfg, seed=synth.SEED)
    pd, po = synth.synth_premises(8, seed=synth.SEED)
    sdt, so = synth.synth_states(1, seed=synth.SEED + 1)
    data = np.concatenate([pd, sdt])
    offsets = np.concatenate([po, po[-1] + so[1:]])
    texts = [s.decode() for s in synth.split_strings(data, offsets)]
    enc = ref.build_hf_encoder(cfg, sd)
    emb = ref.reindex_corpus(enc, tok, texts, 64, 512).numpy()
    sims = emb[8:] @ emb[:8].T
    top3 = np.argsort(-sims, axis=1, kind="stable")[:, :3]
    np.savez_compressed(HERE / "cfg1.npz", embeddings=emb.astype(np.float32), data=data, offsets=offsets,
                        top3=top3, sims=sims.astype(np.float32), weight_seed=synth.SEED,
                        weight_checksum=float(sum(v.double().sum() for v in sd.values())))
    print("wrote", sorted(p.name for p in HERE.iterdir()))


if __name__ == "__main__":
    main()
