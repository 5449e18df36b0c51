"""Shared test helpers (oracle access lives here: tests may import `oracle/`)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import reference_path as ref  # noqa: E402
from reprover_b200 import synth  # noqa: E402

# Acceptance for engine embeddings vs the fp32-"highest" HF oracle:
# no worse than the reference's own GPU dtype (HF bf16 vs fp32: 1.3e-3 / 0.99992).
EMB_MAX_ABS = 2e-3
EMB_MIN_COS = 0.9999


def oracle_embeddings(cfg, sd, data: np.ndarray, offsets: np.ndarray, max_seq_len: int, batch_size: int = 8):
    """fp32 embeddings of the byte strings from the HF-based oracle (CPU)."""
    torch.set_float32_matmul_precision("highest")
    enc = ref.build_hf_encoder(cfg, sd)
    tok = ref.build_hf_tokenizer()
    texts = [s.decode("utf-8") for s in synth.split_strings(data, offsets)]
    return ref.reindex_corpus(enc, tok, texts, batch_size, max_seq_len)


def compare_embeddings(got: torch.Tensor, want: torch.Tensor):
    got = got.float().cpu()
    want = want.float().cpu()
    max_abs = float((got - want).abs().max())
    cos = torch.nn.functional.cosine_similarity(got, want, dim=1)
    return max_abs, float(cos.min())
