"""Host-side logic of the drop-in surface (no GPU): tokenizer mirror, corpus data model and
accessibility masks, synthetic data, shard bounds."""
import json
import pickle
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import reference_path as ref
from reprover_b200 import synth
from reprover_b200 import tokenizer as byt5
from reprover_b200.corpus import Context, Corpus, File, IndexedCorpus, Pos, Premise, PremiseSet, remove_marks
from reprover_b200.dist import shard_bounds

GOLD = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------------------- tokenizer mirror
def test_tokenizer_mirror_matches_golden_probes():
    for p in json.loads((GOLD / "tokenizer_probes.json").read_text()):
        assert byt5.encode_ids(p["text"], p["max_length"]) == p["ids"], p


def test_needs_id_path_only_for_special_literals():
    assert not byt5.needs_id_path("theorem foo : a < b ∧ b > c := by simp")
    assert not byt5.needs_id_path("<a>Nat.add_comm</a> : ∀ n m, n + m = m + n")
    assert not byt5.needs_id_path("<extra_id_125> is not a token, <extra_id_007> neither")
    for s in ("x </s>", "<pad>", "a<unk>b", "<extra_id_0>", "<extra_id_124>"):
        assert byt5.needs_id_path(s)


def test_plain_strings_tokenize_as_bytes_plus_three():
    tok = ref.build_hf_tokenizer()
    data, offsets = synth.synth_premises(10, seed=3, min_len=1, max_len=50)
    for s in synth.split_strings(data, offsets):
        t = s.decode()
        assert not byt5.needs_id_path(t)
        assert tok(t).input_ids == [b + 3 for b in s] + [1]


def test_pad_batch_layout():
    ids, mask = byt5.pad_batch([[5, 6, 1], [7, 1]])
    assert ids.tolist() == [[5, 6, 1], [7, 1, 0]] and mask.tolist() == [[1, 1, 1], [1, 1, 0]]


# ------------------------------------------------------------------------------- data model
def _toy_corpus():
    def prem(path, name, line, code=None):
        return Premise(path, name, Pos(line, 0), Pos(line, 10), code or f"theorem {name} : True := trivial")

    a = File("A.lean", [prem("A.lean", "A.one", 1), prem("A.lean", "A.two", 5)])
    b = File("B.lean", [prem("B.lean", "B.one", 2), prem("B.lean", "B.dup", 4), prem("B.lean", "B.dup", 9)])
    c = File("C.lean", [prem("C.lean", "C.one", 3), prem("C.lean", "C.two", 7), prem("C.lean", "C.three", 11)])
    d = File("D.lean", [])
    return Corpus.from_files([(a, []), (b, ["A.lean"]), (c, ["B.lean"]), (d, [])])


def test_pos_ordering_and_context_invariants():
    assert Pos(1, 5) < Pos(2, 0) and Pos(2, 1) <= Pos(2, 1) and Pos(3, 0) > Pos(2, 9)
    assert Pos.from_any((4, 2)) == Pos(4, 2)
    with pytest.raises(AssertionError):
        Context("A.lean", "thm", Pos(1, 0), "no turnstile here")
    with pytest.raises(AssertionError):
        Context("A.lean", "thm", Pos(1, 0), "⊢ <a>x</a>")
    ctx = Context("A.lean", "thm", (1, 0), "x : Nat\n⊢ x = x")
    assert ctx.serialize() == "x : Nat\n⊢ x = x" and isinstance(ctx.theorem_pos, Pos)
    with pytest.raises(AssertionError):
        Premise("A.lean", "n", Pos(2, 0), Pos(1, 0), "code")
    with pytest.raises(AssertionError):
        Premise("A.lean", "n", Pos(1, 0), Pos(1, 0), "")


def test_premise_serialize_marks_own_name():
    p = Premise("M.lean", "Nat.Foo.add_zero", Pos(1, 0), Pos(2, 0), "theorem add_zero (n : Nat) : n + 0 = n := rfl")
    assert p.serialize() == "theorem <a>Nat.Foo.add_zero</a> (n : Nat) : n + 0 = n := rfl"
    q = Premise("M.lean", "Nat.bar", Pos(1, 0), Pos(2, 0), "def _root_.Nat.bar := 1")
    assert q.serialize() == "def <a>Nat.bar</a> := 1"
    r = Premise("M.lean", "X.y", Pos(1, 0), Pos(2, 0), "instance : Foo := ⟨⟩")
    assert r.serialize() == "instance : Foo := ⟨⟩"
    assert remove_marks(p.serialize()) == "theorem Nat.Foo.add_zero (n : Nat) : n + 0 = n := rfl"


def _serialize_by_regex(full_name: str, code: str) -> str:
    """The reference rule verbatim in behaviour (common.py:93-106): un-escaped name as a regex."""
    import re
    marked = f"<a>{full_name}</a>"
    text = code.replace(f"_root_.{full_name}", marked)
    parts = full_name.split(".")
    for i in range(len(parts)):
        new = re.sub(f"(?<=\\s)«?{'.'.join(parts[i:])}»?", marked, text)
        if new != text:
            return new
    return text


def test_premise_serialize_fast_path_equals_the_regex_rule():
    """`Premise.serialize` skips the per-premise regex compile for ordinary names; the scan it uses
    instead must reproduce `re.sub` exactly — overlaps, « » quoting, the any-character dots, the
    whitespace look-behind, newlines, repeated and adjacent occurrences, suffix fallback."""
    rng = np.random.default_rng(7)
    comps = ["Nat", "add", "add_comm", "a", "ab", "foo'", "β", "x1", "Nat", "comm"]
    glue = [" ", "\n", "\t", ".", "x", "«", "»", "(", ":", "  ", "\u00a0", "_", "a", "Nat", "add", "comm", "ab", "β"]
    cases = [
        ("Nat.add_comm", "theorem Nat.add_comm : a + b = b + a"),
        ("Nat.add_comm", "namespace Nat\ntheorem add_comm (n m : Nat) : n + m = m + n := by\n  exact add_comm n m"),
        ("Nat.add_comm", "theorem «Nat.add_comm» and NatXadd_comm and Nat\nadd_comm"),
        ("a.a", " a.a.a aXa a a"),
        ("ab", " abab ab«ab» «ab» ab"),
        ("Foo.bar", "lemma _root_.Foo.bar : True"),
        ("Foo.bar", "no match here"),
        ("x", "x x  x\nx"),
    ]
    for _ in range(4000):
        k = int(rng.integers(1, 4))
        name = ".".join(comps[int(i)] for i in rng.integers(0, len(comps), k))
        pieces = []
        for _ in range(int(rng.integers(1, 14))):
            r = rng.random()
            if r < 0.35:
                pieces.append(name if rng.random() < 0.5 else name.split(".", 1)[-1])
            elif r < 0.5:
                pieces.append(name.replace(".", glue[int(rng.integers(0, len(glue)))]))
            else:
                pieces.append(glue[int(rng.integers(0, len(glue)))])
        cases.append((name, "".join(pieces) or "x"))
    for name, code in cases:
        p = Premise("A.lean", name, Pos(1, 0), Pos(2, 0), code)
        assert p.serialize() == _serialize_by_regex(name, code), (name, code)
    # names with regex metacharacters / quotes keep the regex path (same rule by construction)
    for name, code in [("Foo.«bar baz»", "def Foo.«bar baz» := 1"), ("a+b", " aab a+b"), ("f(x", "def f(x")]:
        try:
            want = _serialize_by_regex(name, code)
        except Exception as e:  # the reference raises on an invalid pattern; so do we
            with pytest.raises(type(e)):
                Premise("A.lean", name, Pos(1, 0), Pos(2, 0), code).serialize()
            continue
        assert Premise("A.lean", name, Pos(1, 0), Pos(2, 0), code).serialize() == want


def test_corpus_accessibility_and_mask():
    corpus = _toy_corpus()
    assert len(corpus) == 8 and corpus.num_files == 4 and "C.lean" in corpus
    assert set(corpus.get_dependencies("C.lean")) == {"A.lean", "B.lean"}
    assert corpus.file_range("B.lean") == (2, 5)
    assert corpus.locate_premise("B.lean", Pos(4, 3)).full_name == "B.dup"
    for path in ("A.lean", "B.lean", "C.lean", "D.lean"):
        for pos in (Pos(0, 0), Pos(4, 10), Pos(7, 10), Pos(100, 0)):
            acc = corpus.get_accessible_premises(path, pos)
            want = np.array([p in acc for p in corpus.all_premises])
            assert np.array_equal(corpus.accessible_mask(path, pos), want), (path, pos)
            words = corpus.accessible_mask_words(path, pos)
            bits = np.unpackbits(words.view(np.uint8), bitorder="little")[: len(corpus)].astype(bool)
            assert np.array_equal(bits, want)
    # membership is by (path, full_name): once B.dup@4 is visible, its later duplicate tests True too
    m = corpus.accessible_mask("B.lean", Pos(5, 0))
    assert m.tolist() == [True, True, True, True, True, False, False, False]
    assert corpus.get_accessible_premise_indexes("C.lean", Pos(7, 10)) == [0, 1, 2, 3, 4, 5, 6]


def test_corpus_rejects_forward_imports_and_duplicates():
    a = File("A.lean", [])
    with pytest.raises(AssertionError):
        Corpus.from_files([(a, ["Z.lean"])])
    with pytest.raises(AssertionError):
        Corpus.from_files([(a, []), (File("A.lean", []), [])])


def test_corpus_jsonl_loader_and_filtering(tmp_path):
    lines = [
        {"path": "A.lean", "imports": [], "premises": [
            {"full_name": "A.ok", "code": "theorem ok : True := trivial", "start": [1, 0], "end": [1, 9]},
            {"full_name": None, "code": "x", "start": [2, 0], "end": [2, 1]},
            {"full_name": "user__.n.bad", "code": "x", "start": [3, 0], "end": [3, 1]},
            {"full_name": "[mutual]", "code": "x", "start": [4, 0], "end": [4, 1]},
            {"full_name": "A.empty", "code": "", "start": [5, 0], "end": [5, 1]}]},
        {"path": "B.lean", "imports": ["A.lean"], "premises": [
            {"full_name": "B.x", "code": "def x := 1", "start": [1, 0], "end": [1, 5]}]},
    ]
    f = tmp_path / "corpus.jsonl"
    f.write_text("\n".join(json.dumps(l) for l in lines))
    corpus = Corpus(str(f))
    assert [p.full_name for p in corpus.all_premises] == ["A.ok", "B.x"]
    assert corpus.get_dependencies("B.lean") == ["A.lean"]


def test_indexed_corpus_contract_and_pickle(tmp_path):
    corpus = _toy_corpus()
    emb = torch.zeros(len(corpus), 8)
    idx = IndexedCorpus(corpus, emb)
    with pytest.raises(AssertionError):
        IndexedCorpus(corpus, torch.zeros(3, 8))
    path = tmp_path / "index.pickle"
    path.write_bytes(pickle.dumps(idx))
    back = pickle.loads(path.read_bytes())
    assert len(back.corpus) == len(corpus) and back.embeddings.shape == (8, 8)
    assert back.corpus.accessible_mask("C.lean", Pos(7, 10)).tolist() == corpus.accessible_mask("C.lean", Pos(7, 10)).tolist()


def test_premise_set_semantics():
    s = PremiseSet()
    p1 = Premise("A.lean", "x", Pos(1, 0), Pos(1, 1), "c")
    p2 = Premise("A.lean", "x", Pos(9, 0), Pos(9, 1), "other")
    s.add(p1)
    assert p2 in s and len(s) == 1
    s.update([Premise("B.lean", "y", Pos(1, 0), Pos(1, 1), "c")])
    assert len(s) == 2 and {p.full_name for p in s} == {"x", "y"}


# ------------------------------------------------------------------------------- synth / sharding
def test_synthetic_data_is_deterministic_and_in_spec():
    d1, o1 = synth.synth_premises(1000, seed=synth.SEED)
    d2, o2 = synth.synth_premises(1000, seed=synth.SEED)
    assert np.array_equal(d1, d2) and np.array_equal(o1, o2)
    lens = np.diff(o1)
    assert lens.min() >= 16 and lens.max() <= 511 and abs(lens.mean() - 263.5) < 15
    assert d1.min() >= 0x20 and d1.max() <= 0x7E and (d1 != 0x3C).all()
    sdat, soff = synth.synth_states(5)
    for s in synth.split_strings(sdat, soff):
        assert s.decode().startswith("⊢ ")
    sd = synth.random_t5_state_dict(synth.tiny_config(1), seed=1)
    sd2 = synth.random_t5_state_dict(synth.tiny_config(1), seed=1)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert sd["encoder.block.0.layer.0.SelfAttention.q.weight"].shape == (384, 1472)


def test_shard_bounds_partition_rows():
    for n, w in ((10, 3), (1_600_000, 8), (5, 8), (0, 2)):
        b = shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = np.diff(b)
        assert (sizes >= 0).all() and sizes.max() - sizes.min() <= 1


# ------------------------------------------------------------------------------- metrics / compat
def test_recall_and_mrr_matches_reference_arithmetic():
    from reprover_b200.evaluation import recall_and_mrr

    P = [Premise("A.lean", f"p{i}", Pos(i + 1, 0), Pos(i + 1, 1), "c") for i in range(6)]
    retrieved = [[P[0], P[1], P[2]], [P[3], P[4], P[5]], [P[0], P[2], P[4]]]
    positives = [[P[1], P[5]], [], [P[5]]]
    recall, mrr, n = recall_and_mrr(positives, retrieved, 3)
    # example 0: hits after j=1,2,3 -> 0, 1, 1 of 2 positives; example 1 skipped; example 2: no hit
    assert n == 2
    assert recall == pytest.approx([0.0, 25.0, 25.0])
    assert mrr == pytest.approx((0.5 + 0.0) / 2)


def test_reference_index_pickle_is_readable(tmp_path):
    """An IndexedCorpus pickled by the reference (classes `common.*`, `lean_dojo.Pos`, networkx graph)
    loads through reprover_b200.compat without either module being importable."""
    import subprocess
    import sys
    import textwrap

    out = tmp_path / "ref_index.pickle"
    script = textwrap.dedent(f"""
        import pickle, sys, types
        from dataclasses import dataclass, field
        import networkx as nx, torch
        ld = types.ModuleType("lean_dojo"); sys.modules["lean_dojo"] = ld
        common = types.ModuleType("common"); sys.modules["common"] = common
        @dataclass(frozen=True, order=True)
        class Pos:
            line_nb: int
            column_nb: int
        Pos.__module__ = "lean_dojo"; ld.Pos = Pos
        @dataclass(unsafe_hash=True)
        class Premise:
            path: str
            full_name: str
            start: Pos = field(repr=False)
            end: Pos = field(repr=False, compare=False)
            code: str = field(compare=False)
        @dataclass(frozen=True)
        class File:
            path: str
            premises: list = field(repr=False, compare=False)
        class Corpus:
            pass
        @dataclass(frozen=True)
        class IndexedCorpus:
            corpus: Corpus
            embeddings: torch.Tensor
        for c in (Premise, File, Corpus, IndexedCorpus):
            c.__module__ = "common"; setattr(common, c.__name__, c)
        fa = File("A.lean", [Premise("A.lean", "A.x", Pos(1, 0), Pos(2, 0), "def x := 1")])
        fb = File("B.lean", [Premise("B.lean", "B.y", Pos(3, 0), Pos(4, 0), "def y := 2"),
                             Premise("B.lean", "B.z", Pos(5, 0), Pos(6, 0), "def z := 3")])
        g = nx.DiGraph(); g.add_node("A.lean", file=fa); g.add_node("B.lean", file=fb); g.add_edge("B.lean", "A.lean")
        corpus = Corpus(); corpus.transitive_dep_graph = nx.transitive_closure_dag(g)
        corpus.all_premises = fa.premises + fb.premises; corpus.imported_premises_cache = {{}}
        pickle.dump(IndexedCorpus(corpus, torch.arange(12.).reshape(3, 4)), open(r"{out}", "wb"))
    """)
    subprocess.run([sys.executable, "-c", script], check=True)
    from reprover_b200.compat import load_reference_index

    idx = load_reference_index(str(out))
    assert [p.full_name for p in idx.corpus.all_premises] == ["A.x", "B.y", "B.z"]
    assert idx.corpus.get_dependencies("B.lean") == ["A.lean"] and idx.embeddings.shape == (3, 4)
    assert idx.corpus.accessible_mask("B.lean", Pos(4, 5)).tolist() == [True, True, False]
    assert idx.corpus.all_premises[1].end == Pos(4, 0)


class _StubEngine:
    """Stands in for T5EncoderEngine in host-logic tests: an 'embedding' is a pure function of the
    string's bytes, and every call is recorded."""

    hidden_size = 8

    def __init__(self, max_tokens_per_call):
        self.max_tokens_per_call = max_tokens_per_call
        self.calls = []          # (n_strings, tokens) per encode_strings call
        self.id_calls = []       # batch sizes of encode_ids calls

    @staticmethod
    def _emb(blob: bytes) -> torch.Tensor:
        import hashlib
        h = hashlib.sha256(blob).digest()[:8]
        return torch.tensor([b / 255.0 for b in h], dtype=torch.float32)

    def encode_strings(self, blobs, max_seq_len, out_dtype=torch.float32, out=None):
        self.calls.append((len(blobs), sum(min(len(b) + 1, max_seq_len) for b in blobs)))
        emb = torch.stack([self._emb(b[: max_seq_len - 1]) for b in blobs]).to(out_dtype)
        if out is None:
            return emb
        out.copy_(emb)
        return out

    def encode_ids(self, input_ids, attention_mask, out_dtype=torch.float32):
        self.id_calls.append(int(input_ids.shape[0]))
        rows = []
        for ids, m in zip(input_ids.tolist(), attention_mask.tolist()):
            rows.append(self._emb(bytes(("ids:" + ",".join(str(i) for i, k in zip(ids, m) if k)).encode())))
        return torch.stack(rows).to(out_dtype)


def _stub_retriever(budget, max_seq_len=64):
    from reprover_b200.retriever import B200PremiseRetriever
    r = object.__new__(B200PremiseRetriever)   # host logic only: no CUDA engine behind it
    r.encoder = _StubEngine(budget)
    r.device = torch.device("cpu")
    r.dtype = torch.float32
    r.max_seq_len = max_seq_len
    r.num_retrieved = 100
    r.corpus = None
    r.corpus_embeddings = None
    r.embeddings_staled = True
    return r


def test_encode_texts_groups_by_token_budget_and_keeps_row_order():
    """Host streaming logic of `encode_texts` / `reindex_corpus` (reference retrieval/model.py:183-210):
    groups never exceed the engine's token budget, every group but the last is as full as the next
    string allows, strings with special-token literals go through the ids path, rows come back in
    input order whatever the grouping — and a lazily serialised corpus is walked exactly once."""
    rng = np.random.default_rng(3)
    texts = []
    for i in range(300):
        n = int(rng.integers(1, 120))
        t = "".join(chr(int(c)) for c in rng.integers(97, 123, n))
        if i in (5, 140, 141, 299):
            t += " <extra_id_3> tail"          # host-tokenised path
        if i == 77:
            t = "a < b and x <a> marked </a>"    # '<' without a special literal stays on the byte path
        texts.append(t)
    want = None
    for budget in (64, 257, 1000, 10**9):
        r = _stub_retriever(budget)
        got = r.encode_texts(texts, batch_size=3)
        assert got.shape == (300, 8)
        if want is None:
            want = got.clone()
            # reference values straight from the stub, row by row
            for i in (0, 77, 150, 298):
                assert torch.equal(got[i], _StubEngine._emb(texts[i].encode()[:63]))
        assert torch.equal(got, want), budget
        eng = r.encoder
        assert all(tok <= budget for _, tok in eng.calls)
        assert sum(n for n, _ in eng.calls) == 296 and sum(eng.id_calls) == 4 and max(eng.id_calls) <= 3
        if budget == 10**9:
            assert len(eng.calls) == 1
        elif budget >= 257:
            # every group but the last could not have taken the next string (<= 64 tokens)
            assert all(tok > budget - 64 for _, tok in eng.calls[:-1])

    # reindex_corpus: lazy serialisation, one pass, same rows as encoding the serialised strings
    prem = [Premise("A.lean", f"A.t{i}", Pos(i + 1, 0), Pos(i + 1, 5), f"theorem t{i} : {texts[i]}") for i in range(120)]
    corpus = Corpus.from_files([(File("A.lean", prem), [])])
    calls = {"n": 0}
    orig = Premise.serialize

    def counting(self):
        calls["n"] += 1
        return orig(self)

    r = _stub_retriever(500)
    r.load_corpus(corpus)
    assert r.embeddings_staled
    Premise.serialize = counting
    try:
        r.reindex_corpus(batch_size=4)
    finally:
        Premise.serialize = orig
    assert calls["n"] == 120 and not r.embeddings_staled
    assert torch.equal(r.corpus_embeddings, _stub_retriever(10**9).encode_texts([p.serialize() for p in corpus.all_premises]))
    before = r.corpus_embeddings
    r.reindex_corpus(batch_size=4)             # fresh index: no-op (reference :185-186)
    assert r.corpus_embeddings is before


def test_import_closure_bitsets_and_old_pickles():
    """Transitive imports are kept as bit sets over the file order; a corpus pickled with the earlier
    representation (sets of paths under `_deps`) is upgraded on load and answers identically."""
    rng = np.random.default_rng(5)
    files = []
    for f in range(60):
        imports = sorted({f"S/F{int(i)}.lean" for i in rng.integers(0, f, size=min(f, 3))}) if f else []
        prem = [Premise(f"S/F{f}.lean", f"S.F{f}.l{j}", Pos(10 * j + 1, 0), Pos(10 * j + 5, 0), f"theorem l{j} : True")
                for j in range(int(rng.integers(1, 6)))]
        files.append((File(f"S/F{f}.lean", prem), imports))
    c = Corpus.from_files(files)
    # closure by definition
    direct = {f.path: set(imps) for f, imps in files}
    def closure(p, seen=None):
        seen = set() if seen is None else seen
        for d in direct[p]:
            if d not in seen:
                seen.add(d)
                closure(d, seen)
        return seen
    order = [f.path for f, _ in files]
    for p in order:
        assert c.get_dependencies(p) == [q for q in order if q in closure(p)]
    # masks: packed words == boolean mask == membership in the reference-style accessible set
    for p in (order[0], order[17], order[-1]):
        pos = Pos(25, 0)
        b = c.accessible_mask(p, pos)
        w = c.accessible_mask_words(p, pos)
        assert (np.unpackbits(w.view(np.uint8), bitorder="little")[: len(c)].astype(bool) == b).all()
        acc = c.get_accessible_premises(p, pos)
        assert [q in acc for q in c.all_premises] == b.tolist()
    # old-format state
    state = c.__getstate__()
    old = {k: v for k, v in state.items() if k not in ("_order", "_index", "_dep_bits")}
    old["_deps"] = {p: frozenset(closure(p)) for p in order}
    c_old = Corpus.__new__(Corpus)
    c_old.__setstate__(old)
    for p in (order[3], order[-1]):
        assert c_old.get_dependencies(p) == c.get_dependencies(p)
        assert (c_old.accessible_mask_words(p, Pos(99, 0)) == c.accessible_mask_words(p, Pos(99, 0))).all()
    c2 = pickle.loads(pickle.dumps(c))
    assert c2.get_dependencies(order[-1]) == c.get_dependencies(order[-1])


def test_index_written_by_the_reference_code_loads(tmp_path):
    """`tests/golden/reference_indexed_corpus.pickle` was written by the reference's own classes
    (`common.IndexedCorpus(corpus, embeddings.float().cpu())`, retrieval/index.py:37-40; generator:
    tests/golden/make_reference_retriever_golden.py).  `load_corpus(path)` must accept it without
    `common` / `lean_dojo` being importable and end up with the same corpus and embeddings."""
    gold = Path(__file__).resolve().parent / "golden"
    meta = json.loads((gold / "reference_retriever_cfg1.json").read_text())
    want_emb = np.load(gold / "reference_retriever_cfg1.npz")["corpus_embeddings"]
    r = _stub_retriever(10**9)
    r.load_corpus(str(gold / "reference_indexed_corpus.pickle"))
    assert not r.embeddings_staled
    assert r.corpus_embeddings.dtype == torch.float32 and r.corpus_embeddings.device.type == "cpu"
    assert np.array_equal(r.corpus_embeddings.numpy(), want_emb)
    want = [(l["path"], p["full_name"], p["code"], tuple(p["start"]), tuple(p["end"]))
            for l in meta["corpus_lines"] for p in l["premises"]]
    got = [(p.path, p.full_name, p.code, (p.start.line_nb, p.start.column_nb), (p.end.line_nb, p.end.column_nb))
           for p in r.corpus.all_premises]
    assert got == want
    assert r.corpus.get_dependencies("Gold/F1.lean") == ["Gold/F0.lean"] and r.corpus.get_dependencies("Gold/F0.lean") == []
    # the converted corpus answers accessibility like one built from the jsonl
    jsonl = tmp_path / "corpus.jsonl"
    jsonl.write_text("\n".join(json.dumps(l) for l in meta["corpus_lines"]))
    fresh = Corpus(str(jsonl))
    for path, pos in [("Gold/F1.lean", Pos(8, 0)), ("Gold/F0.lean", Pos(27, 0)), ("Gold/F1.lean", Pos(999, 0))]:
        assert (r.corpus.accessible_mask_words(path, pos) == fresh.accessible_mask_words(path, pos)).all()


def _toy_files():
    def P(path, i):
        return Premise(path, f"N.{path[:-5]}.l{i}", Pos(10 * i + 1, 0), Pos(10 * i + 5, 0), f"theorem l{i} : True := trivial")
    return [(File("A.lean", [P("A.lean", 0), P("A.lean", 1)]), []),
            (File("Empty.lean", []), []),                       # a file without premises ...
            (File("B.lean", [P("B.lean", 0)]), ["A.lean", "Empty.lean"]),   # ... imported by a later one
            (File("C.lean", [P("C.lean", 0), P("C.lean", 1)]), ["B.lean"])]


def test_index_export_in_the_reference_layout_roundtrips(tmp_path):
    """`dump_reference_index` writes GLOBALs `common.*` / `lean_dojo.Pos` (what a stock checkout unpickles,
    retrieval/model.py:81-85) without leaving those modules behind, and `load_reference_index` reads the file
    back to an equal corpus — including a premise-less file that a later file imports (the order the compat
    loader used to get wrong)."""
    import io
    import pickletools
    import sys

    from reprover_b200.compat import dump_reference_index, load_reference_index

    corpus = Corpus.from_files(_toy_files())
    emb = torch.arange(5 * 8, dtype=torch.float32).reshape(5, 8)
    buf = io.BytesIO()
    dump_reference_index(corpus, emb.to(torch.bfloat16), buf)
    assert "common" not in sys.modules and "lean_dojo" not in sys.modules
    names = {arg for op, arg, _ in pickletools.genops(buf.getvalue()) if op.name == "SHORT_BINUNICODE"}
    assert {"common", "lean_dojo", "IndexedCorpus", "Corpus", "File", "Premise", "Pos", "networkx.classes.digraph"} <= names
    assert not any(str(n).startswith("reprover_b200") for n in names)
    with pytest.raises((ModuleNotFoundError, AttributeError)):
        pickle.loads(buf.getvalue())                       # only loadable where the reference's classes exist ...
    back = load_reference_index(buf.getvalue())            # ... or through the compat loader
    assert back.embeddings.dtype == torch.float32 and torch.equal(back.embeddings, emb.to(torch.bfloat16).float())
    assert [(p.path, p.full_name, p.start, p.end, p.code) for p in back.corpus.all_premises] == \
           [(p.path, p.full_name, p.start, p.end, p.code) for p in corpus.all_premises]
    assert [f.path for f in back.corpus.files] == ["A.lean", "Empty.lean", "B.lean", "C.lean"]
    for path in ("A.lean", "B.lean", "C.lean"):
        assert back.corpus.get_dependencies(path) == corpus.get_dependencies(path)
        assert (back.corpus.accessible_mask_words(path, Pos(12, 0)) == corpus.accessible_mask_words(path, Pos(12, 0))).all()
    # the retriever's load_corpus takes the same file
    path = tmp_path / "idx.pickle"
    path.write_bytes(buf.getvalue())
    r = _stub_retriever(10**9)
    r.load_corpus(str(path))
    assert not r.embeddings_staled and len(r.corpus) == 5 and "Empty.lean" in r.corpus


def test_compat_file_order_falls_back_to_a_topological_sort():
    """A reference pickle whose graph nodes are not in import order still converts (imports first, premise
    order kept); a cycle is reported."""
    import networkx as nx

    from reprover_b200 import compat

    fa = compat._RefFile(); fa.__dict__.update(path="A.lean", premises=[])
    pb = compat._RefPremise()
    pos = compat._RefPos(); pos.__dict__.update(line_nb=1, column_nb=0)
    pb.__dict__.update(path="B.lean", full_name="B.x", start=pos, end=pos, code="def x := 1")
    fb = compat._RefFile(); fb.__dict__.update(path="B.lean", premises=[pb])
    g = nx.DiGraph()
    g.add_node("B.lean", file=fb)          # importer first: not an import order
    g.add_node("A.lean", file=fa)
    g.add_edge("B.lean", "A.lean")
    ref_corpus = compat._RefCorpus(); ref_corpus.__dict__.update(transitive_dep_graph=g, all_premises=[pb])
    c = compat.convert_corpus(ref_corpus)
    assert [f.path for f in c.files] == ["A.lean", "B.lean"] and c.get_dependencies("B.lean") == ["A.lean"]
    g.add_edge("A.lean", "B.lean")
    with pytest.raises(ValueError, match="cycle"):
        compat.convert_corpus(ref_corpus)


def test_corpus_embeddings_assignment_drops_the_index_handle():
    class _Handle:
        closed = False

        def close(self):
            self.closed = True

    r = _stub_retriever(10**9)
    h = _Handle()
    r._index_handle, r._index_source = h, (1, 0)
    same = r.corpus_embeddings
    r.corpus_embeddings = same                     # same object: nothing to invalidate
    assert r._index_handle is h and not h.closed
    r.corpus_embeddings = torch.zeros(2, 8)
    assert r._index_handle is None and h.closed


def test_checkpoint_and_k_errors_are_explicit(tmp_path):
    from reprover_b200.engine import load_hf_checkpoint, required_weight_keys
    from reprover_b200.retrieval_ops import _check_k

    with pytest.raises(FileNotFoundError, match="neither a local checkpoint directory nor a hub snapshot"):
        load_hf_checkpoint("no-such-org/no-such-retriever-checkpoint")
    (tmp_path / "weights.bin").write_bytes(b"x")
    with pytest.raises(FileNotFoundError, match="DIRECTORY"):
        load_hf_checkpoint(str(tmp_path / "weights.bin"))
    (tmp_path / "ck").mkdir()
    with pytest.raises(FileNotFoundError, match="config.json"):
        load_hf_checkpoint(str(tmp_path / "ck"))
    keys = required_weight_keys({"num_layers": 2})
    assert len(keys) == 2 + 2 * 9 and "encoder.block.1.layer.1.DenseReluDense.wo.weight" in keys
    for bad in (0, -3, 1025, 2.5):
        with pytest.raises(ValueError, match="premises per query"):
            _check_k(bad)
    _check_k(1), _check_k(1024)
