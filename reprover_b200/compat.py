"""Reading indexes written by the *reference* (`retrieval/index.py:37-40`).

A reference index is a pickle of `common.IndexedCorpus(corpus: common.Corpus, embeddings)`, where
`common.Corpus` holds a networkx `transitive_dep_graph` (node attribute "file" = `common.File`) and
`all_premises` of `common.Premise` with `lean_dojo.Pos` positions (common.py:181-219).  Neither
`common` nor `lean_dojo` is importable next to this package, so the unpickler below maps those
classes onto light stand-ins and rebuilds a `reprover_b200.corpus.Corpus` (file order = order of
first appearance in `all_premises`, then remaining nodes; imports = graph successors, which are
already transitive).  networkx must be importable to decode the graph object itself.
"""
from __future__ import annotations

import io
import pickle
from typing import Any, Dict, List

from .corpus import Corpus, File, IndexedCorpus, Pos, Premise


class _Bag:
    """Stand-in that just records the pickled attribute dict."""

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):  # dataclass with slots
            state = {**(state[0] or {}), **state[1]}
        self.__dict__.update(state)


class _RefPos(_Bag):
    pass


class _RefPremise(_Bag):
    pass


class _RefFile(_Bag):
    pass


class _RefCorpus(_Bag):
    pass


class _RefIndexedCorpus(_Bag):
    pass


_MAP = {
    ("common", "IndexedCorpus"): _RefIndexedCorpus,
    ("common", "Corpus"): _RefCorpus,
    ("common", "File"): _RefFile,
    ("common", "Premise"): _RefPremise,
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if (module, name) in _MAP:
            return _MAP[(module, name)]
        if name == "Pos" and module.split(".")[0] == "lean_dojo":
            return _RefPos
        return super().find_class(module, name)


def _pos(p: Any) -> Pos:
    return Pos(int(p.line_nb), int(p.column_nb))


def _premise(p: Any) -> Premise:
    return Premise(p.path, p.full_name, _pos(p.start), _pos(p.end), p.code)


def convert_corpus(ref: Any) -> Corpus:
    """`common.Corpus` stand-in -> `reprover_b200.corpus.Corpus` with the same `all_premises` order."""
    g = ref.transitive_dep_graph
    order: List[str] = []
    seen = set()
    for p in ref.all_premises:
        if p.path not in seen:
            seen.add(p.path)
            order.append(p.path)
    order += [n for n in g.nodes if n not in seen]
    # a file may only import files that precede it; a topological order of the DAG guarantees that
    # while the premise order inside each file is kept.  The reference builds all_premises in file
    # (= topological) order (common.py:202-215), so `order` normally already satisfies it.
    rank = {path: i for i, path in enumerate(order)}
    for path in order:
        for dep in g.successors(path):
            if rank[dep] > rank[path]:
                raise ValueError(f"reference corpus is not in import order: {path} imports {dep}")
    files = []
    for path in order:
        premises = [_premise(p) for p in g.nodes[path]["file"].premises]
        files.append((File(path, premises), list(g.successors(path))))
    corpus = Corpus.from_files(files)
    got = [(p.path, p.full_name, p.start) for p in corpus.all_premises]
    want = [(p.path, p.full_name, _pos(p.start)) for p in ref.all_premises]
    if got != want:
        raise ValueError("converted corpus does not reproduce the reference's all_premises order")
    return corpus


def load_reference_index(path_or_bytes) -> IndexedCorpus:
    """Load an `IndexedCorpus` pickle produced by the reference and convert it."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        fh = io.BytesIO(path_or_bytes)
    else:
        fh = open(path_or_bytes, "rb")
    with fh:
        ref = _Unpickler(fh).load()
    return IndexedCorpus(convert_corpus(ref.corpus), ref.embeddings)
