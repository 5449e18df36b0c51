"""Golden vectors from the REFERENCE'S OWN CODE for the host data model of the path.

    python tests/golden/make_reference_goldens.py        # needs /root/reference (this container only)

Imports `/root/reference/common.py` unmodified and runs its `Premise.serialize`, `File.from_data`,
`Corpus` (jsonl loading, transitive imports, accessibility) and `Corpus.get_nearest_premises` on a
small synthetic corpus; the inputs and the reference's outputs go to
`tests/golden/reference_host_model.json`, which `tests/test_oracle_cpu.py` replays against
`reprover_b200.corpus` and `oracle/reference_path.py` on any machine.

`common.py` imports packages that are not installed here (lean_dojo, pytorch_lightning, deepspeed).
None of the exercised code touches them except `lean_dojo.Pos`, so they are replaced by inert
stub modules and `Pos` by a stand-in with lean_dojo's fields and ordering (line_nb, column_nb;
lexicographic `<`/`<=`, hashable, iterable) — stated here because it is the one piece of this
generator that is not the reference's own code.
"""
import json
import sys
import tempfile
import types
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import torch

OUT = Path(__file__).resolve().parent / "reference_host_model.json"


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

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    inert = type("Inert", (), {})
    mod("lean_dojo", Pos=Pos)
    mod("pytorch_lightning", LightningModule=inert, Trainer=inert)
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
        logger = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None, debug=lambda *a, **k: None)
        mod("loguru", logger=logger)
    return Pos


def make_corpus_lines():
    """A small corpus that exercises every branch of the host model."""
    rng = np.random.default_rng(3407)
    alphabet = [c for c in range(0x20, 0x7F) if c != 0x3C]
    lines = []
    n_files = 9
    for f in range(n_files):
        path = f"Gold/F{f}.lean"
        # F0, F1 roots; F2 <- F0; F3 <- F1, F2; F4 <- F3 (transitively F0..F2); F5 <- F0; F6 <- F4, F5; F7 root; F8 <- F7, F6
        imports = {2: [0], 3: [1, 2], 4: [3], 5: [0], 6: [4, 5], 8: [7, 6]}.get(f, [])
        prem = []
        for j in range(int(rng.integers(6, 13))):
            body = bytes(rng.choice(alphabet, size=int(rng.integers(5, 60))).tolist()).decode()
            ns = ["Gold", f"F{f}", "Sub"][: int(rng.integers(1, 4))]
            short = f"lemma_{j}" + ("'" if j % 5 == 4 else "")
            full = ".".join(ns + [short])
            style = j % 6
            if style == 0:
                code = f"theorem {full} : {body}"                       # fully qualified spelling
            elif style == 1:
                code = f"theorem {short} (h : {body}) : True := by\n  exact {short} h"   # shortest suffix, twice
            elif style == 2:
                code = f"lemma «{'.'.join(ns[1:] + [short])}» : {body}"   # quoted partial suffix
            elif style == 3:
                code = f"def _root_.{full} := {body}"                   # _root_ spelling
            elif style == 4:
                code = f"instance : Foo := ⟨{body}⟩"                    # name does not occur
            else:
                code = f"theorem x{short} y.{short} ({short}) : {body}"  # occurrences without leading whitespace
            prem.append({"full_name": full, "code": code, "start": [20 * j + 1, 2], "end": [20 * j + 9, 40]})
        if f == 2:   # entries File.from_data drops
            prem.insert(1, {"full_name": None, "code": "x", "start": [500, 0], "end": [501, 0]})
            prem.insert(3, {"full_name": "Gold.user__.n.foo", "code": "x", "start": [502, 0], "end": [503, 0]})
            prem.insert(4, {"full_name": "Gold.empty", "code": "", "start": [504, 0], "end": [505, 0]})
            prem.insert(5, {"full_name": "[Gold.a, Gold.b]", "code": "mutual", "start": [506, 0], "end": [507, 0]})
        if f == 5:   # a second premise with an already used (path, full_name): PremiseSet keeps one entry
            prem.append({"full_name": prem[0]["full_name"], "code": "theorem dup : True", "start": [900, 0], "end": [901, 0]})
        lines.append({"path": path, "imports": [f"Gold/F{i}.lean" for i in imports], "premises": prem})
    return lines


def main():
    Pos = _stub_modules()
    sys.path.insert(0, "/root/reference")
    import common as refc  # the reference, unmodified

    lines = make_corpus_lines()
    with tempfile.TemporaryDirectory() as tmp:
        jsonl = Path(tmp) / "corpus.jsonl"
        jsonl.write_text("\n".join(json.dumps(l) for l in lines))
        corpus = refc.Corpus(str(jsonl))

    index_of = {id(p): i for i, p in enumerate(corpus.all_premises)}
    out = {
        "generator": "tests/golden/make_reference_goldens.py (reference common.py imported unmodified; see its docstring)",
        "corpus_lines": lines,
        "all_premises": [[p.path, p.full_name, list(p.start), list(p.end)] for p in corpus.all_premises],
        "serialize": [p.serialize() for p in corpus.all_premises],
        "dependencies": {l["path"]: sorted(corpus.get_dependencies(l["path"])) for l in lines},
        "num_premises": {l["path"]: corpus.num_premises(l["path"]) for l in lines},
    }
    contexts = []
    for path, pos in [("Gold/F0.lean", (1, 0)), ("Gold/F0.lean", (75, 0)), ("Gold/F3.lean", (49, 40)), ("Gold/F3.lean", (49, 39)),
                      ("Gold/F5.lean", (10_000, 0)), ("Gold/F6.lean", (120, 5)), ("Gold/F8.lean", (64, 0)), ("Gold/F7.lean", (29, 40))]:
        acc = corpus.get_accessible_premises(path, Pos(*pos))
        contexts.append({
            "path": path, "pos": list(pos),
            "accessible_indexes": corpus.get_accessible_premise_indexes(path, Pos(*pos)),
            "member_of_accessible_set": [i for i, p in enumerate(corpus.all_premises) if p in acc],
            "located": (lambda p: None if p is None else index_of[id(p)])(corpus.locate_premise(path, Pos(*pos))),
        })
    out["contexts"] = contexts

    # get_nearest_premises on the CPU in fp32, exactly as the reference computes it
    torch.manual_seed(3407)
    n, d, k = len(corpus.all_premises), 24, 7
    E = torch.nn.functional.normalize(torch.randn(n, d), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(len(contexts), d), dim=1)
    ctxs = [refc.Context(c["path"], "Gold.some_theorem", Pos(*c["pos"]), "h : p\n⊢ q") for c in contexts]
    nearest = {"k": k, "E": E.tolist(), "Q": Q.tolist(), "results": []}
    for j, ctx in enumerate(ctxs):
        try:
            prem, scores = corpus.get_nearest_premises(E, [ctx], Q[j:j + 1], k)
            nearest["results"].append({"indices": [index_of[id(p)] for p in prem[0]], "scores": scores[0]})
        except ValueError:
            nearest["results"].append({"raises": "ValueError"})
    out["nearest"] = nearest
    # the same search on bf16-valued embeddings (what the GPU path holds): pins the C oracle's
    # `rpx_oracle_sim_topk` (bf16 operands, fp64 accumulation, accessibility bitmask) to the reference walk
    E16, Q16 = E.bfloat16().float(), Q.bfloat16().float()
    nearest16 = {"k": k, "results": []}
    for j, ctx in enumerate(ctxs):
        try:
            prem, scores = corpus.get_nearest_premises(E16, [ctx], Q16[j:j + 1], k)
            nearest16["results"].append({"indices": [index_of[id(p)] for p in prem[0]], "scores": scores[0]})
        except ValueError:
            nearest16["results"].append({"raises": "ValueError"})
    out["nearest_bf16"] = nearest16
    # constructor invariants (AssertionError or not) and PremiseSet / equality / hashing semantics
    def outcome(fn):
        try:
            fn()
            return "ok"
        except AssertionError:
            return "AssertionError"
        except Exception as e:  # anything else the reference lets through
            return type(e).__name__

    ctor_cases = [
        ["Context", ["A.lean", "A.t", [1, 0], "h : p\n⊢ q"]],
        ["Context", ["A.lean", "A.t", [1, 0], "no turnstile"]],
        ["Context", ["A.lean", "A.t", [1, 0], "⊢ <a>x</a>"]],
        ["Context", ["A.lean", "A.t", [1, 0], "⊢ x </a>"]],
        ["Context", [7, "A.t", [1, 0], "⊢ q"]],
        ["Context", ["A.lean", None, [1, 0], "⊢ q"]],
        ["Context", ["A.lean", "A.t", None, "⊢ q"]],
        ["Premise", ["A.lean", "A.x", [1, 0], [2, 0], "def x := 1"]],
        ["Premise", ["A.lean", "A.x", [2, 0], [2, 0], "def x := 1"]],
        ["Premise", ["A.lean", "A.x", [2, 1], [2, 0], "def x := 1"]],
        ["Premise", ["A.lean", "A.x", [1, 0], [2, 0], ""]],
        ["Premise", ["A.lean", "A.x", [1, 0], [2, 0], None]],
        ["Premise", ["A.lean", 5, [1, 0], [2, 0], "def x := 1"]],
        ["Premise", ["A.lean", "A.x", None, [2, 0], "def x := 1"]],
    ]

    def build(kind, a):
        if kind == "Context":
            return refc.Context(a[0], a[1], None if a[2] is None else Pos(*a[2]), a[3])
        return refc.Premise(a[0], a[1], None if a[2] is None else Pos(*a[2]), None if a[3] is None else Pos(*a[3]), a[4])

    out["constructors"] = [{"kind": k, "args": a, "outcome": outcome(lambda k=k, a=a: build(k, a))} for k, a in ctor_cases]

    P = lambda path, name, s, e, code: refc.Premise(path, name, Pos(*s), Pos(*e), code)   # noqa: E731
    a1 = P("A.lean", "A.x", (1, 0), (2, 0), "def x := 1")
    a2 = P("A.lean", "A.x", (1, 0), (9, 9), "other code")        # end / code do not take part in ==
    a3 = P("A.lean", "A.x", (5, 0), (6, 0), "def x := 1")        # start does
    b1 = P("B.lean", "A.x", (1, 0), (2, 0), "def x := 1")
    ps = refc.PremiseSet()
    ps.update([a1, b1])
    ps.add(a3)                                                   # same (path, full_name): replaces a1
    out["premise_semantics"] = {
        "a1==a2": a1 == a2, "a1==a3": a1 == a3, "hash(a1)==hash(a2)": hash(a1) == hash(a2),
        "len": len(ps), "a1 in": a1 in ps, "a2 in": a2 in ps, "a3 in": a3 in ps, "b1 in": b1 in ps,
        "iter": [[p.path, p.full_name, list(p.start)] for p in ps],
        "remove_marks": refc.remove_marks("x <a>Nat.add</a> y </a><a>"),
        "context_serialize": refc.Context("A.lean", "A.t", Pos(1, 0), "h : p\n⊢ q").serialize(),
        "context_eq_ignores_pos": refc.Context("A.lean", "A.t", Pos(1, 0), "⊢ q") == refc.Context("A.lean", "A.t", Pos(9, 9), "⊢ q"),
    }
    # Premise.serialize on adversarial (name, code) pairs: un-escaped dots, repeated / overlapping /
    # adjacent occurrences, « » quoting, look-behind at string start and after newlines / NBSP, suffix
    # fallback, names with regex metacharacters (some are invalid patterns: the exception type is recorded)
    rng = np.random.default_rng(11)
    comps = ["Nat", "add", "add_comm", "a", "ab", "foo'", "β", "x1", "comm", "a+b", "f(x", "x*", "[b]", "«q r»", "p.q"]
    glue = [" ", "\n", "\t", ".", "x", "«", "»", "(", ":", "  ", "\u00a0", "_", "a", "Nat", "add", "comm", "ab", "β", "+", "*"]
    ser_cases = []
    for _ in range(400):
        name = ".".join(comps[int(i)] for i in rng.integers(0, len(comps), int(rng.integers(1, 4))))
        pieces = []
        for _ in range(int(rng.integers(1, 12))):
            r = rng.random()
            if r < 0.35:
                pieces.append(name if rng.random() < 0.5 else name.split(".", 1)[-1])
            elif r < 0.5:
                pieces.append(name.replace(".", glue[int(rng.integers(0, len(glue)))]))
            elif r < 0.55:
                pieces.append("_root_." + name)
            else:
                pieces.append(glue[int(rng.integers(0, len(glue)))])
        code = "".join(pieces) or "x"
        try:
            res = {"out": refc.Premise("A.lean", name, Pos(1, 0), Pos(2, 0), code).serialize()}
        except Exception as e:
            res = {"raises": type(e).__name__}
        ser_cases.append({"full_name": name, "code": code, **res})
    out["serialize_cases"] = ser_cases
    OUT.write_text(json.dumps(out, indent=1, ensure_ascii=False))
    print(f"wrote {OUT}: {n} premises, {len(contexts)} contexts")


if __name__ == "__main__":
    main()
