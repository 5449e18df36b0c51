"""Parity of the fused similarity + top-k path (rpx_sim_topk / rpx_topk_merge through the C ABI)
against the C oracle (`oracle/rpx_oracle.c`): indices and fp64 scores bit-exact under the
(score desc, index asc) contract; fp32 scores == float32(fp64 score)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from reprover_b200 import _native, synth
from reprover_b200.retrieval_ops import sim_topk, topk_merge

pytestmark = pytest.mark.gpu


def _check(Q, E, k, mask_words=None, idx_offset=0):
    dev_mask = None
    if mask_words is not None:
        dev_mask = torch.from_numpy(mask_words.view(np.int32)).to(Q.device)
    s32, idx, cnt, s64 = sim_topk(Q, E, k, access_mask=dev_mask, idx_offset=idx_offset, want_scores64=True)
    torch.cuda.synchronize()
    ws, wi, wc = c_oracle.sim_topk(c_oracle.bf16_bits(Q), c_oracle.bf16_bits(E), k, mask_words, idx_offset)
    gi, gs, gc = idx.cpu().numpy(), s64.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(gc, wc), (gc[:8], wc[:8])
    bad = np.argwhere(gi != wi)
    assert bad.size == 0, f"{len(bad)} index mismatches, first at {bad[:5].tolist()}: got {gi[tuple(bad[0])]} want {wi[tuple(bad[0])]}"
    assert np.array_equal(gs, ws)  # includes -inf padding
    assert np.array_equal(s32.cpu().numpy(), ws.astype(np.float32))
    return gi, gs


def _unit(n, d, seed, dev):
    return synth.random_unit_rows(n, d, seed, dev)


def test_small_exact(rpx_lib, cuda_device):
    _check(_unit(5, 128, 1, cuda_device), _unit(1000, 128, 2, cuda_device), 10)


def test_ragged_multi_qtile_byt5_dim(rpx_lib, cuda_device):
    # 300 queries = 3 query blocks (last one ragged), 20001 premises (ragged last tile), k = 100
    _check(_unit(300, 1472, 3, cuda_device), _unit(20001, 1472, 4, cuda_device), 100, idx_offset=1_000_000)


def test_single_query_like_reference_retrieve(rpx_lib, cuda_device):
    # the reference's retrieve(): Q = 1, k = 100 (retrieval/model.py:338-375)
    _check(_unit(1, 1472, 5, cuda_device), _unit(50_000, 1472, 6, cuda_device), 100)


def test_access_mask_and_short_rows(rpx_lib, cuda_device):
    nq, n, k = 40, 9000, 100
    rng = np.random.default_rng(0)
    m = rng.random((nq, n)) < 0.5
    m[0, :] = False            # nothing accessible
    m[1, :] = False
    m[1, [5, 77, 8999]] = True  # fewer than k accessible -> count 3, tail idx -1
    m[2, :] = True
    words = np.zeros((nq, (n + 31) // 32 * 32), dtype=bool)
    words[:, :n] = m
    words = np.packbits(words.reshape(nq, -1, 8), axis=2, bitorder="little").reshape(nq, -1).view("<u4").copy()
    gi, _ = _check(_unit(nq, 256, 7, cuda_device), _unit(n, 256, 8, cuda_device), k, mask_words=words)
    assert (gi[0] == -1).all() and (gi[1][3:] == -1).all()
    for q in range(3, nq):
        assert m[q, gi[q]].all()


def test_duplicates_tie_break_by_index(rpx_lib, cuda_device):
    base = _unit(50, 192, 9, cuda_device)
    E = base.repeat(60, 1)                      # every row appears 60 times -> massive exact ties
    Q = _unit(9, 192, 10, cuda_device)
    gi, gs = _check(Q, E, 100)
    for q in range(9):
        for r in range(99):
            assert gs[q, r] > gs[q, r + 1] or (gs[q, r] == gs[q, r + 1] and gi[q, r] < gi[q, r + 1])


def test_all_rows_identical(rpx_lib, cuda_device):
    E = _unit(1, 128, 11, cuda_device).repeat(5000, 1)
    gi, _ = _check(_unit(3, 128, 12, cuda_device), E, 100)
    assert (gi == np.arange(100)[None, :]).all()


@pytest.mark.parametrize("ascending", [True, False])
def test_monotone_scores_stress_compaction(rpx_lib, cuda_device, ascending):
    # scores strictly increasing (worst case: every premise enters the list) or decreasing with index
    n, d = 30_000, 128
    u = _unit(1, d, 13, cuda_device).float()
    scale = torch.linspace(0.05, 1.0, n, device=cuda_device)
    if not ascending:
        scale = scale.flip(0)
    E = (scale[:, None] * u).to(torch.bfloat16)
    Q = torch.cat([u, -u, _unit(2, d, 14, cuda_device).float()]).to(torch.bfloat16)
    _check(Q, E, 100)


@pytest.mark.parametrize("k", [1, 3, 101, 200])
def test_k_range(rpx_lib, cuda_device, k):
    _check(_unit(17, 256, 15, cuda_device), _unit(3000, 256, 16, cuda_device), k)


def test_k_larger_than_corpus_and_empty_corpus(rpx_lib, cuda_device):
    gi, _ = _check(_unit(4, 128, 17, cuda_device), _unit(37, 128, 18, cuda_device), 100)
    assert (gi[:, 37:] == -1).all() and (np.sort(gi[:, :37], axis=1) == np.arange(37)).all()
    s, i, c = sim_topk(_unit(4, 128, 17, cuda_device), torch.empty(0, 128, dtype=torch.bfloat16, device=cuda_device), 5)
    assert (i.cpu() == -1).all() and (c.cpu() == 0).all()


def test_rejects_unsupported(rpx_lib, cuda_device):
    Q, E = _unit(2, 128, 1, cuda_device), _unit(10, 128, 2, cuda_device)
    with pytest.raises(ValueError, match="k=5000"):
        sim_topk(Q, E, 5000)   # validated at the Python boundary with a clear message (k <= 1024)
    with pytest.raises(TypeError):
        sim_topk(Q.float(), E.float(), 5)


def test_sharded_equals_single(rpx_lib, cuda_device):
    """Row-sharded index + rpx_topk_merge == one-shot top-k on the concatenated index."""
    nq, d, k, R = 70, 1472, 100, 4
    Q = _unit(nq, d, 19, cuda_device)
    shards = [_unit(5000 + 17 * r, d, 20 + r, cuda_device) for r in range(R)]
    offs = np.concatenate([[0], np.cumsum([s.shape[0] for s in shards])])
    parts = [sim_topk(Q, shards[r], k, idx_offset=int(offs[r]), want_scores64=True) for r in range(R)]
    s64 = torch.stack([p[3] for p in parts])
    idx = torch.stack([p[1] for p in parts])
    ms, mi, mc, ms64 = topk_merge(s64, idx)
    ws, wi, wc = c_oracle.topk_merge(s64.cpu().numpy(), idx.cpu().numpy())
    assert np.array_equal(mi.cpu().numpy(), wi) and np.array_equal(ms64.cpu().numpy(), ws)
    one = sim_topk(Q, torch.cat(shards), k, want_scores64=True)
    assert torch.equal(one[1], mi) and torch.equal(one[3], ms64) and torch.equal(one[0], ms)


def test_full_size_cfg3_properties(rpx_lib, cuda_device):
    """BASELINE config 3 shape (1024 x 200k x 1472, k=100): checked against an fp64 torch matmul
    on the same device (size-independent properties: sortedness, scores reproduce, set equality)."""
    nq, n, d, k = 1024, 200_000, 1472, 100
    Q = _unit(nq, d, 31, cuda_device)
    E = _unit(n, d, 32, cuda_device)
    s32, idx, cnt, s64 = sim_topk(Q, E, k, want_scores64=True)
    torch.cuda.synchronize()
    assert (cnt == k).all()
    assert (s64[:, :-1] >= s64[:, 1:]).all()
    tie = s64[:, :-1] == s64[:, 1:]
    assert (idx[:, :-1][tie] < idx[:, 1:][tie]).all()
    # reference ranking in fp64, 64 queries at a time
    for q0 in range(0, nq, 64):
        S = Q[q0:q0 + 64].double() @ E.double().t()
        top = torch.topk(S, k + 8, dim=1)
        kth = top.values[:, k - 1:k]
        got_scores = torch.gather(S, 1, idx[q0:q0 + 64])
        assert torch.allclose(got_scores, s64[q0:q0 + 64], rtol=0, atol=1e-12)
        # every returned score is >= the true k-th best (up to fp64 summation-order noise)
        assert (got_scores >= kth - 1e-12).all()
        same = (torch.sort(idx[q0:q0 + 64], dim=1).values == torch.sort(top.indices[:, :k], dim=1).values).all(1)
        # rows may differ only where the k-th / (k+1)-th scores coincide to ~1e-13
        gap = (top.values[:, k - 1] - top.values[:, k]).abs()
        assert (same | (gap < 1e-12)).all()
        del S


def test_more_queries_than_one_launch_covers(rpx_lib, cuda_device):
    """nq > 128 * #SMs: the query set is processed in several launches (one CTA per query block,
    no cross-CTA list merging) — results must not depend on the split."""
    n_sms = torch.cuda.get_device_properties(cuda_device).multi_processor_count
    nq = 128 * n_sms + 257
    _check(_unit(nq, 64, 41, cuda_device), _unit(3001, 64, 42, cuda_device), 5)


def test_odd_number_of_query_blocks(rpx_lib, cuda_device):
    _check(_unit(3 * 128 + 5, 128, 43, cuda_device), _unit(70_000, 128, 44, cuda_device), 100)
