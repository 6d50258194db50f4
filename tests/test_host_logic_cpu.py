"""Host-side logic that needs no GPU: the paged-KV block bookkeeping (mirror of
T/tensorrt_llm/runtime/kv_cache_manager.py) and the numpy restatement of the sampling kernels, pinned by Random123's
known-answer vectors for Philox4x32-10 and by the sampling distributions it must produce."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ref_ops as R  # noqa: E402


def test_kv_cache_manager_bookkeeping():
    from trtllm_llama_b200.runtime import GenerationSequence, KVCacheManager
    m = KVCacheManager(blocks=12, tokens_per_block=16, max_blocks_per_seq=4)
    a, b = GenerationSequence(0, 0), GenerationSequence(1, 1)
    m.add_sequence(a, 15)       # 15 + 1 positions -> one block
    m.add_sequence(b, 16)       # 16 + 1 -> two blocks
    t = m.get_block_table(2).numpy()
    assert (t[0] >= 0).sum() == 1 and (t[1] >= 0).sum() == 2 and len(m.free_blocks) == 9
    assert m.step([False, False])                       # a: len 15 -> crosses into block 1; b: 16 -> 17, no new block
    t = m.get_block_table(2).numpy()
    assert (t[0] >= 0).sum() == 2 and (t[1] >= 0).sum() == 2
    for _ in range(14):
        m.step([False, False])
    assert m.step([False, True])                        # b finishes: its blocks return to the free list, a is batch idx 0
    assert len(m.sequences) == 1 and m.sequences[0].get_batch_idx() == 0 and len(m.free_blocks) == 12 - len(m.allocated[a])
    with pytest.raises(ValueError):
        KVCacheManager(blocks=4, tokens_per_block=24, max_blocks_per_seq=2)


def test_philox_known_answer_vectors():
    """Random123 kat_vectors for philox4x32 with 10 rounds (the generator behind curand's Philox and tb_sample)."""
    assert R.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert R.philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert R.philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = np.array([R.sampling_uniform(7, s, b) for s in range(40) for b in range(50)], np.float64)
    assert (u > 0).all() and (u <= 1).all() and abs(u.mean() - 0.5) < 0.02 and len(np.unique(u)) == len(u)


def test_sampling_restatement_distributions():
    """top-k / top-p / top-k + top-p of the restated reference on a 6-token distribution (chi-square against the
    renormalised masses the reference's walk implies: the last kept token only gets the mass below the threshold)."""
    base = np.full(64, -20.0, np.float32)
    probs = [0.4, 0.25, 0.15, 0.1, 0.06, 0.04]
    toks = [3, 9, 17, 21, 40, 63]
    base[toks] = np.log(np.array(probs, np.float32))
    n = 2000
    u = np.array([R.sampling_uniform(99, 0, b) for b in range(n)], np.float32)
    for top_k, top_p, kept in ((4, 1.0, 4), (0, 0.7, 3), (6, 0.6, 2), (1, 1.0, 1)):
        ids, _ = R.sample_top_k_top_p(np.tile(base, (n, 1)), top_k, top_p, 1.0, u)
        counts = np.array([(ids == t).sum() for t in toks[:kept]], np.float64)
        assert counts.sum() == n
        mass = np.array(probs[:kept], np.float64)
        limit = top_p * (sum(probs[:top_k]) if top_k > 0 else 1.0)
        mass[-1] = limit - mass[:-1].sum()
        expected = n * mass / limit
        assert float(((counts - expected) ** 2 / np.maximum(expected, 1e-9)).sum()) < 25.0, (top_k, top_p, counts, expected)
    # temperature -> 0 sharpens to greedy
    ids, _ = R.sample_top_k_top_p(np.tile(base, (50, 1)), 6, 1.0, 1e-3, u[:50])
    assert (ids == 3).all()
