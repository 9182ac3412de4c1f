"""Host-side logic of the hot path's callers (CPU): replay buffers, key plumbing, env descriptors."""
import numpy as np
import torch

from gcbfplus_b200.trainer.buffer import MaskedReplayBuffer


def _flat(n_graphs, tag, N=3):
    """n_graphs records whose content identifies (tag, index)."""
    base = torch.arange(n_graphs, dtype=torch.float32) + 100.0 * tag
    return {"agent": base[:, None, None].expand(n_graphs, N, 2).clone(), "goal": torch.zeros(n_graphs, N, 2),
            "hits": torch.zeros(n_graphs, N, 4, 2), "safe": torch.ones(n_graphs, N, dtype=torch.uint8),
            "unsafe": torch.zeros(n_graphs, N, dtype=torch.uint8)}


def test_rollout_buffer_is_a_fifo_of_whole_rollouts():
    """gcbfplus/trainer/buffer.py:66-85: append keeps the newest `size` rollouts; sampling returns whole
    rollouts (T consecutive graphs of one stored rollout), with replacement."""
    T, size = 4, 3
    buf = MaskedReplayBuffer(size=size)
    for tag in range(5):                                   # 5 appends of 1 rollout each: tags 2, 3, 4 survive
        buf.append_rollouts(_flat(T, tag), 1, T)
    assert buf.n_items == size and buf.length == size * T
    np.random.seed(0)
    out = buf.sample_rollouts(64)
    a = out["agent"][:, 0, 0].reshape(64, T)
    tags = (a[:, 0] // 100).long()
    assert set(tags.tolist()) == {2, 3, 4}
    assert torch.equal(a - 100.0 * tags[:, None].float(), torch.arange(T, dtype=torch.float32).expand(64, T))


def test_unsafe_buffer_keeps_only_masked_graphs_and_caps_at_size():
    """gcbf_plus.py:246-251: unsafe_buffer.append(rollout[unsafe_multi_mask]); capacity counts graphs."""
    buf = MaskedReplayBuffer(size=5)
    mask = torch.tensor([1, 0, 1, 1, 0, 0, 1, 0], dtype=torch.bool)
    buf.append_graphs(_flat(8, 0), mask)
    assert buf.length == 4
    assert buf.get_data(np.arange(4))["agent"][:, 0, 0].tolist() == [0.0, 2.0, 3.0, 6.0]
    buf.append_graphs(_flat(8, 1), mask)                   # 8 masked graphs in total -> newest 5 kept
    assert buf.length == 5
    assert buf.get_data(np.arange(5))["agent"][:, 0, 0].tolist() == [6.0, 100.0, 102.0, 103.0, 106.0]
    np.random.seed(1)
    s = buf.sample_graphs(200)["agent"][:, 0, 0]
    assert set(s.tolist()) <= {6.0, 100.0, 102.0, 103.0, 106.0} and len(set(s.tolist())) == 5


def test_env_descriptor_thresholds_are_exact_fp32_images():
    """The host rounds python-float constants where JAX's weak typing does and precomputes the sqrt-free
    thresholds: (sqrtf(x) < R) == (x < comm_sq_thr) for every fp32 x around R^2."""
    from gcbfplus_b200 import _lib
    R = np.float32(0.5)
    thr = np.float32(_lib.sqrt_threshold(float(R)))
    x = np.float32(R) * np.float32(R)
    cand = [np.nextafter(x, np.float32(0), dtype=np.float32), x, np.nextafter(x, np.float32(1), dtype=np.float32)]
    for _ in range(40):
        cand.append(np.nextafter(cand[-1], np.float32(1), dtype=np.float32))
        cand.insert(0, np.nextafter(cand[0], np.float32(0), dtype=np.float32))
    for c in cand:
        assert (np.sqrt(np.float32(c)) < R) == (np.float32(c) < thr), (c, thr)


def test_epoch_label_counts_match_a_per_minibatch_loop():
    """algo/train.py _minibatch_counts: the (n_unsafe, n_safe, n_agents) triples of every minibatch of an epoch, computed
    on the device without a loop (per-graph counts -> permutation -> cumulative sums at the np.array_split points), equal
    the per-minibatch counts gcbf_plus.py:377,384 uses as denominators; and the counts of the two halves of every
    minibatch (what two ranks hold) add up to it -- the identity behind ONE count all-reduce per epoch."""
    from gcbfplus_b200.algo.train import _minibatch_counts
    rng = np.random.default_rng(3)
    n, N, n_mb = 53, 7, 5
    batch = {"safe": torch.from_numpy((rng.uniform(size=(n, N)) < 0.5).astype(np.uint8)),
             "unsafe": torch.from_numpy((rng.uniform(size=(n, N)) < 0.2).astype(np.uint8))}
    idx = torch.from_numpy(rng.permutation(n))
    splits = np.array_split(np.arange(n), n_mb)
    bounds = np.concatenate([[0], np.cumsum([len(m) for m in splits])]).astype(np.int64)
    got = _minibatch_counts(batch, idx, bounds)
    assert got.shape == (n_mb, 4)
    for i, mb in enumerate(splits):
        sel = idx[torch.from_numpy(mb)]
        want = [float(batch["unsafe"][sel].sum()), float(batch["safe"][sel].sum()), float(len(mb) * N), 0.0]
        assert got[i].tolist() == want, (i, got[i].tolist(), want)
    # shards: rank r holds graphs [lo_r, hi_r) of every minibatch's selection -> counts add up
    for i, mb in enumerate(splits):
        sel = idx[torch.from_numpy(mb)]
        half = len(sel) // 2
        parts = [_minibatch_counts({k: v[s] for k, v in batch.items()}, torch.arange(len(s)), np.array([0, len(s)]))
                 for s in (sel[:half], sel[half:])]
        assert torch.equal(parts[0][0] + parts[1][0], got[i])
