"""CPU, world_size 2, gloo: the sharding scheme of the train step.  Each rank takes a
contiguous shard of the minibatch, the label counts are all-reduced first (global denominators),
every rank evaluates its shard's loss/gradient with those denominators, and ONE all-reduce of the
packed gradient reproduces the single-process full-minibatch gradient.  The arithmetic is the
oracle's (no GPU here); the host-side pieces under test are gcbfplus_b200.dist."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle_env, oracle_obstacles, oracle_params, product_obstacles, random_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    env_id, N, B, area, n_obs = "DoubleIntegrator", 4, 5, 1.2, 2
    agent, goal, obs = random_scene(env_id, N, B, area, n_obs, seed=13)
    pobs = product_obstacles(env_id, obs, device="cpu")
    oenv = oracle_env(env_id, N, area, n_obs, dtype=torch.float64)
    ap, cp = oracle_params(env_id, dtype=torch.float64)
    graphs = [oenv.sparsify(oenv.get_graph(torch.from_numpy(agent[g]).double(), torch.from_numpy(goal[g]).double(),
                                           oracle_obstacles(pobs.packed.numpy()[g], torch.float64))) for g in range(B)]
    rng = np.random.default_rng(0)
    unsafe = torch.stack([oenv.unsafe_mask(g) for g in graphs])
    safe = (~unsafe) & torch.from_numpy(rng.uniform(size=(B, N)) < 0.5)
    u_qp = torch.from_numpy(rng.normal(size=(B, N, 2)))
    return oenv, ap, cp, graphs, safe, unsafe, u_qp


def _grads(oenv, ap, cp, graphs, safe, unsafe, u_qp, denoms=None):
    from oracle.algo import gcbf_plus_loss
    cp = {k: v.clone().requires_grad_(True) for k, v in cp.items()}
    ap = {k: v.clone().requires_grad_(True) for k, v in ap.items()}
    total, _ = gcbf_plus_loss(oenv, cp, ap, graphs, safe, unsafe, u_qp, coef_action=0.05, coef_h_dot=0.3, denoms=denoms)
    gs = torch.autograd.grad(total, list(cp.values()) + list(ap.values()), allow_unused=True)
    flat = [g.reshape(-1) if g is not None else torch.zeros_like(p).reshape(-1)
            for g, p in zip(gs, list(cp.values()) + list(ap.values()))]
    return torch.cat(flat + [total.detach().reshape(1)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from gcbfplus_b200 import dist as gd
    r, _, w = gd.init_from_env(backend="gloo")
    oenv, ap, cp, graphs, safe, unsafe, u_qp = _problem()
    lo, hi = gd.shard_bounds(len(graphs), r, w)
    counts = torch.tensor([float(unsafe[lo:hi].sum()), float(safe[lo:hi].sum()), float((hi - lo) * safe.shape[1])],
                          dtype=torch.float64)
    gd.allreduce_sum_(counts)                                  # exchange (1): global denominators
    packed = _grads(oenv, ap, cp, graphs[lo:hi], safe[lo:hi], unsafe[lo:hi], u_qp[lo:hi], denoms=counts.tolist())
    gd.allreduce_sum_(packed)                                  # exchange (2): one packed all-reduce
    if r == 0:
        torch.save({"packed": packed, "counts": counts}, out)
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from gcbfplus_b200.dist import shard_bounds
    for n in (1, 5, 16, 257):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_sharded_gradient_equals_full_batch(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    oenv, ap, cp, graphs, safe, unsafe, u_qp = _problem()
    want = _grads(oenv, ap, cp, graphs, safe, unsafe, u_qp)
    assert got["counts"].tolist() == [float(unsafe.sum()), float(safe.sum()), float(safe.numel())]
    torch.testing.assert_close(got["packed"], want, atol=1e-12, rtol=1e-9)
    assert float(want[:-1].abs().max()) > 0


def test_env_sharded_reset_keys_reproduce_the_unsharded_batch():
    """rank r of W rolls out environments [lo, hi) of the global batch with the same per-env threefry
    keys the single-process run uses (trainer.py:134-136), so the union of the shards is the unsharded batch."""
    import numpy as np
    from gcbfplus_b200.dist import shard_bounds
    from gcbfplus_b200.env import make_env
    from gcbfplus_b200.utils import jrandom as jr
    env = make_env("DoubleIntegrator", 6, area_size=2.5, num_obs=3, device="cpu")
    key_x0, _ = jr.split(jr.PRNGKey(3))
    keys = jr.split(key_x0, 6)
    reset_keys = jr.split(keys, 2)[:, 0]

    def sample(k):
        obstacles, k2 = env._sample_obstacles(k)
        s, g = env._sample_agents_goals(k2, obstacles.packed.numpy())
        return obstacles.packed.numpy(), s, g
    full = sample(reset_keys)
    for world in (2, 3):
        parts = []
        for rank in range(world):
            lo, hi = shard_bounds(6, rank, world)
            parts.append(sample(reset_keys[lo:hi]))
        for q in range(3):
            assert np.array_equal(np.concatenate([p[q] for p in parts], axis=0), full[q])
