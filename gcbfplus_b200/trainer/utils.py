"""gcbfplus/trainer/utils.py: rollout() and test metrics on the B200 rollout engine."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .data import Rollout
from .rollout import RolloutEngine


def rollout(env, engine: RolloutEngine, actor_params, keys, n_envs: Optional[int] = None) -> Rollout:
    """trainer/utils.py:25-55 for a batch of environments (the reference vmaps it over per-env keys,
    trainer/trainer.py:84-87): key_x0, _ = split(key); reset(key_x0); then T closed-loop steps of (algo.step,
    env.step).  keys: uint32 [E, 2] rollout keys, or an int seed / single key expanded with split(key, E)."""
    from ..utils import jrandom as jr
    n_envs = n_envs or engine.E
    keys = jr.as_key(keys) if not isinstance(keys, np.ndarray) else keys
    if keys.ndim == 1:
        keys = jr.split(keys, n_envs)
    assert keys.shape[0] == n_envs
    g0 = env.reset(jr.split(keys, 2)[:, 0])
    engine.set_params(actor_params)
    engine.set_initial(g0.agent, g0.goal, g0.obstacle)
    engine.run()
    return engine.result()


def eval_metrics(env, ro: Rollout) -> dict:
    """trainer/trainer.py:105-123 eval block."""
    total_reward = ro.rewards.sum(dim=-1)
    b, Tp1, N, sd = ro.agent.shape
    from .. import _lib
    import ctypes as C
    agent = ro.agent.reshape(b * Tp1, N, sd).contiguous()
    goal = ro.goal[:, None].expand(b, Tp1, N, sd).reshape(b * Tp1, N, sd).contiguous()
    d = env.desc(b * Tp1, 0, edge_cap=1)
    fin = torch.empty(b * Tp1 * N, dtype=torch.uint8, device=agent.device)
    _lib.check(env.lib.gcbf_masks(C.byref(d), _lib.ptr(agent), _lib.ptr(goal), None, None, None, None, _lib.ptr(fin),
                                  None, env._stream()), "gcbf_masks")
    finish = fin.reshape(b, Tp1, N)[:, :-1].amax(dim=1).float().mean()
    return {
        "eval/reward": float(total_reward.mean()), "eval/reward_final": float(ro.rewards[:, -1].mean()),
        "eval/cost": float(ro.costs.sum(dim=-1).mean()),
        "eval/unsafe_frac": float((ro.costs.amax(dim=-1) >= 1e-6).float().mean()), "eval/finish": float(finish),
        "reward_min": float(total_reward.min()), "reward_max": float(total_reward.max()),
    }


def test_rates(env, ro: Rollout):
    """test.py:147-198: per-episode safe / finish / success rates from collision & finish masks over
    the T+1 graphs.  Returns (rates [b,3], is_unsafe [b,N], is_finish [b,N])."""
    from ..env.base import RolloutResult
    g = {"agent": ro.agent.transpose(0, 1).contiguous(), "goal": ro.goal, "hits": ro.hits.transpose(0, 1).contiguous(),
         "obstacle": ro.obstacle}
    res = RolloutResult(g, None, None, None, None, {})
    col, fin = env.rollout_masks(res)                      # [T+1, b, N]
    is_unsafe = col.amax(dim=0).float()
    is_finish = fin.amax(dim=0).float()
    rates = torch.stack([1 - is_unsafe.mean(dim=1), is_finish.mean(dim=1), ((1 - is_unsafe) * is_finish).mean(dim=1)],
                        dim=1)
    return rates.cpu().numpy(), is_unsafe.cpu().numpy(), is_finish.cpu().numpy()


def get_bb_cbf(algo, env, agent: torch.Tensor, goal: torch.Tensor, hits: torch.Tensor, agent_id: int,
               x_dim: int = 0, y_dim: int = 1, n_mesh: int = 20, params=None):
    """gcbfplus/trainer/utils.py:149-168 get_bb_cbf for ONE graph (agent / goal [N, sd], hits [N, R, pd]): the CBF of
    agent `agent_id` on an n_mesh x n_mesh grid of its own (x_dim, y_dim) position over [0, area]^2, everything
    else frozen -- same topology, all edge features recomputed with the norm clip (env.add_edge_feats), exactly what
    the reference's vmap(vmap(add_edge_feats)) does.  One batched get_cbf call over the n_mesh^2 copies.
    Returns (b_xs [n_mesh], b_ys [n_mesh], bb_h [n_mesh, n_mesh]) with bb_h[i, j] = h at (b_xs[j], b_ys[i])."""
    from ..utils.graph import SwarmGraph
    N, dev = env.num_agents, env.device
    M = n_mesh * n_mesh
    base = env.get_graph(agent[None].contiguous(), goal[None].contiguous(), None, hits=hits[None].contiguous())
    n_e = base.n_edge                                            # host sync: sizes the tiled edge lists
    base.check_overflow()
    b_xs = torch.from_numpy(np.linspace(0.0, env.area_size, n_mesh).astype(np.float32)).to(dev)
    b_ys = b_xs.clone()
    states = agent[None].repeat(M, 1, 1).contiguous()
    states[:, agent_id, x_dim] = b_xs[None, :].expand(n_mesh, n_mesh).reshape(-1)     # meshgrid 'xy': X[i, j] = xs[j]
    states[:, agent_id, y_dim] = b_ys[:, None].expand(n_mesh, n_mesh).reshape(-1)     #                Y[i, j] = ys[i]
    off_a = (torch.arange(M, device=dev, dtype=torch.int32) * N)[:, None]
    off_e = (torch.arange(M, device=dev, dtype=torch.int32) * n_e)[:, None]
    recv, src = base.edge_recv[:n_e][None], base.edge_src[:n_e][None]
    tiled = SwarmGraph(env, states, goal[None].expand(M, -1, -1).contiguous(), None,
                       hits[None].expand(M, -1, -1, -1).contiguous(),
                       (base.row_start[None] + off_e).reshape(-1).contiguous(), base.row_deg.repeat(M).contiguous(),
                       (recv + off_a).reshape(-1).contiguous(),
                       torch.where(src >= 0, src + off_a, src.expand(M, -1)).reshape(-1).contiguous(),
                       torch.tensor([M * n_e, 0, 0, 0], dtype=torch.int32, device=dev), clip_all=True)
    h = algo.get_cbf(tiled, params)                              # [M, N, 1]
    return b_xs, b_ys, h[:, agent_id, 0].reshape(n_mesh, n_mesh)


def cbf_contours(algo, env, ro: Rollout, episode: int, agent_id: int, n_mesh: int = 20):
    """test.py:125-131 get_bb_cbf_fn over the T+1 graphs of one episode -> (Tb_x [T+1, n], Tb_y [T+1, n],
    Tbb_h [T+1, n, n]) as NumPy arrays (the reference feeds them to its video renderer; here they are saved)."""
    xs, ys, hs = [], [], []
    for t in range(ro.agent.shape[1]):
        x, y, h = get_bb_cbf(algo, env, ro.agent[episode, t], ro.goal[episode], ro.hits[episode, t], agent_id,
                             n_mesh=n_mesh)
        xs.append(x), ys.append(y), hs.append(h)
    return torch.stack(xs).cpu().numpy(), torch.stack(ys).cpu().numpy(), torch.stack(hs).cpu().numpy()
