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
