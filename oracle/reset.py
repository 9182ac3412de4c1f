"""Oracle for env.reset (TEST INFRASTRUCTURE ONLY): the reference's scenario sampler, one environment at a
time, scalar Python, independent of gcbfplus_b200/utils/jrandom.py.

Restated from gcbfplus/env/double_integrator.py:83-112 (SingleIntegrator :76-105, DubinsCar :72-100,
LinearDrone :91-116) and gcbfplus/env/utils.py:134-226 (get_node_goal_rng); random draws follow jax.random's
public threefry algorithm (jax/_src/prng.py, 0.4.x, `jax_threefry_partitionable` off): see
gcbfplus_b200/utils/jrandom.py for the description.  Pinned by known answers: Random123's Threefry-2x32-20
vectors and the values jax prints for split(PRNGKey(0)), split(PRNGKey(42)), uniform(PRNGKey(0))
(tests/test_oracle.py).  PARITY UNPINNED beyond those: the reference's reset itself cannot run here (no JAX).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np

M32 = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
# jax_threefry_partitionable (False: the 0.4.x default the reference was written on; True: default from JAX 0.5.0).
# Partitionable layout (jax/_src/prng.py _threefry_split_foldlike / _threefry_random_bits_partitionable): element i
# uses the counter pair (hi(i), lo(i)) of its 64-bit row-major index; split keeps the output pair, random bits
# are y0 ^ y1.  Tests flip this together with gcbfplus_b200.utils.jrandom.set_partitionable.
PARTITIONABLE = False


def threefry2x32(k0: int, k1: int, c0: int, c1: int) -> Tuple[int, int]:
    ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0, x1 = (c0 + ks[0]) & M32, (c1 + ks[1]) & M32
    for g in range(5):
        for r in _ROT[g % 2]:
            x0 = (x0 + x1) & M32
            x1 = ((x1 << r) | (x1 >> (32 - r))) & M32
            x1 ^= x0
        x0 = (x0 + ks[(g + 1) % 3]) & M32
        x1 = (x1 + ks[(g + 2) % 3] + g + 1) & M32
    return x0, x1


def _bits(key: Tuple[int, int], n: int) -> List[int]:
    """threefry_2x32(key, iota(n)): counters split in halves (zero padded when n is odd)."""
    cnt = list(range(n)) + ([0] if n % 2 else [])
    h = len(cnt) // 2
    out0, out1 = [], []
    for a, b in zip(cnt[:h], cnt[h:]):
        y0, y1 = threefry2x32(key[0], key[1], a, b)
        out0.append(y0)
        out1.append(y1)
    return (out0 + out1)[:n]


def prng_key(seed: int) -> Tuple[int, int]:
    return ((seed >> 32) & M32, seed & M32)


def split(key: Tuple[int, int], num: int = 2) -> List[Tuple[int, int]]:
    if PARTITIONABLE:
        return [threefry2x32(key[0], key[1], 0, i) for i in range(num)]
    b = _bits(key, 2 * num)
    return [(b[2 * i], b[2 * i + 1]) for i in range(num)]


def uniform(key: Tuple[int, int], shape: Tuple[int, ...], minval: float, maxval: float) -> np.ndarray:
    n = int(np.prod(shape)) if len(shape) else 1
    f = np.float32
    if PARTITIONABLE:
        pairs = [threefry2x32(key[0], key[1], 0, i) for i in range(n)]
        bits = np.array([a ^ b for a, b in pairs], dtype=np.uint32)
    else:
        bits = np.array(_bits(key, n), dtype=np.uint32)
    fl = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - f(1.0)
    lo, hi = f(minval), f(maxval)
    return np.maximum(lo, fl * (hi - lo) + lo).astype(f).reshape(shape)


def inside_obstacles(p: np.ndarray, obs: Optional[dict], r: float) -> bool:
    """env/obstacle.py:53-96 (Rectangle) / :234-270 (Sphere) for one point, fp32."""
    f = np.float32
    r = f(r)
    if obs is None or len(obs["center"]) == 0:
        return False
    if "radius" in obs:
        d = np.sqrt(((p[None, :] - obs["center"]) ** 2).sum(-1, dtype=f))
        return bool((d <= obs["radius"] + r).any())
    rel = p[None, :] - obs["center"]
    c, s = obs["cos"], obs["sin"]
    xx = np.abs(rel[:, 0] * c + rel[:, 1] * s) - obs["width"] / f(2)
    yy = np.abs(rel[:, 0] * s - rel[:, 1] * c) - obs["height"] / f(2)
    is_in = ((xx < r) & (yy < 0)) | ((xx < 0) & (yy < r)) | ((xx > 0) & (yy > 0) & (np.sqrt(xx ** 2 + yy ** 2) < r))
    return bool(is_in.any())


def get_node_goal_rng(key, side_length: float, dim: int, obs: Optional[dict], n: int, min_dist: float,
                      max_travel: Optional[float] = None):
    """env/utils.py:134-226."""
    f = np.float32
    max_iter = 1024
    min_dist = f(min_dist)
    states = np.zeros((n, dim), dtype=f)
    goals = np.zeros((n, dim), dtype=f)
    agent_id, this_key = 0, key

    def dmin(all_pts, p):
        return np.sqrt(((all_pts - p[None, :]) ** 2).sum(-1)).min()

    while agent_id < n:
        agent_key, goal_key, this_key = split(this_key, 3)
        cand = uniform(agent_key, (dim,), 0, side_length)
        it_a, k = 0, agent_key
        while (dmin(states, cand) <= min_dist or inside_obstacles(cand, obs, min_dist)) and it_a < max_iter:
            use, k = split(k, 2)
            it_a += 1
            cand = uniform(use, (dim,), 0, side_length)
        states[agent_id] = cand
        if max_travel is None:
            g = uniform(goal_key, (dim,), 0, side_length)
        else:
            g = uniform(goal_key, (dim,), 0, max_travel) + cand
        it_g, k = 0, goal_key
        while True:
            bad = dmin(goals, g) <= min_dist or inside_obstacles(g, obs, min_dist)
            bad = bad or bool((g < 0).any() or (g > f(side_length)).any())
            if max_travel is not None:
                bad = bad or bool(np.sqrt(((g - cand) ** 2).sum()) > f(max_travel))
            if not bad or it_g >= max_iter:
                break
            use, k = split(k, 2)
            it_g += 1
            if max_travel is None:
                g = uniform(use, (dim,), 0, side_length)
            else:
                g = uniform(use, (dim,), -max_travel, max_travel) + cand
        goals[agent_id] = g
        agent_id += 1
        if it_a >= max_iter or it_g >= max_iter:
            agent_id = 0
            states[:] = 0
            goals[:] = 0
    return states, goals


def reset(env_id: str, key, n_agents: int, area_size: float, n_obs: int, obs_len_range, radius: float,
          max_travel: Optional[float] = None):
    """-> dict(agent [N, sd], goal [N, sd], obstacle parameter arrays)."""
    f = np.float32
    L = area_size
    lo, hi = obs_len_range
    obstacle_key, key = split(key, 2)
    if env_id == "LinearDrone":
        pos = uniform(obstacle_key, (n_obs, 3), 0, L)
        r_key, key = split(key, 2)
        rad = uniform(r_key, (n_obs,), lo / 2, hi / 2)
        obs = {"center": pos, "radius": rad}
        dim, sd = 3, 6
    else:
        pos = uniform(obstacle_key, (n_obs, 2), 0, L)
        length_key, key = split(key, 2)
        ln = uniform(length_key, (n_obs, 2), lo, hi)
        theta_key, key = split(key, 2)
        th = uniform(theta_key, (n_obs,), 0, 2 * math.pi)
        obs = {"center": pos, "width": ln[:, 0], "height": ln[:, 1], "theta": th, "cos": np.cos(th).astype(f),
               "sin": np.sin(th).astype(f)}
        dim, sd = 2, (2 if env_id == "SingleIntegrator" else 4)
    states, goals = get_node_goal_rng(key, L, dim, obs, n_agents, 4 * radius, max_travel)
    agent = np.zeros((n_agents, sd), dtype=f)
    goal = np.zeros((n_agents, sd), dtype=f)
    agent[:, :dim], goal[:, :dim] = states, goals
    if env_id == "DubinsCar":
        theta_key, key = split(key, 2)
        agent[:, 2] = uniform(theta_key, (n_agents,), -math.pi, math.pi)
        goal[:, 2] = np.arctan2(goal[:, 1] - agent[:, 1], goal[:, 0] - agent[:, 0])
    return {"agent": agent, "goal": goal, "obs": obs}
