"""Shared scene builders for the parity tests (product objects <-> oracle objects)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ENVS_2D = ["SingleIntegrator", "DoubleIntegrator", "DubinsCar"]
ENVS = ENVS_2D + ["LinearDrone"]


def random_scene(env_id: str, n_agents: int, n_graphs: int, area: float, n_obs: int, seed: int,
                 vel_scale: float = 0.4):
    """Random (not collision-free) states: dense enough that neighbours, hits and label
    thresholds are exercised.  Returns numpy arrays."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dims = {"SingleIntegrator": (2, 2), "DoubleIntegrator": (4, 2), "DubinsCar": (4, 2), "LinearDrone": (6, 3)}
    sd, pd = dims[env_id]
    agent = np.zeros((n_graphs, n_agents, sd), dtype=np.float32)
    goal = np.zeros((n_graphs, n_agents, sd), dtype=np.float32)
    agent[..., :pd] = rng.uniform(0, area, size=(n_graphs, n_agents, pd))
    goal[..., :pd] = rng.uniform(0, area, size=(n_graphs, n_agents, pd))
    if env_id == "DubinsCar":
        agent[..., 2] = rng.uniform(-np.pi, np.pi, size=(n_graphs, n_agents))
        agent[..., 3] = rng.uniform(-0.8, 0.8, size=(n_graphs, n_agents))
        goal[..., 2] = np.arctan2(goal[..., 1] - agent[..., 1], goal[..., 0] - agent[..., 0])
    elif sd > pd:
        agent[..., pd:] = rng.uniform(-vel_scale, vel_scale, size=(n_graphs, n_agents, sd - pd))
    obs = {}
    if pd == 2:
        obs = dict(center=rng.uniform(0, area, size=(n_graphs, n_obs, 2)),
                   width=rng.uniform(0.1, 0.6, size=(n_graphs, n_obs)),
                   height=rng.uniform(0.1, 0.6, size=(n_graphs, n_obs)),
                   theta=rng.uniform(0, 2 * np.pi, size=(n_graphs, n_obs)))
    else:
        obs = dict(center=rng.uniform(0, area, size=(n_graphs, n_obs, 3)),
                   radius=rng.uniform(0.075, 0.3, size=(n_graphs, n_obs)))
    return agent, goal, obs


def product_env(env_id, n_agents, area, n_obs, n_rays=None, device="cuda"):
    from gcbfplus_b200.env import make_env
    return make_env(env_id, n_agents, area_size=area, num_obs=n_obs, n_rays=n_rays, device=device)


def product_obstacles(env_id, obs, device="cuda"):
    from gcbfplus_b200.env.obstacle import Rectangle, Sphere
    if "radius" in obs:
        return Sphere.create(obs["center"], obs["radius"], device=device)
    return Rectangle.create(obs["center"], obs["width"], obs["height"], obs["theta"], device=device)


def oracle_env(env_id, n_agents, area, n_obs, n_rays=None, dtype=torch.float32):
    from oracle.envs import OracleEnv
    p = {"n_obs": n_obs}
    if n_rays is not None:
        p["n_rays"] = n_rays
    return OracleEnv(env_id, n_agents, area, params=p, dtype=dtype)


def oracle_obstacles(packed_g: np.ndarray, dtype=torch.float32):
    """Oracle obstacle object for one graph from the product's packed array (identical inputs)."""
    from oracle.geometry import Rectangle, Sphere
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=dtype)
    if packed_g.shape[0] == 0:
        return None
    if packed_g.shape[1] == 4:
        return Sphere(t(packed_g[:, :3]), t(packed_g[:, 3]))
    return Rectangle(center=t(packed_g[:, 0:2]), width=t(packed_g[:, 2] * 2), height=t(packed_g[:, 3] * 2),
                     theta=t(np.zeros(packed_g.shape[0])), points=t(packed_g[:, 6:14].reshape(-1, 4, 2)),
                     cos=t(packed_g[:, 4]), sin=t(packed_g[:, 5]))


def oracle_params(env_id, dtype=torch.float32):
    from oracle.nn import to_torch, unflatten_params
    z = np.load(os.path.join(GOLDEN, f"params_{env_id}.npz"))
    a = {k[6:]: z[k] for k in z.files if k.startswith("actor:")}
    c = {k[4:]: z[k] for k in z.files if k.startswith("cbf:")}
    return to_torch(unflatten_params(a), dtype), to_torch(unflatten_params(c), dtype)


def product_algo(env, env_id=None, seed=0):
    from gcbfplus_b200.algo import make_algo
    algo = make_algo("gcbf+", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=env.num_agents, gnn_layers=1, batch_size=256,
                     buffer_size=512, horizon=32, lr_actor=1e-5, lr_cbf=1e-5, alpha=1.0, eps=0.02, inner_epoch=8,
                     loss_action_coef=1e-4, loss_unsafe_coef=1.0, loss_safe_coef=1.0, loss_h_dot_coef=0.01,
                     max_grad_norm=2.0, seed=seed)
    if env_id is not None:
        algo.load_npz(os.path.join(GOLDEN, f"params_{env_id}.npz"))
    return algo


def edge_sets_product(graph, g: int, n_agents: int):
    """Per receiver: list of sender codes in stored order, for graph g."""
    rs = graph.row_start.cpu().numpy()
    rd = graph.row_deg.cpu().numpy()
    src = graph.edge_src.cpu().numpy()
    recv = graph.edge_recv.cpu().numpy()
    out = []
    for i in range(n_agents):
        a = g * n_agents + i
        codes = src[rs[a]: rs[a] + rd[a]].tolist()
        assert all(r == a for r in recv[rs[a]: rs[a] + rd[a]])
        out.append([c - g * n_agents if c >= 0 else c for c in codes])
    return out


def edge_sets_oracle(og, n_agents: int, n_hits: int):
    """Same from the oracle's sparsified graph: agent j -> j, goal -> -1, hit k -> -2-k."""
    recv = og.receivers.numpy()
    send = og.senders.numpy()
    out = [[] for _ in range(n_agents)]
    goal, agents, hits = [[] for _ in range(n_agents)], [[] for _ in range(n_agents)], [[] for _ in range(n_agents)]
    for r, s in zip(recv, send):
        if s < n_agents:
            agents[r].append(int(s))
        elif s < 2 * n_agents:
            assert s - n_agents == r
            goal[r].append(-1)
        else:
            k = int(s - 2 * n_agents - r * n_hits)
            assert 0 <= k < n_hits
            hits[r].append(-2 - k)
    for i in range(n_agents):
        out[i] = goal[i] + sorted(agents[i]) + sorted(hits[i], reverse=True)
    return out


def probe_reference_stack() -> dict:
    """Can the reference itself run here?  Re-probed on every call (never cached): its JAX stack and a
    driver-provided install under baseline/_ref.  Used by tests/test_reference_goldens.py and by
    `bench.py --impl reference` (which records the outcome in its JSON line)."""
    import importlib
    mods = {}
    for m in ("jax", "flax", "jraph", "optax"):
        try:
            importlib.import_module(m)
            mods[m] = "ok"
        except Exception as e:                       # noqa: BLE001 -- any failure means "cannot run the reference"
            mods[m] = f"missing ({type(e).__name__})"
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    has_ref = os.path.isdir(os.path.join(ref_dir, "gcbfplus"))
    return {"modules": mods, "baseline_ref_installed": has_ref,
            "runnable": all(v == "ok" for v in mods.values())}
