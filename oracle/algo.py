"""Oracle GCBF+ algorithm pieces (TEST INFRASTRUCTURE ONLY), restated from
gcbfplus/algo/gcbf_plus.py (safe_mask :160-174, act/step :176-186, update_inner
:354-447, update_tgt :188-191), trainer/utils.py (rollout :25-55,
compute_norm_and_clip :66-75) and test.py:184-198 (rates).  Gradients come from
torch autograd over the restated loss (float64 capable) -- the reference uses
jax.value_and_grad over the same expression.

optax semantics restated from their public definitions:
adamw(lr, b1=.9, b2=.999, eps=1e-8, weight_decay=1e-3):
  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  mhat = m/(1-b1^t); vhat = v/(1-b2^t)
  p <- p - lr * (mhat / (sqrt(vhat) + eps) + wd * p)
apply_if_finite: if any grad entry is non-finite the update is skipped and the
inner optimizer state is not advanced.  incremental_update(new, old, tau) =
tau*new + (1-tau)*old.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .envs import Graph, OracleEnv
from .nn import net_forward


def act(env: OracleEnv, actor_p, g: Graph) -> torch.Tensor:
    """gcbf_plus.py:176-180: 2 * pi(g) + u_ref(g) (not clipped here)."""
    return 2 * net_forward(actor_p, g, "actor") + env.u_ref(g.agent, g.goal)


def get_cbf(cbf_p, g: Graph) -> torch.Tensor:
    """gcbf.py:209-212 -> [N, 1]."""
    return net_forward(cbf_p, g, "cbf")


def safe_mask_horizon(unsafe_mask: np.ndarray, horizon: int) -> np.ndarray:
    """gcbf_plus.py:160-174.  unsafe_mask [T, N] bool for one rollout -> safe [T, N]:
    safe[t] = no unsafe in [t, t+horizon] ... and safe[0] forced to 1."""
    T = unsafe_mask.shape[0]
    safe = np.ones_like(unsafe_mask, dtype=bool)
    for i in range(T):
        start = 0 if i < horizon else i - horizon
        safe[start:i + 1] = ((1 - unsafe_mask[i].astype(np.int64))[None, :] * safe[start:i + 1]).astype(bool)
        safe[0] = True
    return safe


def gcbf_plus_loss(env: OracleEnv, cbf_p, actor_p, graphs: List[Graph], safe_mask: torch.Tensor,
                   unsafe_mask: torch.Tensor, u_qp: torch.Tensor, *, alpha: float = 1.0, eps: float = 0.02,
                   coef_action: float = 1e-4, coef_unsafe: float = 1.0, coef_safe: float = 1.0,
                   coef_h_dot: float = 0.01, denoms=None) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """gcbf_plus.py:362-431 `get_loss` for one minibatch.
    graphs: B graphs; safe/unsafe_mask [B, N] bool; u_qp [B, N, nu].
    denoms = (n_unsafe, n_safe, n_agents) overrides the local counts: with the GLOBAL counts the
    losses of the shards of a minibatch add up to the full-minibatch loss."""
    dtype = u_qp.dtype
    cbf_ng = {k: v.detach() for k, v in cbf_p.items()}               # stop_gradient(cbf_params)
    h = torch.stack([get_cbf(cbf_p, g).squeeze(-1) for g in graphs]).reshape(-1)
    sm = safe_mask.reshape(-1)
    um = unsafe_mask.reshape(-1)
    one = torch.ones_like(h)
    # unsafe region h < 0
    unsafe_ratio = um.to(dtype).mean()
    h_unsafe = torch.where(um, h, -one * eps * 2)
    n_us = um.sum().to(dtype) if denoms is None else torch.tensor(float(denoms[0]), dtype=dtype)
    n_sf = sm.sum().to(dtype) if denoms is None else torch.tensor(float(denoms[1]), dtype=dtype)
    n_tot = float(h.numel()) if denoms is None else float(denoms[2])
    loss_unsafe = torch.relu(h_unsafe + eps).sum() / (n_us + 1e-6)
    acc_unsafe = ((torch.where(um, h, one) < 0).sum().to(dtype) + 1e-6) / (um.sum().to(dtype) + 1e-6)
    # safe region h > 0
    h_safe = torch.where(sm, h, one * eps * 2)
    loss_safe = torch.relu(-h_safe + eps).sum() / (n_sf + 1e-6)
    acc_safe = ((torch.where(sm, h, -one) > 0).sum().to(dtype) + 1e-6) / (sm.sum().to(dtype) + 1e-6)
    # actions and next graphs
    actions = [act(env, actor_p, g) for g in graphs]
    nxt = [env.forward_graph(g, a) for g, a in zip(graphs, actions)]
    h_next = torch.stack([get_cbf(cbf_p, g).squeeze(-1) for g in nxt]).reshape(-1)
    h_dot = (h_next - h) / env.dt
    h_next_ng = torch.stack([get_cbf(cbf_ng, g).squeeze(-1) for g in nxt]).reshape(-1)
    h_dot_ng = (h_next_ng - h.detach()) / env.dt
    labeled = um | sm
    v = torch.relu(-h_dot - alpha * h + eps)
    v_ng = torch.relu(-h_dot_ng - alpha * h + eps)
    loss_h_dot = torch.where(labeled, v, v_ng).sum() / n_tot
    acc_h_dot = ((h_dot + alpha * h) > 0).to(dtype).mean()
    action = torch.stack(actions)
    loss_action = ((action - u_qp) ** 2).sum(dim=-1).sum() / n_tot
    total = coef_action * loss_action + coef_unsafe * loss_unsafe + coef_safe * loss_safe + coef_h_dot * loss_h_dot
    info = {"loss/action": loss_action, "loss/unsafe": loss_unsafe, "loss/safe": loss_safe,
            "loss/h_dot": loss_h_dot, "loss/total": total, "acc/unsafe": acc_unsafe, "acc/safe": acc_safe,
            "acc/h_dot": acc_h_dot, "acc/unsafe_data_ratio": unsafe_ratio}
    return total, info


def compute_norm_and_clip(grads: Dict[str, torch.Tensor], max_norm: float):
    """trainer/utils.py:62-75: g * max_norm / max(max_norm, ||g||)."""
    g_norm = torch.sqrt(sum((g * g).sum() for g in grads.values()))
    denom = torch.maximum(torch.tensor(max_norm, dtype=g_norm.dtype), g_norm)
    return {k: (g / denom) * max_norm for k, g in grads.items()}, g_norm


class AdamW:
    """optax.apply_if_finite(optax.adamw(lr, weight_decay=1e-3), 1_000_000)
    (gcbf_plus.py:109-110, 127-128)."""

    def __init__(self, params: Dict[str, torch.Tensor], lr: float, wd: float = 1e-3,
                 b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.lr, self.wd, self.b1, self.b2, self.eps = lr, wd, b1, b2, eps
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        if not all(bool(torch.isfinite(g).all()) for g in grads.values()):
            return params
        self.t += 1
        out = {}
        for k, p in params.items():
            g = grads[k]
            self.m[k] = self.b1 * self.m[k] + (1 - self.b1) * g
            self.v[k] = self.b2 * self.v[k] + (1 - self.b2) * g * g
            mhat = self.m[k] / (1 - self.b1 ** self.t)
            vhat = self.v[k] / (1 - self.b2 ** self.t)
            out[k] = p - self.lr * (mhat / (torch.sqrt(vhat) + self.eps) + self.wd * p)
        return out


def train_step(env, cbf_p, actor_p, opt_cbf: AdamW, opt_actor: AdamW, graphs, safe_mask, unsafe_mask, u_qp,
               max_grad_norm: float = 2.0, **loss_kw):
    """One `update_fn` of gcbf_plus.py:356-441.  Returns (cbf_p, actor_p, info, raw grads)."""
    cbf_p = {k: v.detach().clone().requires_grad_(True) for k, v in cbf_p.items()}
    actor_p = {k: v.detach().clone().requires_grad_(True) for k, v in actor_p.items()}
    total, info = gcbf_plus_loss(env, cbf_p, actor_p, graphs, safe_mask, unsafe_mask, u_qp, **loss_kw)
    names_c, names_a = list(cbf_p), list(actor_p)
    gs = torch.autograd.grad(total, [cbf_p[k] for k in names_c] + [actor_p[k] for k in names_a], allow_unused=True)
    gc = {k: (g if g is not None else torch.zeros_like(cbf_p[k])) for k, g in zip(names_c, gs[:len(names_c)])}
    ga = {k: (g if g is not None else torch.zeros_like(actor_p[k])) for k, g in zip(names_a, gs[len(names_c):])}
    gc_c, n_c = compute_norm_and_clip(gc, max_grad_norm)
    ga_c, n_a = compute_norm_and_clip(ga, max_grad_norm)
    new_c = opt_cbf.step({k: v.detach() for k, v in cbf_p.items()}, gc_c)
    new_a = opt_actor.step({k: v.detach() for k, v in actor_p.items()}, ga_c)
    info = {k: float(v) for k, v in info.items()}
    info["grad_norm/cbf"], info["grad_norm/actor"] = float(n_c), float(n_a)
    return new_c, new_a, info, (gc, ga)


def polyak(new: Dict[str, torch.Tensor], old: Dict[str, torch.Tensor], tau: float = 0.5):
    """gcbf_plus.py:188-191 / optax.incremental_update."""
    return {k: tau * new[k] + (1 - tau) * old[k] for k in new}


# --------------------------------------------------------------------------- rollout + metrics
def rollout(env: OracleEnv, actor_p, agent0, goal0, obstacles, T: Optional[int] = None, sparse: bool = True):
    """trainer/utils.py:25-55 closed loop from given initial conditions (reset is not
    restated against JAX's PRNG).  Returns dict of stacked arrays."""
    T = T or env.max_episode_steps
    g = env.get_graph(agent0, goal0, obstacles)
    states, actions, rewards, costs, unsafe, collide, finish, lidars = [], [], [], [], [], [], [], []
    with torch.no_grad():
        for _ in range(T):
            gs = env.sparsify(g) if sparse else g
            a = act(env, actor_p, gs)
            states.append(g.agent)
            lidars.append(g.states[2 * env.num_agents:-1].reshape(env.num_agents, env.n_hits, -1))
            unsafe.append(env.unsafe_mask(g))
            collide.append(env.collision_mask(g))
            finish.append(env.finish_mask(g))
            g, r, c = env.step(g, a)
            actions.append(a)
            rewards.append(r)
            costs.append(c)
        states.append(g.agent)
        collide.append(env.collision_mask(g))
        finish.append(env.finish_mask(g))
    return {"states": torch.stack(states), "actions": torch.stack(actions), "rewards": torch.stack(rewards),
            "costs": torch.stack(costs), "unsafe": torch.stack(unsafe), "collision": torch.stack(collide),
            "finish": torch.stack(finish), "lidar": torch.stack(lidars), "final_graph": g}


def rates(collision: np.ndarray, finish: np.ndarray):
    """test.py:184-186 for one episode: collision/finish [T+1, N] bool ->
    (safe_rate, finish_rate, success_rate)."""
    unsafe = collision.max(axis=0)
    fin = finish.max(axis=0)
    return float(1 - unsafe.mean()), float(fin.mean()), float(((1 - unsafe) * fin).mean())


def get_bb_cbf(env: OracleEnv, cbf_p, g: Graph, agent_id: int, x_dim: int = 0, y_dim: int = 1, n_mesh: int = 20):
    """trainer/utils.py:149-168: h of agent `agent_id` over an n_mesh x n_mesh grid of its (x, y), other states
    frozen, topology of `g`, edge features through add_edge_feats.  bb_h[i, j] is at (b_xs[j], b_ys[i])."""
    b_xs = np.linspace(0.0, env.area_size, n_mesh).astype(np.float32)
    b_ys = np.linspace(0.0, env.area_size, n_mesh).astype(np.float32)
    out = np.zeros((n_mesh, n_mesh), dtype=np.float32)
    with torch.no_grad():
        for i in range(n_mesh):
            for j in range(n_mesh):
                st = g.states[:-1].clone()
                st[agent_id, x_dim] = float(b_xs[j])
                st[agent_id, y_dim] = float(b_ys[i])
                out[i, j] = float(get_cbf(cbf_p, env.add_edge_feats(g, st))[agent_id, 0])
    return b_xs, b_ys, out
