"""Oracle for the GCBF+ action labels u_qp (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Restated from gcbfplus/algo/gcbf_plus.py:299-352 (`get_qp_action`) and :193-211 (`get_b_u_qp`,
chunked over the batch with the TARGET cbf parameters).

PARITY UNPINNED.  The reference solves the QP with JaxProxQP (third-party `jaxproxqp`, unpinned git
dependency in requirements.txt, absent from /root/reference and not installable here).  ProxQP's
published problem form is

    min_x  1/2 x^T H x + g^T x   s.t.  C x <= b,  l_box <= x <= u_box

(the reference's own comment drops the 1/2, but only the 1/2 form makes the unconstrained minimiser
u = u_ref, which is what the label is for).  H is positive definite, so the minimiser is UNIQUE: any
exact solver returns the same u up to its tolerance, and a candidate can be certified by its KKT
residual alone (`kkt_residual`).  That is how this oracle is pinned: `solve_qp_dual` (float64, an
accelerated projected-gradient ascent on the dual) is cross-checked against SciPy's SLSQP active-set
solver on small swarms and against the KKT conditions at every size (tests/test_oracle.py).

Problem data for one graph with N agents, x = [u (N*nu) | r (N)]:
    H = diag(1 ... 1 | 10 ... 10),  g = [-u_ref | 1000 ... 1000]
    C = -[Lg_h | I],  b = Lf_h + 0.1 * alpha * h,   -u_lim <= u <= u_lim,  r >= 0
with h = cbf(add_edge_feats(graph, x)), h_x = dh/dx (dense [N, N, sd] here), Lf_h = h_x . f(x),
Lg_h = h_x . g(x) for the control-affine dynamics xdot = f(x) + g(x) u.

Dual (the form the CUDA kernel iterates on): for multipliers lam >= 0 of C x <= b the inner minimisers
are closed-form because H is diagonal,
    u(lam) = clip(u_ref + Lg_h^T lam, -u_lim, u_lim),   r(lam) = max(0, (lam - 1000) / 10),
and the dual gradient is the constraint residual  -Lg_h u(lam) - r(lam) - b.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from .algo import get_cbf
from .envs import Graph, OracleEnv

RELAX_PENALTY = 1e3   # gcbf_plus.py:302
RELAX_WEIGHT = 10.0   # gcbf_plus.py:331
H_SCALE = 0.1         # gcbf_plus.py:334  (alpha * 0.1 * h)


def control_affine_dyn(env: OracleEnv, agent: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """f [N, sd], g [N, sd, nu].  single_integrator.py:231-238, double_integrator.py:266-273,
    dubins_car.py:243-254 (note g[2,0] = 10 here although the step uses 20), linear_drone.py:255-262."""
    N = agent.shape[0]
    dt = agent.dtype
    if env.env_id == "SingleIntegrator":
        f = torch.zeros_like(agent)
        g = torch.eye(2, dtype=dt)
    elif env.env_id == "DoubleIntegrator":
        f = torch.cat([agent[:, 2:], torch.zeros(N, 2, dtype=dt)], dim=1)
        g = torch.cat([torch.zeros(2, 2, dtype=dt), torch.eye(2, dtype=dt) / env.params["m"]], dim=0)
    elif env.env_id == "DubinsCar":
        f = torch.stack([torch.cos(agent[:, 2]) * agent[:, 3], torch.sin(agent[:, 2]) * agent[:, 3],
                         torch.zeros(N, dtype=dt), torch.zeros(N, dtype=dt)], dim=1)
        g = torch.cat([torch.zeros(2, 2, dtype=dt), torch.tensor([[10.0, 0.0], [0.0, 1.0]], dtype=dt)], dim=0)
    else:
        A = torch.tensor(env._A, dtype=dt)
        f = agent @ A.T
        g = torch.tensor(env._B, dtype=dt)
    return f, g[None].expand(N, -1, -1)


def qp_data(env: OracleEnv, cbf_p: Dict[str, torch.Tensor], g: Graph, alpha: float = 1.0) -> Dict[str, np.ndarray]:
    """gcbf_plus.py:310-334: h, its Jacobian, Lie derivatives and the QP right-hand side (numpy float64)."""
    N = env.num_agents
    rest = g.states[N:]

    def h_aug(agent_state: torch.Tensor) -> torch.Tensor:
        new_graph = env.add_edge_feats(g, torch.cat([agent_state, rest[: g.states.shape[0] - N]], dim=0))
        return get_cbf(cbf_p, new_graph).squeeze(-1)

    x = g.states[:N].detach().clone()
    h = h_aug(x).detach()
    h_x = torch.autograd.functional.jacobian(h_aug, x)            # [N, N, sd]
    f, gm = control_affine_dyn(env, x)
    Lf_h = torch.einsum("ijx,jx->i", h_x, f)
    Lg_h = torch.einsum("ijx,jxu->iju", h_x, gm).reshape(N, -1)   # [N, N*nu]
    u_ref = env.u_ref(g.agent, g.goal).reshape(-1)
    b = Lf_h + alpha * H_SCALE * h
    to = lambda t: t.detach().to(torch.float64).numpy()
    return {"h": to(h), "h_x": to(h_x), "Lf_h": to(Lf_h), "Lg_h": to(Lg_h), "u_ref": to(u_ref), "b": to(b),
            "u_lim": float(env.action_lim()[1][0])}


def primal_from_dual(Lg: np.ndarray, u_ref: np.ndarray, u_lim: float, lam: np.ndarray):
    u = np.clip(u_ref + Lg.T @ lam, -u_lim, u_lim)
    r = np.maximum(0.0, (lam - RELAX_PENALTY) / RELAX_WEIGHT)
    return u, r


def solve_qp_dual(Lg: np.ndarray, b: np.ndarray, u_ref: np.ndarray, u_lim: float, *, tol: float = 1e-11,
                  max_iter: int = 400000):
    """Accelerated projected-gradient ascent on the dual with gradient restart (float64).
    Returns (u [N*nu], r [N], lam [N], iterations)."""
    N = Lg.shape[0]
    s = 1.0 / np.sqrt((Lg * Lg).sum(1) + 1.0 / RELAX_WEIGHT)      # row scaling (preconditioner)
    Ls = Lg * s[:, None]
    lip = np.linalg.norm(Ls, 2) ** 2 + (s * s).max() / RELAX_WEIGHT
    step = 1.0 / lip
    mu = np.zeros(N)
    y = mu.copy()
    t = 1.0
    it = 0
    for it in range(1, max_iter + 1):
        lam = s * y
        u, r = primal_from_dual(Lg, u_ref, u_lim, lam)
        grad = s * (-Lg @ u - r - b)
        mu_new = np.maximum(0.0, y + step * grad)
        res = np.abs(mu_new - y).max() / step
        if np.dot(grad, mu_new - mu) < 0:                          # gradient restart
            t = 1.0
            y = mu_new.copy()
        else:
            t_new = 0.5 * (1 + np.sqrt(1 + 4 * t * t))
            y = mu_new + (t - 1) / t_new * (mu_new - mu)
            t = t_new
        mu = mu_new
        if res < tol:
            break
    lam = s * mu
    u, r = primal_from_dual(Lg, u_ref, u_lim, lam)
    return u, r, lam, it


def kkt_residual(Lg: np.ndarray, b: np.ndarray, u_ref: np.ndarray, u_lim: float, u: np.ndarray, r: np.ndarray,
                 lam: np.ndarray) -> Dict[str, float]:
    """KKT certificate of (u, r, lam) for the strictly convex QP (all zero <=> optimal)."""
    u_s, r_s = primal_from_dual(Lg, u_ref, u_lim, lam)
    slack = -Lg @ u - r - b                                        # C x - b  (must be <= 0)
    return {"stationarity_u": float(np.abs(u - u_s).max()), "stationarity_r": float(np.abs(r - r_s).max()),
            "primal": float(np.maximum(slack, 0).max()), "dual": float(np.maximum(-lam, 0).max()),
            "complementarity": float(np.abs(lam * slack).max())}


def solve_qp_slsqp(Lg: np.ndarray, b: np.ndarray, u_ref: np.ndarray, u_lim: float):
    """Independent active-set solve of the primal (SciPy SLSQP); small N only."""
    from scipy.optimize import minimize
    N, nun = Lg.shape
    hdiag = np.concatenate([np.ones(nun), RELAX_WEIGHT * np.ones(N)])
    gvec = np.concatenate([-u_ref, RELAX_PENALTY * np.ones(N)])
    C = -np.concatenate([Lg, np.eye(N)], axis=1)
    x0 = np.concatenate([np.clip(u_ref, -u_lim, u_lim), np.maximum(0.0, -(Lg @ np.clip(u_ref, -u_lim, u_lim)) - b)])
    res = minimize(lambda x: 0.5 * x @ (hdiag * x) + gvec @ x, x0, jac=lambda x: hdiag * x + gvec, method="SLSQP",
                   bounds=[(-u_lim, u_lim)] * nun + [(0.0, None)] * N,
                   constraints=[{"type": "ineq", "fun": lambda x: b - C @ x, "jac": lambda x: -C}],
                   options={"ftol": 1e-15, "maxiter": 2000})
    return res.x[:nun], res.x[nun:], res


def get_qp_action(env: OracleEnv, cbf_p: Dict[str, torch.Tensor], g: Graph, alpha: float = 1.0):
    """gcbf_plus.py:299-352 -> (u_opt [N, nu], r [N], lam [N], data)."""
    d = qp_data(env, cbf_p, g, alpha)
    u, r, lam, _ = solve_qp_dual(d["Lg_h"], d["b"], d["u_ref"], d["u_lim"])
    return u.reshape(env.num_agents, -1), r, lam, d
