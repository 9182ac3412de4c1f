"""Oracle environments (TEST INFRASTRUCTURE ONLY): SingleIntegrator, DoubleIntegrator,
DubinsCar, LinearDrone restated from gcbfplus/env/*.py on torch-CPU tensors.

One un-batched environment per call (the reference is written for one env and
vmapped, trainer/trainer.py:84-87).  Two graph forms are produced:

* ``get_graph``  -- the reference's dense padded layout, node order
  [agents | goals | hit nodes | pad], edges = concatenated dense blocks with
  masked edges redirected to the pad node (utils/graph.py:35-44, 209-244);
* ``sparsify``   -- the same graph with the masked edges dropped (what the
  CUDA path computes).  tests/test_oracle_dense_sparse.py proves both give the
  same agent outputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Optional

import numpy as np
import scipy.linalg
import torch

from .geometry import (Rectangle, Sphere, exact_sqrt, get_lidar_all, inside_obstacles, ray_table_2d,
                       ray_table_3d)

AGENT, GOAL, OBS = 0, 1, 2


def lqr(A, B, Q, R):
    """env/utils.py:24-46."""
    X = scipy.linalg.solve_discrete_are(A, B, Q, R)
    return scipy.linalg.inv(B.T @ X @ B + R) @ (B.T @ X @ A)


@dataclass
class Graph:
    """Restatement of utils/graph.py:47-186 GraphsTuple (single graph)."""
    nodes: torch.Tensor       # [n_nodes, 3]
    edges: torch.Tensor       # [n_edges, ed]
    states: torch.Tensor      # [n_nodes, sd]
    receivers: torch.Tensor   # [n_edges] int64
    senders: torch.Tensor     # [n_edges] int64
    node_type: torch.Tensor   # [n_nodes] int64 (pad = -1)
    agent: torch.Tensor       # env_states.agent [N, sd]
    goal: torch.Tensor        # env_states.goal [N, sd]
    obstacles: object         # env_states.obstacle
    n_agents: int
    n_hits: int               # hit nodes per agent (R)

    def type_states(self, type_idx: int) -> torch.Tensor:
        """utils/graph.py:126-139 (rows of `states` whose node_type == type_idx)."""
        return self.states[self.node_type == type_idx]


class OracleEnv:
    """Generic restatement; env-specific pieces are selected on ``env_id``.

    Reference: env/single_integrator.py, env/double_integrator.py,
    env/dubins_car.py, env/linear_drone.py, env/base.py:86-88 (clip_action),
    env/__init__.py:23-46 (dt = 0.03, max_step = 256).
    """

    DEFAULT_PARAMS = {
        "SingleIntegrator": {"car_radius": 0.05, "comm_radius": 0.5, "n_rays": 32,
                             "obs_len_range": [0.1, 0.6], "n_obs": 8},
        "DoubleIntegrator": {"car_radius": 0.05, "comm_radius": 0.5, "n_rays": 32,
                             "obs_len_range": [0.1, 0.5], "n_obs": 8, "m": 0.1},
        "DubinsCar": {"car_radius": 0.05, "comm_radius": 0.5, "n_rays": 16,
                      "obs_len_range": [0.1, 0.6], "n_obs": 8},
        "LinearDrone": {"drone_radius": 0.05, "comm_radius": 0.5, "n_rays": 32,
                        "obs_len_range": [0.15, 0.3], "n_obs": 4},
    }
    DIMS = {  # state_dim, edge_dim, action_dim, pos_dim
        "SingleIntegrator": (2, 2, 2, 2), "DoubleIntegrator": (4, 4, 2, 2),
        "DubinsCar": (4, 4, 2, 2), "LinearDrone": (6, 6, 3, 3),
    }

    def __init__(self, env_id: str, num_agents: int, area_size: float, params: Optional[dict] = None,
                 dt: float = 0.03, max_step: int = 256, dtype=torch.float32):
        assert env_id in self.DIMS, env_id
        self.env_id = env_id
        self.num_agents = num_agents
        self.area_size = area_size
        self.dt = dt
        self.max_episode_steps = max_step
        self.dtype = dtype
        self.params = dict(self.DEFAULT_PARAMS[env_id])
        if params:
            self.params.update(params)
        self.state_dim, self.edge_dim, self.action_dim, self.pos_dim = self.DIMS[env_id]
        self.node_dim = 3
        self.r = self.params.get("car_radius", self.params.get("drone_radius"))
        self.comm_radius = self.params["comm_radius"]
        nb = self.params["n_rays"]
        if env_id == "LinearDrone":
            self.ray_table = ray_table_3d(nb, self.comm_radius, dtype)
            self.n_hits = 16                                    # linear_drone.py:73
        else:
            self.ray_table = ray_table_2d(nb, self.comm_radius, dtype)
            self.n_hits = min(nb, 32)                           # env/utils.py:49 max_returns=32
        self.K = None
        sd, nu = self.state_dim, self.action_dim
        if env_id == "SingleIntegrator":                        # single_integrator.py:53-59
            A = np.zeros((sd, sd), dtype=np.float32) * dt + np.eye(sd)
            B = np.array([[1.0, 0.0], [0.0, 1.0]]) * dt
            self.K = lqr(A, B, np.eye(sd) * 2, np.eye(nu))
        elif env_id == "DoubleIntegrator":                      # double_integrator.py:53-66
            A = np.zeros((sd, sd), dtype=np.float32)
            A[0, 2] = 1.0
            A[1, 3] = 1.0
            A = A * dt + np.eye(sd)
            m = self.params["m"]
            B = np.array([[0.0, 0.0], [0.0, 0.0], [1.0 / m, 0.0], [0.0, 1.0 / m]]) * dt
            self.K = lqr(A, B, np.eye(sd) * 5, np.eye(nu))
        elif env_id == "LinearDrone":                           # linear_drone.py:55-72
            A = np.zeros((sd, sd))
            A[0, 3] = A[1, 4] = A[2, 5] = 1.0
            A[3, 3] = A[4, 4] = -1.1
            A[5, 5] = -6.0
            B = np.zeros((sd, nu))
            B[3, 0] = B[4, 1] = B[5, 2] = 10.0
            self._A, self._B = A, B
            self.K = lqr(scipy.linalg.expm(A * dt), B, np.diag([5e1, 5e1, 5e1, 1.0, 1.0, 1.0]), np.eye(nu))

    # ------------------------------------------------------------------ helpers
    def _c(self, v: float) -> torch.Tensor:
        """python double -> working dtype (JAX weak-type rounding point)."""
        return torch.tensor(float(v), dtype=self.dtype)

    @staticmethod
    def _norm(x: torch.Tensor) -> torch.Tensor:
        """jnp.linalg.norm(axis=-1) = sqrt(sum of squares), left-to-right."""
        acc = x[..., 0] * x[..., 0]
        for k in range(1, x.shape[-1]):
            acc = acc + x[..., k] * x[..., k]
        return exact_sqrt(acc)

    @staticmethod
    def _matvec(x: torch.Tensor, M: torch.Tensor, zero_init: bool = False) -> torch.Tensor:
        """x @ M.T with an explicit left-to-right sum of separately rounded products
        (the reference's XLA dot has no specified order; this is the order the CUDA path uses)."""
        cols = []
        for r in range(M.shape[0]):
            acc = torch.zeros_like(x[:, 0]) if zero_init else None
            for c in range(M.shape[1]):
                t = x[:, c] * M[r, c]
                acc = t if acc is None else acc + t
            cols.append(acc)
        return torch.stack(cols, dim=-1)

    def state_lim(self):
        inf = float("inf")
        lim = {"SingleIntegrator": [inf, inf], "DoubleIntegrator": [inf, inf, 0.5, 0.5],
               "DubinsCar": [inf, inf, inf, 0.8], "LinearDrone": [inf, inf, inf, 0.5, 0.5, 0.5]}[self.env_id]
        up = torch.tensor(lim, dtype=self.dtype)
        return -up, up

    def action_lim(self):
        a = 3.0 if self.env_id == "DubinsCar" else 1.0          # dubins_car.py:317-326
        up = torch.ones(self.action_dim, dtype=self.dtype) * a
        return -up, up

    def clip_state(self, s):
        lo, up = self.state_lim()
        return torch.minimum(torch.maximum(s, lo), up)

    def clip_action(self, a):
        lo, up = self.action_lim()
        return torch.minimum(torch.maximum(a, lo), up)

    def edge_state(self, states: torch.Tensor) -> torch.Tensor:
        """Dubins: (x, y, v cos th, v sin th) (dubins_car.py:260-264); others identity."""
        if self.env_id != "DubinsCar":
            return states
        v = torch.stack([states[:, 3] * torch.cos(states[:, 2]), states[:, 3] * torch.sin(states[:, 2])], dim=-1)
        return torch.cat([states[:, :2], v], dim=-1)

    def _clip_pos(self, feats: torch.Tensor) -> torch.Tensor:
        """Norm clip of the position part (double_integrator.py:239-244 / 279-284)."""
        pd = self.pos_dim
        sq = feats[:, 0] * feats[:, 0]
        for k in range(1, pd):
            sq = sq + feats[:, k] * feats[:, k]
        feats_norm = exact_sqrt(1e-6 + sq)[:, None]
        cr = self._c(self.comm_radius)
        safe = torch.maximum(feats_norm, cr)
        coef = torch.where(feats_norm > cr, cr / safe, torch.ones_like(feats_norm))
        return torch.cat([feats[:, :pd] * coef, feats[:, pd:]], dim=-1)

    # ------------------------------------------------------------------ graph construction
    def lidar(self, agent: torch.Tensor, obstacles) -> torch.Tensor:
        """get_graph's LiDAR part (double_integrator.py:300-310): [N, R, sd], zero padded."""
        hits = get_lidar_all(agent[:, : self.pos_dim], obstacles, self.ray_table, self.n_hits)
        pad = self.state_dim - self.pos_dim
        if pad:
            hits = torch.cat([hits, torch.zeros(*hits.shape[:-1], pad, dtype=hits.dtype)], dim=-1)
        return hits

    def get_graph(self, agent: torch.Tensor, goal: torch.Tensor, obstacles,
                  lidar: Optional[torch.Tensor] = None) -> Graph:
        """double_integrator.py:223-264 + 288-320 (+ SI/Dubins/LD twins) + graph.py:209-244."""
        N, R, sd = self.num_agents, self.n_hits, self.state_dim
        if lidar is None:
            lidar = self.lidar(agent, obstacles)
        lidar_data = lidar.reshape(N * R, sd)
        n_real = 2 * N + N * R
        pad_id = n_real
        nodes = torch.zeros(n_real + 1, 3, dtype=self.dtype)
        nodes[:N, 2] = 1
        nodes[N:2 * N, 1] = 1
        nodes[2 * N:n_real, 0] = 1
        node_type = torch.full((n_real + 1,), -1, dtype=torch.int64)
        node_type[:N] = AGENT
        node_type[N:2 * N] = GOAL
        node_type[2 * N:n_real] = OBS
        states = torch.cat([agent, goal, lidar_data, -torch.ones(1, sd, dtype=self.dtype)], dim=0)

        pd = self.pos_dim
        es_agent = self.edge_state(agent)
        pos = agent[:, :pd]
        # agent - agent block
        pos_diff = pos[:, None, :] - pos[None, :, :]
        dist = self._norm(pos_diff) + torch.eye(N, dtype=self.dtype) * (self.comm_radius + 1)
        aa_feats = es_agent[:, None, :] - es_agent[None, :, :]
        aa_mask = dist < self._c(self.comm_radius)
        ids_agent = torch.arange(N)
        feats, recv, send = [], [], []

        def add_block(f, mask, ids_r, ids_s):
            f = f.reshape(-1, f.shape[-1])
            mask = mask.reshape(-1)
            rr = ids_r[:, None].expand(len(ids_r), len(ids_s)).reshape(-1)
            ss = ids_s[None, :].expand(len(ids_r), len(ids_s)).reshape(-1)
            feats.append(f)
            recv.append(torch.where(mask, rr, torch.full_like(rr, pad_id)))
            send.append(torch.where(mask, ss, torch.full_like(ss, pad_id)))

        add_block(aa_feats, aa_mask, ids_agent, ids_agent)
        # agent - goal block (features clipped)
        ids_goal = torch.arange(N, 2 * N)
        if self.env_id == "DubinsCar":                                   # dubins_car.py:212-227
            g_feats = torch.cat([agent[:, :2] - goal[:, :2], es_agent[:, 2:]], dim=-1)
            g_feats = self._clip_pos(g_feats)
            for i in range(N):
                add_block(g_feats[i][None, None, :], torch.ones(1, 1, dtype=torch.bool), ids_agent[i:i + 1],
                          ids_goal[i:i + 1])
        else:
            ag = agent[:, None, :] - goal[None, :, :]
            ag = self._clip_pos(ag.reshape(N * N, sd)).reshape(N, N, sd)
            add_block(ag, torch.eye(N, dtype=torch.bool), ids_agent, ids_goal)
        # agent - hit blocks
        ids_obs = torch.arange(2 * N, 2 * N + N * R)
        thr = self._c(self.comm_radius - 1e-1)
        for i in range(N):
            idh = slice(i * R, (i + 1) * R)
            lidar_pos = pos[i][None, :] - lidar_data[idh, :pd]
            lidar_feats = es_agent[i][None, :] - lidar_data[idh, :]
            active = self._norm(lidar_pos) < thr
            add_block(lidar_feats[None, :, :], active[None, :], ids_agent[i:i + 1], ids_obs[idh])
        return Graph(nodes, torch.cat(feats), states, torch.cat(recv), torch.cat(send), node_type,
                     agent, goal, obstacles, N, R)

    @staticmethod
    def sparsify(g: Graph) -> Graph:
        """Drop masked edges (recv == pad) -- the CUDA path's edge set."""
        pad_id = g.nodes.shape[0] - 1
        keep = g.receivers != pad_id
        return replace(g, edges=g.edges[keep], receivers=g.receivers[keep], senders=g.senders[keep])

    def add_edge_feats(self, g: Graph, states: torch.Tensor) -> Graph:
        """double_integrator.py:275-286 (Dubins :256-270, LD :265-276): recompute *all*
        edge features from `states` with the position norm-clip; topology unchanged.
        `states` has 2N+NR rows (no pad row): out-of-range gathers clamp like JAX."""
        es = self.edge_state(states)
        last = es.shape[0] - 1
        r = torch.clamp(g.receivers, max=last)
        s = torch.clamp(g.senders, max=last)
        feats = self._clip_pos(es[r] - es[s])
        st = states if states.shape[0] == g.states.shape[0] else torch.cat([states, g.states[-1:]], 0)
        return replace(g, edges=feats, states=st)

    # ------------------------------------------------------------------ control / dynamics
    def u_ref(self, agent: torch.Tensor, goal: torch.Tensor) -> torch.Tensor:
        if self.env_id == "DubinsCar":
            return self._u_ref_dubins(agent, goal)
        # double_integrator.py:332-338 (SI :297-303, LD :322-328)
        error = goal - agent
        nrm = self._norm(error)[:, None]
        error_max = torch.abs(error / nrm * self._c(self.comm_radius))
        error = torch.minimum(torch.maximum(error, -error_max), error_max)
        K = torch.tensor(np.asarray(self.K), dtype=self.dtype)
        return self.clip_action(self._matvec(error, K))

    def _u_ref_dubins(self, agent, goal):
        """dubins_car.py:328-379."""
        pi = math.pi
        pos_diff = agent[:, :2] - goal[:, :2]
        k_omega, k_v, k_a = 1.0, 2.3, 2.5
        dist = self._norm(pos_diff)
        theta_t = torch.remainder(torch.atan2(-pos_diff[:, 1], -pos_diff[:, 0]), 2 * pi)
        theta = torch.remainder(agent[:, 2], 2 * pi)
        theta_diff = theta_t - theta
        omega = torch.zeros(agent.shape[0], dtype=self.dtype)
        dot = (-pos_diff[:, 0]) * torch.cos(theta) + (-pos_diff[:, 1]) * torch.sin(theta)
        theta_between = torch.acos(torch.clamp(dot / (dist + 0.0001), -1, 1))
        c1 = (theta_diff < pi) & (theta_diff >= 0)
        omega = torch.where(c1 & (theta <= pi), k_omega * theta_between, omega)
        omega = torch.where((~c1) & (theta <= pi), -k_omega * theta_between, omega)
        c2 = (theta_diff > -pi) & (theta_diff <= 0)
        omega = torch.where(c2 & (theta > pi), -k_omega * theta_between, omega)
        omega = torch.where((~c2) & (theta > pi), k_omega * theta_between, omega)
        omega = torch.clamp(omega, -5.0, 5.0)
        nrm = exact_sqrt(1e-6 + (pos_diff[:, 0] * pos_diff[:, 0] + pos_diff[:, 1] * pos_diff[:, 1]))[:, None]
        cr = self._c(self.comm_radius)
        coef = torch.where(nrm > cr, cr / torch.maximum(nrm, cr), torch.ones_like(nrm))
        pd2 = coef * pos_diff
        a = -k_a * agent[:, 3] + k_v * self._norm(pd2)
        return torch.stack([omega, a], dim=-1)

    def stop_mask(self, agent, goal):
        """dubins_car.py:483-487."""
        return self._norm(agent[:, :2] - goal[:, :2]) < self._c(self.r * 0.5)

    def agent_xdot(self, agent, action):
        if self.env_id == "SingleIntegrator":                    # single_integrator.py:104-109
            return action
        if self.env_id == "DoubleIntegrator":                    # double_integrator.py:137-143
            return torch.cat([agent[:, 2:], action / self._c(self.params["m"])], dim=1)
        if self.env_id == "DubinsCar":                           # dubins_car.py:112-122
            return torch.stack([torch.cos(agent[:, 2]) * agent[:, 3], torch.sin(agent[:, 2]) * agent[:, 3],
                                action[:, 0] * 20.0, action[:, 1]], dim=1)
        A = torch.tensor(self._A, dtype=self.dtype)              # linear_drone.py:130-134
        B = torch.tensor(self._B, dtype=self.dtype)
        return self._matvec(agent, A, zero_init=True) + self._matvec(action, B, zero_init=True)

    def agent_step_euler(self, agent, action, goal=None):
        """double_integrator.py:128-135; Dubins :104-110 (with stop mask)."""
        x_dot = self.agent_xdot(agent, action)
        dt = self._c(self.dt)
        if self.env_id == "DubinsCar":
            stop = self.stop_mask(agent, goal).to(self.dtype)
            x_dot = x_dot * (1 - stop)[:, None]
            return self.clip_state(agent + x_dot * dt)
        if self.env_id == "LinearDrone":
            return self.clip_state(agent + x_dot * dt)
        return self.clip_state(x_dot * dt + agent)

    def forward_graph(self, g: Graph, action: torch.Tensor) -> Graph:
        """double_integrator.py:340-354: same topology, next agent states, all edge
        features recomputed with add_edge_feats (hits and goals frozen)."""
        N = self.num_agents
        agent = g.states[:N]
        goal = g.states[N:2 * N]
        obs_states = g.states[2 * N:2 * N + N * g.n_hits]
        nxt = self.agent_step_euler(agent, self.clip_action(action), goal)
        return self.add_edge_feats(g, torch.cat([nxt, goal, obs_states], dim=0))

    def get_cost(self, agent, obstacles):
        """double_integrator.py:183-198."""
        pos = agent[:, : self.pos_dim]
        N = self.num_agents
        dist = self._norm(pos[:, None, :] - pos[None, :, :]) + torch.eye(N, dtype=self.dtype) * 1e6
        collision = (self._c(self.r * 2) > dist).any(dim=1)
        cost = collision.to(self.dtype).mean()
        return cost + inside_obstacles(pos, obstacles, r=self.r).to(self.dtype).mean()

    def step(self, g: Graph, action: torch.Tensor):
        """double_integrator.py:145-181: returns (next graph, reward, cost)."""
        action = self.clip_action(action)
        nxt = self.agent_step_euler(g.agent, action, g.goal)
        diff = action - self.u_ref(g.agent, g.goal)
        reward = -((self._norm(diff) ** 2).mean())
        cost = self.get_cost(g.agent, g.obstacles)
        return self.get_graph(nxt, g.goal, g.obstacles), reward, cost

    # ------------------------------------------------------------------ masks
    def _agent_dist(self, pos, diag_add):
        return self._norm(pos[:, None, :] - pos[None, :, :]) + torch.eye(pos.shape[0], dtype=self.dtype) * diag_add

    def collision_mask(self, g: Graph):
        """double_integrator.py:419-434 (SI: = unsafe_mask, single_integrator.py:360-361)."""
        pos = g.agent[:, : self.pos_dim]
        dist = self._agent_dist(pos, self.r * 2 + 1)
        unsafe_agent = (dist < self._c(self.r * 2)).any(dim=1)
        return unsafe_agent | inside_obstacles(pos, g.obstacles, self.r)

    def finish_mask(self, g: Graph):
        """double_integrator.py:436-440."""
        pd = self.pos_dim
        return self._norm(g.agent[:, :pd] - g.goal[:, :pd]) < self._c(self.r * 2)

    def safe_mask(self, g: Graph):
        """double_integrator.py:356-374 (gcbf-v0 labels; kept for completeness)."""
        pos = g.agent[:, : self.pos_dim]
        k_a, k_o = (2.5, 1.5) if self.env_id == "SingleIntegrator" else (4, 2)
        dist = self._agent_dist(pos, self.r * 2 + 1)
        safe_agent = (dist > self._c(self.r * k_a)).all(dim=1)
        return safe_agent & ~inside_obstacles(pos, g.obstacles, self.r * k_o)

    def unsafe_mask(self, g: Graph):
        """SI single_integrator.py:343-358; DI double_integrator.py:376-417;
        Dubins dubins_car.py:421-464; LD linear_drone.py:367-383."""
        pd, r, N = self.pos_dim, self.r, self.num_agents
        pos = g.agent[:, :pd]
        if self.env_id == "SingleIntegrator":
            return self.collision_mask(g)
        if self.env_id == "LinearDrone":
            dist = self._agent_dist(pos, r * 2 + 1)
            return (dist < self._c(r * 2.5)).any(dim=1) | inside_obstacles(pos, g.obstacles, r * 1.5)
        agent_pos_diff = pos[None, :, :] - pos[:, None, :]
        agent_dist = self._norm(agent_pos_diff) + torch.eye(N, dtype=self.dtype) * (r * 2 + 1)
        unsafe_agent = (agent_dist < self._c(r * 2)).any(dim=1)
        r_obs = r if self.env_id == "DoubleIntegrator" else r * 1.5
        collision = unsafe_agent | inside_obstacles(pos, g.obstacles, r_obs)
        R = g.n_hits
        obs_pos = g.states[2 * N:2 * N + N * R, :pd]
        obs_pos_diff = obs_pos[None, :, :] - pos[:, None, :]
        obs_dist = self._norm(obs_pos_diff)
        pos_diff = torch.cat([agent_pos_diff, obs_pos_diff], dim=1)
        warn = torch.cat([agent_dist < self._c(3 * r), obs_dist < self._c(2 * r)], dim=1)
        pos_vec = pos_diff / (self._norm(pos_diff)[..., None] + 0.0001)
        if self.env_id == "DoubleIntegrator":
            vel = g.agent[:, 2:]
            speed = self._norm(vel)[:, None]
            heading = (vel / (speed + 0.0001))[:, None, :]
        else:
            heading = torch.stack([torch.cos(g.agent[:, 2]), torch.sin(g.agent[:, 2])], dim=1)[:, None, :]
        inner = (pos_vec * heading)[..., 0] + (pos_vec * heading)[..., 1]
        th_a = torch.atan2(self._c(r * 2).expand_as(agent_dist), exact_sqrt(agent_dist ** 2 - 4 * r ** 2))
        th_o = torch.atan2(self._c(r).expand_as(obs_dist), exact_sqrt(obs_dist ** 2 - r ** 2))
        theta = torch.cat([th_a, th_o], dim=1)
        lidar_mask = torch.block_diag(*[torch.ones(1, R)] * N).bool()
        valid = torch.cat([torch.ones(N, N, dtype=torch.bool), lidar_mask], dim=-1)
        warn = warn & valid
        unsafe_dir = (warn & (inner > torch.cos(theta))).any(dim=1)
        return collision | unsafe_dir
